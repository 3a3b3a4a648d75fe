"""Lines of one kernel where ptxas fused more multiply-adds than the compiler front end did (SASS DFMA > PTX fma).
usage: fma_ptx_vs_sass.py file.ptx all.sass kernel_substring"""
import re, sys, collections
ptx, sass, key = sys.argv[1:4]
P = collections.Counter(); S = collections.Counter()
inside = False; cur = None; files = {}
for ln in open(ptx):
    if ln.startswith('.visible .entry') or ln.startswith('.entry'):
        inside = key in ln; continue
    m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', ln)
    if m: files[int(m.group(1))] = m.group(2).split('/')[-1]
    if not inside: continue
    m = re.match(r'\s*\.loc\s+(\d+)\s+(\d+)\s+\d+', ln)
    if m: cur = (int(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*(?:@!?%p\d+\s+)?(fma\.rn|mul|add|sub)\.f64', ln)
    if m and cur: P[(cur, m.group(1))] += 1
P2 = collections.Counter()
for (loc, op), n in P.items(): P2[((files.get(loc[0]), loc[1]), op)] += n
inside = False; cur = None
for ln in open(sass):
    if ln.startswith('//---') and '.text.' in ln:
        inside = key in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(DFMA|DMUL|DADD)\b', ln)
    if m and cur: S[(cur, m.group(1))] += 1
lines = sorted(set(k[0] for k in P2) | set(k[0] for k in S), key=lambda t: (str(t[0]), t[1]))
tot = 0
for loc in lines:
    pf, pm, pa = P2.get((loc, 'fma.rn'), 0), P2.get((loc, 'mul'), 0), P2.get((loc, 'add'), 0) + P2.get((loc, 'sub'), 0)
    sf, sm, sa = S.get((loc, 'DFMA'), 0), S.get((loc, 'DMUL'), 0), S.get((loc, 'DADD'), 0)
    if sf != pf:
        tot += sf - pf
        print('%s:%d  PTX fma/mul/add %d/%d/%d   SASS %d/%d/%d' % (loc[0], loc[1], pf, pm, pa, sf, sm, sa))
print('extra fused by ptxas:', tot)
