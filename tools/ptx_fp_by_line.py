"""Per source line, count the FP64 fma / mul / add / sub instructions of two kernels in a PTX file (built with
-lineinfo) and print the lines where they differ: that is where the compiler contracted differently.
usage: ptx_fp_by_line.py file.ptx entry_substring_A entry_substring_B"""
import re, sys, collections
ptx, ka, kb = sys.argv[1:4]
files = {}
def scan(key):
    out = collections.Counter()
    inside = False
    cur = None
    for ln in open(ptx):
        if ln.startswith('.visible .entry') or ln.startswith('.entry'):
            inside = key in ln
            continue
        m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', ln)
        if m: files[int(m.group(1))] = m.group(2).split('/')[-1]
        if not inside: continue
        m = re.match(r'\s*\.loc\s+(\d+)\s+(\d+)\s+\d+(.*)', ln)
        if m:
            # inlined_at chains: keep the innermost location
            cur = (int(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*(?:@%p\d+\s+)?(fma\.rn|mul|add|sub|neg|rcp\.approx\.ftz)\.f64', ln)
        if m and cur: out[(cur, m.group(1))] += 1
    return out
A, B = scan(ka), scan(kb)
keys = sorted(set(k[0] for k in A) | set(k[0] for k in B))
tot = collections.Counter()
for loc in keys:
    ca = {op: A.get((loc, op), 0) for op in ('fma.rn', 'mul', 'add', 'sub', 'neg')}
    cb = {op: B.get((loc, op), 0) for op in ('fma.rn', 'mul', 'add', 'sub', 'neg')}
    for op in ca: tot['A ' + op] += ca[op]; tot['B ' + op] += cb[op]
    if ca != cb:
        print(files.get(loc[0], loc[0]), loc[1], 'A', ca, 'B', cb)
print(dict(tot))
