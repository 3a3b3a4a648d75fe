"""Per source line, count DFMA / DMUL / DADD in the SASS of two kernels (nvdisasm -g -c output) and print the lines
where they differ: where ptxas fused multiply-adds differently.
usage: sass_fp_by_line.py all.sass kernel_substring_A kernel_substring_B"""
import re, sys, collections
f, ka, kb = sys.argv[1:4]
def scan(key):
    out = collections.Counter()
    inside = False
    cur = None
    for ln in open(f):
        if ln.startswith('//---') and '.text.' in ln:
            inside = key in ln
            continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(DFMA|DMUL|DADD)\b', ln)
        if m and cur: out[(cur, m.group(1))] += 1
    return out
A, B = scan(ka), scan(kb)
keys = sorted(set(k[0] for k in A) | set(k[0] for k in B))
tot = collections.Counter()
for loc in keys:
    ca = tuple(A.get((loc, op), 0) for op in ('DFMA', 'DMUL', 'DADD'))
    cb = tuple(B.get((loc, op), 0) for op in ('DFMA', 'DMUL', 'DADD'))
    for i, op in enumerate(('DFMA', 'DMUL', 'DADD')): tot['A ' + op] += ca[i]; tot['B ' + op] += cb[i]
    if ca != cb: print(loc[0], loc[1], 'A fma/mul/add', ca, 'B', cb)
print(dict(tot))
