"""Compare the FP64 expression trees of two kernels in a PTX file, per source line: every arithmetic instruction
becomes a string built from its opcode and, recursively (depth-limited), the defining instructions of its
operands; loads are leaves labelled by their immediate offset.  Lines whose multisets differ are printed.
usage: ptx_expr_shapes.py file.ptx entry_substring_A entry_substring_B [depth]"""
import re, sys, collections
ptx, ka, kb = sys.argv[1:4]
DEPTH = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ARITH = ('fma.rn.f64', 'mul.f64', 'add.f64', 'sub.f64', 'neg.f64', 'abs.f64', 'rcp.approx.ftz.f64', 'mul.rn.f64', 'add.rn.f64', 'sub.rn.f64')
def scan(key):
    inside = False; cur = None; files = {}
    defs = {}          # reg -> list of (op, operands)
    insts = []         # (line, op, dst, operands)
    for ln in open(ptx):
        if ln.startswith('.visible .entry') or ln.startswith('.entry'):
            inside = key in ln; continue
        m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', ln)
        if m: files[int(m.group(1))] = m.group(2).split('/')[-1]
        if not inside: continue
        m = re.match(r'\s*\.loc\s+(\d+)\s+(\d+)\s+\d+', ln)
        if m: cur = (int(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*(?:@!?%p\d+\s+)?([a-z0-9_.]+)\s+(%fd\d+),\s*(.*);', ln)
        if not m: continue
        op, dst, rest = m.groups()
        ops = [o.strip() for o in rest.split(',')]
        if op.startswith('ld.'):
            mm = re.search(r'\[([^\]]+)\]', rest)
            off = re.search(r'\+(-?\d+)', mm.group(1)) if mm else None
            lab = 'x'
            defs.setdefault(dst, []).append((lab, []))
        else:
            defs.setdefault(dst, []).append((op, ops))
            if op in ARITH: insts.append((cur, op, dst, ops))
    def shape(reg, d):
        if not reg.startswith('%fd'): return 'c' if reg.startswith('0d') else 'x'
        dl = defs.get(reg)
        if not dl: return 'x'
        if len(dl) > 1: return 'x'
        op, ops = dl[0]
        if not ops: return 'x'
        if op not in ARITH or d == 0: return 'v' if op in ARITH else 'x'
        return '%s(%s)' % (op.split('.')[0], ','.join(shape(o, d - 1) for o in ops))
    out = collections.Counter()
    for cur, op, dst, ops in insts:
        out[(files.get(cur[0]) if cur else None, cur[1] if cur else 0, '%s(%s)' % (op.split('.')[0], ','.join(shape(o, DEPTH) for o in ops)))] += 1
    return out
A, B = scan(ka), scan(kb)
lines = sorted(set((k[0], k[1]) for k in A) | set((k[0], k[1]) for k in B), key=lambda t: (str(t[0]), t[1]))
for f, l in lines:
    a = {k[2]: n for k, n in A.items() if k[0] == f and k[1] == l}
    b = {k[2]: n for k, n in B.items() if k[0] == f and k[1] == l}
    if a != b:
        print('---- %s:%d' % (f, l))
        for s in sorted(set(a) | set(b)):
            if a.get(s, 0) != b.get(s, 0): print('   A %d  B %d   %s' % (a.get(s, 0), b.get(s, 0), s))
