"""Join an ncu SASS source page (csv) with nvdisasm line info: instructions executed and stall
samples per CUDA source line / per function.  usage: ncu_by_line.py prof.ncu-rep lib.so [kernel-regex]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s > /dev/null" % (tmp, os.path.abspath(lib)), shell=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}
cur = None
insec = False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        insec = "mpc_ipm_kernel" in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[hi]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
per_line = collections.Counter(); samp_line = collections.Counter(); n_static = collections.Counter()
tot = ts = 0
for r in rows[hi + 1:]:
    if len(r) <= ii: continue
    a = int(r[ia], 16)
    if base is None: base = a
    off = a - base
    key = line_of.get(off, (None, ""))[0]
    n = int(r[ii] or 0); s = int(r[isamp] or 0)
    per_line[key] += n; samp_line[key] += s; n_static[key] += 1
    tot += n; ts += s
srcfile = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", "mpc_kernel.cuh")
text = open(srcfile).read().splitlines() if os.path.exists(srcfile) else []
print("total warp-instructions %d, samples %d, static SASS %d" % (tot, ts, sum(n_static.values())))
print("---- top lines by executed instructions")
for key, n in per_line.most_common(45):
    t = text[key[1] - 1].strip()[:90] if key and key[0] == "mpc_kernel.cuh" and key[1] <= len(text) else ""
    print("%5.2f%% inst %5.2f%% stall  static %4d  %s  %s" % (100.0 * n / tot, 100.0 * samp_line[key] / max(ts, 1), n_static[key], key, t))
# by function: map line -> enclosing "__device__" function name
fn_of_line = {}
curfn = "?"
for i, l in enumerate(text, 1):
    m = re.search(r"__device__[^;(]*?(\w+)\s*\(", l)
    if m and "{" in l or (m and not l.strip().endswith(";")):
        curfn = m.group(1)
    fn_of_line[i] = curfn
per_fn = collections.Counter(); samp_fn = collections.Counter(); stat_fn = collections.Counter()
for key, n in per_line.items():
    f = fn_of_line.get(key[1], "?") if key and key[0] == "mpc_kernel.cuh" else str(key[0] if key else None)
    per_fn[f] += n; samp_fn[f] += samp_line[key]; stat_fn[f] += n_static[key]
print("---- by function")
for f, n in per_fn.most_common(30):
    print("%5.2f%% inst %5.2f%% stall  static %5d  %s" % (100.0 * n / tot, 100.0 * samp_fn[f] / max(ts, 1), stat_fn[f], f))
