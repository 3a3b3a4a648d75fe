"""Join an ncu SASS source page (csv) with nvdisasm line info: instructions executed, lane utilisation and
stall samples per CUDA source line / per device function.
usage: ncu_by_line.py prof.ncu-rep lib.so kernel-substring [source.cuh] [n-lines]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
srcname = sys.argv[4] if len(sys.argv) > 4 else "mpc_kernel.cuh"
ntop = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s > /dev/null" % (tmp, os.path.abspath(lib)), shell=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}
cur = None
insec = False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        insec = kname in ln
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
his = [i for i, r in enumerate(rows) if "Address" in r]
hi = his[0]                                   # first launch of the report
hdr = rows[hi]
if len(his) > 1: rows = rows[:his[1]]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ith = hdr.index("Thread Instructions Executed")
ilsb = hdr.index("stall_long_sb")
base = None
C = collections.Counter
per_line, samp_line, n_static, thr_line, lsb_line = C(), C(), C(), C(), C()
op_count = C()
tot = ts = tthr = 0
for r in rows[hi + 1:]:
    if len(r) <= ii: continue
    a = int(r[ia], 16)
    if base is None: base = a
    off = a - base
    key, text = line_of.get(off, (None, ""))
    n = int(r[ii] or 0); s = int(r[isamp] or 0); t = int(r[ith] or 0)
    per_line[key] += n; samp_line[key] += s; n_static[key] += 1; thr_line[key] += t; lsb_line[key] += int(r[ilsb] or 0)
    op = r[hdr.index("Source")].split()
    op = [x for x in op if not x.startswith("@")]
    op_count[op[0].split(".")[0] if op else "?"] += n
    tot += n; ts += s; tthr += t
srcfile = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", srcname)
text = open(srcfile).read().splitlines() if os.path.exists(srcfile) else []
print("total warp-instructions %d, thread-instructions %d (%.1f lanes/inst), samples %d, static SASS %d" % (tot, tthr, tthr / max(tot, 1), ts, sum(n_static.values())))
print("---- opcode mix (warp instructions)")
print("  ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in op_count.most_common(24)))
print("---- top lines by stall samples")
for key, s in samp_line.most_common(ntop):
    n = per_line[key]
    t = text[key[1] - 1].strip()[:90] if key and key[0] == srcname and key[1] <= len(text) else ""
    print("%5.2f%% inst %5.2f%% stall (%4.1f%% lsb) lanes %4.1f static %4d  %s  %s" % (100.0 * n / tot, 100.0 * s / max(ts, 1), 100.0 * lsb_line[key] / max(ts, 1), thr_line[key] / max(n, 1), n_static[key], key, t))
fn_of_line = {}
curfn = "?"
for i, l in enumerate(text, 1):
    m = re.search(r"__device__[^;(]*?(\w+)\s*\(", l) or re.search(r"__global__.*?(\w+)\s*\(", l)
    if m and not l.strip().endswith(";"):
        curfn = m.group(1)
    fn_of_line[i] = curfn
per_fn, samp_fn, stat_fn, thr_fn, lsb_fn = C(), C(), C(), C(), C()
for key, n in per_line.items():
    f = fn_of_line.get(key[1], "?") if key and key[0] == srcname else str(key[0] if key else None)
    per_fn[f] += n; samp_fn[f] += samp_line[key]; stat_fn[f] += n_static[key]; thr_fn[f] += thr_line[key]; lsb_fn[f] += lsb_line[key]
print("---- by function")
for f, n in per_fn.most_common(30):
    print("%5.2f%% inst %5.2f%% stall (%4.1f%% lsb) lanes %4.1f  static %5d  %s" % (100.0 * n / tot, 100.0 * samp_fn[f] / max(ts, 1), 100.0 * lsb_fn[f] / max(ts, 1), thr_fn[f] / max(n, 1), stat_fn[f], f))
