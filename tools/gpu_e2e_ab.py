"""A/B of mpc_solve_batch_host on the bench batch: copies back after the last launch (MPC_TAIL_LATE_COPY) against
copies beside the final launch + mpc_patch_outputs_kernel (default).  Pinned host buffers, host wall clock per call."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["stable"]))
b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
N = cfg.N
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).pin_memory().numpy()
ins = [pin(b[k]) for k in ("state", "coeffs", "yaw_lo", "yaw_hi")]
z = lambda shape, dt: torch.zeros(*shape, dtype=dt).pin_memory().numpy()
res, tx, ty, st, it = z((9, B), torch.float64), z((N, B), torch.float64), z((N, B), torch.float64), z((B,), torch.int32), z((B,), torch.int32)
S = mpc.Solver(cfg, 0)
L = mpc.lib()
p = lambda a: a.ctypes.data
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
def call():
    rc = L.mpc_solve_batch_host(S._h, B, p(ins[0]), p(ins[1]), p(ins[2]), p(ins[3]), None, None, None, p(res), p(tx), p(ty), None, p(st), p(it))
    assert rc == 0
for late in (True, False, True, False):
    S.set_tail(16, 3, resume_min=8192, late_copy=late)
    for _ in range(3):
        call()
    ts = []
    for k in range(30):
        flush.fill_(k & 0xFF); torch.cuda.synchronize()
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    ts.sort()
    print("late_copy=%d  B=%d  median %.3f ms  mean %.3f ms  min %.3f ms  -> %.2f M solves/s  parked by launch %s  csum=%.6f"
          % (late, B, 1e3 * ts[len(ts) // 2], 1e3 * sum(ts) / len(ts), 1e3 * ts[0], B / (sum(ts) / len(ts)) / 1e6, S.tail_counts(4), res[8].sum()))
