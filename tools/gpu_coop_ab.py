"""Coop-kernel A/B of library builds (MPC_B200_LIB=...): single-solve latency (native loop), B = 1024 / 8192 coop batches, the 64K chain."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["stable"])); cd = cfg.as_dict()
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
name = os.path.basename(os.environ.get("MPC_B200_LIB", "default"))
b = mpc.workloads.batch_perturbed_states(65536, 0, cd)
S = mpc.Solver(cfg, 0)
p50, p99 = S.measure_solve_latency(b["state"][:1000], b["coeffs"][:1000], b["yaw_lo"][:1000], b["yaw_hi"][:1000], 1000, 200)
line = "%s solve_one native p50 %.1f us p99 %.1f us;" % (name, p50, p99)
for B, kind in ((1024, mpc.KERNEL_COOP), (8192, mpc.KERNEL_COOP), (65536, mpc.KERNEL_AUTO)):
    ins = [up(b["state"][:B]), up(b["coeffs"][:B]), up(b["yaw_lo"][:B]), up(b["yaw_hi"][:B])]
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    S.set_kernel(kind)
    ts = []
    for r in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, *ins, res, None, None, None, st, it); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    line += "  B=%d %.3f ms (csum %.6f)" % (B, np.median(ts[2:]), res.sum().item())
print(line, flush=True)
S.close()
