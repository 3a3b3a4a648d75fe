"""Tail experiment: kernel time vs iteration cap and vs batch size (lane kernel)."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
def setup(B, max_iter):
    cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
    cfg.max_iter = max_iter
    S = mpc.Solver(cfg, 0)
    S.set_kernel(mpc.KERNEL_LANE)
    b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    N = cfg.N
    return S, B, [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])], [torch.zeros(9, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
def run(S, B, ins, outs, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for mi in (12, 16, 20, 30, 3000):
    S, B, ins, outs = setup(65536, mi)
    ms = run(S, B, ins, outs)
    it = outs[5].cpu().numpy()
    print('max_iter=%4d  B=65536  %.3f ms  %.0f solves/s  ok=%.4f  sum_iters=%d' % (mi, ms, B / ms * 1e3, (outs[4] == 1).float().mean().item(), it.sum()))
for B in (4096, 16384, 33152, 66304, 132608, 265216, 1048576):
    S, B, ins, outs = setup(B, 3000)
    ms = run(S, B, ins, outs, 3)
    print('B=%7d  %.3f ms  %.0f solves/s  %.1f ns/solve' % (B, ms, B / ms * 1e3, ms * 1e6 / B))
