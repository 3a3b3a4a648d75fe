"""Config 3 (N x dt grid) diagnosis: uniform-(N, dt) batches one by one, then the ragged mix, lane kernel."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 18724
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
PAIRS = [(10, .1), (20, .1), (30, .1), (40, .1), (10, .05), (20, .05), (30, .05), (40, .05), (50, .05), (10, .02), (20, .02), (30, .02), (40, .02), (50, .02)]
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
cfg0 = mpc.config_from_json_text(json.dumps(js))
b = mpc.workloads.batch_perturbed_states(B, 0, cfg0.as_dict())
ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
res = torch.zeros(9, B, dtype=torch.float64, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
tot = 0.0
for N, dt in PAIRS:
    cfg = mpc.config_from_json_text(json.dumps(dict(js, N=N, dt=dt)))
    S = mpc.Solver(cfg, 0)
    ms = timed(lambda: S.solve_batch_device(B, *ins, res, None, None, None, st, it))
    tot += ms
    i = it.cpu().numpy(); s = st.cpu().numpy()
    print('N=%2d dt=%.2f B=%d  %.2f ms  ok=%.4f  iters mean %.1f p50 %d p99 %d max %d  status %s  ns/stage-iter %.3f' % (
        N, dt, B, ms, (s == 1).mean(), i.mean(), np.percentile(i, 50), np.percentile(i, 99), i.max(), np.bincount(s).tolist(), ms * 1e6 / (N * i.sum())))
    S.close()
print('sum of uniform launches %.1f ms for %d problems' % (tot, B * len(PAIRS)))
