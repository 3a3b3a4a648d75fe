"""GPU check of the lane kernel: parity vs the CPU oracle and vs the warp kernel, then a timing sweep."""
import json, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
from oracle import pyoracle as po
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
cd = cfg.as_dict()
ocfg = po.make_config(po.load_config_dict(rd['configs']['stable']))
S = mpc.Solver(cfg, 0)
B = 4096
b = mpc.workloads.batch_perturbed_states(B, 0, cd)
S.set_kernel(mpc.KERNEL_LANE)
g = S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
S.set_kernel(mpc.KERNEL_WARP)
w = S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
probs = po.problems_from_arrays(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
c = po.solve_batch(ocfg, probs, 16)
for name, r in (('lane', g), ('warp', w)):
    print(name, 'status', np.bincount(r['status']), 'cpu', np.bincount(c['status']))
    print(name, 'iters equal frac', (r['iters'] == c['iters']).mean(), 'mean', r['iters'].mean(), 'cpu mean', c['iters'].mean())
    d = np.abs(r['result'] - c['result']); rel = d[:, 8] / np.abs(c['result'][:, 8])
    print(name, 'max abs diff first 8', d[:, :8].max(), 'max rel cost', rel.max(), 'traj', np.abs(r['traj_x'] - c['traj_x']).max(), np.abs(r['traj_y'] - c['traj_y']).max())
    bad = np.nonzero((d[:, :8].max(axis=1) > 1e-6) | (rel > 1e-8))[0]
    print(name, 'n bad', bad.size, bad[:10])
    for i in bad[:3]:
        print(i, r['status'][i], c['status'][i], r['iters'][i], c['iters'][i], r['result'][i], c['result'][i])
print('lane vs warp max diff', np.abs(g['result'] - w['result']).max(), 'iters equal', (g['iters'] == w['iters']).mean())

# ---- timing sweep on the bench workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = mpc.workloads.batch_perturbed_states(B, 0, cd)
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
state, coeffs, ylo, yhi = up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])
N = cfg.N
result = torch.zeros(9, B, dtype=torch.float64, device=dev)
tx = torch.zeros(N, B, dtype=torch.float64, device=dev); ty = torch.zeros(N, B, dtype=torch.float64, device=dev)
status = torch.zeros(B, dtype=torch.int32, device=dev); iters = torch.zeros(B, dtype=torch.int32, device=dev)
def run(reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, state, coeffs, ylo, yhi, result, tx, ty, None, status, iters); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for thr, ctas in ((0, 0), (256, 1), (224, 1), (192, 1), (160, 1), (128, 1), (128, 2), (96, 2), (64, 3), (32, 7)):
    S.set_kernel(mpc.KERNEL_LANE, thr, ctas)
    ms = run()
    print('lane threads=%d ctas/sm=%d  B=%d  %.3f ms  %.0f solves/s  ok=%.4f iters=%.2f' % (thr, ctas, B, ms, B / ms * 1e3, (status == 1).float().mean().item(), iters.float().mean().item()))
S.set_kernel(mpc.KERNEL_WARP)
ms = run(2)
print('warp B=%d  %.3f ms  %.0f solves/s' % (B, ms, B / ms * 1e3))
