"""Where do the solo kernel and the lane kernel part ways?  B = 1, N = 50, max_iter = 0, 1, 2, ...: primal and dual outputs."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = 4
for m in range(0, 5):
    cfg = mpc.config_from_json_text(json.dumps(dict(js, N=N, dt=0.02)))
    cfg.max_iter = m
    S = mpc.Solver(cfg, 0)
    b = mpc.workloads.batch_perturbed_states(B, 5, cfg.as_dict())
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    out = {}
    for kind, name in ((mpc.KERNEL_LANE, 'lane'), (mpc.KERNEL_SOLO, 'solo')):
        S.set_kernel(kind); S.set_tail(0, 0)
        res = torch.zeros(9, B, dtype=torch.float64, device=dev)
        full = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev)
        lam = torch.zeros(6 * N, B, dtype=torch.float64, device=dev)
        zl = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev); zu = torch.zeros_like(zl)
        st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
        S.set_dual_outputs(lam, zl, zu)
        S.solve_batch_device(B, *ins, res, None, None, full, st, it); torch.cuda.synchronize()
        out[name] = dict(res=res.cpu().numpy(), full=full.cpu().numpy(), lam=lam.cpu().numpy(), zl=zl.cpu().numpy(), zu=zu.cpu().numpy(), it=it.cpu().numpy())
    S.close()
    msg = []
    for k in ('res', 'full', 'lam', 'zl', 'zu'):
        d = np.abs(out['lane'][k] - out['solo'][k])
        nz = np.argwhere(d > 0)
        msg.append('%s max %.2e n=%d first %s' % (k, d.max(), len(nz), nz[0].tolist() if len(nz) else '-'))
    print('max_iter=%d iters %s | %s' % (m, out['lane']['it'].tolist(), ' | '.join(msg)), flush=True)
