"""Solo kernel check: N = 50 problems, solo vs lane kernel (bit-identical), timing at small batch sizes and B = 1."""
import json, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
for N, dt in ((50, 0.02), (40, 0.05), (25, 0.05), (10, 0.1)):
    cfg = mpc.config_from_json_text(json.dumps(dict(js, N=N, dt=dt)))
    S = mpc.Solver(cfg, 0)
    for B in (1, 64, 512, 2048):
        b = mpc.workloads.batch_perturbed_states(B, 5, cfg.as_dict())
        ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
        out = {}
        for kind, name in ((mpc.KERNEL_LANE, 'lane'), (mpc.KERNEL_SOLO, 'solo'), (mpc.KERNEL_COOP, 'coop')):
            if kind == mpc.KERNEL_COOP and N > 32: continue
            S.set_kernel(kind); S.set_tail(0, 0)
            res = torch.zeros(9, B, dtype=torch.float64, device=dev)
            st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); S.solve_batch_device(B, *ins, res, None, None, None, st, it); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = (best, res.clone(), st.clone(), it.clone())
        ref = out['lane']
        msg = ' '.join('%s %.3f ms%s' % (k, v[0], '' if all(torch.equal(a, b) for a, b in zip(v[1:], ref[1:])) else ' (DIFFERS %.2e)' % (v[1] - ref[1]).abs().max().item()) for k, v in out.items())
        print('N=%d dt=%.2f B=%4d iters max %3d  %s' % (N, dt, B, ref[3].max().item(), msg), flush=True)
    S.close()
