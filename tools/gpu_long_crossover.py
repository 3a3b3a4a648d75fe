"""N > 32: coop kernel vs lane kernel (with tail packing) by batch size -- where AUTO should switch."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
for N, dt in ((50, 0.02), (40, 0.05)):
    cfg = mpc.config_from_json_text(json.dumps(dict(js, N=N, dt=dt)))
    S = mpc.Solver(cfg, 0)
    for B in (1024, 2048, 4096, 8192, 16384, 32768):
        b = mpc.workloads.batch_perturbed_states(B, 5, cfg.as_dict())
        ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
        msg = []
        for kind, name in ((mpc.KERNEL_LANE, 'lane+tail'), (mpc.KERNEL_COOP, 'coop')):
            S.set_kernel(kind)
            res = torch.zeros(9, B, dtype=torch.float64, device=dev)
            st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
            best = 1e9
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); S.solve_batch_device(B, *ins, res, None, None, None, st, it); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            msg.append('%s %.2f ms' % (name, best))
        print('N=%d dt=%.2f B=%5d iters max %3d  %s' % (N, dt, B, it.max().item(), '  '.join(msg)), flush=True)
    S.close()
