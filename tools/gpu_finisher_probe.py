"""How much of the tail is the few longest problems?  64K batch with the iteration cap lowered: total time by cap."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = 65536
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
for cap in (3000, 40, 30, 24, 20, 16, 13):
    cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
    cfg.max_iter = cap
    b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    S = mpc.Solver(cfg, 0)
    for park, resume in ((16, 2), (16, 0), (0, 0)):
        S.set_tail(park, resume)
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_batch_device(B, *ins, res, None, None, None, st, it); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('max_iter=%4d park=%2d resume=%d  %.3f ms  iters mean %.2f max %d  hit cap %d' % (cap, park, resume, best, it.float().mean().item(), it.max().item(), (st == 2).sum().item()), flush=True)
    S.close()
