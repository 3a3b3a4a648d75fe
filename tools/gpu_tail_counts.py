"""Records parked by every launch of the chain, and total time, for a few tail settings on the 64K batch."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = 65536
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
res = torch.zeros(9, B, dtype=torch.float64, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
S = mpc.Solver(cfg, 0)
for park, resume in ((16, 2), (16, 1), (16, 0), (8, 0), (8, 1), (24, 0), (24, 1), (24, 2), (31, 1), (31, 2), (12, 0)):
    S.set_tail(park, resume)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, *ins, res, None, None, None, st, it); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('park=%2d resume=%d  %.3f ms  parked per launch %s' % (park, resume, best, S.tail_counts(resume + 1)), flush=True)
