"""Migration of long runners from the lane kernel to the coop kernel: results must not change; timing vs the iteration threshold."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
cd = cfg.as_dict()
dev = torch.device('cuda:0')
S = mpc.Solver(cfg, 0)
for B in [int(x) for x in (sys.argv[1:] or ['65536', '1048576'])]:
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    N = cfg.N
    ref = None
    for hi in (0, 8, 10, 11, 12, 13, 14, 16, 20, 30):
        S.set_handoff(hi)
        outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res = (outs[0].cpu().numpy(), outs[1].cpu().numpy(), outs[4].cpu().numpy(), outs[5].cpu().numpy())
        if ref is None:
            ref = res
        same = all(np.array_equal(a, b_) for a, b_ in zip(ref, res))
        print('B=%7d handoff_iter=%2d  %.3f ms  %.0f solves/s  identical to no-handoff: %s  ok=%.4f' % (B, hi, best, B / best * 1e3, same, (res[2] == 1).mean()))
