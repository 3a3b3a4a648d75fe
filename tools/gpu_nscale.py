"""How much does L2 residency of the per-lane rows matter?  Same batch, horizons N = 4..10 (state per lane ~ N * 624 B)."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
B = 65536
for N in (4, 5, 6, 7, 8, 9, 10):
    cfg = mpc.config_from_json_text(json.dumps(dict(rd['configs']['stable'], N=N)))
    cd = cfg.as_dict()
    S = mpc.Solver(cfg, 0); S.set_kernel(mpc.KERNEL_LANE); S.set_handoff(0)
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), None, None, None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
    for mi in (3000, 12):
        cfg.max_iter = mi; S.set_config(cfg)
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        it = outs[5].cpu().numpy().astype(np.float64)
        trips = (it + 2).sum()
        print('N=%2d max_iter=%4d  %.3f ms  stage-trips %.3g  ns per stage-trip %.2f  resident state %.0f MB' % (N, mi, best, trips * N, best * 1e6 / (trips * N), 224 * 148 * N * 78 * 8 / 1e6))
    S.close()
