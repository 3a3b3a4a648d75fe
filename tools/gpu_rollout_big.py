"""Closed loop with more vehicles than the coop/lane crossover: the lane-kernel chain inside mpc_rollout must give the
same trajectories as the coop kernel forced."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['fast']))
cd = cfg.as_dict()
wx, wy = up(np.array(rd['waypoints']['x'])), up(np.array(rd['waypoints']['y']))
V, T = 16384, 30
b = mpc.workloads.batch_perturbed_states(V, 3, cd)
out = {}
for kind, name in ((mpc.KERNEL_COOP, 'coop'), (mpc.KERNEL_AUTO, 'auto')):
    S = mpc.Solver(cfg, 0)
    S.set_kernel(kind)
    veh = up(np.stack([b['px'], b['py'], b['psi'], np.clip(b['v'], 8, 30), np.zeros(V), np.zeros(V)]))
    seg = up(b['segment'].astype(np.int32))
    pending = torch.zeros(2, V, dtype=torch.float64, device=dev)
    rec = torch.zeros(T, 8, V, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = S.launches
    e0.record(); S.rollout_device(V, T, wx, wy, veh, seg, pending, 0.1, 0.02, rec); e1.record(); torch.cuda.synchronize()
    print('%s: %.2f ms per control step, %d launches per step, status ok %.4f' % (name, e0.elapsed_time(e1) / T, (S.launches - l0) // T, (rec[:, 6] == 1).float().mean().item()))
    out[name] = rec.clone()
    S.close()
print('identical', torch.equal(out['coop'], out['auto']))
