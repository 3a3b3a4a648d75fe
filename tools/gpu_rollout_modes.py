"""Closed loop (config-fast, 100 ms latency): one persistent launch against three launches per control step, by fleet size.
Decides MPC_ROLLOUT_PERSISTENT_MAX."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['fast']))
cd = cfg.as_dict()
wx, wy = up(np.array(rd['waypoints']['x'])), up(np.array(rd['waypoints']['y']))
S = mpc.Solver(cfg, 0)
for V in (256, 1024, 2048, 2368, 3072, 4096, 4736, 8192, 16384):
    b = mpc.workloads.batch_perturbed_states(V, 3, cd)
    veh0 = np.stack([b['px'], b['py'], b['psi'], np.clip(b['v'], 8, 30), np.zeros(V), np.zeros(V)])
    seg0 = b['segment'].astype(np.int32)
    line = 'V=%5d T=%d ' % (V, T)
    recs = {}
    for mode, nm in ((1, 'per-step'), (2, 'persistent')):
        S.set_rollout_mode(mode)
        best = 1e9
        for rep in range(2):
            veh, seg = up(veh0), up(seg0)
            pending = torch.zeros(2, V, dtype=torch.float64, device=dev)
            rec = torch.zeros(T, 8, V, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.rollout_device(V, T, wx, wy, veh, seg, pending, 0.1, 0.02, rec); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        recs[mode] = rec.cpu().numpy()
        line += ' %s %.4f ms/step' % (nm, best / T)
    print(line + '  same bits %s' % np.array_equal(recs[1], recs[2]), flush=True)
