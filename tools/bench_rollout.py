"""BASELINE config 5: closed-loop rollouts (config-fast, 100 ms latency), V vehicles x T steps on one GPU; also the
warp/lane kernel crossover at rollout-sized batches."""
import json, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
up = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['fast']))
cd = cfg.as_dict()
wx, wy = up(np.array(rd['waypoints']['x'])), up(np.array(rd['waypoints']['y']))
S = mpc.Solver(cfg, 0)
if 'crossover' in sys.argv:
    for B in (1, 256, 1024, 2048, 4096, 8192, 12288, 16384, 32768):
        b = mpc.workloads.batch_perturbed_states(B, 0, cd)
        ins = [up(b['state'].T), up(b['coeffs'].T), up(b['yaw_lo']), up(b['yaw_hi'])]
        outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), None, None, None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
        line = 'B=%6d ' % B
        for kind, nm in ((mpc.KERNEL_LANE, 'lane'), (mpc.KERNEL_COOP, 'coop')):
            S.set_kernel(kind)
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            line += ' %s %.3f ms' % (nm, best)
        print(line + '  iters max %d' % outs[5].max().item())
    S.set_kernel(mpc.KERNEL_AUTO)
for V, T in ((1024, 200), (8192, 200)):
    b = mpc.workloads.batch_perturbed_states(V, 3, cd)
    veh = up(np.stack([b['px'], b['py'], b['psi'], np.clip(b['v'], 8, 30), np.zeros(V), np.zeros(V)]))
    seg = up(b['segment'].astype(np.int32))
    pending = torch.zeros(2, V, dtype=torch.float64, device=dev)
    rec = torch.zeros(T, 8, V, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S.rollout_device(V, T, wx, wy, veh, seg, pending, 0.1, 0.02, rec); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    r = rec.cpu().numpy()
    ok = (r[:, 6] == 1).mean()
    print('rollout V=%d T=%d: %.1f ms  %.3f ms/step  %.0f vehicle-steps/s  status ok %.4f  mean iters %.2f  |cte| median %.3f  p90 %.3f  final speed median %.1f m/s'
          % (V, T, ms, ms / T, V * T / ms * 1e3, ok, r[:, 7].mean(), np.median(np.abs(r[-1, 0])), np.percentile(np.abs(r[-1, 0]), 90), np.median(r[-1, 2])))
