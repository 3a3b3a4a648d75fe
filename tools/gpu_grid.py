import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
cd = cfg.as_dict()
dev = torch.device('cuda:0')
S = mpc.Solver(cfg, 0)
B = 65536
b = mpc.workloads.batch_perturbed_states(B, 0, cd)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
N = cfg.N
outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for thr in (160, 192, 224, 256):
    for hi in (11, 12, 13, 14):
        S.set_kernel(mpc.KERNEL_LANE, thr, 1); S.set_handoff(hi)
        ts = []
        for k in range(8):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('threads=%d handoff=%d  min %.3f  median %.3f ms' % (thr, hi, min(ts), float(np.median(ts))))
