"""ncu driver: the coop kernel on a rollout-sized batch."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["stable"]))
b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
ins = [up(b["state"]), up(b["coeffs"]), up(b["yaw_lo"]), up(b["yaw_hi"])]
N = cfg.N
outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
S = mpc.Solver(cfg, 0)
S.set_kernel(mpc.KERNEL_COOP)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
    print("coop B=%d %.3f ms iters max %d mean %.2f" % (B, e0.elapsed_time(e1), outs[5].max().item(), outs[5].float().mean().item()))
