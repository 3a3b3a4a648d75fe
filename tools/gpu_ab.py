"""A/B of library builds (MPC_B200_LIB=...): configs[1] workload, median ms per batch (L2 flushed between launches) and a
checksum of the results."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["stable"]))
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in [int(x) for x in sys.argv[1:]] or [65536, 1 << 20]:
    b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
    ins = [up(b["state"]), up(b["coeffs"]), up(b["yaw_lo"]), up(b["yaw_hi"])]
    N = cfg.N
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    tx = torch.zeros(N, B, dtype=torch.float64, device=dev); ty = torch.zeros(N, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    S = mpc.Solver(cfg, 0)
    ts = []
    for r in range(9 if B <= 65536 else 5):
        flush.fill_(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, *ins, res, tx, ty, None, st, it); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%s B=%d median %.3f ms min %.3f  ok=%.4f iters=%.3f csum=%.6f parked %s" % (os.path.basename(os.environ.get("MPC_B200_LIB", "default")), B, np.median(ts[2:]), min(ts),
          (st == 1).float().mean().item(), it.float().mean().item(), res.sum().item(), S.tail_counts(4)), flush=True)
    S.close()
