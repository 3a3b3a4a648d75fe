"""Small run of every kernel for compute-sanitizer (memcheck / racecheck): coop, lane + migration, warp, run_batch, rollout."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["fast"]))
cd = cfg.as_dict()
S = mpc.Solver(cfg, 0)
b = mpc.workloads.batch_perturbed_states(600, 5, cd)
args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
out = {}
for kind, nm in ((mpc.KERNEL_COOP, "coop"), (mpc.KERNEL_LANE, "lane"), (mpc.KERNEL_WARP, "warp")):
    S.set_kernel(kind)
    out[nm] = S.solve_batch_host(*args)
    print(nm, "ok", (out[nm]["status"] == 1).mean(), "iters max", out[nm]["iters"].max())
assert np.array_equal(out["coop"]["result"], out["lane"]["result"])
# migration path needs B >= LANE_MIN_BATCH
S.set_kernel(mpc.KERNEL_AUTO)
S.set_handoff(9)
bb = mpc.workloads.batch_perturbed_states(mpc.LANE_MIN_BATCH, 6, cd)
r = S.solve_batch_host(bb["state"], bb["coeffs"], bb["yaw_lo"], bb["yaw_hi"])
print("lane+migration ok", (r["status"] == 1).mean(), "launches", S.launches)
# rollout (coop kernel inside) with per-step records
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
V, T = 64, 5
veh = up(np.stack([b["px"][:V], b["py"][:V], b["psi"][:V], np.clip(b["v"][:V], 8, 30), np.zeros(V), np.zeros(V)]))
seg = up(b["segment"][:V].astype(np.int32))
pend = torch.zeros(2, V, dtype=torch.float64, device=dev)
rec = torch.zeros(T, 8, V, dtype=torch.float64, device=dev)
S.rollout_device(V, T, up(np.array(rd["waypoints"]["x"])), up(np.array(rd["waypoints"]["y"])), veh, seg, pend, 0.1, 0.02, rec)
torch.cuda.synchronize()
print("rollout ok", (rec[:, 6] == 1).float().mean().item())
S.close()
