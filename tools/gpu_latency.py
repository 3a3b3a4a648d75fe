import json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable'])); cd = cfg.as_dict()
b = mpc.workloads.batch_perturbed_states(2048, 0, cd)
S = mpc.Solver(cfg, 0)
ref = S.solve_batch_host(b['state'][:64], b['coeffs'][:64], b['yaw_lo'][:64], b['yaw_hi'][:64])
lat = []
for k in range(1500):
    i = k % 2048
    t0 = time.perf_counter()
    r = S.solve_one(b['state'][i], b['coeffs'][i], b['yaw_lo'][i], b['yaw_hi'][i])
    lat.append(time.perf_counter() - t0)
    if i < 64:
        assert np.array_equal(r['result'], ref['result'][i]) and r['iters'] == ref['iters'][i] and np.array_equal(r['traj_x'], ref['traj_x'][i])
lat = np.array(lat[300:]) * 1e6
print('solve_one p50 %.1f us  p99 %.1f us  min %.1f us' % (np.percentile(lat, 50), np.percentile(lat, 99), lat.min()))
