import json, sys, time, numpy as np
sys.path.insert(0, '.')
import mpc_b200 as mpc
from oracle import pyoracle as po
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
cd = cfg.as_dict()
ocfg = po.make_config(po.load_config_dict(rd['configs']['stable']))
S = mpc.Solver(cfg, 0)
# test.cpp scenario
fx = rd['test_cpp_fixtures'][0]
state, coeffs, ylo, yhi, ex = po.preprocess(po.load_config_dict(rd['configs']['stable']), (fx['x'], fx['y'], fx['psi'], fx['v']), fx['ptsx'], fx['ptsy'])
r = S.solve_one(state, coeffs, ylo, yhi)
o = po.solve(ocfg, po.make_problem(state, coeffs, ylo, yhi))
print('gpu', r['status'], r['iters'], r['result'])
print('cpu', o['status'], o['iters'], o['result'])
print('maxdiff', np.abs(r['result'] - o['result']).max())
B = 4096
b = mpc.workloads.batch_perturbed_states(B, 0, cd)
t = time.time(); g = S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi']); print('gpu batch host', time.time() - t)
probs = po.problems_from_arrays(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
t = time.time(); c = po.solve_batch(ocfg, probs, 16); print('cpu batch', time.time() - t)
print('status gpu', np.bincount(g['status']), 'cpu', np.bincount(c['status']))
print('iters equal frac', (g['iters'] == c['iters']).mean(), 'gpu mean', g['iters'].mean(), 'cpu mean', c['iters'].mean())
d = np.abs(g['result'] - c['result']); rel = d[:, 8] / np.abs(c['result'][:, 8])
print('max abs diff first 8', d[:, :8].max(), 'max rel cost', rel.max())
bad = np.nonzero((d[:, :8].max(axis=1) > 1e-6) | (rel > 1e-8))[0]
print('n bad', bad.size, bad[:10])
for i in bad[:5]:
    print(i, g['status'][i], c['status'][i], g['iters'][i], c['iters'][i], g['result'][i], c['result'][i])
print('traj diff', np.abs(g['traj_x'] - c['traj_x']).max(), np.abs(g['traj_y'] - c['traj_y']).max())
