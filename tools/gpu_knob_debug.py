"""Which problems differ between the GPU and the oracle under a forced option (debugging aid)."""
import json, sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
from oracle import pyoracle as po
knob, val = sys.argv[1], float(sys.argv[2])
val = int(val) if knob != 'tiny_step_tol' else val
rd = mpc.workloads.reference_data()
js = dict(rd['configs']['stable'], N=30, dt=0.1)
cd = po.load_config_dict(js)
b = mpc.workloads.batch_perturbed_states(200, 1, cd)
B = 200
probs = po.problems_from_arrays(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
res = (po.OrcResult * B)()
ocfg = po.make_config(cd, **{knob: val})
po.lib().orc_solve_batch(C.byref(ocfg), probs, B, res, 16)
cfg = mpc.config_from_json_text(json.dumps(js)); setattr(cfg, knob, val)
S = mpc.Solver(cfg, 0); S.set_kernel(mpc.KERNEL_COOP)
g = S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
for i in range(B):
    d = np.abs(g['result'][i, :8] - np.array(res[i].result[:8])).max()
    if d > 1e-6 or g['status'][i] != res[i].status or g['iters'][i] != res[i].iters:
        print(i, 'd %.3g' % d, 'gpu st/it', g['status'][i], g['iters'][i], 'orc st/it', res[i].status, res[i].iters, 'wd', res[i].n_watchdog, 'resto', res[i].n_resto, 'cost gpu %.6f orc %.6f' % (g['result'][i, 8], res[i].result[8]))
