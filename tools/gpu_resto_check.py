"""Restoration-phase parity on the hard cells of the horizon grid: GPU (lane chain and coop kernel) against the
oracle, no status mask.  Usage: gpu_resto_check.py [B]"""
import json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
from oracle import pyoracle as po
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rd = mpc.workloads.reference_data()
for N, dt in ((20, 0.1), (30, 0.1), (40, 0.1), (50, 0.05), (10, 0.1)):
    js = dict(rd['configs']['stable'], N=N, dt=dt)
    cfg = mpc.config_from_json_text(json.dumps(js)); cd = po.load_config_dict(js)
    b = mpc.workloads.batch_perturbed_states(B, 1, cd)
    args = (b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
    t = time.time()
    c = po.solve_batch(po.make_config(cd), po.problems_from_arrays(*args), os.cpu_count())
    tc = time.time() - t
    S = mpc.Solver(cfg, 0)
    for kind, name in ((mpc.KERNEL_COOP, 'coop'), (mpc.KERNEL_LANE, 'lane')):
        S.set_kernel(kind)
        t = time.time()
        g = S.solve_batch_host(*args)
        tg = time.time() - t
        same_st = (g['status'] == c['status'])
        same_it = (g['iters'] == c['iters'])
        d = np.abs(g['result'][:, :8] - c['result'][:, :8]).max(axis=1)
        dc = np.abs(g['result'][:, 8] - c['result'][:, 8]) / np.maximum(1.0, np.abs(c['result'][:, 8]))
        print('N=%d dt=%.2f %s: oracle status %s (%.1fs)  gpu status %s (%.3fs)  status equal %.4f  iters equal %.4f  max|d| %.3g (p99 %.3g)  max rel cost %.3g  iters max gpu %d oracle %d' % (
            N, dt, name, dict(zip(*np.unique(c['status'], return_counts=True))), tc, dict(zip(*np.unique(g['status'], return_counts=True))), tg,
            same_st.mean(), same_it.mean(), d.max(), np.percentile(d, 99), dc.max(), g['iters'].max(), c['iters'].max()))
        bad = np.nonzero(~same_st | (d > 1e-4))[0]
        if bad.size:
            print('   mismatches', bad[:10], 'gpu st', g['status'][bad][:10], 'orc st', c['status'][bad][:10], 'gpu it', g['iters'][bad][:10], 'orc it', c['iters'][bad][:10], 'd', d[bad][:10])
    S.close()
