"""Reproducer for status 13 in ragged launches (debugging aid): hard cell N=30 dt=0.1 run (a) uniform with its own template,
(b) as N_per under a cfg.N=50 handle (NS=64 template), (c) mixed with N=50 dt=0.05; lane chain and coop kernel."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
cd = mpc.config_from_json_text(json.dumps(js)).as_dict()
b = mpc.workloads.batch_perturbed_states(B, 1, cd)
args = (b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
def run(cfgN, Nper, dtper, kind, tail=None, label=''):
    S = mpc.Solver(mpc.config_from_json_text(json.dumps(dict(js, N=cfgN, dt=0.1))), 0)
    S.set_kernel(kind)
    if tail: S.set_tail(*tail)
    outs = []
    for r in range(2):
        g = S.solve_batch_host(*args, N_per=Nper, dt_per=dtper)
        outs.append(g)
    S.close()
    st = outs[1]['status']
    print('%-46s status %s  repeat-identical %s' % (label, dict(zip(*[x.tolist() for x in np.unique(st, return_counts=True)])),
          np.array_equal(outs[0]['status'], outs[1]['status']) and np.array_equal(outs[0]['result'], outs[1]['result'])), flush=True)
    return outs[1]
ref = run(30, None, None, mpc.KERNEL_COOP, label='uniform N=30 (NS=32) coop')
g = run(30, None, None, mpc.KERNEL_LANE, label='uniform N=30 (NS=32) lane chain')
print('   lane == coop bits', np.array_equal(g['result'], ref['result']))
N30 = np.full(B, 30, dtype=np.int32); dt01 = np.full(B, 0.1)
for kind, nm in ((mpc.KERNEL_COOP, 'coop'), (mpc.KERNEL_LANE, 'lane chain')):
    g = run(50, N30, dt01, kind, label='N_per=30 under cfg.N=50 (NS=64) ' + nm)
    print('   == uniform coop bits', np.array_equal(g['result'], ref['result']), 'bad idx', np.nonzero(g['status'] != ref['status'])[0][:10])
for tail in ((0, 0), (16, 1), (16, 3)):
    g = run(50, N30, dt01, mpc.KERNEL_LANE, tail=tail, label='N_per=30 under cfg.N=50 lane chain tail=%s' % (tail,))
    print('   == uniform coop bits', np.array_equal(g['result'], ref['result']), 'bad', int((g['status'] != ref['status']).sum()))
Nmix = np.where(np.arange(B) % 2 == 0, 30, 50).astype(np.int32); dtmix = np.where(np.arange(B) % 2 == 0, 0.1, 0.05)
for kind, nm in ((mpc.KERNEL_COOP, 'coop'), (mpc.KERNEL_LANE, 'lane chain')):
    g = run(50, Nmix, dtmix, kind, label='mixed 30/50 under cfg.N=50 ' + nm)
    m = Nmix == 30
    print('   N=30 half == uniform coop bits', np.array_equal(g['result'][m], ref['result'][m]), 'bad', int((g['status'][m] != ref['status'][m]).sum()))
