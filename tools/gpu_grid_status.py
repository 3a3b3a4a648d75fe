"""Config 3 (N x dt grid, ragged launch): which problems end without success, are they the same from run to run,
and what do the coop kernel alone and the oracle say about them (debugging aid)."""
import json, sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
from oracle import pyoracle as po
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n_orc = int(sys.argv[2]) if len(sys.argv) > 2 else 24
PAIRS = [(10, .1), (20, .1), (30, .1), (40, .1), (10, .05), (20, .05), (30, .05), (40, .05), (50, .05), (10, .02), (20, .02), (30, .02), (40, .02), (50, .02)]
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
cd = mpc.config_from_json_text(json.dumps(js)).as_dict()
b = mpc.workloads.batch_perturbed_states(B, 1, cd)
pick = np.random.default_rng(1).integers(0, len(PAIRS), B)
Nall = np.array([PAIRS[k][0] for k in pick], dtype=np.int32); dtall = np.array([PAIRS[k][1] for k in pick])
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ins = [up(b['state'].T), up(b['coeffs'].T), up(b['yaw_lo']), up(b['yaw_hi'])]
Np, dtp = up(Nall), up(dtall)
cfg = mpc.config_from_json_text(json.dumps(dict(js, N=50)))
S = mpc.Solver(cfg, 0)
if os.environ.get('KIND'): S.set_kernel(int(os.environ['KIND']))
if os.environ.get('TAIL'): S.set_tail(*[int(x) for x in os.environ['TAIL'].split(',')])
res = torch.zeros(9, B, dtype=torch.float64, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
runs = []
for r in range(3):
    st.zero_(); it.zero_(); res.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.solve_batch_device(B, *ins, res, None, None, None, st, it, N_per=Np, dt_per=dtp)
    e1.record()
    torch.cuda.synchronize()
    print('run', r, '%.2f ms' % e0.elapsed_time(e1), 'parked', S.tail_counts(5))
    runs.append((st.cpu().numpy().copy(), it.cpu().numpy().copy(), res.cpu().numpy().copy()))
    bad = np.nonzero(runs[-1][0] != 1)[0]
    print('run', r, 'status hist', dict(zip(*np.unique(runs[-1][0], return_counts=True))), 'bad', len(bad))
for r in (1, 2):
    print('run', r, 'vs run 0: status equal', np.array_equal(runs[r][0], runs[0][0]), 'iters equal', np.array_equal(runs[r][1], runs[0][1]),
          'result bits equal', np.array_equal(runs[r][2], runs[0][2]), 'n differing', int((runs[r][0] != runs[0][0]).sum()))
s0, i0, r0 = runs[0]
bad = np.nonzero((runs[0][0] != 1) | (runs[1][0] != 1) | (runs[2][0] != 1))[0]
for k in bad[:int(os.environ.get('NSHOW', '60'))]:
    print('  b', k, 'N', Nall[k], 'dt', dtall[k], 'status', [int(x[0][k]) for x in runs], 'iters', [int(x[1][k]) for x in runs])
S.close()
# the same problems alone, uniform horizon, coop kernel and AUTO
for k in bad[:n_orc]:
    n, d = int(Nall[k]), float(dtall[k])
    jsk = dict(js, N=n, dt=d)
    cfgk = mpc.config_from_json_text(json.dumps(jsk))
    Sk = mpc.Solver(cfgk, 0)
    Sk.set_kernel(mpc.KERNEL_COOP)
    g = Sk.solve_batch_host(b['state'][k:k + 1], b['coeffs'][k:k + 1], b['yaw_lo'][k:k + 1], b['yaw_hi'][k:k + 1])
    Sk.close()
    cdk = po.load_config_dict(jsk)
    o = po.solve_batch(po.make_config(cdk), po.problems_from_arrays(b['state'][k:k + 1], b['coeffs'][k:k + 1], b['yaw_lo'][k:k + 1], b['yaw_hi'][k:k + 1]), 1)
    print('  b', k, 'N', n, 'dt', d, 'coop alone st/it', int(g['status'][0]), int(g['iters'][0]), 'oracle st/it', int(o['status'][0]), int(o['iters'][0]),
          'maxdiff %.3g' % np.abs(g['result'][0, :8] - o['result'][0, :8]).max(), flush=True)
