import json, sys, os
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import mpc_b200 as mpc, nlp_numpy as nn
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable'])); cd = cfg.as_dict()
B, N = 65536, cd['N']
b = mpc.workloads.batch_perturbed_states(B, 0, cd)
S = mpc.Solver(cfg, 0)
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
lam = torch.zeros(6 * N, B, dtype=torch.float64, device=dev); zl = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev); zu = torch.zeros_like(zl)
full = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev); res = torch.zeros(9, B, dtype=torch.float64, device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
S.set_dual_outputs(lam, zl, zu)
S.solve_batch_device(B, up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi']), res, None, None, full, st, None)
torch.cuda.synchronize()
ok = st.cpu().numpy() == 1
z, lam, zl, zu = full.cpu().numpy().T[ok], lam.cpu().numpy().T[ok], zl.cpu().numpy().T[ok], zu.cpu().numpy().T[ok]
xl, xu = nn.var_bounds(cd, b['yaw_lo'][ok], b['yaw_hi'][ok])
bl, bu = xl > -1e18, xu < 1e18
cl = np.where(bl, zl * (z - xl), 0.0); cu = np.where(bu, zu * (xu - z), 0.0)
print('compl orig bounds: max', cl.max(), cu.max(), 'p99.9', np.percentile(cl.max(axis=1), 99.9), np.percentile(cu.max(axis=1), 99.9))
i, j = np.unravel_index(np.argmax(cu), cu.shape)
print('worst cu at var', j, 'zu', zu[i, j], 'slack', (xu - z)[i, j], 'xu', xu[i, j])
i, j = np.unravel_index(np.argmax(cl), cl.shape)
print('worst cl at var', j, 'zl', zl[i, j], 'slack', (z - xl)[i, j], 'xl', xl[i, j])
print('zl max', zl.max(), 'zu max', zu.max(), 'lam max', np.abs(lam).max())
