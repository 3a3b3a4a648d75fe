"""BASELINE configs 3 and 4 on one GPU: the N x dt grid (256K problems, submission-report.md:250-265) and the
cost-weight sweep (per-problem weights).  Prints solves/s, status histogram and iteration statistics."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
PAIRS = [(10, .1), (20, .1), (30, .1), (40, .1), (10, .05), (20, .05), (30, .05), (40, .05), (50, .05), (10, .02), (20, .02), (30, .02), (40, .02), (50, .02)]

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def grid(B):
    js = rd['configs']['stable']
    cd = mpc.config_from_json_text(json.dumps(js)).as_dict()
    b = mpc.workloads.batch_perturbed_states(B, 1, cd)
    rng = np.random.default_rng(1)
    pick = rng.integers(0, len(PAIRS), B)
    Np = np.array([PAIRS[k][0] for k in pick], dtype=np.int32)
    dtp = np.array([PAIRS[k][1] for k in pick])
    # (a) grouped by N: one shape-homogeneous launch per horizon, per-problem dt
    tot_ms, ok, its = 0.0, 0, 0
    for N in (10, 20, 30, 40, 50):
        idx = np.nonzero(Np == N)[0]
        n = idx.size
        cfg = mpc.config_from_json_text(json.dumps(dict(js, N=N)))
        S = mpc.Solver(cfg, 0)
        ins = [up(b['state'][idx].T), up(b['coeffs'][idx].T), up(b['yaw_lo'][idx]), up(b['yaw_hi'][idx])]
        res = torch.zeros(9, n, dtype=torch.float64, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev); it = torch.zeros(n, dtype=torch.int32, device=dev)
        dt_d = up(dtp[idx])
        ms = timed(lambda: S.solve_batch_device(n, *ins, res, None, None, None, st, it, dt_per=dt_d))
        tot_ms += ms
        s, i = st.cpu().numpy(), it.cpu().numpy()
        ok += (s == 1).sum(); its += i.sum()
        print('  N=%2d  n=%6d  %.2f ms  %.0f solves/s  ok %.4f  iters mean %.1f p99 %d max %d  status %s' % (N, n, ms, n / ms * 1e3, (s == 1).mean(), i.mean(), np.percentile(i, 99), i.max(), np.bincount(s).tolist()))
        S.close()
    print('config 3 grouped: B=%d  %.1f ms  %.0f solves/s  ok %.4f' % (B, tot_ms, B / tot_ms * 1e3, ok / B))
    # (b) ragged: one launch, per-problem N and dt
    cfg = mpc.config_from_json_text(json.dumps(dict(js, N=50)))
    S = mpc.Solver(cfg, 0)
    ins = [up(b['state'].T), up(b['coeffs'].T), up(b['yaw_lo']), up(b['yaw_hi'])]
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    Nd, dtd = up(Np), up(dtp)
    ms = timed(lambda: S.solve_batch_device(B, *ins, res, None, None, None, st, it, N_per=Nd, dt_per=dtd), 2)
    print('config 3 ragged (one launch): B=%d  %.1f ms  %.0f solves/s  ok %.4f' % (B, ms, B / ms * 1e3, (st == 1).float().mean().item()))
    S.close()

def weights(B):
    js = rd['configs']['stable']
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = cfg.as_dict()
    base = mpc.workloads.batch_perturbed_states(4096, 0, cd)
    rng = np.random.default_rng(2)
    sel = rng.integers(0, 4096, B)
    W = np.tile(np.array(cd['weights']), (B, 1))
    W[:, 3] = np.exp(rng.uniform(np.log(1), np.log(5000), B)); W[:, 4] = np.exp(rng.uniform(np.log(1), np.log(5000), B))
    W[:, 1] = np.exp(rng.uniform(np.log(1), np.log(1000), B)); W[:, 2] = rng.choice([0.01, 0.1, 1, 10, 100], B)
    W[:, 6] = rng.uniform(0, 1e4, B); W[:, 7] = rng.uniform(0, 1e4, B)
    S = mpc.Solver(cfg, 0)
    ins = [up(base['state'][sel].T), up(base['coeffs'][sel].T), up(base['yaw_lo'][sel]), up(base['yaw_hi'][sel])]
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    Wd = up(W.T)
    ms = timed(lambda: S.solve_batch_device(B, *ins, res, None, None, None, st, it, weights=Wd), 2)
    s, i = st.cpu().numpy(), it.cpu().numpy()
    print('config 4 weight sweep: B=%d  %.1f ms  %.0f solves/s  ok %.4f  iters mean %.1f p99 %d max %d  status %s' % (B, ms, B / ms * 1e3, (s == 1).mean(), i.mean(), np.percentile(i, 99), i.max(), np.bincount(s).tolist()))
    S.close()

if __name__ == '__main__':
    grid(int(sys.argv[1]) if len(sys.argv) > 1 else 262144)
    weights(int(sys.argv[2]) if len(sys.argv) > 2 else 131072)
