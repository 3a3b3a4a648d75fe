"""Tail packing sweep (mpc_set_tail / mpc_set_handoff) on BASELINE configs 1, 3, 4: time and bit-identity."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
which = sys.argv[1] if len(sys.argv) > 1 else "134"
only = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]]   # optional: just these (handoff, park, resume, sort) settings
rd = mpc.workloads.reference_data()
js = rd['configs']['stable']
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def sweep(name, S, B, call, settings, reps):
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    ref = None
    for tup in (only or settings):
        hand, park, ph, srt = tup[:4]; rmin = tup[4] if len(tup) > 4 else 0
        S.set_handoff(hand); S.set_tail(park, ph, srt, False, rmin)
        res.zero_(); st.zero_(); it.zero_()
        l0 = S.launches
        ms = timed(lambda: call(S, res, st, it), reps)
        nl = (S.launches - l0) // reps
        cur = (res.clone(), st.clone(), it.clone())
        if ref is None: ref = cur
        same = all(torch.equal(a, b) for a, b in zip(ref, cur))
        print('%s B=%d handoff=%2d park=%2d resume=%d sort=%d rmin=%5d  %8.3f ms  %10.0f solves/s  launches %d  ok=%.4f iters max %d  identical=%s' % (
            name, B, hand, park, ph, srt, rmin, ms, B / ms * 1e3, nl, (st == 1).float().mean().item(), it.max().item(), same), flush=True)
cfg = mpc.config_from_json_text(json.dumps(js))
cd = cfg.as_dict()
if "1" in which:
    B = 65536
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    S = mpc.Solver(cfg, 0)
    sets = [(0, 0, 0, 1), (13, 0, 0, 1)] + [(h, p, r, 1) for h in (13, 0) for p in (4, 8, 12, 16) for r in (0, 1, 2)]
    sweep('cfg1', S, B, lambda S, res, st, it: S.solve_batch_device(B, *ins, res, None, None, None, st, it), sets, 4)
    S.close()
if "4" in which:
    B4 = 131072
    b = mpc.workloads.batch_perturbed_states(4096, 0, cd)
    rng = np.random.default_rng(2)
    sel = rng.integers(0, 4096, B4)
    W = np.tile(np.array(cd["weights"]), (B4, 1))
    W[:, 3] = np.exp(rng.uniform(np.log(1), np.log(5000), B4)); W[:, 4] = np.exp(rng.uniform(np.log(1), np.log(5000), B4))
    W[:, 1] = np.exp(rng.uniform(np.log(1), np.log(1000), B4)); W[:, 2] = rng.choice([0.01, 0.1, 1, 10, 100], B4)
    ins4 = [up(b["state"][sel]), up(b["coeffs"][sel]), up(b["yaw_lo"][sel]), up(b["yaw_hi"][sel])]
    Wd = up(W)
    S = mpc.Solver(cfg, 0)
    sets = [(0, 0, 0, 1), (13, 0, 0, 1), (13, 8, 0, 1), (13, 8, 1, 1), (13, 8, 2, 1), (0, 8, 1, 1), (0, 8, 2, 1), (0, 12, 2, 1), (0, 16, 2, 1), (0, 16, 3, 1)]
    sweep('cfg4', S, B4, lambda S, res, st, it: S.solve_batch_device(B4, *ins4, res, None, None, None, st, it, weights=Wd), sets, 2)
    S.close()
if "3" in which:
    PAIRS = [(10, .1), (20, .1), (30, .1), (40, .1), (10, .05), (20, .05), (30, .05), (40, .05), (50, .05), (10, .02), (20, .02), (30, .02), (40, .02), (50, .02)]
    B3 = 262144
    b = mpc.workloads.batch_perturbed_states(B3, 0, cd)
    ins3 = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    rng = np.random.default_rng(1)
    pick = rng.integers(0, len(PAIRS), B3)
    Np = up(np.array([PAIRS[k][0] for k in pick], dtype=np.int32)); dtp = up(np.array([PAIRS[k][1] for k in pick]))
    cfg3 = mpc.config_from_json_text(json.dumps(dict(js, N=50)))
    S = mpc.Solver(cfg3, 0)
    sets = [(0, 0, 0, 0), (0, 0, 0, 1), (0, 8, 1, 1), (0, 8, 2, 1), (0, 8, 3, 1), (0, 16, 3, 1), (0, 16, 5, 1), (0, 24, 5, 1)]
    sweep('cfg3', S, B3, lambda S, res, st, it: S.solve_batch_device(B3, *ins3, res, None, None, None, st, it, N_per=Np, dt_per=dtp), sets, 1)
    S.close()
