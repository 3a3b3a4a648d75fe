"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle, the committed
golden vectors and size-independent properties.

Tolerances are the north star's: actuations / predicted trajectory 1e-4 absolute, cost 1e-6 relative
(FP64).  In practice the two agree to ~1e-8 because they run the same interior-point iteration with
independent linear algebra (Riccati recursion on the GPU, dense Bunch-Kaufman LDL^T in the oracle)."""
import json
import os

import numpy as np
import pytest

import nlp_numpy as nn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ABS_TOL, REL_TOL = 1e-4, 1e-6


def _assert_parity(g, c, mask=None):
    ok = c["status"] == 1 if mask is None else mask
    assert np.array_equal(g["status"][ok], c["status"][ok])
    d = np.abs(g["result"] - c["result"])[ok]
    assert d[:, :8].max() < ABS_TOL
    assert (d[:, 8] / np.abs(c["result"][ok, 8])).max() < REL_TOL
    assert np.abs(g["traj_x"] - c["traj_x"])[ok].max() < ABS_TOL
    assert np.abs(g["traj_y"] - c["traj_y"])[ok].max() < ABS_TOL
    return d


def test_golden_testcpp_scenarios(solver):
    """run() + 25 x solve() of src/test.cpp:64-111 for the four fixtures, vs tests/golden."""
    gold = json.load(open(os.path.join(GOLD, "oracle_golden.json")))
    for sc in gold["scenarios"]:
        state = np.array(sc["state0"])
        for step in sc["steps"]:
            r = solver.solve_one(state, sc["coeffs"], sc["yaw_lo"], sc["yaw_hi"])
            assert r["status"] == step["status"] == 1
            assert np.abs(r["result"][:8] - np.array(step["result"][:8])).max() < ABS_TOL
            assert r["result"][8] == pytest.approx(step["result"][8], rel=REL_TOL)
            assert np.abs(r["traj_x"] - np.array(step["traj_x"])).max() < ABS_TOL
            assert np.abs(r["traj_y"] - np.array(step["traj_y"])).max() < ABS_TOL
            assert 0 <= r["iters"] - step["iters"] <= 1      # the oracle's count, or one more (never fewer: DESIGN.md section 3)
            state = np.array(step["result"][:6])     # feed the golden state back (no drift accumulation)


def test_testcpp_first_solve_survey_values(solver, po, stable_cd, refdata):
    fx = refdata["test_cpp_fixtures"][0]
    state, coeffs, ylo, yhi, _ = po.preprocess(stable_cd, (fx["x"], fx["y"], fx["psi"], fx["v"]), fx["ptsx"], fx["ptsy"])
    r = solver.solve_one(state, coeffs, ylo, yhi)
    assert r["status"] == 1 and r["iters"] == 10
    assert r["result"][:8] == pytest.approx(
        [2.66806, 0, 0.00241174, 27.1276389, -0.1072304, 0.0283982, 0.002413, 4.4703889], abs=2e-6)
    assert r["result"][8] == pytest.approx(6243.443671, rel=1e-9)


@pytest.mark.parametrize("seed,B", [(0, 4096), (7, 1000), (8, 33), (9, 1)])
def test_batch_vs_oracle(solver, mpc, po, stable_cd, seed, B):
    b = mpc.workloads.batch_perturbed_states(B, seed, stable_cd)
    g = solver.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    c = po.solve_batch(po.make_config(stable_cd), po.problems_from_arrays(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"]), 16)
    assert (c["status"] == 1).mean() > 0.99
    d = _assert_parity(g, c)
    assert np.median(d[:, :8].max(axis=1)) < 1e-9
    # iteration counts: the oracle's on ~99 % of the problems, one more on the rest (the Riccati recursion's rounding residue in
    # the multipliers meets Ipopt's 1e-8 test one iteration later), never fewer
    di = g["iters"].astype(int) - c["iters"].astype(int)
    assert (di == 0).mean() > 0.97 and di.min() >= 0 and di.max() <= 1, np.unique(di, return_counts=True)


@pytest.mark.parametrize("name", ["fast", "no-latency"])
def test_other_shipped_configs(mpc, po, refdata, name, kernel_kind):
    cfg = mpc.config_from_json_text(json.dumps(refdata["configs"][name]))
    cd = po.load_config_dict(refdata["configs"][name])
    b = mpc.workloads.batch_perturbed_states(512, 21, cd)
    S = mpc.Solver(cfg, 0)
    S.set_kernel(kernel_kind)
    g = S.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.close()
    c = po.solve_batch(po.make_config(cd), po.problems_from_arrays(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"]), 16)
    _assert_parity(g, c)


@pytest.mark.parametrize("N,dt", [(2, 0.1), (3, 0.1), (10, 0.05), (10, 0.02), (11, 0.05), (20, 0.05), (21, 0.05),
                                  (25, 0.025), (30, 0.02), (32, 0.05), (40, 0.05), (50, 0.02), (64, 0.02)])
def test_horizon_and_timestep_grid(mpc, po, refdata, N, dt, kernel_kind):
    """N x dt cells of the reference's examples/ grid (submission-report.md:250-265), up to MPC_NMAX = 64."""
    js = dict(refdata["configs"]["stable"], N=N, dt=dt)
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = po.load_config_dict(js)
    b = mpc.workloads.batch_perturbed_states(96, 31, cd)
    S = mpc.Solver(cfg, 0)
    S.set_kernel(kernel_kind)
    g = S.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.close()
    c = po.solve_batch(po.make_config(cd), po.problems_from_arrays(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"]), 16)
    both = (c["status"] == 1) & (g["status"] == 1)
    assert both.mean() > 0.9
    assert (g["status"] == c["status"]).mean() > 0.97
    _assert_parity(g, c, both)


def test_per_problem_weights_and_frozen_accel_terms(solver, mpc, po, stable_cd):
    """Weight sweep (BASELINE config 4): per-problem weights; w_a, w_adot, w_decel must have NO effect."""
    B = 256
    b = mpc.workloads.batch_perturbed_states(B, 41, stable_cd)
    rng = np.random.default_rng(2)
    W = np.tile(np.array(stable_cd["weights"]), (B, 1))
    W[:, 3] = np.exp(rng.uniform(np.log(1), np.log(5000), B))
    W[:, 4] = np.exp(rng.uniform(np.log(1), np.log(5000), B))
    W[:, 1] = np.exp(rng.uniform(np.log(1), np.log(1000), B))
    W[:, 2] = rng.choice([0.01, 0.1, 1, 10, 100], B)
    g = solver.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"], weights=W)
    W2 = W.copy()
    W2[:, 6] = rng.uniform(0, 1e4, B); W2[:, 7] = rng.uniform(0, 1e4, B); W2[:, 8] = rng.uniform(0, 1e4, B); W2[:, 5] = 7.0
    g2 = solver.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"], weights=W2)
    assert np.array_equal(g["result"], g2["result"]) and np.array_equal(g["iters"], g2["iters"])
    # oracle, problem by problem with its own weights
    for i in range(0, B, 8):
        cd = dict(stable_cd, weights=list(W[i]))
        r = po.solve(po.make_config(cd), po.make_problem(b["state"][i], b["coeffs"][i], b["yaw_lo"][i], b["yaw_hi"][i]))
        if r["status"] != 1:
            continue
        assert g["status"][i] == 1
        assert np.abs(g["result"][i, :8] - r["result"][:8]).max() < ABS_TOL
        assert g["result"][i, 8] == pytest.approx(r["result"][8], rel=REL_TOL)


def test_ragged_per_problem_horizon_and_dt(mpc, po, refdata, kernel_kind):
    """Horizon/timestep sweep (BASELINE config 3) in ONE launch: per-problem N and dt."""
    js = dict(refdata["configs"]["stable"], N=30)
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = po.load_config_dict(js)
    B = 192
    b = mpc.workloads.batch_perturbed_states(B, 51, cd)
    rng = np.random.default_rng(1)
    pairs = [(10, 0.1), (20, 0.1), (30, 0.1), (10, 0.05), (20, 0.05), (30, 0.05), (10, 0.02), (20, 0.02), (30, 0.02)]
    pick = rng.integers(0, len(pairs), B)
    Np = np.array([pairs[k][0] for k in pick], dtype=np.int32)
    dtp = np.array([pairs[k][1] for k in pick])
    S = mpc.Solver(cfg, 0)
    S.set_kernel(kernel_kind)
    g = S.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"], N_per=Np, dt_per=dtp)
    S.close()
    n_checked = 0
    for i in range(0, B, 3):
        cdi = dict(cd, N=int(Np[i]), dt=float(dtp[i]))
        r = po.solve(po.make_config(cdi), po.make_problem(b["state"][i], b["coeffs"][i], b["yaw_lo"][i], b["yaw_hi"][i]))
        if r["status"] != 1 or g["status"][i] != 1:
            continue
        n_checked += 1
        assert np.abs(g["result"][i, :8] - r["result"][:8]).max() < ABS_TOL
        assert g["result"][i, 8] == pytest.approx(r["result"][8], rel=REL_TOL)
        assert np.abs(g["traj_x"][i, :Np[i]] - r["z"][:Np[i]]).max() < ABS_TOL
    assert n_checked > 40


def test_full_size_batch_properties(solver, mpc, stable_cd):
    """BASELINE configs[1] at full size (65,536 problems): solver-independent checks of every result
    against the independent numpy statement of the reference NLP."""
    B = 65536
    cd = stable_cd
    N = cd["N"]
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    g = solver.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"], want_full=True)
    ok = g["status"] == 1
    assert ok.mean() > 0.995, np.bincount(g["status"])
    z = g["full"]
    # feasibility of the dynamics (equality constraints of MPC.cpp:116-153)
    c = nn.constraints(cd, b["state"], b["coeffs"], z)
    assert np.abs(c[ok]).max() < 1e-6
    # bounds (MPC.cpp:220-257), honoured exactly after the final clip
    xl, xu = nn.var_bounds(cd, b["yaw_lo"], b["yaw_hi"])
    assert (z[ok] >= xl[ok]).all() and (z[ok] <= xu[ok]).all()
    # returned cost == objective of the returned point; outputs are the documented slices
    f = nn.objective(cd, nn.frozen(cd, b["state"]), z)
    assert np.allclose(g["result"][ok, 8], f[ok], rtol=1e-12)
    assert np.array_equal(g["result"][:, 0], z[:, 1]) and np.array_equal(g["result"][:, 6], z[:, 6 * N])
    assert np.array_equal(g["result"][:, 7], z[:, 7 * N - 1])
    assert np.array_equal(g["traj_x"], z[:, :N]) and np.array_equal(g["traj_y"], z[:, N:2 * N])
    assert np.percentile(g["iters"], 50) <= 11 and g["iters"][ok].max() < 200
    # shard invariance (multi-GPU contract): solving a slice alone gives bit-identical results
    s = slice(12345, 12345 + 4096)
    g2 = solver.solve_batch_host(b["state"][s], b["coeffs"][s], b["yaw_lo"][s], b["yaw_hi"][s])
    assert np.array_equal(g2["result"], g["result"][s]) and np.array_equal(g2["iters"], g["iters"][s])
    # determinism
    g3 = solver.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    assert np.array_equal(g3["result"], g["result"])


def test_edge_cases(solver, mpc, stable_cfg, stable_cd):
    import ctypes as C
    L = mpc.lib()
    # B = 0 is a no-op; NULL required pointers and negative B are EINVAL
    assert L.mpc_solve_batch_host(solver._h, 0, 1, 1, 1, 1, None, None, None, 1, None, None, None, None, None) == 0
    assert L.mpc_solve_batch_host(solver._h, 4, None, 1, 1, 1, None, None, None, 1, None, None, None, None, None) == -1
    assert L.mpc_solve_batch_host(solver._h, -1, 1, 1, 1, 1, None, None, None, 1, None, None, None, None, None) == -1
    # infeasible start (psi0 outside the yaw bounds, |v0| above max speed): terminates with a
    # non-success status like the reference ("Ipopt failed with <int>", MPC.cpp:301), no hang
    st = np.array([[0, 0, 0.5, 10.0, 0.1, 0.0], [0, 0, 0.0, 80.0, 0.1, 0.0], [0, 0, 0, 20.0, 0.0, 0.0]])
    co = np.zeros((3, 5))
    r = solver.solve_batch_host(st, co, np.array([-0.1, -0.1, -0.1]), np.array([0.1, 0.1, 0.1]))
    assert r["status"][0] != 1 and r["status"][1] != 1 and r["status"][2] == 1
    assert np.isfinite(r["result"][2]).all()
    # straight road, centred: steer 0, full throttle (a at its upper bound)
    assert abs(r["result"][2, 6]) < 1e-9 and r["result"][2, 7] == pytest.approx(stable_cd["max_accel"], abs=1e-7)
    # outputs are optional
    out = solver.solve_batch_host(st[2:], co[2:], np.array([-0.1]), np.array([0.1]), want_traj=False)
    assert out["status"][0] == 1


def test_auto_dispatch_and_kernels_agree(mpc, stable_cfg, stable_cd):
    """MPC_KERNEL_AUTO: small batches take the coop kernel, large ones the lane kernel.  Those two run the same
    arithmetic in the same order per problem, so they agree to the last bit (which is what makes migrating a
    problem between them safe)."""
    S = mpc.Solver(stable_cfg, 0)
    b = mpc.workloads.batch_perturbed_states(mpc.LANE_MIN_BATCH, 77, stable_cd)
    args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    auto = S.solve_batch_host(*args)
    S.set_kernel(mpc.KERNEL_LANE)
    lane = S.solve_batch_host(*args)
    S.set_kernel(mpc.KERNEL_COOP)
    coop = S.solve_batch_host(*args)
    assert np.array_equal(auto["result"], lane["result"])          # B >= MPC_LANE_MIN_BATCH
    assert np.array_equal(coop["status"], lane["status"]) and np.array_equal(coop["iters"], lane["iters"])
    assert np.abs(coop["result"] - lane["result"]).max() < 1e-9
    S.set_kernel(mpc.KERNEL_AUTO)
    small = S.solve_batch_host(*(a[:100] for a in args))
    assert np.array_equal(small["result"], coop["result"][:100])   # 100 < MPC_LANE_MIN_BATCH
    # lane grid settings change scheduling only, never results
    S.set_kernel(mpc.KERNEL_LANE, 64, 3)
    lane2 = S.solve_batch_host(*args)
    assert np.array_equal(lane2["result"], lane["result"]) and np.array_equal(lane2["iters"], lane["iters"])
    for gone in (1, 4, 7):   # 1 and 4 were the first-version warp kernel and the solo kernel
        with pytest.raises(mpc.MpcError):
            S.set_kernel(gone)
    S.close()


def test_rare_paths_regularisation_and_long_runs(mpc, po, refdata, kernel_kind):
    """Problems that need Ipopt's inertia correction (delta_w) or many iterations: same answers."""
    cd = po.load_config_dict(refdata["configs"]["stable"])
    cfg = mpc.config_from_json_text(json.dumps(refdata["configs"]["stable"]))
    b = mpc.workloads.batch_perturbed_states(16384, 0, cd)
    S = mpc.Solver(cfg, 0)
    S.set_kernel(kernel_kind)
    g = S.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.close()
    hard = np.argsort(-g["iters"])[:96]                              # the longest runs of the batch
    assert g["iters"][hard].min() >= 14
    sub = {k: b[k][hard] for k in ("state", "coeffs", "yaw_lo", "yaw_hi")}
    import ctypes as C
    probs = po.problems_from_arrays(sub["state"], sub["coeffs"], sub["yaw_lo"], sub["yaw_hi"])
    outs = (po.OrcResult * len(hard))()
    po.lib().orc_solve_batch(C.byref(po.make_config(cd)), probs, len(hard), outs, 16)
    n_reg = sum(o.n_regularized for o in outs)
    assert n_reg > 0                                                  # the inertia-correction path is exercised
    for k, i in enumerate(hard):
        o = outs[k]
        if o.status != 1:
            continue
        assert g["status"][i] == 1
        assert np.abs(g["result"][i, :8] - np.array(o.result[:8])).max() < ABS_TOL
        assert g["result"][i, 8] == pytest.approx(o.result[8], rel=REL_TOL)


def test_migration_between_kernels_is_invisible(mpc, stable_cfg, stable_cd):
    """Large batches: the lane kernel parks long-running problems and the coop kernel finishes them.  Both run
    the same arithmetic on the same state, so the results are bit-identical whatever the threshold -- including
    thresholds that migrate most of the batch or overflow the record buffer."""
    S = mpc.Solver(stable_cfg, 0)
    b = mpc.workloads.batch_perturbed_states(16384, 91, stable_cd)
    args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.set_kernel(mpc.KERNEL_LANE)
    S.set_handoff(0)
    S.set_tail(0, 0)
    ref = S.solve_batch_host(*args, want_full=True)
    for it in (1, 5, 10, 13, 25):
        S.set_handoff(it)
        n0 = S.launches
        got = S.solve_batch_host(*args, want_full=True)
        assert S.launches - n0 == 2                                   # lane kernel + coop finisher
        for k in ("result", "traj_x", "traj_y", "full", "status", "iters"):
            assert np.array_equal(got[k], ref[k]), (it, k)
    with pytest.raises(mpc.MpcError):
        S.set_handoff(-1)
    # tail packing: sparse warps park their problems, resume launches pick them up 32 to a warp, the coop kernel
    # finishes -- any combination, same bits
    for hand, park, resume in ((0, 8, 0), (0, 8, 1), (13, 16, 2), (0, 31, 3), (5, 4, 6)):
        S.set_handoff(hand)
        S.set_tail(park, resume)
        n0 = S.launches
        got = S.solve_batch_host(*args, want_full=True)
        assert S.launches - n0 == 2 + resume
        for k in ("result", "traj_x", "traj_y", "full", "status", "iters"):
            assert np.array_equal(got[k], ref[k]), (hand, park, resume, k)
    # resume launches that find too few parked problems return at once and leave them to the next launch: any
    # threshold -- none skipped, some skipped, all skipped -- same bits, and the counters show who did the work
    for rmin in (0, 1500, 100000):
        S.set_handoff(0)
        S.set_tail(16, 3, True, rmin)
        got = S.solve_batch_host(*args, want_full=True)
        for k in ("result", "traj_x", "traj_y", "full", "status", "iters"):
            assert np.array_equal(got[k], ref[k]), (rmin, k)
        parked = S.tail_counts(4)
        assert parked[0] > 0
        if rmin == 0:
            assert parked[1] > 0
        if rmin == 100000:
            assert parked[1:] == [0, 0, 0]
    with pytest.raises(mpc.MpcError):
        S.set_tail(32, 0)
    with pytest.raises(mpc.MpcError):
        S.set_tail(8, 7)
    S.close()


def test_long_horizon_kernels_agree_bit_for_bit(mpc, stable_cd, refdata):
    """N > 32: the coop kernel gives every lane of a 32-lane group two stages (neighbour values through the shared rows
    instead of shuffles), and a big batch is the lane kernel with resume launches and the coop kernel as the
    finisher.  All of them are the same arithmetic (the
    library is built without implicit multiply-add contraction), so: same bits -- including N = 64, the maximum, N = 33,
    one stage into the second pass, and a horizon of 25 run through the two-stages-per-lane code."""
    import json
    for N, dt, B in ((40, 0.05, 3000), (64, 0.02, 1500), (33, 0.05, 1500)):
        cfg = mpc.config_from_json_text(json.dumps(dict(refdata["configs"]["stable"], N=N, dt=dt)))
        S = mpc.Solver(cfg, 0)
        b = mpc.workloads.batch_perturbed_states(B, 17, cfg.as_dict())
        args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
        S.set_kernel(mpc.KERNEL_LANE)
        S.set_tail(0, 0)
        ref = S.solve_batch_host(*args, want_full=True)
        assert (ref["status"] == 1).mean() > 0.9
        if N == 40:
            assert ref["iters"].max() > 40
        for kind, park, resume in ((mpc.KERNEL_COOP, 0, 0), (mpc.KERNEL_LANE, 8, 0), (mpc.KERNEL_LANE, 16, 2), (mpc.KERNEL_AUTO, 31, 1)):
            S.set_kernel(kind)
            S.set_tail(park, resume, True)
            got = S.solve_batch_host(*args, want_full=True)
            for k in ("result", "traj_x", "traj_y", "full", "status", "iters"):
                assert np.array_equal(got[k], ref[k]), (N, kind, park, resume, k)
        # AUTO on a handful of long-horizon problems is the coop kernel: one launch
        S.set_kernel(mpc.KERNEL_AUTO)
        n0 = S.launches
        few = S.solve_batch_host(*(a[:100] for a in args), want_full=True)
        assert S.launches - n0 == 1
        for k in ("result", "full", "status", "iters"):
            assert np.array_equal(few[k], ref[k][:100]), k
        S.close()
    # per-problem horizons of 10..25 in a solver configured for N = 50: the NS = 64 kernels on short problems
    cfg = mpc.config_from_json_text(json.dumps(dict(refdata["configs"]["stable"], N=50)))
    S = mpc.Solver(cfg, 0)
    B = 2000
    b = mpc.workloads.batch_perturbed_states(B, 29, cfg.as_dict())
    Np = np.random.default_rng(3).choice([10, 16, 25, 32, 34, 50], B).astype(np.int32)
    dtp = np.where(Np > 25, 0.02, 0.05)
    args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.set_kernel(mpc.KERNEL_LANE); S.set_tail(0, 0)
    ref = S.solve_batch_host(*args, N_per=Np, dt_per=dtp)
    S.set_kernel(mpc.KERNEL_COOP)
    got = S.solve_batch_host(*args, N_per=Np, dt_per=dtp)
    for k in ("result", "traj_x", "traj_y", "status", "iters"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    S.close()


def test_ragged_batch_order_is_invisible(mpc, stable_cd, refdata):
    """Batches with per-problem horizons are handed out longest horizon first (a counting sort on the device); the
    order in which lanes pick problems up does not change any result."""
    import json
    cfg = mpc.config_from_json_text(json.dumps(dict(refdata["configs"]["stable"], N=50)))
    S = mpc.Solver(cfg, 0)
    B = 6000
    b = mpc.workloads.batch_perturbed_states(B, 23, cfg.as_dict())
    rng = np.random.default_rng(5)
    Np = rng.choice([10, 17, 25, 33, 50], B).astype(np.int32)
    dtp = rng.choice([0.1, 0.05, 0.02], B)
    dtp[Np >= 33] = 0.02
    args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.set_kernel(mpc.KERNEL_LANE)
    S.set_tail(0, 0, False)
    ref = S.solve_batch_host(*args, N_per=Np, dt_per=dtp)
    for park, resume, srt in ((0, 0, True), (8, 1, True), (16, 2, False)):
        S.set_tail(park, resume, srt)
        got = S.solve_batch_host(*args, N_per=Np, dt_per=dtp)
        for k in ("result", "traj_x", "traj_y", "status", "iters"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (park, resume, srt, k)
    S.close()


def test_full_size_kkt_certificate(mpc, stable_cfg, stable_cd, kernel_kind):
    """Solver-independent proof of optimality at full batch size: with the multipliers the kernels return, every
    successful result of the 65,536-problem batch satisfies the KKT conditions of the reference's NLP as stated
    independently in tests/nlp_numpy.py (complex-step gradient of the Lagrangian, no hand-written derivative):
    stationarity, primal feasibility, dual feasibility, complementarity.  Ipopt's own acceptance test is
    max(scaled errors) <= 1e-8 with scaling s_d, s_c >= 1 and an unscaled dual-infeasibility cap of 1."""
    import torch
    B, N = 65536, stable_cd["N"]
    cd = stable_cd
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    S = mpc.Solver(stable_cfg, 0)
    S.set_kernel(kernel_kind)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    lam = torch.zeros(6 * N, B, dtype=torch.float64, device=dev)
    zl = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev)
    zu = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev)
    full = torch.zeros(8 * N - 2, B, dtype=torch.float64, device=dev)
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    S.set_dual_outputs(lam, zl, zu)
    S.solve_batch_device(B, up(b["state"]), up(b["coeffs"]), up(b["yaw_lo"]), up(b["yaw_hi"]), res, None, None, full, st, None)
    torch.cuda.synchronize()
    S.set_dual_outputs(None, None, None)
    S.close()
    ok = st.cpu().numpy() == 1
    assert ok.mean() > 0.995
    z, lam, zl, zu = full.cpu().numpy().T[ok], lam.cpu().numpy().T[ok], zl.cpu().numpy().T[ok], zu.cpu().numpy().T[ok]
    state, coeffs = b["state"][ok], b["coeffs"][ok]
    fz = {k: v[ok] if isinstance(v, np.ndarray) and v.shape[0] == B else v for k, v in nn.frozen(cd, b["state"]).items()}
    # primal feasibility and bounds
    assert np.abs(nn.constraints(cd, state, coeffs, z)).max() < 1e-7
    xl, xu = nn.var_bounds(cd, b["yaw_lo"][ok], b["yaw_hi"][ok])
    assert (z >= xl).all() and (z <= xu).all()
    # dual feasibility; multipliers only on bounded variables
    assert (zl >= 0).all() and (zu >= 0).all()
    assert (zl[xl < -1e18] == 0).all() and (zu[xu > 1e18] == 0).all()
    # stationarity of the Lagrangian (unscaled; Ipopt scales by s_d >= 1 before comparing with 1e-8)
    g = nn.lagrangian_gradient(cd, fz, state, coeffs, z, lam, zl, zu)
    sd = np.maximum(100.0, (np.abs(lam).sum(axis=1) + zl.sum(axis=1) + zu.sum(axis=1)) / (6 * N + 2 * (4 * N + 4 * (N - 1)) / 2)) / 100.0
    gn = np.abs(g).max(axis=1) / sd
    # Ipopt converges on bounds relaxed by 1e-8*max(1,|b|) and then clips the point to the original bounds
    # (honor_original_bounds): where a steering or speed bound is active that moves the point by <= 5e-9 and the
    # gradient by (2 w_delta + 4 w_ddelta) * 5e-9 ~ 3e-5.  Everywhere else the result is stationary to rounding.
    nl = slice(2 * N, 7 * N - 1)      # psi, v, (cte, epsi: unbounded), delta -- the acceleration enters f and g linearly
    at_bound = (((z - xl) < 1e-7 * np.maximum(1, np.abs(xl)))[:, nl].any(axis=1)
                | ((xu - z) < 1e-7 * np.maximum(1, np.abs(xu)))[:, nl].any(axis=1))
    assert 0.0 < at_bound.mean() < 0.5
    assert gn.max() < 1e-4, gn.max()
    assert gn[~at_bound].max() < 2e-7, gn[~at_bound].max()
    assert np.median(gn) < 1e-10
    # complementarity with the ORIGINAL bounds (a point clipped onto its bound has slack 0; inside, the slack to
    # the original bound is smaller than the one the solver used).  Ipopt's tests: 1e-4 on the unscaled problem
    # (compl_inf_tol) and 1e-8 * s_c, s_c >= 1, on the problem whose objective is scaled by
    # sf = min(1, 100 / ||grad f(xi)||_inf) (gradient-based scaling at the start point xi).
    bounded_l, bounded_u = xl > -1e18, xu < 1e18
    cl = np.where(bounded_l, zl * (z - xl), 0.0).max(axis=1)
    cu = np.where(bounded_u, zu * (xu - z), 0.0).max(axis=1)
    assert max(cl.max(), cu.max()) < 1e-4
    xi = nn.start_point(cd, state)
    g0 = nn.lagrangian_gradient(cd, fz, state, coeffs, xi, np.zeros_like(lam), np.zeros_like(zl), np.zeros_like(zu))
    sf = np.minimum(1.0, 100.0 / np.abs(g0).max(axis=1))
    sc = np.maximum(100.0, sf * (zl.sum(axis=1) + zu.sum(axis=1)) / (4 * N + 4 * (N - 1))) / 100.0
    assert (sf * np.maximum(cl, cu) / sc).max() < 2e-8, (sf * np.maximum(cl, cu) / sc).max()


@pytest.mark.parametrize("N,dt", [(20, 0.05), (32, 0.05)])
def test_migration_longer_horizons(mpc, refdata, po, N, dt):
    """Horizons above 10 use the 20- and 32-stage instantiations (groups of 32 lanes in the coop kernel); the
    lane -> coop migration must stay invisible there too, also with per-problem weights."""
    js = dict(refdata["configs"]["stable"], N=N, dt=dt)
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = po.load_config_dict(js)
    B = mpc.LANE_MIN_BATCH
    b = mpc.workloads.batch_perturbed_states(B, 17, cd)
    rng = np.random.default_rng(3)
    W = np.tile(np.array(cd["weights"]), (B, 1))
    W[:, 3] = np.exp(rng.uniform(np.log(10), np.log(3000), B))
    S = mpc.Solver(cfg, 0)
    args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S.set_kernel(mpc.KERNEL_COOP)
    coop = S.solve_batch_host(*args, weights=W)
    S.set_kernel(mpc.KERNEL_LANE)
    S.set_handoff(0)
    lane = S.solve_batch_host(*args, weights=W)
    S.set_handoff(8)
    mig = S.solve_batch_host(*args, weights=W)
    S.close()
    for k in ("result", "traj_x", "status", "iters"):
        assert np.array_equal(lane[k], mig[k]), k
        assert np.array_equal(lane[k], coop[k]), k
    assert (lane["status"] == 1).mean() > 0.95
    # a few of them against the oracle with their own weights
    for i in range(0, B, B // 12):
        r = po.solve(po.make_config(dict(cd, weights=list(W[i]))), po.make_problem(b["state"][i], b["coeffs"][i], b["yaw_lo"][i], b["yaw_hi"][i]))
        if r["status"] == 1 and lane["status"][i] == 1:
            assert np.abs(lane["result"][i, :8] - r["result"][:8]).max() < ABS_TOL
            assert lane["result"][i, 8] == pytest.approx(r["result"][8], rel=REL_TOL)


def test_pathological_inputs_terminate(solver):
    """NaN / infinite inputs, an empty yaw interval and a huge cross-track error must end with a non-success
    status (the reference prints "Ipopt failed with <int>" and carries on, MPC.cpp:295-303) -- never hang, and
    never disturb the healthy problems solved in the same launch."""
    good = [0, 0, 0.0, 20.0, 0.3, 0.02]
    st = np.array([good, [0, 0, 0.0, np.nan, 0.1, 0.0], [0, 0, 0.0, 20.0, np.inf, 0.0], good, good, [0, 0, 0, 20.0, 1e6, 0.0], good])
    co = np.tile(np.array([0.3, -0.02, 0.001, 0.0, 0.0]), (7, 1))
    co[4, 2] = np.nan
    ylo = np.array([-0.1, -0.1, -0.1, 0.2, -0.1, -0.1, -0.1])
    yhi = np.array([0.4, 0.4, 0.4, -0.2, 0.4, 0.4, 0.4])          # problem 3: empty interval
    r = solver.solve_batch_host(st, co, ylo, yhi)
    assert r["status"][0] == 1 and r["status"][6] == 1
    assert np.array_equal(r["result"][0], r["result"][6])
    for i in (1, 2, 3, 4):
        assert r["status"][i] != 1, (i, r["status"][i])
    assert r["iters"].max() <= 3000


def test_pinned_host_inputs_are_read_in_place(mpc, stable_cfg, stable_cd):
    """mpc_solve_batch_host: inputs in page-locked host memory are read by the kernels directly (no staging copy),
    pageable ones are copied first -- same results, and a mix of both works."""
    import ctypes as C
    import torch
    B = 20000
    b = mpc.workloads.batch_perturbed_states(B, 41, stable_cd)
    S = mpc.Solver(stable_cfg, 0)
    ref = S.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).pin_memory().numpy()
    N = stable_cfg.N
    for pinned in ((True, True, True, True), (True, False, True, False)):
        arrs = [pin(a) if p else np.ascontiguousarray(a.T if a.ndim == 2 else a) for a, p in zip((b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"]), pinned)]
        res = np.zeros((9, B)); tx = np.zeros((N, B)); ty = np.zeros((N, B))
        st = np.zeros(B, dtype=np.int32); it = np.zeros(B, dtype=np.int32)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = mpc.lib().mpc_solve_batch_host(S._h, B, ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(arrs[3]), None, None, None,
                                            ptr(res), ptr(tx), ptr(ty), None, ptr(st), ptr(it))
        assert rc == 0
        assert np.array_equal(res.T, ref["result"]) and np.array_equal(st, ref["status"]) and np.array_equal(it, ref["iters"])
        assert np.array_equal(tx.T, ref["traj_x"])
    S.close()


def test_early_copy_back_gives_the_same_outputs(mpc, stable_cfg, stable_cd):
    """mpc_solve_batch_host with every output array in pinned memory: the device->host copies run beside the final
    launch of the chain and a small kernel rewrites the problems that launch finished (mpc_patch_outputs_kernel).
    Every byte of every output array must equal the plain path (pageable outputs, copies after the last launch) and
    the MPC_TAIL_LATE_COPY setting, call after call."""
    import ctypes as C
    import torch
    B = 65536
    b = mpc.workloads.batch_perturbed_states(B, 0, stable_cd)
    N = stable_cfg.N
    S = mpc.Solver(stable_cfg, 0)
    ref = S.solve_batch_host(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"], want_full=True)   # pageable outputs
    assert S.launches == 5
    ins = [np.ascontiguousarray(a.T if a.ndim == 2 else a) for a in (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])]
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    pinz = lambda shape, dt: torch.zeros(*shape, dtype=dt).pin_memory().numpy()
    for late, want_full, launches in ((False, True, 6), (False, False, 6), (True, True, 5), (False, True, 6)):
        S.set_tail(16, 3, resume_min=8192, late_copy=late)
        res, tx, ty = pinz((9, B), torch.float64), pinz((N, B), torch.float64), pinz((N, B), torch.float64)
        full = pinz((8 * N - 2, B), torch.float64) if want_full else None
        st, it = pinz((B,), torch.int32), pinz((B,), torch.int32)
        for a in (res, tx, ty, full):
            if a is not None:
                a.fill(np.nan)
        st.fill(-7); it.fill(-7)
        before = S.launches
        rc = mpc.lib().mpc_solve_batch_host(S._h, B, ptr(ins[0]), ptr(ins[1]), ptr(ins[2]), ptr(ins[3]), None, None, None,
                                            ptr(res), ptr(tx), ptr(ty), ptr(full) if want_full else None, ptr(st), ptr(it))
        assert rc == 0
        assert S.launches - before == launches, (late, S.launches - before)      # + the patch kernel on the early path
        assert np.array_equal(res.T, ref["result"]) and np.array_equal(st, ref["status"]) and np.array_equal(it, ref["iters"])
        assert np.array_equal(tx.T, ref["traj_x"]) and np.array_equal(ty.T, ref["traj_y"])
        if want_full:
            assert np.array_equal(full.T, ref["full"])
    assert sum(S.tail_counts(4)) > 0          # the final launch had problems to finish, so the patch kernel had work
    S.close()


def test_repeated_runs_give_the_same_bits(mpc, stable_cfg, stable_cd):
    """Which lane picks which problem up, which warps go sparse first and which launch of the chain finishes a
    problem all depend on the timing of atomics; none of it may show in the results."""
    import torch
    B = 65536
    b = mpc.workloads.batch_perturbed_states(B, 0, stable_cd)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    ins = [up(b["state"]), up(b["coeffs"]), up(b["yaw_lo"]), up(b["yaw_hi"])]
    S = mpc.Solver(stable_cfg, 0)          # default settings: AUTO kernel, tail packing on
    ref = None
    for rep in range(6):
        res = torch.zeros(9, B, dtype=torch.float64, device=dev)
        tx = torch.zeros(stable_cfg.N, B, dtype=torch.float64, device=dev)
        st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
        S.solve_batch_device(B, *ins, res, tx, None, None, st, it)
        torch.cuda.synchronize()
        if rep == 0:
            ref = (res, tx, st, it)
            assert S.launches == 5 and sum(S.tail_counts(4)) > 0       # main + 3 resume + final launches; problems were parked
        else:
            assert all(torch.equal(a, b_) for a, b_ in zip(ref, (res, tx, st, it))), rep
    S.close()


def test_c_abi_multi_gpu_entry_point(mpc, stable_cfg, stable_cd, refdata, tmp_path):
    """mpc_create_multi / mpc_solve_batch_multi: contiguous shards over several handles (one per device; here every
    visible device, and three handles dealt over them so that the sharding is exercised on a one-GPU box too), from
    Python and from a C++ program that uses nothing but include/mpc_b200.h.  Same bits as one handle."""
    import subprocess
    import torch
    nd = torch.cuda.device_count()
    b = mpc.workloads.batch_perturbed_states(20011, 5, stable_cd)      # odd size: uneven shards
    args = (b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"])
    S = mpc.Solver(stable_cfg, 0)
    ref = S.solve_batch_host(*args, want_full=True)
    S.close()
    for devices in ([0], list(range(nd)), [k % nd for k in range(3)]):
        M = mpc.MultiSolver(stable_cfg, devices)
        assert M.n_devices == len(devices)
        got = M.solve_batch_host(*args, want_full=True)
        M.close()
        for k in ("result", "traj_x", "traj_y", "full", "status", "iters"):
            assert np.array_equal(got[k], ref[k]), (devices, k)
    # ragged batch through the multi entry point: rows beyond a problem's horizon stay untouched
    Np = np.random.default_rng(2).choice([4, 7, 10], 3000).astype(np.int32)
    S = mpc.Solver(stable_cfg, 0)
    a3 = tuple(a[:3000] for a in args)
    ref = S.solve_batch_host(*a3, N_per=Np)
    S.close()
    M = mpc.MultiSolver(stable_cfg, [k % nd for k in range(2)])
    got = M.solve_batch_host(*a3, N_per=Np)
    M.close()
    for k in ("result", "traj_x", "traj_y", "status", "iters"):
        assert np.array_equal(got[k], ref[k]), k
    assert (got["traj_x"][Np == 4][:, 4:] == 0).all()
    with pytest.raises(mpc.MpcError):
        mpc.MultiSolver(stable_cfg, [])
    # the C++ caller
    exe = tmp_path / "test_multi_gpu"
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "carnd-mpc-project_b200")
    root = os.path.dirname(pkg)
    subprocess.run(["g++", "-O2", "-std=c++11", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include", "-o", str(exe),
                    os.path.join(root, "tests", "cpp", "test_multi_gpu.cpp"), "-L", pkg, "-lmpc_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                    "-Wl,-rpath," + pkg, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    cfgp = tmp_path / "config-stable.json"
    cfgp.write_text(json.dumps(refdata["configs"]["stable"]))
    for nh in (1, max(2, nd)):
        out = subprocess.run([str(exe), str(cfgp), "30000", str(nh)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        f = out.stdout.split()
        assert f[f.index("mismatches") + 1] == "0" and float(f[f.index("ok_frac") + 1]) > 0.99, out.stdout
