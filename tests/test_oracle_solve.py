"""The oracle's interior-point solve: KKT certificates, an independent SciPy solve, golden vectors."""
import json
import os

import numpy as np
import pytest

import nlp_numpy as nn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kkt_certificate(po, cfg, p, r):
    """Solver-independent proof that (z, lam, zl, zu) is a KKT point of the reference NLP."""
    z = r["z"]
    xl, xu, gl, gu, xi = po.bounds(cfg, p)
    g = po.eval_grad(cfg, p, z)
    J = po.eval_jac(cfg, p, z)
    c = po.eval_g(cfg, p, z) - gl
    has_l, has_u = xl > -1e19, xu < 1e19
    stat = g + J.T @ r["lam"] - np.where(has_l, r["zl"], 0) + np.where(has_u, r["zu"], 0)
    sl = np.where(has_l, z - xl, 1.0)
    su = np.where(has_u, xu - z, 1.0)
    return {
        "stationarity": np.abs(stat).max(), "feasibility": np.abs(c).max(),
        "bound_violation": max(0.0, (-sl[has_l]).max(), (-su[has_u]).max()),
        "complementarity": max(np.abs(sl * r["zl"])[has_l].max(), np.abs(su * r["zu"])[has_u].max()),
        "min_mult": min(r["zl"][has_l].min(), r["zu"][has_u].min()),
    }


def _testcpp(po, cd, refdata, k=0):
    fx = refdata["test_cpp_fixtures"][k]
    return po.preprocess(cd, (fx["x"], fx["y"], fx["psi"], fx["v"]), fx["ptsx"], fx["ptsy"])


def test_testcpp_first_solve_matches_survey_values(po, stable_cd, refdata):
    """SURVEY.md App. C (indicative 1e-5 values from two throw-away solvers) and the reference's
    own plot examples/10-01-2.png (CTE -0.107, ePsi 0.0284, delta 0.0024, v 27.13)."""
    cfg = po.make_config(stable_cd)
    state, coeffs, ylo, yhi, ex = _testcpp(po, stable_cd, refdata)
    assert ex["order"] == 2 and ex["fit_err"] == pytest.approx(0.121, abs=1e-3)
    assert coeffs[:3] == pytest.approx([-0.1765561, -0.02599232, 0.00302917], abs=1e-7)
    assert (ylo, yhi) == (-0.1, pytest.approx(0.4843454, abs=1e-6))
    r = po.solve(cfg, po.make_problem(state, coeffs, ylo, yhi))
    assert r["status"] == 1 and r["iters"] == 10 and r["n_regularized"] == 0
    assert r["result"][:8] == pytest.approx(
        [2.66806, 0, 0.00241174, 27.1276389, -0.1072304, 0.0283982, 0.002413, 4.4703889], abs=2e-6)
    assert r["obj"] == pytest.approx(6243.443671, rel=1e-9)
    N = cfg.N
    assert r["z"][6 * N:7 * N - 1] == pytest.approx(
        [0.002413, 0.005368, 0.008670, 0.011130, 0.012390, 0.012540, 0.011898, 0.010908, 0.010113], abs=2e-6)
    assert np.all(np.abs(r["z"][7 * N - 1:] - stable_cd["max_accel"]) < 1e-9)   # a at its upper bound


def test_kkt_certificates(po, stable_cd, refdata, mpc):
    cfg = po.make_config(stable_cd)
    b = mpc.workloads.batch_perturbed_states(48, 11, stable_cd)
    for i in range(48):
        p = po.make_problem(b["state"][i], b["coeffs"][i], b["yaw_lo"][i], b["yaw_hi"][i])
        r = po.solve(cfg, p)
        assert r["status"] == 1
        k = kkt_certificate(po, cfg, p, r)
        assert k["stationarity"] < 1e-6, (i, k)
        assert k["feasibility"] < 2e-7, (i, k)   # honor_original_bounds clips by <= 5e-7 after convergence
        assert k["bound_violation"] == 0.0, (i, k)
        assert k["complementarity"] < 1e-6, (i, k)
        assert k["min_mult"] >= 0.0


@pytest.mark.parametrize("case", ["testcpp", 3, 8, 21])
def test_same_local_optimum_as_scipy_slsqp(po, stable_cd, refdata, mpc, case):
    """Independent algorithm (SLSQP, cold start at xi) on the independent numpy NLP statement."""
    from scipy.optimize import minimize
    cd = stable_cd
    cfg = po.make_config(cd)
    if case == "testcpp":
        state, coeffs, ylo, yhi, _ = _testcpp(po, cd, refdata)
    else:
        b = mpc.workloads.batch_perturbed_states(32, 5, cd)
        state, coeffs, ylo, yhi = b["state"][case], b["coeffs"][case], b["yaw_lo"][case], b["yaw_hi"][case]
    r = po.solve(cfg, po.make_problem(state, coeffs, ylo, yhi))
    assert r["status"] == 1
    fz = nn.frozen(cd, state[None])
    xl, xu = nn.var_bounds(cd, np.array([ylo]), np.array([yhi]))
    bnds = [(None if l <= -1e19 else l, None if u >= 1e19 else u) for l, u in zip(xl[0], xu[0])]
    fun = lambda z: nn.objective(cd, fz, z[None])[0]
    con = lambda z: nn.constraints(cd, state[None], coeffs[None], z[None])[0]
    res = minimize(fun, nn.start_point(cd, state[None])[0], method="SLSQP", bounds=bnds,
                   constraints=[{"type": "eq", "fun": con}], options={"ftol": 1e-15, "maxiter": 1000})
    assert np.abs(con(res.x)).max() < 1e-7
    assert res.fun == pytest.approx(r["obj"], rel=1e-6)
    assert np.abs(res.x - r["z"]).max() < 1e-4


def test_golden_scenarios(po, stable_cd, refdata):
    """Committed golden vectors (tests/golden/oracle_golden.json, made by make_oracle_golden.py once
    the oracle agreed with SciPy and passed its KKT certificates)."""
    with open(os.path.join(GOLD, "oracle_golden.json")) as f:
        gold = json.load(f)
    cfg = po.make_config(stable_cd)
    for sc in gold["scenarios"]:
        state = np.array(sc["state0"])
        coeffs, ylo, yhi = np.array(sc["coeffs"]), sc["yaw_lo"], sc["yaw_hi"]
        for step in sc["steps"]:
            r = po.solve(cfg, po.make_problem(state, coeffs, ylo, yhi))
            assert r["status"] == step["status"]
            assert np.allclose(r["result"][:8], step["result"][:8], rtol=0, atol=1e-9)
            assert r["result"][8] == pytest.approx(step["result"][8], rel=1e-10)
            state = r["result"][:6].copy()
