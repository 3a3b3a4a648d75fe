"""The oracle's NLP pieces against an independent numpy statement of FG_eval and finite differences."""
import numpy as np
import pytest

import nlp_numpy as nn


def _problem(po, cd, seed):
    rng = np.random.default_rng(seed)
    state = np.array([rng.normal(0, 2), rng.normal(0, 1), rng.uniform(-0.3, 0.3), rng.uniform(-5, 45),
                      rng.uniform(-2, 2), rng.uniform(-0.3, 0.3)])
    coeffs = np.array([rng.normal(0, 1), rng.normal(0, 0.1), rng.normal(0, 0.01), rng.normal(0, 1e-4),
                       rng.normal(0, 1e-6)])
    return state, coeffs, -0.1, 0.4


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("N", [2, 3, 10, 25])
def test_f_and_g_match_numpy_statement(po, stable_cd, seed, N):
    cd = dict(stable_cd, N=N)
    cfg = po.make_config(cd)
    state, coeffs, ylo, yhi = _problem(po, cd, seed)
    p = po.make_problem(state, coeffs, ylo, yhi)
    rng = np.random.default_rng(100 + seed)
    z = rng.normal(0, 1, 8 * N - 2)
    fz = nn.frozen(cd, state[None])
    f_np = nn.objective(cd, fz, z[None])[0]
    assert po.eval_f(cfg, p, z) == pytest.approx(f_np, rel=1e-13)
    xl, xu, gl, gu, xi = po.bounds(cfg, p)
    c_np = nn.constraints(cd, state[None], coeffs[None], z[None])[0]
    assert np.allclose(po.eval_g(cfg, p, z) - gl, c_np, rtol=0, atol=1e-12)
    assert np.array_equal(gl, gu)
    xl_np, xu_np = nn.var_bounds(cd, np.array([ylo]), np.array([yhi]))
    assert np.array_equal(xl, xl_np[0]) and np.array_equal(xu, xu_np[0])
    assert np.array_equal(xi, nn.start_point(cd, state[None])[0])


def test_frozen_branches_follow_the_start_point(po, stable_cd):
    """MPC.cpp:72,79,87,89 are decided at xi: panic weights / v_ref / negative-speed term only via stage 0."""
    cd = stable_cd
    cfg = po.make_config(cd)
    w = cd["weights"]
    for cte0, epsi0, psi0, v0 in [(0.5, 0.05, 0.0, 10.0), (0.9, 0.2, 0.06, -1.0), (-0.8, -0.1, 0.31, 3.0)]:
        state = np.array([0, 0, psi0, v0, cte0, epsi0], dtype=float)
        p = po.make_problem(state, [0, 0, 0], -0.1, 0.1)
        N = cd["N"]
        import ctypes as C
        wc, we, vref, nvw = (np.zeros(N) for _ in range(4))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        po.lib().orc_frozen(C.byref(cfg), C.byref(p), dp(wc), dp(we), dp(vref), dp(nvw))
        assert wc[0] == (w[0] if abs(cte0) < cd["cte_panic"] else w[11])
        assert we[0] == (w[10] if abs(epsi0) > cd["epsi_panic"] else w[1])
        assert nvw[0] == (w[9] if v0 < 0 else 0.0)
        assert np.all(wc[1:] == w[0]) and np.all(we[1:] == w[1]) and np.all(nvw[1:] == 0)
        assert np.all(vref[1:] == min(cd["steer_speeds"][0], cd["max_speed"]))
        fz = nn.frozen(cd, state[None])
        assert np.array_equal(vref, fz["vref"][0])


def test_accel_weights_have_no_effect(po, stable_cd):
    """submission-report.md:317-319: the acceleration weight 'has no visible impact' -- the frozen
    branch a>0 is false at xi, so weights 6,7,8 must not change f at all."""
    cd = stable_cd
    state, coeffs, ylo, yhi = _problem(po, cd, 5)
    p = po.make_problem(state, coeffs, ylo, yhi)
    z = np.random.default_rng(7).normal(0, 1, 8 * cd["N"] - 2)
    f0 = po.eval_f(po.make_config(cd), p, z)
    cd2 = dict(cd, weights=list(cd["weights"]))
    cd2["weights"][6], cd2["weights"][7], cd2["weights"][8] = 1e4, 1e4, 1.0
    assert po.eval_f(po.make_config(cd2), p, z) == f0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gradient_jacobian_hessian_vs_finite_differences(po, stable_cd, seed):
    cd = dict(stable_cd, N=6)
    N = cd["N"]
    cfg = po.make_config(cd)
    state, coeffs, ylo, yhi = _problem(po, cd, seed)
    p = po.make_problem(state, coeffs, ylo, yhi)
    rng = np.random.default_rng(seed)
    z = rng.normal(0, 0.5, 8 * N - 2)
    lam = rng.normal(0, 1, 6 * N)
    n = z.size
    h = 1e-6
    g = po.eval_grad(cfg, p, z)
    J = po.eval_jac(cfg, p, z)
    H = po.eval_hess(cfg, p, z, 0.7, lam)
    g_fd = np.zeros(n)
    J_fd = np.zeros_like(J)
    H_fd = np.zeros_like(H)
    for i in range(n):
        e = np.zeros(n)
        e[i] = h
        g_fd[i] = (po.eval_f(cfg, p, z + e) - po.eval_f(cfg, p, z - e)) / (2 * h)
        J_fd[:, i] = (po.eval_g(cfg, p, z + e) - po.eval_g(cfg, p, z - e)) / (2 * h)
        gl_p = 0.7 * po.eval_grad(cfg, p, z + e) + po.eval_jac(cfg, p, z + e).T @ lam
        gl_m = 0.7 * po.eval_grad(cfg, p, z - e) + po.eval_jac(cfg, p, z - e).T @ lam
        H_fd[:, i] = (gl_p - gl_m) / (2 * h)
    assert np.allclose(g, g_fd, rtol=1e-6, atol=1e-6)
    assert np.allclose(J, J_fd, rtol=1e-6, atol=1e-7)
    assert np.allclose(H, H_fd, rtol=1e-5, atol=1e-5)
    assert np.allclose(H, H.T)
    # Jacobian sparsity: 25 nnz per transition + 6
    assert np.count_nonzero(J) <= 25 * (N - 1) + 6
