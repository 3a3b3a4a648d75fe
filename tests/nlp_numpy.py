"""Independent numpy statement of the reference NLP, written from the text of
/root/reference/src/control/MPC.cpp:50-154 (objective + constraints, frozen branches) and :204-281
(bounds, start point).  It shares no code with oracle/mpc_oracle.c and is vectorised over a batch,
so it serves (a) SciPy cross-checks of the oracle and (b) KKT / feasibility property checks of
GPU results at full batch size.  Variable layout = the reference's (MPC.cpp:189-196)."""
import numpy as np


def speed_target(cd, angle, mx):
    y = np.abs(angle)
    steers, ss = cd["steers"], cd["steer_speeds"]
    out = np.full(np.shape(y), min(ss[-1], mx), dtype=np.float64)
    done = np.zeros(np.shape(y), dtype=bool)
    for i, s in enumerate(steers):
        hit = (~done) & (y <= s)
        val = min(ss[i], mx) if len(ss) > i else min(ss[-1], mx)
        out = np.where(hit, val, out)
        done |= hit
    return out


def frozen(cd, state, weights=None):
    """per-problem frozen constants: dict of arrays [B, N]."""
    B = state.shape[0]
    N = cd["N"]
    w = np.broadcast_to(np.asarray(cd["weights"], dtype=np.float64), (B, 12)) if weights is None else weights
    xi = np.zeros((B, 4, N))           # psi, v, cte, epsi at the start point
    xi[:, 0, 0], xi[:, 1, 0], xi[:, 2, 0], xi[:, 3, 0] = state[:, 2], state[:, 3], state[:, 4], state[:, 5]
    wc = np.where(np.abs(xi[:, 2]) < cd["cte_panic"], w[:, 0:1], w[:, 11:12])
    we = np.where(np.abs(xi[:, 3]) > cd["epsi_panic"], w[:, 10:11], w[:, 1:2])
    vref = speed_target(cd, xi[:, 0], cd["max_speed"])
    nvw = np.where(xi[:, 1] < 0, w[:, 9:10], 0.0)
    return {"wc": wc, "we": we, "vref": vref, "nvw": nvw, "w": w}


def split(z, N):
    o = np.cumsum([0, N, N, N, N, N, N, N - 1])
    return [z[..., o[k]:o[k] + (N if k < 6 else N - 1)] for k in range(8)]


def objective(cd, fz, z):
    """z [B, 8N-2] -> f [B] (MPC.cpp:68-114 with the branches frozen at xi)."""
    N = cd["N"]
    x, y, psi, v, cte, epsi, delta, a = split(z, N)
    w = fz["w"]
    f = np.sum(fz["wc"] * cte ** 2 + fz["we"] * epsi ** 2 + w[:, 2:3] * (v - fz["vref"]) ** 2
               + fz["nvw"] * v ** 2, axis=1)
    f += np.sum(w[:, 3:4] * delta ** 2, axis=1)
    f += np.sum(w[:, 4:5] * (delta[:, 1:] - delta[:, :-1]) ** 2, axis=1)
    return f


def polyval(c, x):
    r = np.zeros_like(x)
    for i in range(c.shape[1] - 1, -1, -1):
        r = r * x + c[:, i:i + 1]
    return r


def polyder(c, x):
    r = np.zeros_like(x)
    for i in range(c.shape[1] - 1, 0, -1):
        r = r * x + i * c[:, i:i + 1]
    return r


def constraints(cd, state, coeffs, z, dt=None):
    """g(z) - gl, [B, 6N] in the reference's row order (MPC.cpp:116-153, 261-281)."""
    N = cd["N"]
    dt = cd["dt"] if dt is None else dt
    Lf = cd["Lf"]
    x, y, psi, v, cte, epsi, delta, a = split(z, N)
    x0, y0, p0, v0, e0 = x[:, :-1], y[:, :-1], psi[:, :-1], v[:, :-1], epsi[:, :-1]
    vdt = v0 * dt
    pn = p0 + delta * vdt / Lf
    rows = [
        np.concatenate([x[:, :1] - state[:, 0:1], x[:, 1:] - (x0 + np.cos(p0) * vdt)], axis=1),
        np.concatenate([y[:, :1] - state[:, 1:2], y[:, 1:] - (y0 + np.sin(p0) * vdt)], axis=1),
        np.concatenate([psi[:, :1] - state[:, 2:3], psi[:, 1:] - pn], axis=1),
        np.concatenate([v[:, :1] - state[:, 3:4], v[:, 1:] - (v0 + a * dt)], axis=1),
        np.concatenate([cte[:, :1] - state[:, 4:5],
                        cte[:, 1:] - ((polyval(coeffs, x0) - y0) + np.sin(e0) * vdt)], axis=1),
        np.concatenate([epsi[:, :1] - state[:, 5:6],
                        epsi[:, 1:] - (pn - np.arctan(polyder(coeffs, x0)))], axis=1),
    ]
    return np.concatenate(rows, axis=1)


def var_bounds(cd, yaw_lo, yaw_hi):
    """xl, xu [B, 8N-2] (MPC.cpp:220-257); +-1e19 = unbounded."""
    N = cd["N"]
    B = yaw_lo.shape[0]
    big = 1.0e19
    one = np.ones((B, N))
    xl = np.concatenate([-big * one, -big * one, yaw_lo[:, None] * one, -cd["max_speed"] * one,
                         -big * one, -big * one, -cd["max_steering"] * one[:, :N - 1],
                         cd["max_decel"] * one[:, :N - 1]], axis=1)
    xu = np.concatenate([big * one, big * one, yaw_hi[:, None] * one, cd["max_speed"] * one,
                         big * one, big * one, cd["max_steering"] * one[:, :N - 1],
                         cd["max_accel"] * one[:, :N - 1]], axis=1)
    return xl, xu


def start_point(cd, state):
    N = cd["N"]
    xi = np.zeros((state.shape[0], 8 * N - 2))
    for k in range(6):
        xi[:, k * N] = state[:, k]
    return xi


def lagrangian_gradient(cd, fz, state, coeffs, z, lam, zl, zu):
    """grad_z [ f(z) + lam^T g(z) ] - zl + zu, [B, 8N-2], by complex-step differentiation of the statements
    above (exact to rounding: f and g are analytic) -- no hand-written derivative is trusted here."""
    B, n = z.shape
    h = 1e-30
    out = np.zeros((B, n))
    zc = z.astype(np.complex128)
    for k in range(n):
        zk = zc.copy()
        zk[:, k] += 1j * h
        L = objective(cd, fz, zk) + np.sum(lam * constraints(cd, state, coeffs, zk), axis=1)
        out[:, k] = L.imag / h
    return out - zl + zu
