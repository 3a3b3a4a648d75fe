"""Synthetic MPC workloads -- host-side numpy, deterministic given a seed.

Turns random poses on the lake-track centre line into NLP inputs exactly the way the reference's
``MPC::run`` does before it calls ``MPC::solve`` (/root/reference/src/control/MPC.cpp:327-356):
global->vehicle transform (Vehicle.cpp:105-114), adaptive-order polynomial fit
(RoadGeometry.cpp:18-39, utils.cpp:10-29), cte/epsi (MPC.cpp:334-336) and the yaw bounds
(MPC.cpp:339-352).  The arrays it returns are the inputs of ``mpc_solve_batch``.
"""
import json
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(_HERE, "data", "reference_data.json")   # config-*.json knob sets, lake-track waypoints, test.cpp fixtures
NCOEF = 5


def reference_data():
    with open(_DATA) as f:
        return json.load(f)


def mph2mps(mph):
    return mph * 1609.34 / 3600.0


def _fit_batch(tx, ty, max_order, max_err):
    """Adaptive-order least-squares fit for a batch: tx, ty [B, P] -> coeffs [B, 5], order [B], err [B]."""
    B = tx.shape[0]
    coeffs = np.zeros((B, NCOEF))
    order_out = np.zeros(B, dtype=np.int32)
    err_out = np.zeros(B)
    todo = np.ones(B, dtype=bool)
    order = 2
    while True:
        idx = np.nonzero(todo)[0]
        if idx.size == 0:
            break
        A = tx[idx, :, None] ** np.arange(order + 1)[None, None, :]
        q, r = np.linalg.qr(A)
        c = np.linalg.solve(r, np.einsum("bpk,bp->bk", q, ty[idx])[..., None])[..., 0]
        fit = np.einsum("bpk,bk->bp", A, c)
        err = np.sum((ty[idx] - fit) ** 2, axis=1)
        coeffs[idx] = 0.0
        coeffs[idx, : order + 1] = c
        order_out[idx] = order
        err_out[idx] = err
        order += 1
        keep = (err > max_err) & (order < max_order)   # do { } while (err > max && order < maxOrder)
        todo[:] = False
        todo[idx[keep]] = True
    return coeffs, order_out, err_out


def _polyder_at(coeffs, x):
    d = np.zeros_like(x)
    for i in range(NCOEF - 1, 0, -1):
        d = d * x + i * coeffs[:, i]
    return d


def preprocess_batch(cfg, px, py, psi, v, wx, wy):
    """MPC::run pre-processing for B vehicles.  wx, wy: [B, P] global waypoints.

    cfg: dict with 'max_fit_order', 'max_fit_error'.  Returns dict of NLP inputs."""
    cs, sn = np.cos(psi)[:, None], np.sin(psi)[:, None]
    vx, vy = wx - px[:, None], wy - py[:, None]
    tx = vx * cs + vy * sn
    ty = vy * cs - vx * sn
    coeffs, order, err = _fit_batch(tx, ty, cfg["max_fit_order"], cfg["max_fit_error"])
    cte = coeffs[:, 0].copy()
    epsi = -np.arctan(coeffs[:, 1])
    xl, xf = tx[:, -1], tx[:, 0]

    def orient(x, direction):
        p = np.arctan(_polyder_at(coeffs, x))
        q = p + math.pi
        q = np.where(q >= math.pi, q - 2 * math.pi, q)
        q = np.where(q < -math.pi, q + 2 * math.pi, q)
        return np.where(direction < 0, q, p)

    myc = (orient(xl, xl) - orient(np.zeros_like(xl), xl)) * (xl - xf) / xl
    yaw_lo = np.where(myc < 0, myc, -0.1)
    yaw_hi = np.where(myc < 0, 0.1, myc)
    B = px.shape[0]
    state = np.zeros((B, 6))
    state[:, 3] = v
    state[:, 4] = cte
    state[:, 5] = epsi
    return {"state": state, "coeffs": coeffs, "yaw_lo": yaw_lo, "yaw_hi": yaw_hi,
            "fit_order": order, "fit_err": err, "myc": myc}


def batch_perturbed_states(B, seed, cfg, n_pts=6):
    """B perturbed poses along the lake track -> NLP inputs (config 2)."""
    rd = reference_data()
    wx_all = np.asarray(rd["waypoints"]["x"])
    wy_all = np.asarray(rd["waypoints"]["y"])
    W = wx_all.shape[0]
    rng = np.random.default_rng(seed)
    j = rng.integers(0, W, size=B)
    t = rng.random(B)
    dpsi = rng.uniform(-0.15, 0.15, size=B)
    lat = rng.uniform(-1.2, 1.2, size=B)
    v = rng.uniform(5.0, 45.0, size=B)
    j1 = (j + 1) % W
    sx, sy = wx_all[j1] - wx_all[j], wy_all[j1] - wy_all[j]
    heading = np.arctan2(sy, sx)
    # point on the segment, shifted sideways (left of travel = +)
    px = wx_all[j] + t * sx - lat * np.sin(heading)
    py = wy_all[j] + t * sy + lat * np.cos(heading)
    psi = heading + dpsi
    win = (j[:, None] + np.arange(n_pts)[None, :]) % W
    out = preprocess_batch(cfg, px, py, psi, v, wx_all[win], wy_all[win])
    out.update({"px": px, "py": py, "psi": psi, "v": v, "segment": j})
    return out
