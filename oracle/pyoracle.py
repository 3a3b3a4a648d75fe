"""ctypes binding of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  PARITY UNPINNED (see oracle/mpc_oracle.h): the reference ships no golden vectors and
Ipopt/CppAD cannot be installed here.

Also holds the Python restatement of Config::load (/root/reference/src/utils/Config.cpp:31-87) used
to build an ``orc_config`` from a config-*.json dict, and the ``MPC::run`` pre-processing
(/root/reference/src/control/MPC.cpp:327-356) used to turn (pose, waypoints) into NLP inputs.
"""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmpc_oracle.so")

NMAX, NCOEF, NTAB = 64, 5, 16


class OrcConfig(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("n_steers", C.c_int), ("n_steer_speeds", C.c_int), ("max_iter", C.c_int),
        ("dt", C.c_double), ("Lf", C.c_double), ("cte_panic", C.c_double), ("epsi_panic", C.c_double),
        ("max_speed", C.c_double), ("max_steering", C.c_double), ("max_accel", C.c_double),
        ("max_decel", C.c_double), ("weights", C.c_double * 12), ("steers", C.c_double * NTAB),
        ("steer_speeds", C.c_double * NTAB), ("tol", C.c_double), ("mu_init", C.c_double),
        ("max_soc", C.c_int), ("obj_scaling", C.c_int), ("watchdog_trigger", C.c_int),
        ("filter_reset_trigger", C.c_int), ("tiny_step_tol", C.c_double),
    ]


class OrcProblem(C.Structure):
    _fields_ = [("state", C.c_double * 6), ("coeffs", C.c_double * NCOEF),
                ("yaw_lo", C.c_double), ("yaw_hi", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("iters", C.c_int), ("n_regularized", C.c_int), ("n_soc", C.c_int),
        ("n_backtrack", C.c_int), ("obj", C.c_double), ("result", C.c_double * 9),
        ("kkt_error", C.c_double), ("z", C.c_double * (8 * NMAX)), ("lam", C.c_double * (6 * NMAX)),
        ("zl", C.c_double * (8 * NMAX)), ("zu", C.c_double * (8 * NMAX)),
        ("n_resto", C.c_int), ("n_resto_iter", C.c_int), ("n_watchdog", C.c_int), ("n_tiny", C.c_int), ("n_filter_reset", C.c_int), ("n_filter_max", C.c_int),
    ]


def build(force=False):
    """Compile oracle/mpc_oracle.c with gcc (idempotent)."""
    src = os.path.join(HERE, "mpc_oracle.c")
    hdr = os.path.join(HERE, "mpc_oracle.h")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    subprocess.check_call(["make", "-C", HERE, "-B", "libmpc_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        L.orc_config_defaults.argtypes = [C.POINTER(OrcConfig)]
        L.orc_eval_f.restype = C.c_double
        L.orc_eval_f.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), dp]
        for nm in ("orc_eval_grad", "orc_eval_g", "orc_eval_jac"):
            getattr(L, nm).argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), dp, dp]
        L.orc_eval_hess.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), dp, C.c_double, dp, dp]
        L.orc_bounds.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), dp, dp, dp, dp, dp]
        L.orc_frozen.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), dp, dp, dp, dp]
        L.orc_solve.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), C.POINTER(OrcResult)]
        L.orc_solve_batch.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), C.c_int,
                                      C.POINTER(OrcResult), C.c_int]
        L.orc_solve_batch_compact.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcProblem), C.c_int,
                                              dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        _lib = L
    return _lib


# --------------------------------------------------------------------------------------------
# Config::load restated (Config.cpp:31-87)
# --------------------------------------------------------------------------------------------
def mph2mps(mph):  # utils.h:11-13
    return mph * 1609.34 / 3600.0


def load_config_dict(js):
    """config-*.json dict -> dict of SI-unit values, exactly Config.cpp:39-86."""
    c = {}
    c["N"] = int(js["N"])
    c["dt"] = float(js["dt"])
    c["max_accel"] = mph2mps(js["max acceleration"])
    c["max_decel"] = mph2mps(js["max deceleration"])
    c["max_steering"] = js["max steering"] * math.pi / 180
    c["max_speed"] = mph2mps(js["max speed"])
    scale = c["max_speed"] / mph2mps(100.0)
    c["latency"] = int(js["latency"])
    c["lookahead"] = c["latency"] * 1.0e-3
    c["max_fit_order"] = int(js["max polynomial fitting order"])
    c["max_fit_error"] = float(js["max polynomial fitting error"])
    c["ipopt_timeout"] = float(js["ipopt timeout"])
    c["Lf"] = float(js["Lf"])
    c["epsi_panic"] = float(js["epsi panic"])
    c["cte_panic"] = float(js["cte panic"])
    c["steer_adjust_thresh"] = float(js["steer adjustment threshold"])
    c["steer_adjust_ratio"] = min(max(float(js["steer adjustment ratio"]), 0.0), 0.1)
    c["weights"] = [float(w) for w in js["weights"]]
    assert len(c["weights"]) > 11
    c["steers"] = [float(s) for s in js["steers"]]

    def conv(v):
        return min(mph2mps(v), c["max_speed"]) if scale <= 1 else mph2mps(v) * scale

    c["steer_speeds"] = [conv(v) for v in js["steer speeds"]]
    c["yaw_changes"] = [float(s) for s in js["yaw changes"]]
    c["yaw_change_speeds"] = [conv(v) for v in js["yaw change speeds"]]
    return c


def make_config(cd, **over):
    """dict from load_config_dict -> OrcConfig (solver knobs at Ipopt defaults unless overridden)."""
    cfg = OrcConfig()
    lib().orc_config_defaults(C.byref(cfg))
    cfg.N = cd["N"]
    cfg.dt, cfg.Lf = cd["dt"], cd["Lf"]
    cfg.cte_panic, cfg.epsi_panic = cd["cte_panic"], cd["epsi_panic"]
    cfg.max_speed, cfg.max_steering = cd["max_speed"], cd["max_steering"]
    cfg.max_accel, cfg.max_decel = cd["max_accel"], cd["max_decel"]
    for i in range(12):
        cfg.weights[i] = cd["weights"][i]
    cfg.n_steers = len(cd["steers"])
    cfg.n_steer_speeds = len(cd["steer_speeds"])
    for i, v in enumerate(cd["steers"]):
        cfg.steers[i] = v
    for i, v in enumerate(cd["steer_speeds"]):
        cfg.steer_speeds[i] = v
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def make_problem(state, coeffs, yaw_lo, yaw_hi):
    p = OrcProblem()
    for i in range(6):
        p.state[i] = float(state[i])
    for i in range(NCOEF):
        p.coeffs[i] = float(coeffs[i]) if i < len(coeffs) else 0.0
    p.yaw_lo, p.yaw_hi = float(yaw_lo), float(yaw_hi)
    return p


def problems_from_arrays(state, coeffs, yaw_lo, yaw_hi):
    """state [B,6], coeffs [B,5], yaw_lo/hi [B] -> ctypes array of OrcProblem."""
    B = state.shape[0]
    arr = (OrcProblem * B)()
    buf = np.frombuffer(arr, dtype=np.float64).reshape(B, 13)
    buf[:, 0:6] = state
    buf[:, 6:11] = coeffs
    buf[:, 11] = yaw_lo
    buf[:, 12] = yaw_hi
    return arr


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def eval_f(cfg, p, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    return lib().orc_eval_f(C.byref(cfg), C.byref(p), _dp(z))


def eval_grad(cfg, p, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    g = np.zeros(8 * cfg.N - 2)
    lib().orc_eval_grad(C.byref(cfg), C.byref(p), _dp(z), _dp(g))
    return g


def eval_g(cfg, p, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    g = np.zeros(6 * cfg.N)
    lib().orc_eval_g(C.byref(cfg), C.byref(p), _dp(z), _dp(g))
    return g


def eval_jac(cfg, p, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    J = np.zeros((6 * cfg.N, 8 * cfg.N - 2))
    lib().orc_eval_jac(C.byref(cfg), C.byref(p), _dp(z), _dp(J))
    return J


def eval_hess(cfg, p, z, sigma, lam):
    z = np.ascontiguousarray(z, dtype=np.float64)
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    n = 8 * cfg.N - 2
    H = np.zeros((n, n))
    lib().orc_eval_hess(C.byref(cfg), C.byref(p), _dp(z), sigma, _dp(lam), _dp(H))
    return H


def bounds(cfg, p):
    n, m = 8 * cfg.N - 2, 6 * cfg.N
    xl, xu, xi = np.zeros(n), np.zeros(n), np.zeros(n)
    gl, gu = np.zeros(m), np.zeros(m)
    lib().orc_bounds(C.byref(cfg), C.byref(p), _dp(xl), _dp(xu), _dp(gl), _dp(gu), _dp(xi))
    return xl, xu, gl, gu, xi


def solve(cfg, p):
    r = OrcResult()
    lib().orc_solve(C.byref(cfg), C.byref(p), C.byref(r))
    n, m = 8 * cfg.N - 2, 6 * cfg.N
    return {
        "status": r.status, "iters": r.iters, "n_regularized": r.n_regularized, "n_soc": r.n_soc,
        "n_backtrack": r.n_backtrack, "obj": r.obj, "result": np.array(r.result[:9]),
        "kkt_error": r.kkt_error, "z": np.array(r.z[:n]), "lam": np.array(r.lam[:m]),
        "zl": np.array(r.zl[:n]), "zu": np.array(r.zu[:n]),
        "n_resto": r.n_resto, "n_resto_iter": r.n_resto_iter, "n_watchdog": r.n_watchdog, "n_tiny": r.n_tiny, "n_filter_reset": r.n_filter_reset, "n_filter_max": r.n_filter_max,
    }


def solve_batch(cfg, probs, n_threads=1):
    """probs: ctypes array of OrcProblem -> dict of numpy arrays (compact outputs)."""
    B = len(probs)
    N = cfg.N
    res = np.zeros((B, 9))
    tx, ty = np.zeros((B, N)), np.zeros((B, N))
    st, it = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    lib().orc_solve_batch_compact(C.byref(cfg), probs, B, _dp(res), _dp(tx), _dp(ty),
                                  st.ctypes.data_as(ip), it.ctypes.data_as(ip), n_threads)
    return {"result": res, "traj_x": tx, "traj_y": ty, "status": st, "iters": it}


# --------------------------------------------------------------------------------------------
# MPC::run pre-processing restated (MPC.cpp:327-356, RoadGeometry.cpp:18-39,57-61, utils.cpp:10-29)
# --------------------------------------------------------------------------------------------
def polyfit_qr(x, y, order):
    """utils.cpp:10-29: Vandermonde + Householder QR least squares (numpy lstsq is QR/SVD based;
    agreement with Eigen's householderQr is to rounding, which is all the NLP needs)."""
    A = np.vander(np.asarray(x, dtype=np.float64), order + 1, increasing=True)
    q, r = np.linalg.qr(A)
    return np.linalg.solve(r, q.T @ np.asarray(y, dtype=np.float64))


def road_fit(x, y, max_order, max_err):
    """RoadGeometry::fit, RoadGeometry.cpp:18-39: order 2,3,.. while err > max_err and order < max_order."""
    order = 2
    while True:
        c = polyfit_qr(x, y, order)
        order += 1
        err = float(np.sum((np.asarray(y) - np.polynomial.polynomial.polyval(np.asarray(x), c)) ** 2))
        if not (err > max_err and order < max_order):
            break
    return c, err


def preprocess(cd, pose, ptsx, ptsy):
    """(x, y, psi, v) + global waypoints -> (state, coeffs[5], yaw_lo, yaw_hi, extras) per MPC::run."""
    px, py, psi, v = pose
    cs, sn = math.cos(psi), math.sin(psi)
    vx = np.asarray(ptsx, dtype=np.float64) - px
    vy = np.asarray(ptsy, dtype=np.float64) - py
    tx = vx * cs + vy * sn          # Vehicle.cpp:105-114
    ty = vy * cs - vx * sn
    c, err = road_fit(tx, ty, cd["max_fit_order"], cd["max_fit_error"])
    coeffs = np.zeros(NCOEF)
    coeffs[: len(c)] = c
    cte = float(np.polynomial.polynomial.polyval(0.0, c))    # MPC.cpp:334
    epsi = -math.atan(c[1])                                  # MPC.cpp:336

    def orient(xx):
        d = sum(i * c[i] * xx ** (i - 1) for i in range(1, len(c)))
        return math.atan(d)

    # computeOrientationChange(0, x_last): dir = x_last - 0; negative dir adds pi to both -> cancels
    # unless normalizeAngle wraps; restate exactly (RoadGeometry.cpp:41-61)
    def orient_dir(xx, direction):
        p = orient(xx)
        if direction < 0:
            p = p + math.pi
            while p >= math.pi:
                p -= 2 * math.pi
            while p < -math.pi:
                p += 2 * math.pi
        return p

    xl_, xf_ = float(tx[-1]), float(tx[0])
    myc = (orient_dir(xl_, xl_ - 0.0) - orient_dir(0.0, xl_ - 0.0)) * (xl_ - xf_) / xl_   # MPC.cpp:339
    if myc < 0:                                              # MPC.cpp:345-352
        yaw_lo, yaw_hi = myc, 0.1
    else:
        yaw_lo, yaw_hi = -0.1, myc
    state = np.array([0.0, 0.0, 0.0, v, cte, epsi])
    return state, coeffs, yaw_lo, yaw_hi, {"fit_err": err, "order": len(c) - 1, "myc": myc,
                                           "tx": tx, "ty": ty}
