"""ctypes view of oracle/_ref/libmpc_ref.so: the reference's OWN controller sources (MPC.cpp, Vehicle.cpp,
RoadGeometry.cpp, utils.cpp, Config.cpp, compiled unmodified from /root/reference by
oracle/ref_shim/Makefile) against the CppAD / Ipopt stand-ins.  TEST INFRASTRUCTURE ONLY.

The library keeps the reference's global mutable Config statics, so it is not thread-safe: one solve
at a time."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libmpc_ref.so")
_lib = None


def available():
    return os.path.exists(LIB)


def build():
    """Only possible where /root/reference exists (the build container)."""
    if not os.path.isdir("/root/reference"):
        return available()
    out = subprocess.run(["make", "-C", os.path.join(HERE, "ref_shim")], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + out.stdout[-3000:] + out.stderr[-3000:])
    return True


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ref_config_load.argtypes = [C.c_char_p]
        L.ref_config_get.argtypes = [dp] * 6 + [ip]
        L.ref_config_set_weights.argtypes = [dp, C.c_int]
        L.ref_config_set_horizon.argtypes = [C.c_int, C.c_double]
        L.ref_solve.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, dp, dp, dp, ip, ip]
        L.ref_run.argtypes = [C.c_double] * 6 + [dp, dp, C.c_int, dp, dp, dp, dp, ip, dp, dp, ip, ip]
        L.ref_vehicle_move.argtypes = [dp, C.c_double, C.c_double]
        L.ref_compute_throttle.argtypes = [C.c_double] * 4
        L.ref_compute_throttle.restype = C.c_double
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def config_load(js):
    """Config::load on a config-*.json given as a dict (written to a temp file)."""
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(js, f)
        path = f.name
    try:
        if lib().ref_config_load(path.encode()) != 0:
            raise RuntimeError("Config::load failed")
    finally:
        os.unlink(path)


def config_get():
    scal, w = np.zeros(16), np.zeros(12)
    tabs = [np.zeros(16) for _ in range(4)]
    cnt = (C.c_int * 4)()
    lib().ref_config_get(_dp(scal), _dp(w), *[_dp(t) for t in tabs], cnt)
    names = ["N", "dt", "Lf", "cte_panic", "epsi_panic", "max_speed", "max_steering", "max_accel", "max_decel",
             "max_fit_order", "max_fit_error", "latency", "lookahead", "ipopt_timeout", "steer_adjust_thresh",
             "steer_adjust_ratio"]
    d = dict(zip(names, scal.tolist()))
    d["N"], d["max_fit_order"], d["latency"] = int(d["N"]), int(d["max_fit_order"]), int(d["latency"])
    d["weights"] = w.tolist()
    for k, t, n in zip(["steers", "steer_speeds", "yaw_changes", "yaw_change_speeds"], tabs, cnt):
        d[k] = t[:n].tolist()
    return d


def set_weights(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    lib().ref_config_set_weights(_dp(w), len(w))


def set_horizon(N, dt):
    lib().ref_config_set_horizon(int(N), float(dt))


def solve(state, coeffs, yaw_lo, yaw_hi, N):
    """MPC::solve with roadGeometry.polynomial = coeffs (trailing zeros kept: same polynomial)."""
    st = np.ascontiguousarray(state, dtype=np.float64)
    co = np.ascontiguousarray(coeffs, dtype=np.float64)
    res, tx, ty = np.zeros(9), np.zeros(N), np.zeros(N)
    status, iters = C.c_int(0), C.c_int(0)
    lib().ref_solve(_dp(st), _dp(co), len(co), float(yaw_lo), float(yaw_hi), _dp(res), _dp(tx), _dp(ty),
                    C.byref(status), C.byref(iters))
    return {"result": res, "traj_x": tx, "traj_y": ty, "status": status.value, "iters": iters.value}


def run(pose, ptsx, ptsy, N, steering=0.0, accel=0.0):
    """MPC::run: (x, y, psi, v) + waypoints -> 8 outputs, trajectory, fitted coefficients, yaw bounds."""
    x = np.array(ptsx, dtype=np.float64)
    y = np.array(ptsy, dtype=np.float64)
    res, tx, ty, co = np.zeros(8), np.zeros(N), np.zeros(N), np.zeros(5)
    nco, status, iters = C.c_int(0), C.c_int(0), C.c_int(0)
    ylo, yhi = C.c_double(0), C.c_double(0)
    lib().ref_run(float(pose[0]), float(pose[1]), float(pose[2]), float(pose[3]), float(steering), float(accel), _dp(x),
                  _dp(y), len(x), _dp(res), _dp(tx), _dp(ty), _dp(co), C.byref(nco), C.byref(ylo), C.byref(yhi),
                  C.byref(status), C.byref(iters))
    return {"result": res, "traj_x": tx, "traj_y": ty, "coeffs": co, "ncoef": nco.value, "yaw_lo": ylo.value,
            "yaw_hi": yhi.value, "ptsx": x, "ptsy": y, "status": status.value, "iters": iters.value}
