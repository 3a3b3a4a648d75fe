"""Writes tests/golden/ref_closed_loop.json: closed-loop trajectories in which every piece of controller and plant
arithmetic is the REFERENCE'S OWN compiled code (oracle/_ref/libmpc_ref.so): MPC::run (frame transform, adaptive fit,
yaw bounds, solve, steering adjustment, acceleration clamp -- MPC.cpp:327-382), Vehicle::move as the plant
(Vehicle.cpp:145-168) and Vehicle::computeThrottle (Vehicle.cpp:81-103).  What is restated here is only the glue of
the message handler that cannot be compiled without uWebSockets (src/mpc_main.cpp:113-214: psi normalisation, the
acceleration estimate (throttle - v/50)*6, the latency move by lookahead + tau, the one-step actuation delay) and the
6-waypoint window the simulator would send.  Build container only:
    make -C oracle/ref_shim && python tests/golden/make_ref_closed_loop.py"""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po, pyref as pr  # noqa: E402
import mpc_b200 as mpc  # noqa: E402  (workload generator only)

rd = mpc.workloads.reference_data()
wx, wy = np.array(rd["waypoints"]["x"]), np.array(rd["waypoints"]["y"])
W = len(wx)
dp = C.POINTER(C.c_double)


def ref_move(x, y, psi, v, steering, accel, length, dt):
    p6 = np.array([x, y, psi, v, steering, accel])
    pr.lib().ref_vehicle_move(p6.ctypes.data_as(dp), float(length), float(dt))
    return p6[0], p6[1], p6[2], p6[3]


def normalize_angle(a):   # utils.h:86-92
    while a >= math.pi:
        a -= 2 * math.pi
    while a < -math.pi:
        a += 2 * math.pi
    return a


def rollout(js, cd, veh, seg, T, dt_ctrl, tau):
    x, y, psi, v, steer, thr = veh
    pending = (0.0, 0.0)
    rec = []
    for _ in range(T):
        cs, sn = math.cos(psi), math.sin(psi)
        for _g in range(8):
            j = (seg + 1) % W
            if (wx[j] - x) * cs + (wy[j] - y) * sn > 0.0:
                break
            seg = j
        win = [(seg + i) % W for i in range(6)]
        px, py, pp, pv = x, y, normalize_angle(psi), v                      # mpc_main.cpp:127
        accel_est = (thr - v / 50.0) * 6                                    # :156
        if cd["latency"]:
            px, py, pp, pv = ref_move(px, py, pp, pv, steer, accel_est, cd["Lf"], cd["lookahead"] + tau)   # :157-159
        r = pr.run((px, py, pp, pv), wx[win], wy[win], js["N"], steering=steer, accel=accel_est)          # :167
        o = r["result"]
        throttle = pr.lib().ref_compute_throttle(float(o[5]), float(o[3]), cd["max_accel"], cd["max_decel"])   # :174
        d_cmd = o[4] * cd["max_steering"]
        d_apply, t_apply = d_cmd, throttle
        if cd["latency"]:
            d_apply, t_apply = pending
            pending = (d_cmd, throttle)
        a_plant = (t_apply - v / 50.0) * 6
        x, y, psi, v = ref_move(x, y, psi, v, d_apply, a_plant, cd["Lf"], dt_ctrl)
        steer, thr = d_apply, throttle
        rec.append([0.0, 0.0, pv, float(o[4]), float(throttle), 0.0, float(r["status"]), float(r["iters"])])
        # cte / epsi at the (latency-compensated) pose are the start state of the solve: recover them from the fit
        c = r["coeffs"]
        rec[-1][0] = float(c[0])
        rec[-1][1] = float(-math.atan(c[1]))
    return rec, [x, y, psi, v, steer, thr], seg


out = {"generator": "tests/golden/make_ref_closed_loop.py (oracle/_ref/libmpc_ref.so)", "cases": []}
for name, tau, V, T in (("fast", 0.02, 6, 150), ("stable", 0.0, 3, 80), ("no-latency", 0.0, 3, 80)):
    js = rd["configs"][name]
    pr.config_load(js)
    cd = po.load_config_dict(js)
    b = mpc.workloads.batch_perturbed_states(V, 3, cd)
    for i in range(V):
        veh0 = [float(b["px"][i]), float(b["py"][i]), float(b["psi"][i]), float(np.clip(b["v"][i], 8, 30)), 0.0, 0.0]
        seg0 = int(b["segment"][i])
        rec, veh, seg = rollout(js, cd, list(veh0), seg0, T, 0.1, tau)
        out["cases"].append({"config": name, "tau": tau, "T": T, "dt_ctrl": 0.1, "veh0": veh0, "seg0": seg0, "rec": rec, "veh": [float(t) for t in veh], "seg": seg})
        st = np.array(rec)[:, 6]
        print(name, i, "status ok %.3f" % (st == 1).mean(), "final v %.3f cte %.3f" % (veh[3], rec[-1][0]))
pr.config_load(rd["configs"]["stable"])
json.dump(out, open(os.path.join(HERE, "ref_closed_loop.json"), "w"))
print("wrote ref_closed_loop.json", os.path.getsize(os.path.join(HERE, "ref_closed_loop.json")), "bytes")
