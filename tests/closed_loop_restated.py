"""CPU restatement of one vehicle's closed loop as include/mpc_b200.h documents it for mpc_rollout -- the message
handler of /root/reference/src/mpc_main.cpp:113-214 with the reference's own kinematic plant (Vehicle::move)
standing in for the simulator.  TEST INFRASTRUCTURE: numpy + the CPU oracle (oracle/pyoracle.py)."""
import math

import numpy as np


def normalize_angle(a):   # utils.h:86-92
    while a >= math.pi:
        a -= 2 * math.pi
    while a < -math.pi:
        a += 2 * math.pi
    return a


def vehicle_move(x, y, psi, v, steering, accel, length, dt):   # Vehicle.cpp:145-168
    dist = v * dt
    return x + dist * math.cos(psi), y + dist * math.sin(psi), psi + steering * dist / length, v + accel * dt


def compute_throttle(cd, accel, target):   # Vehicle.cpp:81-103
    keep = target / cd["max_speed"]
    if accel >= 0:
        return keep if accel < 0.001 else min(1.0, keep + (1 - keep) * accel / cd["max_accel"])
    if accel <= -15:
        return -1.0
    if accel < -10:
        return -0.95 - (1 - 0.95) * accel / cd["max_decel"]
    if accel < -5:
        return -0.9 - (1 - 0.9) * accel / cd["max_decel"]
    return -0.85 - (1 - 0.85) * accel / cd["max_decel"]


def table_limit(keys, vals, angle, mx):   # Vehicle.cpp:34-48, 66-79
    y = abs(angle)
    for i, k in enumerate(keys):
        if y <= k:
            return min(vals[i] if len(vals) > i else vals[-1], mx)
    return min(vals[-1], mx)


def rollout(po, cd, track_x, track_y, veh, seg, pending, T, dt_ctrl=0.1, tau_solve=0.0):
    """veh = [x, y, psi, v, steer, throttle]; returns (records [T][8], veh, seg, pending)."""
    cfg = po.make_config(cd)
    W = len(track_x)
    x, y, psi, v, steer, thr = veh
    rec = []
    for _ in range(T):
        cs, sn = math.cos(psi), math.sin(psi)
        for _g in range(8):
            j = (seg + 1) % W
            if (track_x[j] - x) * cs + (track_y[j] - y) * sn > 0.0:
                break
            seg = j
        win = [(seg + i) % W for i in range(6)]
        px, py, pp, pv = x, y, normalize_angle(psi), v
        accel_est = (thr - v / 50.0) * 6
        if cd["latency"]:
            px, py, pp, pv = vehicle_move(px, py, pp, pv, steer, accel_est, cd["Lf"], cd["lookahead"] + tau_solve)
        state, coeffs, ylo, yhi, ex = po.preprocess(cd, (px, py, pp, pv), np.asarray(track_x)[win], np.asarray(track_y)[win])
        myc = ex["myc"]
        max_speed = table_limit(cd["yaw_changes"], cd["yaw_change_speeds"], myc, cd["max_speed"])
        target = table_limit(cd["steers"], cd["steer_speeds"], steer, max_speed)
        r = po.solve(cfg, po.make_problem(state, coeffs, ylo, yhi))
        res = r["result"]
        sa = res[6]
        if abs(myc) > cd["steer_adjust_thresh"]:
            sa += cd["steer_adjust_ratio"] * myc
        accel = min(res[7], target - pv)
        sv = min(max(sa / cd["max_steering"], -1.0), 1.0)
        throttle = compute_throttle(cd, accel, res[3])
        d_cmd = sv * cd["max_steering"]
        d_apply, t_apply = d_cmd, throttle
        if cd["latency"]:
            d_apply, t_apply = pending
            pending = (d_cmd, throttle)
        a_plant = (t_apply - v / 50.0) * 6
        x, y, psi, v = vehicle_move(x, y, psi, v, d_apply, a_plant, cd["Lf"], dt_ctrl)
        steer, thr = d_apply, throttle
        rec.append([state[4], state[5], pv, sv, throttle, res[8], float(r["status"]), float(r["iters"])])
    return np.array(rec), [x, y, psi, v, steer, thr], seg, pending
