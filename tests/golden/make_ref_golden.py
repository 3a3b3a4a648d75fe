"""Writes tests/golden/ref_golden.json: input/output vectors of the REFERENCE'S OWN controller code
(MPC::run, MPC::solve, FG_eval, Config::load -- /root/reference/src/control/MPC.cpp, src/model/*.cpp,
src/utils/*.cpp compiled unmodified into oracle/_ref/libmpc_ref.so against the CppAD/Ipopt stand-ins of
oracle/ref_shim).  Build container only:   make -C oracle/ref_shim && python tests/golden/make_ref_golden.py

Cases: (a) src/test.cpp's scenario (run() + 25 x solve(), test.cpp:64-111) for its four fixtures;
(b) MPC::run from perturbed poses on lake-track windows, all three shipped configs; (c) MPC::solve with
modified Config::weights -- including the acceleration weights, which the recorded tape ignores;
(d) MPC::solve over the N x dt grid of examples/; (f) MPC::solve on long-horizon cells (N = 20, 30, 40 at dt = 0.1)
for problems whose line search runs below alpha_min, i.e. that ENTER Ipopt's feasibility restoration phase."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po, pyref as pr  # noqa: E402
import mpc_b200 as mpc  # noqa: E402  (workload generator only)

rd = json.load(open(os.path.join(ROOT, "carnd-mpc-project_b200", "data", "reference_data.json")))
wx, wy = np.array(rd["waypoints"]["x"]), np.array(rd["waypoints"]["y"])
out = {"generator": "tests/golden/make_ref_golden.py (oracle/_ref/libmpc_ref.so)", "testcpp": [], "run": [],
       "weights": [], "grid": []}
L = lambda a: np.asarray(a).tolist()

# (a) test.cpp scenario, config-stable
js = rd["configs"]["stable"]
pr.config_load(js)
N = js["N"]
for fx in rd["test_cpp_fixtures"]:
    r = pr.run((fx["x"], fx["y"], fx["psi"], fx["v"]), fx["ptsx"], fx["ptsy"], N)
    sc = {"name": fx["name"], "run": {k: L(r[k]) if not np.isscalar(r[k]) else r[k] for k in r}, "steps": []}
    st = np.array([r["result"][0], r["result"][1], r["result"][2], r["result"][3], r["result"][6], r["result"][7]])
    for k in range(25):
        s = pr.solve(st, r["coeffs"][:r["ncoef"]], r["yaw_lo"], r["yaw_hi"], N)
        sc["steps"].append({"state": L(st), "status": s["status"], "iters": s["iters"], "result": L(s["result"]),
                            "traj_x": L(s["traj_x"]), "traj_y": L(s["traj_y"])})
        st = s["result"][:6].copy()
    out["testcpp"].append(sc)
    print("testcpp", fx["name"], [s["status"] for s in sc["steps"]])

# (b) MPC::run on perturbed lake-track poses, every shipped config
for name in ("stable", "fast", "no-latency"):
    js = rd["configs"][name]
    pr.config_load(js)
    cd = po.load_config_dict(js)
    b = mpc.workloads.batch_perturbed_states(40, 100 + len(name), cd)
    for i in range(40):
        win = (b["segment"][i] + np.arange(6)) % len(wx)
        pose = (b["px"][i], b["py"][i], b["psi"][i], b["v"][i])
        r = pr.run(pose, wx[win], wy[win], js["N"])
        out["run"].append({"config": name, "pose": L(pose), "ptsx": L(wx[win]), "ptsy": L(wy[win]), "status": r["status"],
                           "iters": r["iters"], "result": L(r["result"]), "traj_x": L(r["traj_x"]), "traj_y": L(r["traj_y"]),
                           "coeffs": L(r["coeffs"]), "ncoef": r["ncoef"], "yaw_lo": r["yaw_lo"], "yaw_hi": r["yaw_hi"],
                           "ptsx_vehicle": L(r["ptsx"]), "ptsy_vehicle": L(r["ptsy"])})
    print("run", name, np.bincount([c["status"] for c in out["run"] if c["config"] == name]))

# (c) weight sweep through MPC::solve
js = rd["configs"]["stable"]
pr.config_load(js)
cd = po.load_config_dict(js)
b = mpc.workloads.batch_perturbed_states(24, 7, cd)
rng = np.random.default_rng(5)
for i in range(24):
    w = np.array(js["weights"], dtype=float)
    w[3] = np.exp(rng.uniform(np.log(1), np.log(5000))); w[4] = np.exp(rng.uniform(np.log(1), np.log(5000)))
    w[1] = np.exp(rng.uniform(np.log(1), np.log(1000))); w[2] = rng.choice([0.01, 0.1, 1, 10, 100])
    w[6], w[7], w[8] = rng.uniform(0, 1e4, 3)      # acceleration-related weights: dead in the recorded tape
    pr.set_weights(w)
    s = pr.solve(b["state"][i], b["coeffs"][i], b["yaw_lo"][i], b["yaw_hi"][i], js["N"])
    out["weights"].append({"weights": L(w), "state": L(b["state"][i]), "coeffs": L(b["coeffs"][i]), "yaw_lo": b["yaw_lo"][i],
                           "yaw_hi": b["yaw_hi"][i], "status": s["status"], "iters": s["iters"], "result": L(s["result"])})
pr.set_weights(js["weights"])
print("weights", np.bincount([c["status"] for c in out["weights"]]))

# (d) horizon / timestep grid (submission-report.md:250-265)
for (Ng, dtg) in [(10, 0.1), (20, 0.1), (30, 0.1), (10, 0.05), (20, 0.05), (30, 0.05), (40, 0.05), (50, 0.05), (10, 0.02),
                  (30, 0.02), (50, 0.02)]:
    pr.set_horizon(Ng, dtg)
    for i in range(3):
        s = pr.solve(b["state"][i], b["coeffs"][i], b["yaw_lo"][i], b["yaw_hi"][i], Ng)
        out["grid"].append({"N": Ng, "dt": dtg, "state": L(b["state"][i]), "coeffs": L(b["coeffs"][i]), "yaw_lo": b["yaw_lo"][i],
                            "yaw_hi": b["yaw_hi"][i], "status": s["status"], "iters": s["iters"], "result": L(s["result"]),
                            "traj_x": L(s["traj_x"]), "traj_y": L(s["traj_y"])})
    print("grid", Ng, dtg, [c["status"] for c in out["grid"][-3:]], [c["iters"] for c in out["grid"][-3:]])
pr.set_horizon(js["N"], js["dt"])

# (f) problems that enter the restoration phase (long horizons at dt = 0.1: the fit is extrapolated, the line search
# stalls).  Picked with the plain-C oracle's counters, solved through the reference's own MPC::solve.
import ctypes  # noqa: E402
out["resto"] = []
for (Ng, dtg, want) in [(20, 0.1, 3), (30, 0.1, 5), (40, 0.1, 4)]:
    jsg = dict(js, N=Ng, dt=dtg)
    cdg = po.load_config_dict(jsg)
    bb = mpc.workloads.batch_perturbed_states(160, 1, cdg)
    probs = po.problems_from_arrays(bb["state"], bb["coeffs"], bb["yaw_lo"], bb["yaw_hi"])
    res = (po.OrcResult * 160)()
    ocfg = po.make_config(cdg)
    po.lib().orc_solve_batch(ctypes.byref(ocfg), probs, 160, res, 8)
    idx = [i for i in range(160) if res[i].n_resto > 0][:want]
    pr.set_horizon(Ng, dtg)
    for i in idx:
        s = pr.solve(bb["state"][i], bb["coeffs"][i], bb["yaw_lo"][i], bb["yaw_hi"][i], Ng)
        out["resto"].append({"N": Ng, "dt": dtg, "state": L(bb["state"][i]), "coeffs": L(bb["coeffs"][i]), "yaw_lo": bb["yaw_lo"][i],
                             "yaw_hi": bb["yaw_hi"][i], "status": s["status"], "iters": s["iters"], "result": L(s["result"]),
                             "traj_x": L(s["traj_x"]), "traj_y": L(s["traj_y"]),
                             "oracle_n_resto": int(res[i].n_resto), "oracle_n_resto_iter": int(res[i].n_resto_iter),
                             "oracle_iters": int(res[i].iters), "oracle_status": int(res[i].status)})
        print("resto", Ng, dtg, i, "ref status", s["status"], "iters", s["iters"], "| oracle iters", res[i].iters, "resto calls", res[i].n_resto)
pr.set_horizon(js["N"], js["dt"])

# (e) plant and actuation map: Vehicle::move, Vehicle::computeThrottle (Vehicle.cpp:81-103,145-168)
out["plant"] = {"move": [], "throttle": []}
rng = np.random.default_rng(11)
for name in ("stable", "fast"):
    js = rd["configs"][name]
    pr.config_load(js)
    cd = po.load_config_dict(js)
    for _ in range(20):
        pose = [rng.uniform(-200, 200), rng.uniform(-200, 200), rng.uniform(-4, 4), rng.uniform(0, 70), rng.uniform(-0.4, 0.4), rng.uniform(-9, 5)]
        dt = float(rng.choice([0.02, 0.1, 0.12]))
        p6 = np.array(pose)
        pr.lib().ref_vehicle_move(p6.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)), cd["Lf"], dt)
        out["plant"]["move"].append({"config": name, "pose": pose, "dt": dt, "moved": L(p6[:4])})
    for accel in [-20, -15, -12, -10, -7, -5, -3, -0.5, 0.0, 0.0005, 0.001, 0.5, 2.0, 4.4, 6.0]:
        target = float(rng.uniform(0, 60))
        t = pr.lib().ref_compute_throttle(float(accel), target, cd["max_accel"], cd["max_decel"])
        out["plant"]["throttle"].append({"config": name, "accel": float(accel), "target": target, "throttle": t})
pr.config_load(rd["configs"]["stable"])
json.dump(out, open(os.path.join(HERE, "ref_golden.json"), "w"), indent=0)
print("wrote ref_golden.json", os.path.getsize(os.path.join(HERE, "ref_golden.json")), "bytes")
