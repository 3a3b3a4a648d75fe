"""Writes tests/golden/oracle_golden.json: the test.cpp scenario (run() + 25 x solve() feeding the
step-1 state back, /root/reference/src/test.cpp:64-111) for the active fixture and the four
commented-out ones (test.cpp:18-43), solved by the CPU oracle.  Run after tests/test_oracle_solve.py
(SciPy agreement + KKT certificates) is green:   python tests/golden/make_oracle_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

rd = json.load(open(os.path.join(ROOT, "carnd-mpc-project_b200", "data", "reference_data.json")))
cd = po.load_config_dict(rd["configs"]["stable"])
cfg = po.make_config(cd)
out = {"config": "stable", "scenarios": []}
for fx in rd["test_cpp_fixtures"]:
    state, coeffs, ylo, yhi, ex = po.preprocess(cd, (fx["x"], fx["y"], fx["psi"], fx["v"]), fx["ptsx"], fx["ptsy"])
    sc = {"name": fx["name"], "state0": state.tolist(), "coeffs": coeffs.tolist(), "yaw_lo": ylo, "yaw_hi": yhi,
          "fit_order": ex["order"], "steps": []}
    st = state
    for k in range(26):
        r = po.solve(cfg, po.make_problem(st, coeffs, ylo, yhi))
        sc["steps"].append({"status": r["status"], "iters": r["iters"], "result": r["result"].tolist(),
                            "traj_x": r["z"][:cfg.N].tolist(), "traj_y": r["z"][cfg.N:2 * cfg.N].tolist()})
        st = r["result"][:6].copy()
    out["scenarios"].append(sc)
    print(fx["name"], [s["status"] for s in sc["steps"]], [s["iters"] for s in sc["steps"]])
json.dump(out, open(os.path.join(HERE, "oracle_golden.json"), "w"), indent=1)
