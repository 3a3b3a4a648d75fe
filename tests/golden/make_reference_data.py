"""Regenerates carnd-mpc-project_b200/data/reference_data.json from the read-only reference checkout.

Run in the build container only (/root/reference does not exist on the GPU box):
    python tests/golden/make_reference_data.py
It captures DATA the reference ships (no source code): the three config-*.json knob sets, the
lake-track centre-line waypoints and the hard-coded fixtures of src/test.cpp:18-50.
"""
import csv
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "carnd-mpc-project_b200", "data", "reference_data.json")

data = {"configs": {}, "waypoints": {"x": [], "y": []}, "test_cpp_fixtures": []}
for name in ("stable", "fast", "no-latency"):
    with open(os.path.join(REF, "config-%s.json" % name)) as f:
        data["configs"][name] = json.load(f)
with open(os.path.join(REF, "lake_track_waypoints.csv")) as f:
    rd = csv.reader(f)
    next(rd)
    for row in rd:
        if row:
            data["waypoints"]["x"].append(float(row[0]))
            data["waypoints"]["y"].append(float(row[1]))
# src/test.cpp:45-50 (active) and :18-43 (commented-out alternatives); (ptsx, ptsy, x, y, psi, v)
data["test_cpp_fixtures"] = [
    {"name": "active_45_50",
     "ptsx": [-145.1165, -158.3417, -164.3164, -169.3365, -175.4917, -176.9617],
     "ptsy": [4.339378, -17.42898, -30.18062, -42.84062, -66.52898, -76.85062],
     "x": -146.7283, "y": 1.660802, "psi": 4.125825, "v": 26.6806},
    {"name": "alt_18_23",
     "ptsx": [-134.97, -145.1165, -158.3417, -164.3164, -169.3365, -175.4917],
     "ptsy": [18.404, 4.339378, -17.42898, -30.18062, -42.84062, -66.52898],
     "x": -146.8912, "y": 2.129487, "psi": 0.4009452, "v": 13.87815},
    {"name": "alt_25_31",
     "ptsx": [-164.3164, -169.3365, -175.4917, -176.9617, -176.8864, -175.0817],
     "ptsy": [-30.18062, -42.84062, -66.52898, -76.85062, -90.64063, -100.3206],
     "x": -166.0726, "y": -29.59644, "psi": 4.088, "v": 30.62756},
    # the third commented-out pose (test.cpp:33-36: x=-144.7913 y=3.767814 psi=0.03732295 v=10.32361) has no
    # waypoint list of its own and faces against the track for either neighbouring list; it is not a
    # well-posed scenario and is left out.
    {"name": "alt_38_43",
     "ptsx": [-61.09, -78.29172, -93.05002, -107.7717, -123.3917, -134.97],
     "ptsy": [92.88499, 78.73102, 65.34102, 50.57938, 33.37102, 18.404],
     "x": -61.97283, "y": 93.53992, "psi": 3.857562, "v": 33.06046},
]
with open(OUT, "w") as f:
    json.dump(data, f, indent=1)
print("wrote", OUT, len(data["waypoints"]["x"]), "waypoints")
