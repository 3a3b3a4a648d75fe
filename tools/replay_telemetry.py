#!/usr/bin/env python
"""Replay recorded simulator traffic through the controller without the socket:
    python tools/replay_telemetry.py [-config f | -fast | -stable | -latency ms | -speed mph] [--tau 0.02] [--plot] < messages.txt
One SocketIO text per input line ('42["telemetry",{...}]'); prints what src/mpc_main.cpp would send back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc

args = sys.argv[1:]
tau, plot = 0.0, False
if "--tau" in args:
    i = args.index("--tau"); tau = float(args[i + 1]); del args[i:i + 2]
if "--plot" in args:
    args.remove("--plot"); plot = True
cfg, path = mpc.config_from_cli(args, os.environ.get("MPC_CONFIG_DIR", ".."))
print("config:", path, file=sys.stderr)
S = mpc.Solver(cfg, 0)
thr = 0.0
for line in sys.stdin:
    line = line.strip()
    if not line:
        continue
    reply, thr = S.telemetry_step(line, thr, tau, plot)
    if reply:
        print(reply)
