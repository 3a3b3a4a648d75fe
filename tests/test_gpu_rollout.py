"""GPU tests of the control-step entry points on the device: mpc_run_batch (MPC::run for a batch) against the
vectors of the reference's own MPC::run, and mpc_rollout (closed loop, BASELINE config 5) against the CPU
restatement tests/closed_loop_restated.py (oracle solve + the reference's plant)."""
import json
import os

import numpy as np
import pytest

import closed_loop_restated as clr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ABS_TOL = 1e-4


def _dev(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def test_run_batch_matches_reference_run(mpc, refdata, kernel_kind):
    import torch
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_golden.json")))
    for name in refdata["configs"]:
        cfg = mpc.config_from_json_text(json.dumps(refdata["configs"][name]))
        cases = [c for c in gold["run"] if c["config"] == name]
        B, N = len(cases), cfg.N
        pose = _dev(np.array([c["pose"] for c in cases]).T)
        ptsx = _dev(np.array([c["ptsx"] for c in cases]).T)
        ptsy = _dev(np.array([c["ptsy"] for c in cases]).T)
        out8 = torch.zeros(8, B, dtype=torch.float64, device="cuda")
        tx = torch.zeros(N, B, dtype=torch.float64, device="cuda")
        ty = torch.zeros(N, B, dtype=torch.float64, device="cuda")
        co = torch.zeros(5, B, dtype=torch.float64, device="cuda")
        pxv = torch.zeros(6, B, dtype=torch.float64, device="cuda")
        pyv = torch.zeros(6, B, dtype=torch.float64, device="cuda")
        st = torch.zeros(B, dtype=torch.int32, device="cuda")
        it = torch.zeros(B, dtype=torch.int32, device="cuda")
        S = mpc.Solver(cfg, 0)
        S.set_kernel(kernel_kind)
        S.run_batch_device(B, pose, ptsx, ptsy, 6, out8, traj_x=tx, traj_y=ty, coeffs_out=co, ptsx_v=pxv, ptsy_v=pyv,
                           status=st, iters=it)
        torch.cuda.synchronize()
        S.close()
        o, tx, co, pxv, st = out8.cpu().numpy().T, tx.cpu().numpy().T, co.cpu().numpy().T, pxv.cpu().numpy().T, st.cpu().numpy()
        for i, c in enumerate(cases):
            assert st[i] == c["status"]
            assert np.allclose(pxv[i], c["ptsx_vehicle"], rtol=0, atol=1e-10)
            assert np.allclose(co[i], c["coeffs"], rtol=1e-7, atol=1e-10)
            if c["status"] == 1:
                assert np.abs(o[i] - np.array(c["result"])).max() < ABS_TOL
                assert np.abs(tx[i] - np.array(c["traj_x"])).max() < ABS_TOL


@pytest.mark.parametrize("name,tau", [("fast", 0.02), ("no-latency", 0.0), ("stable", 0.0)])
def test_rollout_matches_cpu_restatement(mpc, po, refdata, name, tau):
    import torch
    js = refdata["configs"][name]
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = po.load_config_dict(js)
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    V, T = 12, 30
    b = mpc.workloads.batch_perturbed_states(V, 3, cd)
    veh0 = np.stack([b["px"], b["py"], b["psi"], np.clip(b["v"], 8, 30), np.zeros(V), np.zeros(V)])
    seg0 = b["segment"].astype(np.int32)
    veh, seg = _dev(veh0), _dev(seg0)
    pending = torch.zeros(2, V, dtype=torch.float64, device="cuda")
    rec = torch.zeros(T, 8, V, dtype=torch.float64, device="cuda")
    S = mpc.Solver(cfg, 0)
    S.rollout_device(V, T, _dev(wx), _dev(wy), veh, seg, pending, 0.1, tau, rec)
    torch.cuda.synchronize()
    S.close()
    rec, veh, seg = rec.cpu().numpy(), veh.cpu().numpy(), seg.cpu().numpy()
    for i in range(V):
        r_ref, v_ref, s_ref, _ = clr.rollout(po, cd, wx, wy, list(veh0[:, i]), int(seg0[i]), (0.0, 0.0), T, 0.1, tau)
        ok = r_ref[:, 6] == 1
        assert ok.mean() > 0.9
        assert np.array_equal(rec[:, 6, i], r_ref[:, 6])                       # same status every step
        # cte, epsi, v, steer, throttle along the whole rollout; the loop is stabilising, so the 1e-8
        # per-solve differences do not grow
        assert np.abs(rec[:, :5, i] - r_ref[:, :5]).max() < ABS_TOL
        assert np.abs(rec[ok, 5, i] / r_ref[ok, 5] - 1).max() < 1e-6
        assert np.abs(veh[:, i] - np.array(v_ref)).max() < ABS_TOL
        assert seg[i] == s_ref
    # the vehicles actually drive: they moved ~ v * T * dt along the track and stayed near the centre line
    assert (np.hypot(veh[0] - veh0[0], veh[1] - veh0[1]) > 10).all()


def test_big_fleet_rollout_uses_the_lane_chain_and_agrees_with_the_coop_kernel(mpc, refdata):
    """More vehicles than the coop/lane crossover: every control step of mpc_rollout is then the lane kernel with its
    resume launches and finisher (several launches per step on one stream, counters reset each step); the
    closed-loop record must be the one the coop kernel produces, bit for bit."""
    import torch
    cfg = mpc.config_from_json_text(json.dumps(refdata["configs"]["fast"]))
    cd = cfg.as_dict()
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    V, T = mpc.LANE_MIN_BATCH + 1000, 6
    b = mpc.workloads.batch_perturbed_states(V, 3, cd)
    recs = {}
    for kind in (mpc.KERNEL_COOP, mpc.KERNEL_AUTO):
        S = mpc.Solver(cfg, 0)
        S.set_kernel(kind)
        S.set_rollout_mode(1)     # three launches per control step (fleets this size default to the persistent kernel)
        veh = _dev(np.stack([b["px"], b["py"], b["psi"], np.clip(b["v"], 8, 30), np.zeros(V), np.zeros(V)]))
        seg = _dev(b["segment"].astype(np.int32))
        pending = torch.zeros(2, V, dtype=torch.float64, device="cuda")
        rec = torch.zeros(T, 8, V, dtype=torch.float64, device="cuda")
        n0 = S.launches
        S.rollout_device(V, T, _dev(wx), _dev(wy), veh, seg, pending, 0.1, 0.02, rec)
        torch.cuda.synchronize()
        per_step = (S.launches - n0) // T
        assert per_step == (3 if kind == mpc.KERNEL_COOP else 7)      # pre, solve (1 or 5 launches), post
        recs[kind] = (rec.clone(), veh.clone(), seg.clone())
        S.close()
    for a, c in zip(recs[mpc.KERNEL_COOP], recs[mpc.KERNEL_AUTO]):
        assert torch.equal(a, c)
    assert (recs[mpc.KERNEL_AUTO][0][:, 6] == 1).float().mean().item() > 0.99


def _rollout(mpc, cfg, b, V, T, wx, wy, tau, mode):
    import torch
    veh = _dev(np.stack([b["px"][:V], b["py"][:V], b["psi"][:V], np.clip(b["v"][:V], 8, 30), np.zeros(V), np.zeros(V)]))
    seg = _dev(b["segment"][:V].astype(np.int32))
    pending = torch.zeros(2, V, dtype=torch.float64, device="cuda")
    rec = torch.zeros(T, 8, V, dtype=torch.float64, device="cuda")
    S = mpc.Solver(cfg, 0)
    S.set_rollout_mode(mode)
    n0 = S.launches
    S.rollout_device(V, T, _dev(wx), _dev(wy), veh, seg, pending, 0.1, tau, rec)
    torch.cuda.synchronize()
    launches = S.launches - n0
    S.close()
    return rec, veh, seg, launches


@pytest.mark.parametrize("name,tau,V,T", [("fast", 0.02, 300, 40), ("stable", 0.0, 37, 25), ("fast", 0.02, 5000, 6)])
def test_persistent_rollout_kernel_is_one_launch_and_the_same_bits(mpc, refdata, name, tau, V, T):
    """The closed loop as ONE launch (a lane group owns a vehicle for all T steps) against three launches per control
    step: the same Lane code runs the solve, so the records, the final vehicle states and the waypoint windows are
    bit-identical -- also with more vehicles than resident lane groups (several vehicles per group in turn)."""
    import torch
    cfg = mpc.config_from_json_text(json.dumps(refdata["configs"][name]))
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    b = mpc.workloads.batch_perturbed_states(V, 3, cfg.as_dict())
    per_step = _rollout(mpc, cfg, b, V, T, wx, wy, tau, 1)
    one = _rollout(mpc, cfg, b, V, T, wx, wy, tau, 2)
    auto = _rollout(mpc, cfg, b, V, T, wx, wy, tau, 0)
    # AUTO: one launch up to MPC_ROLLOUT_PERSISTENT_MAX vehicles (beyond, a lane group would take vehicles one after another
    # for all T steps and the launch-per-step path around the throughput kernel is faster: profiles/r02_rollout_modes.txt)
    hdr = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    pmax = int(hdr.split("#define MPC_ROLLOUT_PERSISTENT_MAX")[1].split()[0])
    assert per_step[3] == 3 * T and one[3] == 1 and auto[3] == (1 if V <= pmax else 3 * T)
    for a, c, d in zip(per_step[:3], one[:3], auto[:3]):
        assert torch.equal(a, c) and torch.equal(a, d)
    assert (one[0][:, 6] == 1).float().mean().item() > 0.99


def test_closed_loop_against_the_reference_sources(mpc, refdata):
    """tests/golden/ref_closed_loop.json: closed loops in which MPC::run, the plant (Vehicle::move) and the throttle map
    (Vehicle::computeThrottle) are the reference's own compiled code (tests/golden/make_ref_closed_loop.py); 80-150
    control steps per vehicle, all three shipped configs.  The persistent kernel reproduces them step by step."""
    import torch
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_closed_loop.json")))
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    for name in refdata["configs"]:
        cases = [c for c in gold["cases"] if c["config"] == name]
        assert cases
        cfg = mpc.config_from_json_text(json.dumps(refdata["configs"][name]))
        V, T, tau = len(cases), cases[0]["T"], cases[0]["tau"]
        veh = _dev(np.array([c["veh0"] for c in cases]).T)
        seg = _dev(np.array([c["seg0"] for c in cases], dtype=np.int32))
        pending = torch.zeros(2, V, dtype=torch.float64, device="cuda")
        rec = torch.zeros(T, 8, V, dtype=torch.float64, device="cuda")
        S = mpc.Solver(cfg, 0)
        S.rollout_device(V, T, _dev(wx), _dev(wy), veh, seg, pending, cases[0]["dt_ctrl"], tau, rec)
        torch.cuda.synchronize()
        S.close()
        rec, veh, seg = rec.cpu().numpy(), veh.cpu().numpy(), seg.cpu().numpy()
        for i, c in enumerate(cases):
            ref = np.array(c["rec"])
            assert np.array_equal(rec[:, 6, i], ref[:, 6]) and (ref[:, 6] == 1).all()        # status, every step
            assert np.abs(rec[:, :5, i] - ref[:, :5]).max() < ABS_TOL                         # cte, epsi, v, steer, throttle
            assert np.abs(veh[:, i] - np.array(c["veh"])).max() < ABS_TOL
            assert seg[i] == c["seg"]


def test_thousand_step_rollout_spot_check(mpc, po, refdata):
    """BASELINE config 5 is 1000 control steps per vehicle: 64 vehicles x 1000 steps of config-fast on the persistent
    kernel against the CPU restatement of the loop (oracle solve + the reference's plant formulas), step by step."""
    import concurrent.futures as cf
    import torch
    js = refdata["configs"]["fast"]
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = po.load_config_dict(js)
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    V, T, tau = 64, 1000, 0.02
    b = mpc.workloads.batch_perturbed_states(V, 3, cd)
    rec, veh, seg, launches = _rollout(mpc, cfg, b, V, T, wx, wy, tau, 0)
    assert launches == 1
    rec, veh, seg = rec.cpu().numpy(), veh.cpu().numpy(), seg.cpu().numpy()
    veh0 = np.stack([b["px"], b["py"], b["psi"], np.clip(b["v"], 8, 30), np.zeros(V), np.zeros(V)])
    seg0 = b["segment"].astype(np.int32)
    assert (rec[:, 6] == 1).mean() > 0.995
    po.lib()

    def one(i):
        return clr.rollout(po, cd, wx, wy, list(veh0[:, i]), int(seg0[i]), (0.0, 0.0), T, 0.1, tau)

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:      # the oracle's C code releases the GIL
        refs = list(ex.map(one, range(V)))
    worst = 0.0
    n_same = 0
    for i, (r_ref, v_ref, s_ref, _) in enumerate(refs):
        # a closed loop integrates: compare up to the first step (if any) where the two solvers' statuses differ
        same = rec[:, 6, i] == r_ref[:, 6]
        upto = T if same.all() else int(np.argmin(same))
        n_same += upto == T
        if upto:
            worst = max(worst, np.abs(rec[:upto, :5, i] - r_ref[:upto, :5]).max())
        if upto == T:
            assert np.abs(veh[:, i] - np.array(v_ref)).max() < 1e-3 and seg[i] == s_ref
    assert n_same >= V - 2, n_same
    assert worst < 1e-3, worst


def test_telemetry_replay_matches_restated_handler(mpc, po, refdata):
    """Recorded-style SocketIO text through mpc_telemetry_step == the message handler of mpc_main.cpp:99-214
    restated on the CPU (unit/sign conversions, latency move, MPC::run, throttle), several messages in a row so
    that the previous throttle feeds the next latency compensation."""
    import math
    js = refdata["configs"]["fast"]
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = po.load_config_dict(js)
    ocfg = po.make_config(cd)
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    b = mpc.workloads.batch_perturbed_states(4, 9, cd)
    S = mpc.Solver(cfg, 0)
    assert S.telemetry_step('42["telemetry",null]', 0.0)[0] == '42["manual",{}]'
    assert S.telemetry_step("3", 0.0)[0] == ""
    tau = 0.02
    for i in range(4):
        thr_prev = 0.0
        px, py, psi, v = b["px"][i], b["py"][i], b["psi"][i] + 2 * math.pi, float(np.clip(b["v"][i], 8, 30))
        seg = int(b["segment"][i])
        steer_sim = 0.0
        for step in range(3):
            win = [(seg + k) % len(wx) for k in range(6)]
            body = {"ptsx": wx[win].tolist(), "ptsy": wy[win].tolist(), "x": px, "y": py, "psi": psi, "psi_unity": 0.0,
                    "speed": v * 3600.0 / 1609.34, "steering_angle": steer_sim, "throttle": thr_prev}
            reply, thr_new = S.telemetry_step('42["telemetry",' + json.dumps(body) + ']', thr_prev, tau, with_trajectory=True)
            assert reply.startswith('42["steer",{') and reply.endswith("}]")
            got = json.loads(reply[len('42["steer",'):-1])
            # restated handler
            vv = (v * 3600.0 / 1609.34) * 1609.34 / 3600.0
            p = clr.normalize_angle(psi)
            st = -steer_sim
            x2, y2, p2, v2 = clr.vehicle_move(px, py, p, vv, st, (thr_prev - vv / 50.0) * 6, cd["Lf"], cd["lookahead"] + tau)
            state, coeffs, ylo, yhi, ex = po.preprocess(cd, (x2, y2, p2, v2), wx[win], wy[win])
            r = po.solve(ocfg, po.make_problem(state, coeffs, ylo, yhi))
            assert r["status"] == 1
            myc = ex["myc"]
            target = clr.table_limit(cd["steers"], cd["steer_speeds"], st, clr.table_limit(cd["yaw_changes"], cd["yaw_change_speeds"], myc, cd["max_speed"]))
            sa = r["result"][6] + (cd["steer_adjust_ratio"] * myc if abs(myc) > cd["steer_adjust_thresh"] else 0.0)
            sv = min(max(sa / cd["max_steering"], -1.0), 1.0)
            accel = min(r["result"][7], target - v2)
            thr = clr.compute_throttle(cd, accel, r["result"][3])
            assert got["steering_angle"] == pytest.approx(-sv, abs=1e-6)
            assert got["throttle"] == pytest.approx(thr, abs=1e-6) and thr_new == pytest.approx(thr, abs=1e-6)
            assert np.abs(np.array(got["mpc_x"]) - r["z"][:cd["N"]]).max() < ABS_TOL
            assert np.allclose(got["next_x"], ex["tx"], atol=1e-9)
            # and without trajectories the reference's NULLs serialise as 0
            reply0, _ = S.telemetry_step('42["telemetry",' + json.dumps(body) + ']', thr_prev, tau)
            assert '"mpc_x":0,"mpc_y":0,"next_x":0,"next_y":0' in reply0
            # advance the "simulator" a little for the next message
            px, py, psi, v = clr.vehicle_move(px, py, psi, v, -got["steering_angle"] * cd["max_steering"], (thr - v / 50.0) * 6, cd["Lf"], 0.1)
            steer_sim = got["steering_angle"] * cd["max_steering"]
            thr_prev = thr_new
    S.close()
