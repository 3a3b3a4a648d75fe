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
