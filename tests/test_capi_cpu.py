"""No-GPU checks of the product library: it loads, exports every symbol the header declares,
parses config-*.json exactly like Config::load, and refuses to compute without a device."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(mpc):
    hdr = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = mpc.lib()
    for name in declared:
        assert hasattr(L, name), "missing export %s" % name
    assert declared == set(mpc.EXPORTS)
    assert b"sm_100a" in L.mpc_version()


def test_binding_constants_match_the_header(mpc):
    """The ctypes binding passes kernel kinds and tail flags as literals: they must be the header's."""
    hdr = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    macro = lambda name: int(re.search(r"#define\s+%s\s+(-?\d+)" % name, hdr).group(1))
    for name, val in (("MPC_KERNEL_AUTO", mpc.KERNEL_AUTO), ("MPC_KERNEL_LANE", mpc.KERNEL_LANE), ("MPC_KERNEL_COOP", mpc.KERNEL_COOP)):
        assert macro(name) == val, name
    # Solver.set_tail builds flags as 1 | 4
    assert (macro("MPC_TAIL_SORT_RAGGED"), macro("MPC_TAIL_LATE_COPY")) == (1, 4)


def test_config_struct_size_matches_header(mpc):
    # 4 ints + 8 doubles + 12 + 16 + 16 doubles + tol + 2 ints + tiny_step_tol + 2 ints + 5 doubles + 2 ints + 32 doubles
    assert C.sizeof(mpc.MpcConfig) == 16 + 8 * (8 + 12 + 16 + 16 + 1) + 8 + 8 + 8 + 8 * 5 + 8 + 8 * 32


@pytest.mark.parametrize("name", ["stable", "fast", "no-latency"])
def test_config_load_matches_config_cpp(mpc, po, refdata, name, tmp_path):
    """Config::load conversions (Config.cpp:39-86): product C++ loader == Python restatement."""
    js = refdata["configs"][name]
    path = tmp_path / ("config-%s.json" % name)
    path.write_text(json.dumps(js, indent=2))
    got = mpc.config_from_json_file(str(path)).as_dict()
    want = po.load_config_dict(js)
    for k, v in want.items():
        if isinstance(v, list):
            assert np.array_equal(np.array(got[k][: len(v)]), np.array(v)), k
        else:
            assert got[k] == v, k
    # SURVEY.md App. A.5 numerics for config-stable
    if name == "stable":
        assert got["max_accel"] == pytest.approx(4.4703889, abs=1e-7)
        assert got["max_decel"] == pytest.approx(-8.9407778, abs=1e-7)
        assert got["max_steering"] == pytest.approx(0.4363323, abs=1e-7)
        assert got["max_speed"] == pytest.approx(53.6446667, abs=1e-7)
        assert got["steer_speeds"][0] == got["max_speed"]


def test_config_errors(mpc, tmp_path):
    with pytest.raises(mpc.MpcError, match="MPC_EIO"):
        mpc.config_from_json_file(str(tmp_path / "nope.json"))
    with pytest.raises(mpc.MpcError, match="MPC_EPARSE"):
        mpc.config_from_json_text("{\"N\": 10")
    with pytest.raises(mpc.MpcError, match="MPC_EPARSE"):
        mpc.config_from_json_text("{\"N\": 10, \"dt\": 0.1}")          # missing keys
    bad = {"N": 10, "weights": [1, 2, 3]}
    with pytest.raises(mpc.MpcError, match="MPC_EPARSE"):
        mpc.config_from_json_text(json.dumps(bad))


def test_defaults_are_config_cpp_statics(mpc):
    d = mpc.config_defaults()
    assert (d.N, d.dt, d.Lf, d.max_fit_order) == (25, 0.025, 2.67, 4)       # Config.cpp:5-17
    assert d.max_speed == pytest.approx(100 * 1609.34 / 3600.0)
    assert list(d.weights[:8]) == [100, 100, 1, 1, 1, 5000, 1, 1000]
    assert (d.max_iter, d.tol) == (3000, 1e-8)


def test_no_cpu_fallback(mpc, stable_cfg):
    """Without a CUDA device the product must fail loudly (MPC_ENODEV), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mpc.MpcError, match="MPC_ENODEV"):
        mpc.Solver(stable_cfg, 0)


def test_create_rejects_bad_config(mpc, stable_cfg):
    import copy
    for field, val in [("N", 1), ("N", mpc.NMAX + 1), ("dt", 0.0), ("Lf", -1.0), ("n_steer_speeds", 0)]:
        c = mpc.MpcConfig.from_buffer_copy(stable_cfg)
        setattr(c, field, val)
        with pytest.raises(mpc.MpcError, match="MPC_EINVAL"):
            mpc.Solver(c, 0)


def test_workload_generator_matches_scalar_preprocessing(mpc, po, stable_cd, refdata):
    """workloads.preprocess_batch (vectorised MPC::run pre-processing) == the scalar restatement."""
    b = mpc.workloads.batch_perturbed_states(64, 0, stable_cd)
    wx, wy = np.array(refdata["waypoints"]["x"]), np.array(refdata["waypoints"]["y"])
    for i in range(64):
        win = (b["segment"][i] + np.arange(6)) % len(wx)
        st, co, lo, hi, ex = po.preprocess(stable_cd, (b["px"][i], b["py"][i], b["psi"][i], b["v"][i]), wx[win], wy[win])
        assert ex["order"] == b["fit_order"][i]
        assert np.allclose(co, b["coeffs"][i], rtol=1e-7, atol=1e-9)
        assert np.allclose(st, b["state"][i], rtol=1e-9, atol=1e-9)
        assert (lo, hi) == pytest.approx((b["yaw_lo"][i], b["yaw_hi"][i]), rel=1e-7, abs=1e-9)
    # determinism and the characterisation of SURVEY.md App. C
    b2 = mpc.workloads.batch_perturbed_states(64, 0, stable_cd)
    assert np.array_equal(b["state"], b2["state"]) and np.array_equal(b["coeffs"], b2["coeffs"])
    big = mpc.workloads.batch_perturbed_states(4096, 0, stable_cd)
    assert 0.2 < (np.abs(big["state"][:, 4]) > 0.8).mean() < 0.5
    assert 0.3 < (np.abs(big["state"][:, 5]) > 0.1).mean() < 0.55


def test_command_line_like_mpc_main(mpc, refdata, tmp_path):
    """src/mpc_main.cpp:55-79, 238-246: option parsing, config-file choice, post-load overrides."""
    for name in refdata["configs"]:
        (tmp_path / ("config-%s.json" % name)).write_text(json.dumps(refdata["configs"][name]))
    d = str(tmp_path)
    cfg, f = mpc.config_from_cli([], d)
    assert f.endswith("config-stable.json") and cfg.max_speed == pytest.approx(120 * 1609.34 / 3600)
    cfg, f = mpc.config_from_cli(["-fast"], d)
    assert f.endswith("config-fast.json") and cfg.max_speed == pytest.approx(145 * 1609.34 / 3600)
    cfg, f = mpc.config_from_cli(["-latency", "0"], d)
    assert f.endswith("config-no-latency.json") and cfg.latency_ms == 0
    cfg, f = mpc.config_from_cli(["-latency", "0", "-fast"], d)           # a later option overrides the file choice
    assert f.endswith("config-fast.json") and cfg.latency_ms == 0
    assert cfg.lookahead == pytest.approx(0.1)                             # lookahead keeps the file's value (reference quirk)
    base = mpc.config_from_json_text(json.dumps(refdata["configs"]["stable"]))
    cfg, f = mpc.config_from_cli(["-speed", "60", "-latency", "50"], d)
    assert cfg.latency_ms == 50 and cfg.max_speed == pytest.approx(60 * 1609.34 / 3600)
    assert list(cfg.steer_speeds) == list(base.steer_speeds)               # tables are not rescaled
    cfg, f = mpc.config_from_cli(["-config", str(tmp_path / "config-fast.json")], d)
    assert f.endswith("config-fast.json")
    for bad in (["-bogus"], ["-speed", "fast"], ["-latency"], ["-config"]):
        with pytest.raises(mpc.MpcError, match="MPC_EINVAL"):
            mpc.config_from_cli(bad, d)
    with pytest.raises(mpc.MpcError, match="MPC_EIO"):
        mpc.config_from_cli(["-config", str(tmp_path / "missing.json")], d)


def test_telemetry_message_framing(mpc):
    """hasData() and event dispatch of src/mpc_main.cpp:26-36, 81-97, 215-219."""
    body = {"ptsx": [1.0, 2.0, 3.5, 5.0, 7.0, 9.0], "ptsy": [0.5, 0.25, 0.0, -0.5, -1.0, -2.0], "psi_unity": 4.1, "psi": 3.7,
            "x": -40.62, "y": 108.73, "steering_angle": -0.04, "throttle": 0.3, "speed": 42.5}
    t = mpc.telemetry_parse('42["telemetry",' + json.dumps(body) + ']')
    assert t.kind == 2 and t.npts == 6
    assert (t.x, t.y, t.psi, t.speed_mph, t.steering_angle) == (-40.62, 108.73, 3.7, 42.5, -0.04)
    assert list(t.ptsx[:6]) == body["ptsx"] and list(t.ptsy[:6]) == body["ptsy"]
    assert mpc.telemetry_parse('42["telemetry",null]').kind == 1          # "null" anywhere -> manual driving
    assert mpc.telemetry_parse('42').kind == 0 and mpc.telemetry_parse('2probe').kind == 0
    assert mpc.telemetry_parse('42["steer",{"a":1}]').kind == 0           # other events are ignored
    with pytest.raises(mpc.MpcError, match="MPC_EPARSE"):
        mpc.telemetry_parse('42["telemetry",{"x":1}]')


def test_run_level_tables_are_validated(mpc, stable_cfg, refdata):
    """counts of the run()-level tables outside the fixed-size arrays are argument errors at mpc_create, and
    MPC::run's pre-processing refuses an empty speed-limit table (the reference calls .back() on an empty vector there,
    Vehicle.cpp:66-79) instead of silently using a limit of 0."""
    for field, val in [("n_yaw_changes", mpc.NTAB + 1), ("n_yaw_changes", -1), ("n_yaw_change_speeds", mpc.NTAB + 1), ("n_steers", mpc.NTAB + 1)]:
        c = mpc.MpcConfig.from_buffer_copy(stable_cfg)
        setattr(c, field, val)
        with pytest.raises(mpc.MpcError, match="MPC_EINVAL"):
            mpc.Solver(c, 0)
    fx = refdata["test_cpp_fixtures"][0]
    pose = (fx["x"], fx["y"], fx["psi"], fx["v"])
    ok = mpc.run_prepare(stable_cfg, pose, fx["ptsx"], fx["ptsy"])
    assert ok["aux"].target_speed > 0
    c = mpc.MpcConfig.from_buffer_copy(stable_cfg)
    c.n_yaw_change_speeds = 0
    with pytest.raises(mpc.MpcError, match="MPC_EINVAL"):
        mpc.run_prepare(c, pose, fx["ptsx"], fx["ptsy"])
