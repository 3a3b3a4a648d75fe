#!/usr/bin/env python
"""bench.py -- MPC solves/s on the BASELINE.json workload (N=10 bicycle model, batch 64K per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one pass of the hot path (one `mpc_solve_batch` launch) over one batch of B synthetic
problems: SURVEY.md 8d item 2 = BASELINE.json configs[1] ("batch 64K independent N=10 dt=0.1 solves
from perturbed initial states (cte, epsi, v) on 1 B200"), config-stable knobs.  With N GPUs every
rank solves its own B problems (weak scaling; problems are independent, no collective in the solve,
one result all_gather at the end of the step).  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's CPU path.  Ipopt/CppAD/MUMPS cannot be installed in this
image, so that arm runs the CPU restatement in oracle/ (the one place besides
cpu_baseline where this file executes oracle/) on all host threads, on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "MPC solves/sec (N=10 bicycle, batch 64K per GPU)"
UNIT = "solves/s"


def f_iter(N):
    """Algorithmic FLOPs per interior-point iteration, SURVEY.md 8d / BASELINE.md section 4."""
    return (N - 1) * (1235 + 250) + 30 * (14 * N - 2)


BYTES_PER_SOLVE = lambda N: 104 + 8 * (9 + 2 * N) + 8   # in: 13 doubles; out: result+traj+status,iters


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload(mpc, B, seed):
    rd = mpc.workloads.reference_data()
    cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["stable"]))
    return cfg, mpc.workloads.batch_perturbed_states(B, seed, cfg.as_dict()), rd


def cpu_port(rd, batch, n_problems, threads):
    """The oracle (CPU restatement of the reference path) on the first n_problems of the batch."""
    from oracle import pyoracle as po
    ocfg = po.make_config(po.load_config_dict(rd["configs"]["stable"]))
    s = slice(0, n_problems)
    probs = po.problems_from_arrays(batch["state"][s], batch["coeffs"][s], batch["yaw_lo"][s], batch["yaw_hi"][s])
    t0 = time.perf_counter()
    out = po.solve_batch(ocfg, probs, threads)
    return time.perf_counter() - t0, out


def _ref_worker(job):
    """One worker process of the reference arm: the reference's own sources (oracle/_ref) keep global mutable
    Config statics, so parallelism is by process, one solve at a time per process."""
    js, state, coeffs, ylo, yhi, N = job
    from oracle import pyref as pr
    pr.config_load(js)
    t0 = time.perf_counter()
    ok, its = 0, 0
    for i in range(state.shape[0]):
        r = pr.solve(state[i], coeffs[i], ylo[i], yhi[i], N)
        ok += int(r["status"] == 1); its += r["iters"]
    return time.perf_counter() - t0, ok, its


def ref_build_run(rd, batch, n_problems, procs, pool):
    js = rd["configs"]["stable"]
    idx = np.array_split(np.arange(n_problems), procs)
    jobs = [(js, batch["state"][i], batch["coeffs"][i], batch["yaw_lo"][i], batch["yaw_hi"][i], js["N"]) for i in idx if len(i)]
    t0 = time.perf_counter()
    res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    return wall, sum(r[1] for r in res), sum(r[2] for r in res)


def run_reference(args, rank):
    """The reference arm.  Where oracle/_ref exists (the reference's own MPC.cpp / Vehicle / RoadGeometry / Config
    compiled unmodified against the CppAD / Ipopt stand-ins of oracle/ref_shim -- built in the build container, it
    travels with the snapshot) that is what is timed: MPC::solve as the reference runs it, tape recording and AD
    sweeps per solve included, one process per host core.  Otherwise the plain-C oracle port."""
    if rank != 0:
        return
    import mpc_b200 as mpc   # only for the workload generator / config parser (no GPU use)
    cfg, batch, rd = workload(mpc, args.batch, 0)
    cores = os.cpu_count() or 1
    n_pilot = min(args.batch, 64 * cores)
    cpu_port(rd, batch, n_pilot, cores)
    t_pilot, _ = cpu_port(rd, batch, n_pilot, cores)
    per = t_pilot / n_pilot                                          # wall seconds per solve, all cores busy
    sample = int(max(cores, min(args.batch, 2.0 / max(per, 1e-9))))  # ~2 s of wall per step
    t_port, out_port = cpu_port(rd, batch, sample, cores)
    port_value = sample / t_port
    from oracle import pyref
    kind = "reference" if pyref.available() else "port"
    if kind == "reference":
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(cores)
        w, _, _ = ref_build_run(rd, batch, 2 * cores, cores, pool)          # pilot (also pages the library in)
        w, _, _ = ref_build_run(rd, batch, 2 * cores, cores, pool)
        sample = int(max(cores, min(args.batch, 2.0 / max(w / (2 * cores), 1e-9))))
        for _ in range(args.warmup):
            ref_build_run(rd, batch, sample, cores, pool)
        times, ok, its = [], 0, 0
        for _ in range(args.steps):
            w, ok, its = ref_build_run(rd, batch, sample, cores, pool)
            times.append(w)
        pool.close()
        ok_frac, it_mean = ok / sample, its / sample
        sample_txt = ("first %d of the %d-problem batch per step, %d processes (one per host core); the reference's own "
                      "MPC::solve / FG_eval / Config sources (oracle/_ref, compiled unmodified against the CppAD/Ipopt "
                      "stand-ins of oracle/ref_shim: tape recorded per solve, AD Jacobian/Hessian, dense LDL^T; the "
                      "interior-point core is the oracle's restatement of Ipopt, which is not installable here)" % (sample, args.batch, cores))
    else:
        for _ in range(args.warmup):
            cpu_port(rd, batch, sample, cores)
        times = []
        for _ in range(args.steps):
            t, out = cpu_port(rd, batch, sample, cores)
            times.append(t)
        ok_frac, it_mean = float((out["status"] == 1).mean()), float(out["iters"].mean())
        sample_txt = ("first %d of the %d-problem batch per step, %d host threads, CPU restatement of "
                      "MPC::solve+Ipopt (oracle/mpc_oracle.c; Ipopt/CppAD not installable here)" % (sample, args.batch, cores))
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: batch 64K independent N=10 dt=0.1 solves, perturbed (cte, epsi, v), config-stable",
                       "batch_per_step": sample, "N": cfg.N, "dt": cfg.dt},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_txt},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "status_ok_frac": ok_frac, "iters_mean": it_mean, "host": {"cpu": cpu_model(), "logical_cpus": os.cpu_count()},
            "oracle_port_value": {"value": port_value, "unit": UNIT, "what": "the plain-C restatement (analytic derivatives, no tape) on the same cores, for comparison"}}
    emit(line)


def extras(mpc, torch, dev, rd, local_rank, fp64_peak=None):
    """The other BASELINE.json configs on ONE GPU, informational (the headline stays configs[1]): a 1M batch (the
    steady-state rate without the tail), the N x dt grid in one ragged launch (config 3), a per-problem weight
    sweep (config 4), and closed-loop rollouts with 100 ms latency (config 5, one GPU's share: 1024 vehicles)."""
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    out = {}
    js = rd["configs"]["stable"]
    cfg = mpc.config_from_json_text(json.dumps(js))
    cd = cfg.as_dict()
    # 1M batch
    B = 1 << 20
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    ins = [up(b["state"].T), up(b["coeffs"].T), up(b["yaw_lo"]), up(b["yaw_hi"])]
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    S = mpc.Solver(cfg, local_rank)
    ms = timed(lambda: S.solve_batch_device(B, *ins, res, None, None, None, st, it))
    out["batch_1M"] = {"solves_per_s": B / ms * 1e3, "ms": ms, "status_ok_frac": float((st == 1).float().mean().item()),
                       "fp64_flops_frac_of_measured_peak": None}
    out["batch_1M"]["tflops"] = f_iter(cfg.N) * float(it.sum().item()) / (ms * 1e-3) / 1e12
    if fp64_peak:
        out["batch_1M"]["fp64_flops_frac_of_measured_peak"] = out["batch_1M"]["tflops"] / fp64_peak
    # config 4: per-problem weights, 128K problems
    B4 = 131072
    rng = np.random.default_rng(2)
    sel = rng.integers(0, 4096, B4)
    W = np.tile(np.array(cd["weights"]), (B4, 1))
    W[:, 3] = np.exp(rng.uniform(np.log(1), np.log(5000), B4)); W[:, 4] = np.exp(rng.uniform(np.log(1), np.log(5000), B4))
    W[:, 1] = np.exp(rng.uniform(np.log(1), np.log(1000), B4)); W[:, 2] = rng.choice([0.01, 0.1, 1, 10, 100], B4)
    ins4 = [up(b["state"][sel].T), up(b["coeffs"][sel].T), up(b["yaw_lo"][sel]), up(b["yaw_hi"][sel])]
    Wd = up(W.T)
    ms = timed(lambda: S.solve_batch_device(B4, *ins4, res[:, :B4].contiguous(), None, None, None, st[:B4], it[:B4], weights=Wd), 2)
    out["config4_weight_sweep_128K"] = {"solves_per_s": B4 / ms * 1e3, "ms": ms, "status_ok_frac": float((st[:B4] == 1).float().mean().item()),
                                        "iters_max": int(it[:B4].max().item())}
    S.close()
    # config 3: N x dt grid, one ragged launch of 256K problems
    PAIRS = [(10, .1), (20, .1), (30, .1), (40, .1), (10, .05), (20, .05), (30, .05), (40, .05), (50, .05), (10, .02), (20, .02), (30, .02), (40, .02), (50, .02)]
    B3 = 262144
    rng = np.random.default_rng(1)
    pick = rng.integers(0, len(PAIRS), B3)
    Np = up(np.array([PAIRS[k][0] for k in pick], dtype=np.int32)); dtp = up(np.array([PAIRS[k][1] for k in pick]))
    cfg3 = mpc.config_from_json_text(json.dumps(dict(js, N=50)))
    S3 = mpc.Solver(cfg3, local_rank)
    ins3 = [t[..., :B3].contiguous() for t in ins]
    ms = timed(lambda: S3.solve_batch_device(B3, *ins3, res[:, :B3].contiguous(), None, None, None, st[:B3], it[:B3], N_per=Np, dt_per=dtp), 2)   # best of 2: the first call allocates the record buffers
    out["config3_horizon_grid_256K"] = {"solves_per_s": B3 / ms * 1e3, "ms": ms, "status_ok_frac": float((st[:B3] == 1).float().mean().item()),
                                        "iters_max": int(it[:B3].max().item()), "note": "N in 10..50 x dt in {0.1,0.05,0.02}; long horizons extrapolate the fit and a few percent end without success, as in the reference (SURVEY App. C)"}
    S3.close()
    # config 5: closed loop, config-fast (100 ms latency), 1024 vehicles (one GPU's share of 8192) x 200 steps
    cfgf = mpc.config_from_json_text(json.dumps(rd["configs"]["fast"]))
    cdf = cfgf.as_dict()
    V, T = 1024, 200
    bv = mpc.workloads.batch_perturbed_states(V, 3, cdf)
    veh = up(np.stack([bv["px"], bv["py"], bv["psi"], np.clip(bv["v"], 8, 30), np.zeros(V), np.zeros(V)]))
    seg = up(bv["segment"].astype(np.int32))
    pend = torch.zeros(2, V, dtype=torch.float64, device=dev)
    rec = torch.zeros(T, 8, V, dtype=torch.float64, device=dev)
    wx, wy = up(np.array(rd["waypoints"]["x"])), up(np.array(rd["waypoints"]["y"]))
    S5 = mpc.Solver(cfgf, local_rank)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S5.rollout_device(V, T, wx, wy, veh, seg, pend, 0.1, 0.02, rec); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["config5_closed_loop_1024x200"] = {"vehicle_steps_per_s": V * T / ms * 1e3, "ms_per_control_step": ms / T,
                                           "status_ok_frac": float((rec[:, 6] == 1).float().mean().item()),
                                           "median_abs_cte_final_m": float(rec[-1, 0].abs().median().item())}
    S5.close()
    return out


def cpu_model():
    """CPU model of the box (the host the cpu_baseline / reference arm runs on)."""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line at the first
    collective), so everything but that line goes to stderr: fd 1 is pointed at fd 2, the real stdout is kept aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (1M batch, sweeps, rollouts)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import mpc_b200 as mpc
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the MPC solve has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    cfg, batch, rd = workload(mpc, B, rank)        # rank r solves the seed-r batch (rank 0 = SURVEY's seed 0)
    N = cfg.N
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    state, coeffs, ylo, yhi = up(batch["state"]), up(batch["coeffs"]), up(batch["yaw_lo"]), up(batch["yaw_hi"])
    result = torch.zeros(9, B, dtype=torch.float64, device=dev)
    tx = torch.zeros(N, B, dtype=torch.float64, device=dev)
    ty = torch.zeros(N, B, dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    iters = torch.zeros(B, dtype=torch.int32, device=dev)
    gathered = torch.zeros(world, 9, B, dtype=torch.float64, device=dev) if world > 1 else None
    S = mpc.Solver(cfg, local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        S.solve_batch_device(B, state, coeffs, ylo, yhi, result, tx, ty, None, status, iters)
        if world > 1:   # the only collective: final result gather, after the solve
            dist.all_gather_into_tensor(gathered.view(-1), result.view(-1))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    sync_all()
    smi_index = local_rank
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    if cvd:
        ids = [x.strip() for x in cvd.split(",") if x.strip()]
        if local_rank < len(ids) and ids[local_rank].isdigit():
            smi_index = int(ids[local_rank])
    sampler = ClockSampler(smi_index)
    sampler.start()
    launches0 = S.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)           # L2 flush between timed iterations (outside the per-step events)
        ev[k][0].record()
        step()
        ev[k][1].record()
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = S.launches - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = tot.item() / args.steps
    value = B * world / (ms_per_step * 1e-3)

    it = iters.cpu().numpy()
    st = status.cpu().numpy()
    flops_per_launch = float(f_iter(N)) * float(it.sum())
    kernel_ms = float(np.mean(step_ms))      # the launches of one solve on this rank (all_gather excluded at N=1)

    # ---- e2e: the reference-facing call with HOST buffers (pinned), H2D + solve + D2H per step
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).pin_memory().numpy()
    h_state, h_coef, h_ylo, h_yhi = pin(batch["state"]), pin(batch["coeffs"]), pin(batch["yaw_lo"]), pin(batch["yaw_hi"])
    h_res = torch.zeros(9, B, dtype=torch.float64).pin_memory().numpy()
    h_tx = torch.zeros(N, B, dtype=torch.float64).pin_memory().numpy()
    h_ty = torch.zeros(N, B, dtype=torch.float64).pin_memory().numpy()
    h_st = torch.zeros(B, dtype=torch.int32).pin_memory().numpy()
    h_it = torch.zeros(B, dtype=torch.int32).pin_memory().numpy()
    L = mpc.lib()
    ptr = lambda a: a.ctypes.data

    def e2e_step():
        rc = L.mpc_solve_batch_host(S._h, B, ptr(h_state), ptr(h_coef), ptr(h_ylo), ptr(h_yhi), None, None, None,
                                    ptr(h_res), ptr(h_tx), ptr(h_ty), None, ptr(h_st), ptr(h_it))
        if rc != 0:
            raise SystemExit("mpc_solve_batch_host failed: %d %s" % (rc, L.mpc_last_error().decode()))

    for _ in range(3):
        e2e_step()
    sync_all()
    e2e_t = []
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()                      # returns after the D2H copies completed
        e2e_t.append(time.perf_counter() - t0)
    sync_all()
    e2e_tot = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_tot, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (e2e_tot.item() / args.steps)
    assert np.allclose(h_res, result.cpu().numpy(), rtol=0, atol=0), "host-path result differs from device path"
    h2d = 13 * 8 * B
    d2h = (9 + 2 * N) * 8 * B + 8 * B

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        fp64_peak = mpc.measure_fp64_peak(local_rank)
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this batch, from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("batch") == B and tj.get("N") == N:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
        except Exception:
            pass
        achieved_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_gbs = BYTES_PER_SOLVE(N) * B / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: batch 64K independent N=10 dt=0.1 solves from perturbed (cte, epsi, v), config-stable, seed=rank",
                       "batch_per_gpu": B, "N": N, "dt": cfg.dt, "l2": "flushed between timed steps (256 MiB write)",
                       "sharding": "independent batch shard per rank, no collective in the solve; one all_gather of result[9][B] per step" if world > 1 else "single GPU"},
            "roofline": {"bound": "fp64_fma", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "mpc_lane_kernel<10,1,false> (one problem per lane; dominant) + the launches that finish its tail: up to 3x mpc_lane_kernel<10,1,true> (parked problems, 32 to a warp; each returns at once when at most 8192 are parked) + mpc_coop_resume_kernel<10>; %d launches per step, kernel_ms is their sum" % (launches // args.steps) if B >= mpc.LANE_MIN_BATCH else "mpc_coop_kernel<10> (one problem per group of 16 lanes)",
                         "peak_source": "measured in this run by mpc_measure_fp64_peak (DFMA chains; MEASURED_PEAKS.json has no FP64 figure)",
                         "flops_per_launch": flops_per_launch, "flops_model": "sum_b iters_b * F_iter(N), F_iter(10)=17505 (SURVEY.md 8d)",
                         "kernel_ms": kernel_ms},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                             "bytes_per_solve": BYTES_PER_SOLVE(N),
                             "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_tot.item() / args.steps, "api": "mpc_solve_batch_host (pinned host buffers: inputs read in place over PCIe by the kernels; outputs copied back on a second stream beside the chain's final launch, whose problems a small kernel then rewrites in the host arrays)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "iters": {"mean": float(it.mean()), "p50": float(np.percentile(it, 50)), "p99": float(np.percentile(it, 99)), "max": int(it.max()),
                      "hist": {str(k): int(c) for k, c in enumerate(np.bincount(np.asarray(it, dtype=np.int64).clip(0))) if c}},
            "host": {"cpu": cpu_model(), "logical_cpus": os.cpu_count()},
            "status_ok_frac": float((st == 1).mean()),
        }
        if world == 1 and not args.no_latency:
            one = mpc.Solver(cfg, local_rank)
            lat = []
            for k in range(1200):
                i = k % B
                t0 = time.perf_counter()
                one.solve_one(batch["state"][i], batch["coeffs"][i], batch["yaw_lo"][i], batch["yaw_hi"][i])
                lat.append(time.perf_counter() - t0)
            lat = np.array(lat[200:]) * 1e6
            n50, n99 = one.measure_solve_latency(batch["state"][:1000], batch["coeffs"][:1000], batch["yaw_lo"][:1000], batch["yaw_hi"][:1000], 1000, 200)
            line["latency"] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)),
                               "native_p50_us": n50, "native_p99_us": n99,
                               "what": "mpc_solve_one host call -> result (B=1; inputs and result in mapped pinned host memory, one kernel launch + stream sync); p50/p99 through the Python binding, native_* from a C++ loop (mpc_measure_solve_latency) on the same 1000 problems", "batch_ms": ms_per_step}
            one.close()
        if world == 1 and not args.no_extras:
            line["extras"] = extras(mpc, torch, dev, rd, local_rank, fp64_peak)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            t_pilot, _ = cpu_port(rd, batch, min(B, 4 * cores), cores)
            per = t_pilot / min(B, 4 * cores)
            sample = int(max(cores, min(B, 10.0 / max(per, 1e-9))))          # ~10 s of wall on all cores
            t, out = cpu_port(rd, batch, sample, cores)
            ok = out["status"] == 1
            gres = result.cpu().numpy().T[:sample]
            dmax = float(np.abs(gres - out["result"])[ok][:, :8].max())
            line["cpu_baseline"] = {"value": sample / t, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "first %d problems of the same batch, one solve per host thread (%d threads), %.1f s; CPU restatement "
                                              "oracle/mpc_oracle.c (dense LDL^T), not Ipopt+CppAD+MUMPS" % (sample, cores, t),
                                    "max_abs_diff_vs_gpu": dmax}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
