"""When do the warps / CTAs of the lane kernel's main launch run out of work?  (development build -DMPC_DEBUG_TIMES)"""
import json, sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd["configs"]["stable"]))
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
B = 65536
L = mpc.lib()
M = 4 + 148 * 9 * 2
for seed in (0, 1):
    b = mpc.workloads.batch_perturbed_states(B, seed, cfg.as_dict())
    ins = [up(b["state"]), up(b["coeffs"]), up(b["yaw_lo"]), up(b["yaw_hi"])]
    res = torch.zeros(9, B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    S = mpc.Solver(cfg, 0)
    for r in range(3):
        L.mpc_debug_times(None, 0, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, *ins, res, None, None, None, st, it); e1.record(); torch.cuda.synchronize()
    buf = (C.c_ulonglong * M)()
    L.mpc_debug_times(buf, M, 0)
    t = np.array(buf[:], dtype=np.float64)
    t0 = t[0]
    w = t[4:].reshape(148, 9, 2)[:, :7, :]          # 7 warps of 224 threads
    wd = (w[:, :, 0] - t0) / 1e3                      # warp out of work (us after kernel start)
    we = (w[:, :, 1] - t0) / 1e3                      # warp (= CTA) exit
    cta_exit = we.max(axis=1)
    print("seed %d: chain %.3f ms; main launch ends %.0f us after its start" % (seed, e0.elapsed_time(e1), cta_exit.max()))
    q = lambda x: " ".join("%.0f" % v for v in np.percentile(x, [0, 5, 25, 50, 75, 95, 100]))
    print("  warp out of work  (us, min 5%% 25%% 50%% 75%% 95%% max): %s" % q(wd))
    print("  CTA exit          (us, min 5%% 25%% 50%% 75%% 95%% max): %s" % q(cta_exit))
    idle_warp = (cta_exit[:, None] - wd).mean()
    idle_cta = (cta_exit.max() - cta_exit).mean()
    print("  mean time a warp idles in its CTA after running out of work: %.0f us; mean time an SM idles after its CTA exits, until the launch ends: %.0f us" % (idle_warp, idle_cta))
    S.close()
