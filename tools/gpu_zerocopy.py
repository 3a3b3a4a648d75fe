"""Experiment: the batch kernels reading inputs from / writing outputs to pinned host memory directly (zero copy)
against the staged host path (H2D copies, solve, D2H copies) and the device-resident solve."""
import json, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
b = mpc.workloads.batch_perturbed_states(B, 0, cfg.as_dict())
N = cfg.N
dev = torch.device('cuda:0')
S = mpc.Solver(cfg, 0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).pin_memory()
hin = [pin(b['state']), pin(b['coeffs']), pin(b['yaw_lo']), pin(b['yaw_hi'])]
din = [t.to(dev) for t in hin]
def outs(where):
    mk = (lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt).pin_memory()) if where == 'host' else (lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt, device=dev))
    return [mk(9, B), mk(N, B), mk(N, B), None, mk(B, dt=torch.int32), mk(B, dt=torch.int32)]
hout, dout = outs('host'), outs('dev')
def run(ins, o):
    S.solve_batch_device(B, *ins, *o); torch.cuda.synchronize()
def timeit(fn, reps=8):
    fn(); fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts), 1e3 * np.median(ts)
print('device in, device out      best %.3f ms  median %.3f ms' % timeit(lambda: run(din, dout)))
print('pinned in, device out      best %.3f ms  median %.3f ms' % timeit(lambda: run(hin, dout)))
print('device in, pinned out      best %.3f ms  median %.3f ms' % timeit(lambda: run(din, hout)))
print('pinned in, pinned out      best %.3f ms  median %.3f ms' % timeit(lambda: run(hin, hout)))
hn = [t.numpy() for t in hin]
print('mpc_solve_batch_host       best %.3f ms  median %.3f ms' % timeit(lambda: S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])))
print('identical', torch.equal(hout[0], dout[0].cpu()), torch.equal(hout[1], dout[1].cpu()), torch.equal(hout[4], dout[4].cpu()))
