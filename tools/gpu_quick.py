"""Quick iteration check: lane kernel chain vs coop kernel on 2048 problems (must be bit-identical), then timings and the crossover."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as mpc
rd = mpc.workloads.reference_data()
cfg = mpc.config_from_json_text(json.dumps(rd['configs']['stable']))
cd = cfg.as_dict()
S = mpc.Solver(cfg, 0)
b = mpc.workloads.batch_perturbed_states(2048, 3, cd)
S.set_kernel(mpc.KERNEL_LANE); g = S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
S.set_kernel(mpc.KERNEL_COOP); c = S.solve_batch_host(b['state'], b['coeffs'], b['yaw_lo'], b['yaw_hi'])
print('coop vs lane: max diff %.3g  iters equal %.4f  status ok %.4f' % (np.abs(g['result'] - c['result']).max(), (g['iters'] == c['iters']).mean(), (c['status'] == 1).mean()))
dev = torch.device('cuda:0')
S.set_kernel(mpc.KERNEL_LANE)
for B in [int(x) for x in (sys.argv[1:] or ['4096', '65536', '1048576'])]:
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).to(dev)
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    N = cfg.N
    outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), torch.zeros(N, B, dtype=torch.float64, device=dev), None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('B=%7d  %.3f ms  %.0f solves/s  %.1f ns/solve  csum=%.9g' % (B, best, B / best * 1e3, best * 1e6 / B, outs[0][8].sum().item()))

for B in (1, 256, 1024, 2048, 4096, 8192, 16384):
    b = mpc.workloads.batch_perturbed_states(B, 0, cd)
    ins = [up(b['state']), up(b['coeffs']), up(b['yaw_lo']), up(b['yaw_hi'])]
    outs = [torch.zeros(9, B, dtype=torch.float64, device=dev), None, None, None, torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)]
    line = 'B=%6d ' % B
    for kind, nm in ((mpc.KERNEL_LANE, 'lane'), (mpc.KERNEL_COOP, 'coop')):
        S.set_kernel(kind)
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_batch_device(B, *ins, *outs); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        line += ' %s %.3f ms' % (nm, best)
    print(line + '  iters max %d' % outs[5].max().item())
