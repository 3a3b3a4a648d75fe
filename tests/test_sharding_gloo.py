"""World-size-2 (and 3, uneven) CPU test of the multi-GPU plumbing over gloo: shard by batch index, solve
the shard, gather.  The per-shard solve here is the CPU oracle (this is a test of the sharding logic; on
GPUs the same code runs with the CUDA solve and NCCL -- bench.py --gpus N)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import mpc_b200 as mpc
    from carnd_mpc_project_b200 import sharding
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rd = mpc.workloads.reference_data()
    cd = po.load_config_dict(rd["configs"]["stable"])
    cfg = po.make_config(cd)
    b = mpc.workloads.batch_perturbed_states(B, 5, cd)                 # same seed on every rank

    def solve_fn(state, coeffs, yaw_lo, yaw_hi):
        r = po.solve_batch(cfg, po.problems_from_arrays(state, coeffs, yaw_lo, yaw_hi), 2)
        return {"result": r["result"], "traj_x": r["traj_x"], "status": r["status"], "iters": r["iters"]}

    inputs = {k: b[k] for k in ("state", "coeffs", "yaw_lo", "yaw_hi")}
    out = sharding.solve_sharded(solve_fn, inputs, B, world, rank, dist)
    lo, hi = sharding.shard_bounds(B, rank, world)
    q.put((rank, lo, hi, out["result"], out["traj_x"], out["status"], out["iters"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 48), (3, 50)])
def test_sharded_solve_equals_unsharded(world, B, po, stable_cd):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import mpc_b200 as mpc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = mpc.workloads.batch_perturbed_states(B, 5, stable_cd)
    ref = po.solve_batch(po.make_config(stable_cd), po.problems_from_arrays(b["state"], b["coeffs"], b["yaw_lo"], b["yaw_hi"]), 2)
    covered = np.zeros(B, dtype=int)
    for rank, lo, hi, res, tx, st, it in got:
        covered[lo:hi] += 1
        assert res.shape == (B, 9) and tx.shape == (B, stable_cd["N"])
        assert np.array_equal(res, ref["result"]) and np.array_equal(tx, ref["traj_x"])     # every rank holds the full, identical result
        assert np.array_equal(st, ref["status"]) and np.array_equal(it, ref["iters"])
    assert (covered == 1).all()                                                            # shards partition the batch


def test_shard_bounds_partition():
    from carnd_mpc_project_b200 import sharding
    for B in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_bounds(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
