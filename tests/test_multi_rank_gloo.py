"""world_size-2 test of the multi-GPU plan on CPU (gloo): the scan is split into contiguous shards, every rank linearises
its shard against the replicated map, and ONE all-reduce(sum) of the 30 accumulators per iteration reproduces the
single-rank sums — which is exactly what elm_registration_set_comm arranges with ncclAllReduce on the GPUs.
The per-shard linearisation is done by the oracle here (no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_bounds(n, rank, world):
    """same split as bench.py: contiguous chunks [n*r/W, n*(r+1)/W)"""
    return n * rank // world, n * (rank + 1) // world


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from elimaloc_b200 import synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw = synth.map_u(20_000, 10.0, origin=-2.0)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    T_true = synth.se3([2.0, 3.0, 2.5], [0.01, -0.02, 0.2])
    scan = synth.scan_m(om.export()["pxyz"], 1001, T_true)  # odd size: ragged shards
    T0 = T_true @ synth.canonical_offset()
    lo, hi = shard_bounds(len(scan), rank, world)
    reg = O.Registration()
    res = {}
    for method in (O.P2P, O.VGICP):
        cfg = O.make_config(icp_method=method, **synth.timing_knobs())
        lin = reg.linearize(scan[lo:hi], om, T0, cfg)
        acc = np.zeros(32)
        iu = np.triu_indices(6)
        acc[:21] = lin["JTJ"][iu]
        acc[21:27] = lin["JTr"]
        acc[27], acc[28], acc[29] = lin["residual_sum"], lin["n_corr"], hi - lo
        t = torch.from_numpy(acc)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        full = reg.linearize(scan, om, T0, cfg)
        ref = np.zeros(32)
        ref[:21] = full["JTJ"][iu]
        ref[21:27] = full["JTr"]
        ref[27], ref[28], ref[29] = full["residual_sum"], full["n_corr"], len(scan)
        res[method] = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_accumulators_allreduce_to_the_single_rank_sums():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for method, err in res.items():
        assert err < 1e-12, (method, err)


def test_shard_bounds_cover_the_scan_exactly():
    for n in (0, 1, 7, 1001, 131072):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
