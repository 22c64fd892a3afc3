"""Worker of tests/test_multi_gpu.py: launched by torch.distributed.run, one rank per GPU.

Every rank builds the same map, registers ITS contiguous shard of the same scan with the accumulators all-reduced across
ranks (comm = "peer": in-kernel mailbox exchange over NVLink; "nccl": ncclAllReduce + separate solve) and writes
pose / fitness / iterations and the single-iteration sums to <out>/rank<r>.npz."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    out, comm, method = sys.argv[1], sys.argv[2], int(sys.argv[3])
    import torch
    import torch.distributed as dist
    import elimaloc_b200 as E
    from elimaloc_b200 import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    raw = synth.map_u(60_000, 16.0, origin=-4.0)
    gm = E.VoxelHashMap(1.0, 30, device=local)
    gm.AddPoints(raw)
    gm.CalVoxelCovAll()
    gm.CalPointCovAll(0.4)
    T_true = synth.se3([2.0, 3.0, 2.5], [0.01, -0.02, 0.2])
    scan = synth.scan_m(gm.Pointcloud(), 6001, T_true)  # odd size: ragged shards
    T0 = T_true @ synth.canonical_offset()
    lo, hi = len(scan) * rank // world, len(scan) * (rank + 1) // world
    reg = E.Registration(device=local)
    if comm == "peer":
        reg.peer_setup(dist)
    else:
        ids = [E.Registration.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        reg.set_comm(ids[0], rank, world)
    cfg = E.RegistrationConfig(icp_method=method, max_iteration=8, **synth.timing_knobs())
    res = []
    for _ in range(3):
        T, ok, fit, cov = reg.RunRegister(scan[lo:hi], gm, T0, cfg)
        res.append(T)
    lin = reg.linearize(scan[lo:hi], gm, T0, cfg)
    np.savez(os.path.join(out, f"rank{rank}.npz"), T=np.stack(res), ok=ok, fit=fit, cov=cov, JTJ=lin["JTJ"], JTr=lin["JTr"],
             n_corr=lin["n_corr"], residual_sum=lin["residual_sum"])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
