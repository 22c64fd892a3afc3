"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the scan sharded over 2 / 4 / 8 ranks with the accumulators
all-reduced through peer memory (or NCCL) must give every rank the SAME bits, and the same registration as one GPU over
the whole scan up to the summation order of the shards (1e-9).  All four methods."""
import os
import subprocess
import sys

import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


CASES = [(2, "peer", m) for m in (E.P2P, E.GICP, E.VGICP, E.AVGICP)] + [(2, "nccl", E.P2P), (2, "nccl", E.VGICP)] + \
        [(4, "peer", m) for m in (E.P2P, E.GICP, E.VGICP, E.AVGICP)] + [(8, "peer", m) for m in (E.P2P, E.GICP, E.VGICP, E.AVGICP)] + \
        [(8, "nccl", E.GICP)]


@pytest.mark.parametrize("world,comm,method", CASES)
def test_ranks_match_one_gpu(tmp_path, world, comm, method):
    if E.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() + 7 * method + 31 * world + (3 if comm == "peer" else 0)) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(tmp_path), comm, str(method)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    r0 = np.load(tmp_path / "rank0.npz")
    for r in range(1, world):
        r1 = np.load(tmp_path / f"rank{r}.npz")
        for k in ("T", "fit", "cov", "JTJ", "JTr", "n_corr", "residual_sum"):
            assert np.array_equal(r0[k], r1[k]), (r, k)  # bit-identical on every rank: same sums in the same (rank) order
    assert np.array_equal(r0["T"][0], r0["T"][1]) and np.array_equal(r0["T"][0], r0["T"][2])  # run-to-run reproducible
    # one GPU, whole scan
    raw = synth.map_u(60_000, 16.0, origin=-4.0)
    gm = E.VoxelHashMap(1.0, 30, device=0)
    gm.AddPoints(raw)
    gm.CalVoxelCovAll()
    gm.CalPointCovAll(0.4)
    T_true = synth.se3([2.0, 3.0, 2.5], [0.01, -0.02, 0.2])
    scan = synth.scan_m(gm.Pointcloud(), 6001, T_true)
    T0 = T_true @ synth.canonical_offset()
    reg = E.Registration(device=0)
    cfg = E.RegistrationConfig(icp_method=method, max_iteration=8, **synth.timing_knobs())
    T, ok, fit, cov = reg.RunRegister(scan, gm, T0, cfg)
    lin = reg.linearize(scan, gm, T0, cfg)
    assert lin["n_corr"] == int(r0["n_corr"])
    assert np.abs(lin["JTJ"] - r0["JTJ"]).max() <= 1e-9 * np.abs(lin["JTJ"]).max()
    assert np.abs(T - r0["T"][0]).max() <= 1e-9 * np.abs(T).max()
    assert bool(r0["ok"]) == ok and abs(float(r0["fit"]) - fit) <= 1e-9 * max(1.0, abs(fit))
    if method == E.GICP:
        assert np.abs(cov - r0["cov"]).max() <= 1e-8 * max(np.abs(cov).max(), 1e-300)
