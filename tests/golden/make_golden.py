#!/usr/bin/env python
"""Generates tests/golden/config1_<method>.npz from the oracle (the reference ships no golden vectors and cannot be
built here, so these pin the ORACLE against regressions; parity with the reference itself stays "unpinned").

    python tests/golden/make_golden.py

Inputs are BASELINE config 1 (4096-point Scan-M vs 100 k-raw-point Map-U straddling the origin, 10 forced iterations)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from elimaloc_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def world():
    raw = synth.map_u(100_000, 21.5, origin=-6.0)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    stored = om.export()["pxyz"]
    T_true = synth.se3([4.0, 5.0, 3.5], [0.02, -0.01, 0.3])
    scan = synth.scan_m(stored, 4096, T_true)
    T0 = T_true @ synth.canonical_offset()
    return om, scan, T0


def main():
    om, scan, T0 = world()
    e = om.export()
    np.savez_compressed(os.path.join(HERE, "config1_map_digest.npz"), n_voxels=om.num_voxels(), n_points=om.num_points(),
                        keys_head=e["keys"][:64], counts_head=e["counts"][:64], pxyz_head=e["pxyz"][:64],
                        vcov_head=e["vcov"][:16], pcov_head=e["pcov"][:16], pmean_head=e["pmean"][:16],
                        keys_sum=e["keys"].astype(np.int64).sum(axis=0), pxyz_sum=e["pxyz"].astype(np.float64).sum(axis=0))
    reg = O.Registration()
    for m, name in enumerate(["p2p", "gicp", "vgicp", "avgicp"]):
        cfg = O.make_config(icp_method=m, max_iteration=10, **synth.timing_knobs())
        r = reg.RunRegister(scan, om, T0, cfg)
        cnt, tgt = O.correspondences(om, scan[:256], T0, m, 5.0)
        np.savez_compressed(os.path.join(HERE, f"config1_{name}.npz"), pose=r["pose"], fitness=r["fitness_score"],
                            local_cov=r["local_cov"], is_success=r["is_success"], n_iter=r["n_iter"], JTJ=r["trace"]["JTJ"],
                            JTr=r["trace"]["JTr"], res=r["trace"]["res"], ncorr=r["trace"]["ncorr"],
                            pose_out=r["trace"]["pose_out"], corr_count=cnt, corr_target=tgt)
        print(name, r["n_iter"], r["fitness_score"])


if __name__ == "__main__":
    main()
