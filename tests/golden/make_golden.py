#!/usr/bin/env python
"""Generates tests/golden/config1_<method>.npz and config1_map_digest.npz FROM THE REFERENCE'S OWN SOURCES.

    python tests/golden/make_golden.py          (build container only: needs /root/reference)

The reference ships no golden vectors, so these are outputs of the reference itself run here: oracle/_ref/libref.so is
registration.cpp + voxel_hash_map.{hpp,cpp} compiled unmodified from /root/reference against stand-in third-party headers
(oracle/ref_build/stubs — Eigen3 / oneTBB / PCL are not in this image; the stand-in keeps the reference's control flow and
restates only Eigen's arithmetic kernels, see stubs/mini_eigen.hpp).  The vectors travel; /root/reference does not.

Inputs are BASELINE config 1 (4096-point Scan-M vs 100 k-raw-point Map-U straddling the origin, 10 forced iterations) and are
regenerated from the seeds of elimaloc_b200/synth.py by `world()`, which the tests import as well.

What each array is:
  pose, is_success, fitness, local_cov, n_iter   outputs of Registration::RunRegister (10 forced iterations)
  A, b                                           the system of every iteration as handed to ldlt().solve():
                                                 A = JTJ + lm_lambda * diag(JTJ), b = JTr  (recorded by the stand-in's LDLT)
  JTJ, JTr                                       JTJ = A with its diagonal divided by (1 + lm_lambda); JTr = b
  pose_out[k]                                    RunRegister's result with max_iteration = k + 1  (= pose after iteration k)
  ncorr[k]                                       pairs emitted by the method's search at the pose going into iteration k
  res[k]                                         d_fitness_score_ after iteration k times ncorr[k]  (the residual sum)
  corr_count, corr_target                        the search of the first 256 scan points at the initial pose
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from elimaloc_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LM_LAMBDA = synth.timing_knobs()["lm_lambda"]


def inputs():
    """raw map points, true pose — everything else derives from the STORED points of the built map"""
    return synth.map_u(100_000, 21.5, origin=-6.0), synth.se3([4.0, 5.0, 3.5], [0.02, -0.01, 0.3])


def world(impl=O):
    """(map, scan, initial guess) on the given implementation (oracle by default: what the tests run)."""
    raw, T_true = inputs()
    m = impl.VoxelHashMap(1.0, 30)
    m.AddPoints(raw)
    m.CalVoxelCovAll()
    m.CalPointCovAll(0.4)
    stored = m.export()["pxyz"]
    scan = synth.scan_m(stored, 4096, T_true)
    T0 = T_true @ synth.canonical_offset()
    return m, scan, T0


def main():
    from oracle import reference_build as R
    if not R.sources_present():
        raise SystemExit("the golden vectors are generated from /root/reference; it is not here")
    R.build(force=True)
    rm, scan, T0 = world(R)
    e = rm.export()
    np.savez_compressed(os.path.join(HERE, "config1_map_digest.npz"), n_voxels=rm.num_voxels(), n_points=rm.num_points(),
                        keys_head=e["keys"][:64], counts_head=e["counts"][:64], pxyz_head=e["pxyz"][:64],
                        vcov_head=e["vcov"][:16], pcov_head=e["pcov"][:16], pmean_head=e["pmean"][:16],
                        keys_sum=e["keys"].astype(np.int64).sum(axis=0), pxyz_sum=e["pxyz"].astype(np.float64).sum(axis=0))
    for m, name in enumerate(["p2p", "gicp", "vgicp", "avgicp"]):
        knobs = synth.timing_knobs()
        r = R.Registration().RunRegister(scan, rm, T0, O.make_config(icp_method=m, max_iteration=10, **knobs))
        A, b = r["trace"]["A"], r["trace"]["b"]
        JTJ = A.copy()
        for i in range(6):
            JTJ[:, i, i] = A[:, i, i] / (1.0 + LM_LAMBDA)
        pose_in, pose_out, ncorr, res = T0, [], [], []
        for k in range(r["n_iter"]):
            idx, _ = R.search_pairs(rm, scan, pose_in, m, knobs["max_search_dist"])
            rk = R.Registration().RunRegister(scan, rm, T0, O.make_config(icp_method=m, max_iteration=k + 1, **knobs))
            ncorr.append(float(len(idx)))
            res.append(rk["d_fitness_score"] * len(idx))
            pose_out.append(rk["pose"])
            pose_in = rk["pose"]
        assert np.array_equal(pose_out[-1], r["pose"])
        cnt, tgt = R.correspondences(rm, scan[:256], T0, m, 5.0)
        np.savez_compressed(os.path.join(HERE, f"config1_{name}.npz"), pose=r["pose"], fitness=r["fitness_score"],
                            local_cov=r["local_cov"], is_success=r["is_success"], n_iter=r["n_iter"], A=A, b=b, JTJ=JTJ, JTr=b,
                            res=np.array(res), ncorr=np.array(ncorr), pose_out=np.array(pose_out), corr_count=cnt, corr_target=tgt,
                            generated_by="oracle/_ref/libref.so (reference sources + stand-in Eigen/oneTBB headers)")
        print(name, r["n_iter"], r["is_success"], r["fitness_score"], ncorr)


if __name__ == "__main__":
    main()
