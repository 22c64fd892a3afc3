#!/bin/bash
# usage (under gpurun): bash profiles/sanitize.sh > gpurun_out/sanitize.log 2>&1
# compute-sanitizer memcheck + racecheck + synccheck over a small slice of the GPU parity tests (the reference has no
# sanitizer story at all; the kernels use shared-memory work lists, cp.async, mbarriers and warp-scoped synchronisation).
# ~40x slow-down: keep the slice small.
set -u
SLICE='tests/test_gpu_edges.py::test_quirks_on_the_gpu tests/test_gpu_edges.py::test_exact_ties_and_near_ties tests/test_gpu_edges.py::test_scan_preprocess_matches_oracle tests/test_gpu_edges.py::test_peer_exchange_world_size_one_equals_fused_path'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python -m pytest $SLICE -x -q -m gpu 2>&1 | tail -15
  echo "== exit code $?"
done
