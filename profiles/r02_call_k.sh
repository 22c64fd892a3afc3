#!/bin/bash
# Round 2, GPU call K: warm tests again (voxel-border reuse), the three warm modes for P2P and GICP, 2-GPU is a separate call
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py tests/test_pipeline.py tests/test_shim.py -m gpu -x -q > $O/k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest.log
ELM_WARM_MODE=async timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > $O/k_pytest_async_all_methods.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest_async_all_methods.log
ELM_WARM_MODE=pair timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py -m gpu -x -q > $O/k_pytest_pair_all_methods.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest_pair_all_methods.log
for m in p2p gicp; do for mode in async pair single; do
  ELM_WARM_MODE=$mode timeout 300 python bench.py --method $m --no-cpu-baseline > $O/k_bench_${m}_$mode.json 2> $O/k_bench_${m}_$mode.err
done; done
tail -4 $O/k_pytest.log; tail -3 $O/k_pytest_async_all_methods.log; tail -3 $O/k_pytest_pair_all_methods.log
