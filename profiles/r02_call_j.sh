#!/bin/bash
# Round 2, GPU call J: candidate lists survive voxel changes among positive keys: tests, stragglers per step, A/B of the three warm modes
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py -m gpu -x -q > $O/j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j_pytest.log
ELM_WARM_MODE=async timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py -m gpu -x -q > $O/j_pytest_async_all_methods.log 2>&1; echo "pytest rc=$?" >> $O/j_pytest_async_all_methods.log
for mode in async pair single; do
  ELM_WARM_MODE=$mode timeout 300 python bench.py --method p2p --no-cpu-baseline > $O/j_bench_p2p_$mode.json 2> $O/j_bench_p2p_$mode.err
done
for mode in pair single; do
  ELM_WARM_MODE=$mode timeout 300 python bench.py --method gicp --no-cpu-baseline > $O/j_bench_gicp_$mode.json 2> $O/j_bench_gicp_$mode.err
done
tail -15 $O/j_pytest.log; tail -3 $O/j_pytest_async_all_methods.log
