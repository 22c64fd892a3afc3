#!/bin/bash
# Round 2, GPU call H: the refresh kernel running beside the reuse kernel (ELM_WARM_MODE=async, default for P2P): tests + A/B
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py tests/test_shim.py tests/test_pipeline.py -m gpu -x -q > $O/h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/h_pytest.log
ELM_WARM_MODE=async timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > $O/h_pytest_async_all_methods.log 2>&1; echo "pytest rc=$?" >> $O/h_pytest_async_all_methods.log
for mode in async pair; do
  ELM_WARM_MODE=$mode timeout 300 python bench.py --method p2p --no-cpu-baseline > $O/h_bench_p2p_$mode.json 2> $O/h_bench_p2p_$mode.err
done
for mode in async single; do
  ELM_WARM_MODE=$mode timeout 300 python bench.py --method gicp --no-cpu-baseline > $O/h_bench_gicp_$mode.json 2> $O/h_bench_gicp_$mode.err
done
tail -3 $O/h_pytest.log; tail -3 $O/h_pytest_async_all_methods.log
