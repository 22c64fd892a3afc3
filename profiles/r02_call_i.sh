#!/bin/bash
# Round 2, GPU call I: concurrent refresh v2 (chunks of tiles, completion counter): tests, A/B, wall-clock trace
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py -m gpu -x -q > $O/i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/i_pytest.log
ELM_WARM_MODE=async timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py -m gpu -x -q > $O/i_pytest_async_all_methods.log 2>&1; echo "pytest rc=$?" >> $O/i_pytest_async_all_methods.log
for mode in async pair; do
  ELM_WARM_MODE=$mode timeout 300 python bench.py --method p2p --no-cpu-baseline > $O/i_bench_p2p_$mode.json 2> $O/i_bench_p2p_$mode.err
done
ELM_WARM_MODE=async timeout 300 python bench.py --method gicp --no-cpu-baseline > $O/i_bench_gicp_async.json 2> $O/i_bench_gicp_async.err
ELIMALOC_B200_LIB=elimaloc_b200/lib_trace.so timeout 300 python profiles/trace_async.py > $O/i_trace.txt 2>&1
tail -3 $O/i_pytest.log; tail -3 $O/i_pytest_async_all_methods.log; cat $O/i_trace.txt
