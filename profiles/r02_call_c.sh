#!/bin/bash
# Round 2, GPU call C: A/B of the one-kernel warm iteration (default) against the reuse + refresh pair (ELM_WARM_PAIR=1)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py -m gpu -x -q > $O/c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c_pytest.log
for m in p2p gicp; do
  timeout 300 python bench.py --method $m --no-cpu-baseline > $O/c_bench_${m}_single.json 2> $O/c_bench_${m}_single.err
  ELM_WARM_PAIR=1 timeout 300 python bench.py --method $m --no-cpu-baseline > $O/c_bench_${m}_pair.json 2> $O/c_bench_${m}_pair.err
done
tail -3 $O/c_pytest.log
