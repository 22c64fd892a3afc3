#!/bin/bash
# Round 2, GPU call L: first warm iteration without the reuse kernel (all-queries refresh), async default for both methods
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/l_pytest.log
for m in p2p gicp; do timeout 300 python bench.py --method $m --no-cpu-baseline > $O/l_bench_$m.json 2> $O/l_bench_$m.err; done
tail -4 $O/l_pytest.log
