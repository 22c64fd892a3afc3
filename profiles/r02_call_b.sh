#!/bin/bash
# Round 2, GPU call B: full GPU suite after the VGICP (8-byte candidates) / AVGICP (one kernel, balanced pairs) rewrite + their bench lines
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q --durations=10 > $O/b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/b_pytest.log
for m in vgicp avgicp; do timeout 300 python bench.py --method $m --no-cpu-baseline > $O/b_bench_$m.json 2> $O/b_bench_$m.err; done
timeout 600 python bench.py --config 4 --no-cpu-baseline --steps 30 > $O/b_bench_config4_1gpu.json 2> $O/b_bench_config4_1gpu.err
timeout 500 ncu --set full --clock-control none -k regex:icp_ --launch-skip 26 --launch-count 3 -f -o $O/b_ncu_vgicp python profiles/ncu_driver.py --method vgicp > $O/b_ncu_vgicp.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:icp_ --launch-skip 14 --launch-count 2 -f -o $O/b_ncu_avgicp python profiles/ncu_driver.py --method avgicp > $O/b_ncu_avgicp.log 2>&1
for m in vgicp avgicp; do python profiles/ncu_summary.py --json $O/b_traffic_$m.json $O/b_ncu_$m.ncu-rep > $O/b_ncu_$m.txt 2>&1; done
rm -f $O/b_ncu_vgicp.ncu-rep $O/b_ncu_avgicp.ncu-rep
tail -5 $O/b_pytest.log
