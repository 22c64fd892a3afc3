#!/bin/bash
# Round 2, final GPU call: whole GPU suite, smoke, the bench lines of every method / config and the ncu evidence, all on the final build
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $O/o_pytest.log 2>&1; echo "pytest rc=$?" >> $O/o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/o_smoke.txt 2>&1
timeout 400 python bench.py > $O/o_bench_p2p.json 2> $O/o_bench_p2p.err
for m in gicp vgicp avgicp; do timeout 300 python bench.py --method $m --no-cpu-baseline > $O/o_bench_$m.json 2> $O/o_bench_$m.err; done
timeout 600 python bench.py --config 4 --no-cpu-baseline --steps 30 > $O/o_bench_config4_1gpu.json 2> $O/o_bench_config4_1gpu.err
timeout 400 python bench.py --config 5 --pipeline --steps 50 --warmup 5 > $O/o_bench_pipeline.json 2> $O/o_bench_pipeline.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/o_launches_p2p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/o_launches_bench.log 2>&1
python profiles/launches_summary.py $O/o_launches_p2p.csv > $O/o_launches_p2p.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:icp_ --launch-skip 24 --launch-count 8 -f -o $O/o_ncu_p2p python profiles/ncu_driver.py --method p2p > $O/o_ncu_p2p.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:icp_ --launch-skip 24 --launch-count 8 -f -o $O/o_ncu_gicp python profiles/ncu_driver.py --method gicp > $O/o_ncu_gicp.log 2>&1
for m in p2p gicp; do python profiles/ncu_summary.py --json $O/o_traffic_$m.json $O/o_ncu_$m.ncu-rep > $O/o_ncu_$m.txt 2>&1; done
rm -f $O/o_ncu_gicp.ncu-rep
tail -8 $O/o_pytest.log; cat $O/o_smoke.txt | tail -2
