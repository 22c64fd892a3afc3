#!/bin/bash
# Round 2, GPU call A: bench lines of all four methods + ncu evidence for every kernel of every method (run under gpurun, 1 GPU).
O=gpurun_out; mkdir -p $O
{ nvidia-smi -L; nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; } > $O/a_host.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_smoke.txt 2>&1
timeout 400 python bench.py > $O/a_bench_p2p.json 2> $O/a_bench_p2p.err
for m in gicp vgicp avgicp; do timeout 300 python bench.py --method $m --no-cpu-baseline > $O/a_bench_$m.json 2> $O/a_bench_$m.err; done
timeout 300 python bench.py --no-warm --no-cpu-baseline > $O/a_bench_p2p_nowarm.json 2> $O/a_bench_p2p_nowarm.err
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/a_launches_p2p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/a_launches_bench.log 2>&1
python profiles/launches_summary.py $O/a_launches_p2p.csv > $O/a_launches_p2p.txt 2>&1
# --set full: one step's kernels of every method (skip begin + two whole steps: 2 * (1 + 2 * 6) launches; AVGICP 2 * (1 + 6))
timeout 500 ncu --set full --clock-control none --import-source on -k regex:icp_ --launch-skip 26 --launch-count 9 -f -o $O/a_ncu_p2p python profiles/ncu_driver.py --method p2p > $O/a_ncu_p2p.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:icp_ --launch-skip 26 --launch-count 7 -f -o $O/a_ncu_gicp python profiles/ncu_driver.py --method gicp > $O/a_ncu_gicp.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:icp_ --launch-skip 26 --launch-count 3 -f -o $O/a_ncu_vgicp python profiles/ncu_driver.py --method vgicp > $O/a_ncu_vgicp.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:icp_ --launch-skip 14 --launch-count 2 -f -o $O/a_ncu_avgicp python profiles/ncu_driver.py --method avgicp > $O/a_ncu_avgicp.log 2>&1
for m in p2p gicp vgicp avgicp; do python profiles/ncu_summary.py --json $O/a_traffic_$m.json $O/a_ncu_$m.ncu-rep > $O/a_ncu_$m.txt 2>&1; done
rm -f $O/a_ncu_gicp.ncu-rep $O/a_ncu_vgicp.ncu-rep $O/a_ncu_avgicp.ncu-rep
ls -la $O
