#!/bin/bash
# Round 2, last GPU call: the headline bench lines again after the roofline-kernel selection in bench.py changed (Python only)
O=gpurun_out; mkdir -p $O
timeout 400 python bench.py > $O/p_bench_p2p.json 2> $O/p_bench_p2p.err
timeout 300 python bench.py --method gicp --no-cpu-baseline > $O/p_bench_gicp.json 2> $O/p_bench_gicp.err
tail -c 600 $O/p_bench_p2p.json
