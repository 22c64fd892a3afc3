#!/bin/bash
# Round 2, GPU call F (2 GPUs): multi-GPU parity tests at 2 ranks, the bench's own parity check, config 4 strong at 2 ranks
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "2-" > $O/f_pytest_multi_gpu.log 2>&1; echo "rc=$?" >> $O/f_pytest_multi_gpu.log
timeout 400 $TR --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 3 > $O/f_bench_p2p_n2.json 2> $O/f_bench_p2p_n2.err
timeout 400 $TR --master-port 29612 bench.py --gpus 2 --steps 50 --warmup 3 --scaling strong > $O/f_bench_p2p_n2_strong.json 2> $O/f_bench_p2p_n2_strong.err
timeout 600 $TR --master-port 29613 bench.py --gpus 2 --config 4 --steps 20 --warmup 3 > $O/f_bench_config4_n2.json 2> $O/f_bench_config4_n2.err
tail -3 $O/f_pytest_multi_gpu.log; tail -3 $O/f_bench_*.err
