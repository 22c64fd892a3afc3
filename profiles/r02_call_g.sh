#!/bin/bash
# Round 2, GPU call G (8 GPUs): BASELINE config 4 as specified (VGICP, 262144 x 50M, strong-sharded over 8 B200), config 2 at 8 ranks with
# the bench's parity check, 8-rank parity tests
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29621 bench.py --gpus 8 --config 4 --steps 50 --warmup 3 > $O/g_bench_config4_n8.json 2> $O/g_bench_config4_n8.err
timeout 400 $TR --master-port 29622 bench.py --gpus 8 --steps 50 --warmup 3 > $O/g_bench_p2p_n8.json 2> $O/g_bench_p2p_n8.err
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "8-peer" > $O/g_pytest_multi_gpu.log 2>&1; echo "rc=$?" >> $O/g_pytest_multi_gpu.log
tail -3 $O/g_pytest_multi_gpu.log
