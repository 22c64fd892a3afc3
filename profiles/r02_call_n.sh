#!/bin/bash
# Round 2, GPU call N (8 GPUs): the concurrent refresh at 8 ranks: config 2 weak + strong figure + parity, two 8-rank parity tests
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29822 bench.py --gpus 8 --steps 50 --warmup 3 > $O/n_bench_p2p_n8.json 2> $O/n_bench_p2p_n8.err
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "8-peer-0 or 8-peer-1" > $O/n_pytest_multi_gpu.log 2>&1; echo "rc=$?" >> $O/n_pytest_multi_gpu.log
tail -3 $O/n_pytest_multi_gpu.log
