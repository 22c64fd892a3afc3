#!/bin/bash
# Round 2, GPU call M (2 GPUs): the concurrent refresh with the peer exchange: parity tests at 2 ranks + the bench's parity check
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "2-peer or 2-nccl-0" > $O/m_pytest_multi_gpu.log 2>&1; echo "rc=$?" >> $O/m_pytest_multi_gpu.log
timeout 400 $TR --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 3 > $O/m_bench_p2p_n2.json 2> $O/m_bench_p2p_n2.err
timeout 400 $TR --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 3 --method gicp > $O/m_bench_gicp_n2.json 2> $O/m_bench_gicp_n2.err
tail -3 $O/m_pytest_multi_gpu.log
