#!/bin/bash
# Round 2, GPU call D: device-resident scan chain (tests + config-5 pipeline bench) and the whole GPU suite after the EKF / pre-processing refactor
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_pipeline.py -m gpu -x -q > $O/d_pytest_chain.log 2>&1; echo "rc=$?" >> $O/d_pytest_chain.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest.log
timeout 600 python bench.py --config 5 --pipeline --steps 50 --warmup 5 > $O/d_bench_pipeline.json 2> $O/d_bench_pipeline.err
tail -15 $O/d_pytest_chain.log; tail -3 $O/d_pytest.log; tail -5 $O/d_bench_pipeline.err
