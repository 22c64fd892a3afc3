#!/bin/bash
# Round 2, GPU call E: GPU map build — bit-identity against the host builder, full-size timing, then the whole suite (every map is now built on the GPU)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_map_build.py -m gpu -x -q -s > $O/e_pytest_map_build.log 2>&1; echo "rc=$?" >> $O/e_pytest_map_build.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest.log
timeout 600 python bench.py --config 4 --no-cpu-baseline --steps 20 > $O/e_bench_config4_1gpu.json 2> $O/e_bench_config4_1gpu.err
tail -30 $O/e_pytest_map_build.log; tail -5 $O/e_pytest.log
