#!/bin/bash
# First GPU call of a round, in one box:   (build the variant first, on the CPU:  profiles/build_variant.sh greedy "-DELM_GREEDY_ITEMS")
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash profiles/first_gpu_call.sh'
# 1. the whole GPU test suite (incl. the tests written after round 1's GPU budget was spent)
# 2. the headline bench line
# 3. A/B of the greedy item scheduling (only if elimaloc_b200/lib_greedy.so was built) + its parity slice
# 4. compute-sanitizer slice
# 5. BASELINE config 5 at full size (131 072-point scans, 10 M-point map, 100 scans): GPU arm vs oracle arm
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 400 python bench.py > gpurun_out/bench_p2p.json 2> gpurun_out/bench_p2p.err ; tail -c 1500 gpurun_out/bench_p2p.json
if [ -f elimaloc_b200/lib_greedy.so ]; then
  echo "== A/B greedy" ; bash profiles/quick_bench.sh "" "@elimaloc_b200/lib_greedy.so " "" "@elimaloc_b200/lib_greedy.so " 2>&1 | tee gpurun_out/ab_greedy.txt
  echo "== parity of the greedy variant" ; ELIMALOC_B200_LIB=elimaloc_b200/lib_greedy.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_greedy.txt
fi
echo "== sanitizer" ; timeout 600 bash profiles/sanitize.sh 2>&1 | tail -15 | tee gpurun_out/sanitize.txt
echo "== config 5 at full size" ; timeout 1200 python tests/full_size_pipeline.py 100 gpu,oracle 2>&1 | grep -v "ICP Fitness" | tail -6 | tee gpurun_out/full_size_pipeline.txt
