#!/bin/bash
# Round 2, GPU call Q (about 9 GPU-minutes were left): A/B of the chained warm iterations (ELM_WARM_MODE=chain: flags in HBM instead of
# two grid-completion waits per iteration), of the size of the concurrent refresh grid (ELM_ASYNC_GRID) and of the keep-best variant of
# the reuse kernel; then the whole GPU suite and the bench lines under the winner.
O=gpurun_out; mkdir -p $O
timeout 120 python profiles/ab_chain.py --methods p2p,gicp --sizes 131072,16384 --grids 128,96,80,64,48 --steps 30 > $O/q_ab_default.txt 2> $O/q_ab_default.err
ELIMALOC_B200_LIB=elimaloc_b200/lib_keepbest.so timeout 60 python profiles/ab_chain.py --methods p2p --sizes 131072 --grids 128,80,64 --steps 30 > $O/q_ab_keepbest.txt 2> $O/q_ab_keepbest.err
python profiles/ab_pick.py $O/q_ab_default.txt $O/q_ab_keepbest.txt > $O/q_winner.sh 2> $O/q_winner.err
cat $O/q_winner.sh
source $O/q_winner.sh
timeout 220 python -m pytest tests -m gpu -x -q > $O/q_pytest_winner.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/q_pytest_winner.log
tail -3 $O/q_pytest_winner.log
if [ $rc -ne 0 ] && [ "$ELM_WARM_MODE" = "chain" ]; then
  # the chained mode failed a test: is the grid size alone sound?
  ELM_WARM_MODE=async timeout 100 python -m pytest tests/test_gpu_warm.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q > $O/q_pytest_async_grid.log 2>&1; echo "pytest rc=$?" >> $O/q_pytest_async_grid.log
  tail -3 $O/q_pytest_async_grid.log
fi
timeout 60 python bench.py --no-cpu-baseline > $O/q_bench_p2p_winner.json 2> $O/q_bench_p2p_winner.err
timeout 60 python bench.py --method gicp --no-cpu-baseline > $O/q_bench_gicp_winner.json 2> $O/q_bench_gicp_winner.err
grep -h "it/s" $O/q_ab_default.txt $O/q_ab_keepbest.txt | head -70
python - <<'PY'
import json
for f in ("gpurun_out/q_bench_p2p_winner.json", "gpurun_out/q_bench_gicp_winner.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
