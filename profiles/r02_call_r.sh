#!/bin/bash
# Round 2, GPU call R (the last ~5.9 GPU-minutes): the reduce-scatter block reduction (bit-identical sums, 31 instead of 5 x NACC shuffles
# per warp), 80 refresh blocks by default, the done flag + chunk counter requested together, chained mode with inputs requested ahead of
# the solve (ELM_WARM_MODE=chain).  A/B inside one process, then the whole GPU suite, the headline bench line and smoke on the final build.
O=gpurun_out; mkdir -p $O
timeout 90 python profiles/ab_chain.py --methods p2p,gicp,vgicp,avgicp --sizes 131072,16384 --grids 80,128 --steps 30 > $O/r_ab_default.txt 2> $O/r_ab_default.err
ELIMALOC_B200_LIB=elimaloc_b200/lib_butterfly.so timeout 60 python profiles/ab_chain.py --methods p2p,gicp,vgicp,avgicp --sizes 131072 --grids 80 --modes async --steps 30 > $O/r_ab_butterfly.txt 2> $O/r_ab_butterfly.err
grep -h "it/s\|FAILED" $O/r_ab_default.txt $O/r_ab_butterfly.txt
timeout 200 python -m pytest tests -m gpu -x -q > $O/r_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r_pytest.log
tail -3 $O/r_pytest.log
timeout 90 python bench.py > $O/r_bench_p2p.json 2> $O/r_bench_p2p.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r_smoke.txt 2>&1
tail -1 $O/r_smoke.txt
timeout 40 python bench.py --method gicp --no-cpu-baseline > $O/r_bench_gicp.json 2> $O/r_bench_gicp.err
timeout 40 python bench.py --method vgicp --no-cpu-baseline > $O/r_bench_vgicp.json 2> $O/r_bench_vgicp.err
timeout 40 python bench.py --method avgicp --no-cpu-baseline > $O/r_bench_avgicp.json 2> $O/r_bench_avgicp.err
python - <<'PY'
import json
for m in ("p2p", "gicp", "vgicp", "avgicp"):
    f = f"gpurun_out/r_bench_{m}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(m, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "unreadable", e)
PY
