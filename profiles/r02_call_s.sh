#!/bin/bash
# Round 2, GPU call S (the last 3.1 GPU-minutes): P2P with the reduce-scatter reduction and the table-driven slot scatter (default build)
# against P2P on the butterfly (lib_p2pbutterfly.so = -DELM_P2P_BUTTERFLY; the 29-accumulator kernels use the reduce-scatter in both), then
# the whole GPU suite and the headline bench line on the faster of the two.
O=gpurun_out; mkdir -p $O
timeout 40 python profiles/ab_chain.py --methods p2p --sizes 131072,16384 --grids 80 --modes async --steps 30 > $O/s_ab_default.txt 2> $O/s_ab_default.err
ELIMALOC_B200_LIB=elimaloc_b200/lib_p2pbutterfly.so timeout 40 python profiles/ab_chain.py --methods p2p --sizes 131072,16384 --grids 80 --modes async --steps 30 > $O/s_ab_p2pbutterfly.txt 2> $O/s_ab_p2pbutterfly.err
grep -h "it/s\|FAILED" $O/s_ab_default.txt $O/s_ab_p2pbutterfly.txt
python - > $O/s_winner.sh <<'PY'
import json
def rows(p):
    try:
        return [r for l in open(p) if l.startswith("[{") for r in json.loads(l)]
    except OSError:
        return []
a = [r for r in rows("gpurun_out/s_ab_default.txt") if r.get("n") == 131072 and "failed" not in r]
b = [r for r in rows("gpurun_out/s_ab_p2pbutterfly.txt") if r.get("n") == 131072 and "failed" not in r]
if a and b and a[0]["pose_sha"] == b[0]["pose_sha"] and a[0]["it_per_s"] >= 0.99 * b[0]["it_per_s"]:
    print(f"# winner: the default build ({a[0]['it_per_s']:.0f} vs {b[0]['it_per_s']:.0f} it/s, same pose bits)")
elif b:
    print(f"# winner: lib_p2pbutterfly.so ({b[0]['it_per_s']:.0f} vs {a[0]['it_per_s'] if a else float('nan'):.0f} it/s; pose bits equal: {bool(a) and a[0]['pose_sha'] == b[0]['pose_sha']})")
    print("export ELIMALOC_B200_LIB=elimaloc_b200/lib_p2pbutterfly.so")
else:
    print("# no usable A/B line: the default build")
PY
cat $O/s_winner.sh
source $O/s_winner.sh
timeout 150 python -m pytest tests -m gpu -x -q > $O/s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s_pytest.log
tail -3 $O/s_pytest.log
timeout 60 python bench.py > $O/s_bench_p2p.json 2> $O/s_bench_p2p.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s_bench_p2p.json").read().strip().splitlines()[-1])
    print("p2p", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
except Exception as e:
    print("bench unreadable", e)
PY
