#!/bin/bash
# usage: profiles/quick_bench.sh "<bench args>" ...   — one compact line per argument set (dev helper, run under gpurun)
# an argument set may start with "@<path to another build of the library> " to A/B-test kernel variants
for m in "$@"; do
  lib=""; args="$m"
  if [[ "$m" == @* ]]; then lib="${m%% *}"; lib="${lib#@}"; args="${m#* }"; [[ "$m" != *" "* ]] && args=""; fi
  ELIMALOC_B200_LIB="$lib" timeout 300 python bench.py --no-cpu-baseline $args > /tmp/qb.json 2> /tmp/qb.err || { echo "FAILED[$m]"; tail -5 /tmp/qb.err; continue; }
  python - "$m" <<'PY'
import json, sys
d = json.load(open('/tmp/qb.json')); r = d["roofline"] or {}
print(f"[{sys.argv[1]:44s}] it/s {d['value']:9.0f}  e2e {d['e2e']['value']:9.0f}  search_us {1e3*r.get('kernel_ms_avg',0):7.1f}  acc_us {1e3*r.get('accumulate_kernel_ms_avg',0):6.1f}"
      f"  visited {r.get('map_points_visited_per_search')}  frac {r.get('frac',0):.3f}  exh_GBs {r.get('exhaustive_equivalent_gbs',0):.0f}  launches {d['gpu_launches']}")
PY
done
