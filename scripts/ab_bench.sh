#!/bin/bash
# A/B of library builds on the single-GPU bench: bash scripts/ab_bench.sh "<variant> ..." [bench args]
# (variant v => vectorindex_b200/libvindex_v.so; "b200" = the current build)
V="$1"; shift
for v in $V; do
  VIX_LIB_PATH=$PWD/vectorindex_b200/libvindex_$v.so timeout 600 python bench.py --steps 10 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', 'ms/step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['detail']['stage_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],3), 'recall', round(d['detail']['recall_at_10'],4), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w_max'])"
done
