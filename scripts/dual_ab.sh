#!/bin/bash
# First GPU call for the two-pipeline scan (ivfpq_scan_kernel<G, FILTER, false, 2>, VIX_SCAN_DUAL): correctness against the
# one-pipeline kernel, then time on one GPU (C5) and on one rank's share of an 8-way sharded C5.
#   gpurun --timeout 900 -- 'bash scripts/dual_ab.sh > gpurun_out/dual_ab.log 2>&1'
set -x
VIX_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k two_pipeline 2>&1 | tail -5
timeout 400 python scripts/shard_emul.py 8 0 2>&1 | tail -4
for D in 0 1; do
  VIX_SCAN_DUAL=$D timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('VIX_SCAN_DUAL=$D', 'ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'recall', d['config']['recall_at_10'])"
done
