#!/bin/bash
# runs of the list-major tensor-core scan: parity tests (three times: races show up as run-to-run differences), then the bench A/B
cd /root/repo
for i in 1 2 3; do
VIX_TC_SCAN_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "list_major" 2>&1 | tail -8 > gpurun_out/tcs_test$i.log
done
VIX_TC_SCAN=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/tcs_bench_old.json 2> gpurun_out/tcs_bench_old.err
VIX_TC_SCAN_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/tcs_bench_new.json 2> gpurun_out/tcs_bench_new.err
