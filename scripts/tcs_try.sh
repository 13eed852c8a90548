#!/bin/bash
cd /root/repo
VIX_TC_SCAN_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "list_major" 2>&1 | tail -12 > gpurun_out/tcs_test1.log
if grep -q "7 passed" gpurun_out/tcs_test1.log; then
timeout 900 python scripts/tcs_fullsize_check.py c5 > gpurun_out/tcs_full_c5.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tcs_bench_new.json 2> gpurun_out/tcs_bench_new.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/tcs_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/tcs_ncu.log 2>&1

fi
