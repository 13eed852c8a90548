#!/bin/bash
# round-2 artefacts of the list-major scan on one GPU: tests, default bench + reference arm, launch list, ncu --set full
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/lm_gpu_suite.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/lm_c5.json 2> gpurun_out/lm_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/lm_c5_ref.json 2> gpurun_out/lm_c5_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/lm_smoke.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/lm_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/lm_ncu_list.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_scan_kernel -c 1 \
  -o gpurun_out/lm_scan_c5 -f python bench.py --steps 2 --warmup 3 --profile > gpurun_out/lm_ncu_full.log 2>&1
