#!/bin/bash
# final artefacts of round 2 on one GPU: default bench + reference arm, smoke, launch list, ncu --set full of tc_scan_kernel,
# then the whole GPU suite
cd /root/repo
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/fin_c5.json 2> gpurun_out/fin_c5.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_c5_ref.json 2> gpurun_out/fin_c5_ref.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/fin_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/fin_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_scan_kernel -c 1 \
  -o gpurun_out/fin_scan_c5 -f python bench.py --steps 2 --warmup 3 --profile > gpurun_out/fin_ncu_full.log 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/fin_gpu_suite.log
cat gpurun_out/fin_gpu_suite.log
