#!/bin/bash
# last state of round 2 on one GPU: the whole GPU suite, default bench + reference arm, smoke, launch list
cd /root/repo
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/fin2_gpu_suite.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/fin2_c5.json 2> gpurun_out/fin2_c5.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin2_c5_ref.json 2> gpurun_out/fin2_c5_ref.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/fin2_smoke.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/fin2_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/fin2_ncu_list.log 2>&1
cat gpurun_out/fin2_gpu_suite.log; tail -2 gpurun_out/fin2_smoke.log
