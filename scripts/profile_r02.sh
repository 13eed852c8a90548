#!/bin/bash
# ncu evidence of round 2 (one GPU):  gpurun --timeout 1500 -- 'bash scripts/profile_r02.sh > gpurun_out/profile_r02.log 2>&1'
set -x
# 1. launch list of two timed steps of the default bench (cold-cache, serialised: compare shares)
timeout 300 python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02_launches_c5_n1.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_l.log 2>&1
# 2. the fused scan, one launch, full set
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ivfpq_scan_kernel -c 1 \
    -o gpurun_out/r02_prof_scan_c5 -f python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_s.log 2>&1
# 3. the probe stage kernels (tc_score x2, threshold, rescore)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_score_kernel|rescore_probe" -c 3 \
    -o gpurun_out/r02_prof_probe_c5 -f python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_p.log 2>&1
# 4. the tensor-core PQ encoder at C2
timeout 200 python scripts/c2_time.py > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pq_tc_encode_kernel -s 3 -c 1 \
    -o gpurun_out/r02_prof_pqtc_c2 -f python scripts/c2_time.py > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep
