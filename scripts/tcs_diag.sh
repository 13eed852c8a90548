#!/bin/bash
# where the roles of tc_scan_kernel wait (a -DVIX_TCS_DIAG build of vix_ivfpq_tc.cu, linked as libvindex_diag.so)
cd /root/repo
VIX_TC_SCAN_DEBUG=1 VIX_LIB_PATH=$PWD/vectorindex_b200/libvindex_diag.so timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/tcs_diag.json 2> gpurun_out/tcs_diag.err
grep "vix tc" gpurun_out/tcs_diag.err | tail -8 > gpurun_out/tcs_diag.txt
