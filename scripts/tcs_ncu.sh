#!/bin/bash
# ncu --set full of ONE launch of the list-major scan kernel inside a bench step (after the same command exited 0 without ncu)
cd /root/repo
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_scan_kernel -c 1 \
  -o gpurun_out/tcs_scan_c5 -f python bench.py --steps 2 --warmup 3 --profile > gpurun_out/tcs_ncu_full.log 2>&1
