#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 1200 -- 'bash scripts/first_call_next_round.sh > gpurun_out/first_call.log 2>&1'
# 1. the GPU suite with the opt-in tests switched on (VIX_TEST_EXPERIMENTAL=1: two-pipeline scan == one-pipeline scan bit
#    for bit, cosine CentroidBatchScore / IVF-Flat, 64-row encode CTAs, long rows, insert -> optimize(), trainer edge cases);
# 2. the two-pipeline scan (VIX_SCAN_DUAL) timed on one rank's share of an 8-way sharded C5 and on the single-GPU bench.
# Multi-GPU follow-up (one call, N = 8):
#   gpurun --gpus 8 --timeout 900 -- 'for P in lists replicate; do python -m torch.distributed.run --nnodes=1 \
#     --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 \
#     --partition $P > gpurun_out/bench_c5_n8_$P.json 2> gpurun_out/bench_c5_n8_$P.err; done'
set -x -o pipefail
# the whole GPU suite, opt-in tests included; the two-pipeline kernel (named barriers: a mistake there is a hang, not a
# wrong answer) runs on its own afterwards under a short timeout
VIX_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -rs -k "not two_pipeline" 2>&1 | tail -25
VIX_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k two_pipeline 2>&1 | tail -8 \
    || { echo "two-pipeline test failed or timed out: skipping its timing runs"; exit 0; }
timeout 400 python scripts/shard_emul.py 8 0 2>&1 | tail -4
for D in 0 1; do
  VIX_SCAN_DUAL=$D timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('VIX_SCAN_DUAL=$D', 'ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'recall', d['config']['recall_at_10'])"
done
