set -x
gcc -std=c11 -O1 -I include tests/c/sharded_smoke.c -o /tmp/sharded_smoke -L vectorindex_b200 -lvindex_b200 -Wl,-rpath,$PWD/vectorindex_b200 -lm && timeout 300 /tmp/sharded_smoke 8 > gpurun_out/smoke8.log 2>&1
for P in lists replicate; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --partition $P > gpurun_out/b8_$P.json 2> gpurun_out/b8_$P.err
done
