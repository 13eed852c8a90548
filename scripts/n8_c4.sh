set -x
for P in lists replicate; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --workload c4 --partition $P > gpurun_out/b8_c4_$P.json 2> gpurun_out/b8_c4_$P.err
done
