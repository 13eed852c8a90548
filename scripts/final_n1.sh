set -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_c5.json 2> gpurun_out/f_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_c5_ref.json 2> gpurun_out/f_c5_ref.err
for w in c1 c2 c3 c4; do timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/f_$w.json 2> gpurun_out/f_$w.err; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
