for V in old b200 old b200; do
VIX_LIB_PATH=$PWD/vectorindex_b200/libvindex_$V.so python bench.py --workload c5 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V c5 ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), d['clocks'])"
done
