python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for V in b200 nopf b200 nopf; do
VIX_LIB_PATH=$PWD/vectorindex_b200/libvindex_$V.so python bench.py --workload c5s --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V c5s ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'recall', d['config']['recall_at_10'], 'e2e', round(d['e2e']['ms_per_step'],2))"
done
python bench.py --workload c3 --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3 ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'recall', d['config']['recall_at_10'], 'e2e', round(d['e2e']['ms_per_step'],2))"
