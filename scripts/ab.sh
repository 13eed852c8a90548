timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for W in c5s c3; do python scripts/phases.py $W 2>&1 | tail -1 | cut -c1-60; done
python bench.py --workload c5 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'recall', d['config']['recall_at_10'], 'e2e', round(d['e2e']['ms_per_step'],2), d['config']['build'])"
