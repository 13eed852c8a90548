for W in c5s c3; do
python bench.py --workload $W --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$W ms/step', round(d['ms_per_step'],3), d['config']['stage_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'recall', d['config']['recall_at_10'], 'e2e', round(d['e2e']['ms_per_step'],2))"
done
python bench.py --workload c5s --steps 2 --warmup 3 --profile > gpurun_out/plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5s.csv python bench.py --workload c5s --steps 2 --warmup 3 --profile > gpurun_out/ncu1.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --profile > gpurun_out/ncu1b.log 2>&1
