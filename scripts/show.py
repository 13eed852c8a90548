"""One-line digest of a bench.py JSON line: python scripts/show.py file.json"""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c = d["config"]
print(f"N={d['n_gpus']} {d['value'] / 1e3:.1f} k q/s, {d['ms_per_step']:.3f} ms/step; e2e {d['e2e']['value'] / 1e3:.1f} k q/s "
      f"({d['e2e']['ms_per_step']:.3f} ms); stages {c.get('stage_ms_per_step')}; phases {c.get('sharded_phase_ms')}; "
      f"roofline frac {d['roofline']['frac']:.3f} ({d['roofline']['ms_per_launch']:.3f} ms/launch); recall {c.get('recall_at_10')}; "
      f"clocks {d['clocks']}; launches {d['gpu_launches']}; cpu {d.get('cpu_baseline', {}).get('value')}")
