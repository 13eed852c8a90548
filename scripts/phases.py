"""Phase split of the fused scan kernel on a bench workload: python scripts/phases.py c5s"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from vectorindex_b200 import _lib
cfg = dict(bench.PRESETS[sys.argv[1] if len(sys.argv) > 1 else "c5s"])
dev = torch.device("cuda", 0)
synth = bench.Synth(cfg, dev)
idx, sh, gt, bt = bench.build_index(cfg, synth, 0, 1)
q = synth.queries(cfg["nq"])
for _ in range(3):
    d, i, st = idx.batch_search(q, cfg["k"], stats=True)
tot = st.cycles_prologue + st.cycles_scan + st.cycles_tail
print(f"{sys.argv[1:]} scan {st.ms_scan:.3f} ms probe {st.ms_coarse:.3f} ms; per-CTA phase split: prologue {st.cycles_prologue/tot:.3f} "
      f"scan {st.cycles_scan/tot:.3f} tail {st.cycles_tail/tot:.3f}; cycles/query/SM: {tot/cfg['nq']:.0f} "
      f"(prologue {st.cycles_prologue/cfg['nq']:.0f}, scan {st.cycles_scan/cfg['nq']:.0f}, tail {st.cycles_tail/cfg['nq']:.0f}); "
      f"prologue pieces per query: select {st.cycles_select/cfg['nq']:.0f} probe table {st.cycles_probe_table/cfg['nq']:.0f} "
      f"lut {st.cycles_lut/cfg['nq']:.0f}; merge candidates per query {st.merge_candidates/cfg['nq']:.1f}")
