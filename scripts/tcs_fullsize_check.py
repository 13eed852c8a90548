"""The list-major tensor-core scan against the query-major scan at a bench workload's FULL size (default C5, one GPU):
ids and distance bits of every query of the batch must be identical.

    python scripts/tcs_fullsize_check.py [workload=c5]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from vectorindex_b200 import _lib  # noqa: E402

cfg = dict(bench.PRESETS[sys.argv[1] if len(sys.argv) > 1 else "c5"])
dev = torch.device("cuda", 0)
synth = bench.Synth(cfg, dev)
idx, _, _, tm = bench.build_index(cfg, synth, 0, 1)
q = synth.queries(cfg["nq"])
os.environ["VIX_TC_SCAN"] = "0"
d0, i0 = idx.batch_search(q, cfg["k"])
n0 = _lib.lib().vix_scan_tc_launches()
os.environ["VIX_TC_SCAN"] = "1"
os.environ["VIX_TC_SCAN_DEBUG"] = "1"
d1, i1 = idx.batch_search(q, cfg["k"])
assert _lib.lib().vix_scan_tc_launches() == n0 + 1, "the tensor-core path was not taken"
same_ids = bool(torch.equal(i0, i1))
same_bits = bool(torch.equal(d0.view(torch.int32), d1.view(torch.int32)))
print(f"{cfg['label']}: {q.shape[0]} queries x k={cfg['k']}: ids identical {same_ids}, distance bits identical {same_bits}")
if not (same_ids and same_bits):
    bad = (i0 != i1).any(dim=1).nonzero().flatten()
    print("queries that differ:", bad[:20].tolist(), "of", int(bad.numel()))
    for r in bad[:5].tolist():
        print(r, i0[r].tolist(), i1[r].tolist(), d0[r].tolist(), d1[r].tolist())
    sys.exit(1)
