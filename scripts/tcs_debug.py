"""Debug aid: list-major tensor-core scan vs query-major scan vs oracle on one test shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import oracle
from test_gpu_parity import _make_ivfpq_problem, bits
from vectorindex_b200.index import IVFPQIndex

d, m, n, kc, nq, nprobe, k, kind = 96, 48, 20000, 32, 200, 8, 1, "sift"
xb, q, coarse, cb, norms = _make_ivfpq_problem(oracle, n, d, m, kc, nq, seed=d + m + k + n, sift=kind == "sift", unit=kind == "unit")
idx = IVFPQIndex(d, "euclidean", nlist=kc, nprobe=nprobe, m=m)
idx.set_coarse(coarse); idx.set_codebooks(cb, norms)
idx.batch_insert(xb)
os.environ["VIX_TC_SCAN"] = "0"
d1, i1, p1 = idx.batch_search(q, k, return_probes=True)
os.environ["VIX_TC_SCAN"] = "1"
d2, i2 = idx.batch_search(q, k)
off, codes, lids, asg = idx.export_lists()
od, oi, _ = oracle.ivfpq_search(q, coarse, cb, norms, off, codes, lids, m, 256, nprobe, k, 0)
bad = np.argwhere(i1 != i2)
print("mismatches", len(bad), "list sizes", np.diff(off))
for (qi, j) in bad[:10]:
    a, b = int(i1[qi, j]), int(i2[qi, j])
    print("q", qi, "rank", j, "old", a, float(d1[qi, j]), hex(bits(d1)[qi, j]), "new", b, float(d2[qi, j]), hex(bits(d2)[qi, j]),
          "oracle", int(oi[qi, j]), float(od[qi, j]), "probes", p1[qi], "list(old)", asg[a], "list(new)", asg[b])
