"""Device-resident timings of BASELINE.json configs[0..1] (the parity-test configurations; bench.py measures the IVF-PQ
ones): flat exact L2 search 100k x 128 / 1k queries / k 10, and PQ train + encode 1M x 128, M 16.  CUDA events on the
library's stream, 3 warm-ups, median of 10.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vectorindex_b200 import _lib, datagen, kernels as vk  # noqa: E402

dev = torch.device("cuda", 0)
L = _lib.lib()
_lib.check(L.vix_set_device(0))
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
_lib.check(L.vix_set_stream(C.c_void_p(stream.cuda_stream)))
_lib.check(L.vix_set_async(1))


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {}
# ---- C1
n, d, nq, k = 100_000, 128, 1000, 10
xb = torch.from_numpy(datagen.bench_vectors(n, d, 123)).to(dev)
q = torch.from_numpy(datagen.bench_vectors(nq, d, 321)).to(dev)
ms = timed(lambda: vk.flat_search_f32(q, xb, k, 0))
out["c1_flat_l2_100k_x_128_q1000_k10"] = {
    "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3,
    "algorithmic_tflops": 2.0 * nq * n * d / (ms * 1e-3) / 1e12,
    "note": "two TF32 tensor-core passes (group minima, emission) + exact rescoring; 25.6 GFLOP algorithmic per batch"}
q10 = torch.from_numpy(datagen.bench_vectors(10_000, d, 321)).to(dev)
ms = timed(lambda: vk.flat_search_f32(q10, xb, k, 0))
out["c1_flat_l2_100k_x_128_q10000_k10"] = {"ms_per_batch": ms, "queries_per_s": 10_000 / ms * 1e3,
                                           "algorithmic_tflops": 2.0 * 10_000 * n * d / (ms * 1e-3) / 1e12}
# ---- C2
n, d, m, ks = 1_000_000, 128, 16, 256
x = torch.from_numpy(datagen.bench_vectors(n, d, 123, normalize=False)).to(dev)
cfg = vk.pq_train_cfg(algorithm=0, max_iters=25, sample_n=65536, mode=1)      # GPU Lloyd on a strided sample
t_train = timed(lambda: vk.pq_train_f32(x, m, ks, cfg=cfg), reps=3, warm=1)
cb, norms = vk.pq_train_f32(x, m, ks, cfg=cfg)
ms = timed(lambda: vk.pq_encode_u8_f32_withCSQ(x, cb.reshape(-1), norms.reshape(-1), m, ks))
out["c2_pq_1m_x_128_m16"] = {
    "train_ms_lloyd_25_iters_65536_samples": t_train, "encode_ms": ms, "encode_vectors_per_s": n / ms * 1e3,
    "encode_algorithmic_gbs": (4.0 * n * d + n * m) / (ms * 1e-3) / 1e9,
    "encode_algorithmic_tflops": 2.0 * n * ks * d / (ms * 1e-3) / 1e12,
    "note": "bit-exact scalar-order argmin (pq_encode.c arithmetic); 528 MB in + out, 65.5 GFLOP per pass"}
print(json.dumps(out))
