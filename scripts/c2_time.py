import ctypes as C, os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vectorindex_b200 import _lib, datagen, kernels as vk
dev = torch.device("cuda", 0)
L = _lib.lib(); _lib.check(L.vix_set_device(0)); _lib.check(L.vix_set_async(1))
n, d, m, ks = 1_000_000, 128, 16, 256
x = torch.from_numpy(datagen.bench_vectors(n, d, 123, normalize=False)).to(dev)
rng = np.random.default_rng(0)
cb = torch.from_numpy(rng.uniform(-1, 1, (m, ks, d // m)).astype(np.float32)).to(dev)
norms = (cb * cb).sum(-1).contiguous()
def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
print(os.environ.get("VIX_LIB_PATH", "default"), "encode ms", timed(lambda: vk.pq_encode_u8_f32_withCSQ(x, cb.reshape(-1), norms.reshape(-1), m, ks)))
