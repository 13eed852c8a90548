"""BASELINE.json configs[0..2] at their FULL sizes: the CUDA path against the oracle on the slices the oracle finishes in
seconds, plus size-independent properties over the whole output (sortedness, uniqueness, membership in the probed lists,
batch independence, idempotence).  configs[3..4] (10M x 768, 100M x 96) are exercised by bench.py, which checks recall and
the top-k overlap with the oracle on the full index."""
import numpy as np
import pytest

from test_gpu_parity import bits

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def test_c1_flat_l2_100k_x_128_1k_queries(oracle, vk):
    from vectorindex_b200 import datagen
    n, d, nq, k = 100_000, 128, 1000, 10
    xb = datagen.bench_vectors(n, d, 123)                              # the reference bench's generator / seeds
    q = datagen.bench_vectors(nq, d, 321)
    gd, gi = vk.flat_search_f32(q, xb, k, 0)
    sub = 96
    od, oi, _ = oracle.flat_search(q[:sub], xb, k, 0)
    assert np.array_equal(gi[:sub], oi) and np.array_equal(bits(gd[:sub]), bits(od))
    # whole output: ascending distances, distinct rows, and each distance is the reference kernel's value for that row
    assert (np.diff(gd, axis=1) >= 0).all()
    assert all(np.unique(r).size == k for r in gi)
    rng = np.random.default_rng(0)
    for r in rng.choice(nq, 40, replace=False):
        ref = np.sqrt(oracle.l2sqr_block(q[r], xb[gi[r]]))
        assert np.array_equal(bits(gd[r]), bits(ref.astype(np.float32)))
    # batch independence: a query answers the same alone as inside the batch
    d1, i1 = vk.flat_search_f32(q[500:517], xb, k, 0)
    assert np.array_equal(i1, gi[500:517]) and np.array_equal(bits(d1), bits(gd[500:517]))


def test_c2_pq_encode_1m_x_128_m16(oracle, vk):
    from vectorindex_b200 import datagen
    n, d, m, ks = 1_000_000, 128, 16, 256
    x = datagen.bench_vectors(n, d, 123, normalize=False)
    rc, cb, norms, _ = oracle.pq_train(x[:4000], m, ks, max_iters=3, sample_n=0)
    assert rc == 0
    codes = np.asarray(vk.pq_encode_u8_f32_withCSQ(x, cb.reshape(-1), norms.reshape(-1), m, ks)).reshape(n, m)
    for lo, hi in ((0, 15000), (492_000, 500_000), (985_000, n)):       # oracle (reference C arithmetic) on slices: bit-exact
        ref = oracle.pq_encode_u8(x[lo:hi], cb, m, ks, centroid_sq=norms.reshape(-1))
        assert np.array_equal(codes[lo:hi], ref)
    # a row's code does not depend on the batch it is encoded in; checksum of the whole equals the checksum of the halves
    part = np.asarray(vk.pq_encode_u8_f32_withCSQ(x[333_333:400_001], cb.reshape(-1), norms.reshape(-1), m, ks)).reshape(-1, m)
    assert np.array_equal(part, codes[333_333:400_001])
    halves = [np.asarray(vk.pq_encode_u8_f32_withCSQ(x[a:b], cb.reshape(-1), norms.reshape(-1), m, ks)).reshape(-1, m)
              for a, b in ((0, n // 2), (n // 2, n))]
    assert sum(int(h.astype(np.uint64).sum()) for h in halves) == int(codes.astype(np.uint64).sum())
    # without the precomputed norms the encoder computes them itself (PQEncode.swift:88-92): same codes
    own = np.asarray(vk.pq_encode_u8_f32(x[:50_000], cb.reshape(-1), m, ks)).reshape(-1, m)
    assert np.array_equal(own, codes[:50_000])


def test_c3_ivfpq_1m_x_128_nlist4096_nprobe32_m16_10k_queries(oracle):
    from vectorindex_b200 import datagen
    from vectorindex_b200._lib import KMeansCfg, PQTrainCfg
    from vectorindex_b200.index import IVFPQIndex
    n, d, nlist, nprobe, m, nq, k = 1_000_000, 128, 4096, 32, 16, 10_000, 10
    x = datagen.sift_like(n + nq, d, 4096, 7)
    xb, q = np.ascontiguousarray(x[:n]), np.ascontiguousarray(x[n:])
    idx = IVFPQIndex(d, "euclidean", nlist=nlist, nprobe=nprobe, m=m)
    idx.optimize(xb[:131_072], KMeansCfg(1024, 4, 1e-4, 42, 0, False, 1), PQTrainCfg(0, 6, 1e-4, 1024, 65536, 42, 0, 0, 1))
    idx.batch_insert(xb)
    assert idx.count == n
    gd, gi, gp = idx.batch_search(q, k, return_probes=True)
    coarse = idx.get_coarse()
    cb, norms = idx.get_codebooks()
    off, codes, lids, asg = idx.export_lists()
    # stage-wise parity with the oracle on slices: list assignment and codes bit-exact, probe lists exact, distances 1e-5
    sl = slice(700_000, 712_000)
    oasg, _ = oracle.assign(xb[sl], coarse)
    assert np.array_equal(asg[sl], oasg)
    ocodes = oracle.pq_encode_u8(xb[sl], cb, m, 256, centroid_sq=norms.reshape(-1), coarse=coarse, assign_=oasg)
    pos = np.empty(n, dtype=np.int64)
    pos[lids] = np.arange(n)                                           # ids are the add order 0..n-1
    assert np.array_equal(codes[pos[sl]], ocodes)
    sub = 48
    od, oi, op = oracle.ivfpq_search(q[:sub], coarse, cb, norms, off, codes, lids, m, 256, nprobe, k, 0)
    assert np.array_equal(gp[:sub], op)
    np.testing.assert_allclose(gd[:sub], od, rtol=RTOL)
    assert np.mean([len(set(gi[r]) & set(oi[r])) / k for r in range(sub)]) > 0.99
    # whole output: k results, ascending, distinct ids, every id lives in one of the query's probed lists
    assert (gi >= 0).all() and (np.diff(gd, axis=1) >= 0).all()
    assert all(np.unique(r).size == k for r in gi[::37])
    assert all(np.isin(asg[gi[r]], gp[r]).all() for r in range(0, nq, 13))
    # idempotence and batch independence
    d2, i2 = idx.batch_search(q, k)
    assert np.array_equal(i2, gi) and np.array_equal(bits(d2), bits(gd))
    d3, i3 = idx.batch_search(q[4000:4100], k)
    assert np.array_equal(i3, gi[4000:4100]) and np.array_equal(bits(d3), bits(gd[4000:4100]))
