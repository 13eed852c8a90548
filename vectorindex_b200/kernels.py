"""Kernel-level operator API -- the host-side mirror of the reference's free functions on the search hot
path (same names, argument meaning and error behaviour), each a thin call into ``libvindex_b200.so``.

Reference interfaces (paths under /root/reference/Sources/VectorIndex):
  pq_encode_u8_f32 / pq_encode_residual_u8_f32 [_withCSQ] / u4     Operations/Quantization/PQEncode.swift:66-410
  pq_lut_l2_f32 / pq_lut_batch_l2_f32 / pq_lut_residual_l2_f32     Operations/Quantization/PQLUT.swift:191-465
  adc_scan_u8 / adc_scan_u4                                        Operations/Quantization/ADCScan.swift:99-149
  l2sqr_f32_block / ip_f32_block                                   Operations/Scoring/CABIBridge.swift:5-30
  selectTopK / mergeTopK                                           Operations/Selection/TopK.swift:127, TopKMerge.swift:11
  CentroidBatchScore.run / ivf_select_nprobe_batch_f32             Kernels/CentroidBatchScore.swift:39, IVFSelect.swift:242
  kmeans assignment (_vi_km12_assignAOS)                           Kernels/KMeansMiniBatchKernel.swift:341-359
  kmeansPlusPlusSeed / kmeans_minibatch_f32 / pq_train_f32         Kernels/KMeansSeeding.swift:167, KMeansMiniBatchKernel.swift:401, PQTrain.swift:83

Inputs may be numpy arrays (host) or torch CUDA tensors (device-resident; outputs then live on the same
device and nothing crosses PCIe).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (ADCScanOpts, KMeansCfg, METRIC_IP, METRIC_L2, ORDER_MAX, ORDER_MIN, PQEncodeOpts, PQLutOpts,
                   PQTrainCfg, VectorIndexError, as_input, check, empty_like_input, lib, ptr)

PQ_LAYOUT_AOS, PQ_LAYOUT_SOA_BLOCKED, PQ_LAYOUT_INTERLEAVED_BLOCK = 0, 1, 2


def _shape2(a):
    if a.ndim != 2:
        raise VectorIndexError(-1, "expected a 2-d array [n x d]")
    return int(a.shape[0]), int(a.shape[1])


def _opts_ptr(o):
    return C.byref(o) if o is not None else None


def _void_call(fn, *args):
    """cpq_* entry points return void like the reference; failures surface through vix_last_error()."""
    L = lib()
    L.vix_clear_error()
    fn(*args)
    err = L.vix_last_error()
    if err:
        raise VectorIndexError(-5, err.decode("utf-8", "replace"))


# ------------------------------------------------------------------------------------------------ PQ encode
def _codes_out(x, n, m, layout, B, g, u4):
    if u4:
        return empty_like_input(x, (n, m // 2), np.uint8)
    if layout == PQ_LAYOUT_SOA_BLOCKED:
        return empty_like_input(x, (m * ((n + B - 1) // B) * B,), np.uint8)
    if layout == PQ_LAYOUT_INTERLEAVED_BLOCK:
        return empty_like_input(x, (((n + g - 1) // g) * m * g,), np.uint8)
    return empty_like_input(x, (n, m), np.uint8)


def _encode(name, x, codebooks, m, ks, centroid_sq, coarse, assignments, opts, u4):
    x = as_input(x, np.float32)
    n, d = _shape2(x)
    if m <= 0 or d % m != 0:
        raise VectorIndexError(-1, f"{name}: d ({d}) must be divisible by m ({m})")   # PQEncode.swift:77-78
    codebooks = as_input(codebooks, np.float32)
    layout = int(opts.layout) if opts is not None else 0
    B = opts.soa_block_B if (opts is not None and opts.soa_block_B > 0) else 64
    g = opts.interleave_g if (opts is not None and opts.interleave_g > 0) else 8
    codes = _codes_out(x, n, m, 0 if u4 else layout, B, g, u4)
    if layout != PQ_LAYOUT_AOS and not u4 and isinstance(codes, np.ndarray):
        codes[:] = 0
    args = [ptr(x, np.float32), C.c_int64(n), C.c_int(d), C.c_int(m), C.c_int(ks), ptr(codebooks, np.float32)]
    keep = [x, codebooks]
    if centroid_sq is not None:
        csq = as_input(centroid_sq, np.float32)
        keep.append(csq)
        args.append(ptr(csq, np.float32))
    if coarse is not None:
        co = as_input(coarse, np.float32)
        asg = as_input(assignments, np.int32)
        keep += [co, asg]
        args += [ptr(co, np.float32), ptr(asg, np.int32)]
    args += [ptr(codes, np.uint8), _opts_ptr(opts)]
    _void_call(getattr(lib(), name), *args)
    return codes


def pq_encode_u8_f32(x, codebooks, m, ks=256, opts: PQEncodeOpts | None = None):
    """PQEncode.swift:66 -> cpq_encode_u8_f32 (dot-trick iff ks >= 64 when opts is None)."""
    return _encode("cpq_encode_u8_f32", x, codebooks, m, ks, None, None, None, opts, False)


def pq_encode_u8_f32_withCSQ(x, codebooks, centroid_sq, m, ks=256, opts: PQEncodeOpts | None = None):
    """PQEncode.swift:133 -> cpq_encode_u8_f32_with_csq."""
    return _encode("cpq_encode_u8_f32_with_csq", x, codebooks, m, ks, centroid_sq, None, None, opts, False)


def pq_encode_u4_f32(x, codebooks, m, ks=16, opts: PQEncodeOpts | None = None):
    return _encode("cpq_encode_u4_f32", x, codebooks, m, ks, None, None, None, opts, True)


def pq_encode_residual_u8_f32(x, codebooks, coarse_centroids, assignments, m, ks=256, opts=None):
    """PQEncode.swift:247 -> cpq_encode_residual_u8_f32."""
    return _encode("cpq_encode_residual_u8_f32", x, codebooks, m, ks, None, coarse_centroids, assignments, opts, False)


def pq_encode_residual_u8_f32_withCSQ(x, codebooks, centroid_sq, coarse_centroids, assignments, m, ks=256, opts=None):
    return _encode("cpq_encode_residual_u8_f32_with_csq", x, codebooks, m, ks, centroid_sq, coarse_centroids,
                   assignments, opts, False)


def pq_encode_residual_u4_f32(x, codebooks, coarse_centroids, assignments, m, ks=16, opts=None):
    return _encode("cpq_encode_residual_u4_f32", x, codebooks, m, ks, None, coarse_centroids, assignments, opts, True)


def pq_pack_u4(codes):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = np.empty(codes.size // 2, dtype=np.uint8)
    lib().cpq_pack_u4_bulk(ptr(codes), C.c_int(codes.size), ptr(out))
    return out


def pq_unpack_u4(packed):
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(packed.size * 2, dtype=np.uint8)
    lib().cpq_unpack_u4_bulk(ptr(packed), C.c_int(out.size), ptr(out))
    return out


# ------------------------------------------------------------------------------------------------ scoring
def l2sqr_f32_block(q, xb, xb_norm=None, q_norm=float("nan")):
    q, xb = as_input(q, np.float32), as_input(xb, np.float32)
    n, d = _shape2(xb)
    out = empty_like_input(xb, (n,), np.float32)
    xn = as_input(xb_norm, np.float32)
    check(lib().vix_l2sqr_f32_block(ptr(q, np.float32), ptr(xb, np.float32), C.c_int64(n), C.c_int(d),
                                    ptr(out, np.float32), ptr(xn), C.c_float(q_norm)))
    return out


def ip_f32_block(q, xb):
    q, xb = as_input(q, np.float32), as_input(xb, np.float32)
    n, d = _shape2(xb)
    out = empty_like_input(xb, (n,), np.float32)
    check(lib().vix_ip_f32_block(ptr(q, np.float32), ptr(xb, np.float32), C.c_int64(n), C.c_int(d), ptr(out, np.float32)))
    return out


def row_norms_f32(x):
    x = as_input(x, np.float32)
    n, d = _shape2(x)
    out = empty_like_input(x, (n,), np.float32)
    check(lib().vix_row_norms_f32(ptr(x, np.float32), C.c_int64(n), C.c_int(d), ptr(out, np.float32)))
    return out


def flat_search_f32(queries, xb, k, metric=METRIC_L2):
    """ScoreBlock.run + selectTopK + API distance mapping (FlatIndexOptimized.swift:390-477), batched."""
    queries, xb = as_input(queries, np.float32), as_input(xb, np.float32)
    nq, d = _shape2(queries)
    n, d2 = _shape2(xb)
    if d != d2:
        raise VectorIndexError(-1, f"dimension mismatch: query {d} vs base {d2}")
    kk = max(k, 0)
    dist = empty_like_input(queries, (nq, kk), np.float32)
    ids = empty_like_input(queries, (nq, kk), np.int64)
    if kk == 0:
        return dist, ids
    check(lib().vix_flat_search_f32(ptr(queries, np.float32), C.c_int64(nq), ptr(xb, np.float32), C.c_int64(n),
                                    C.c_int(d), C.c_int(metric), C.c_int(k), ptr(dist, np.float32), ptr(ids, np.int64)))
    return dist, ids


def selectTopK(scores, k, ordering=ORDER_MIN, ids=None):
    scores = as_input(scores, np.float32)
    n = int(scores.shape[0])
    keff = max(0, min(k, n))
    os_ = empty_like_input(scores, (max(keff, 1),), np.float32)
    oi = empty_like_input(scores, (max(keff, 1),), np.int32)
    idp = as_input(ids, np.int32)
    cnt = C.c_int(0)
    check(lib().vix_select_topk_f32(ptr(scores, np.float32), ptr(idp), C.c_int64(n), C.c_int(k), C.c_int(ordering),
                                    ptr(os_, np.float32), ptr(oi, np.int32), C.byref(cnt)))
    return os_[:keff], oi[:keff]


def mergeTopK(scores, ids, k, ordering=ORDER_MIN, lens=None):
    """scores/ids: [batch x nlists x stride] (best->worst per list); returns [batch x k]."""
    scores = as_input(scores, np.float32)
    ids = as_input(ids, np.int64)
    batch, nlists, stride = (int(v) for v in scores.shape)
    ln = as_input(lens, np.int32)
    os_ = empty_like_input(scores, (batch, k), np.float32)
    oi = empty_like_input(scores, (batch, k), np.int64)
    check(lib().vix_merge_topk_f32(ptr(scores, np.float32), ptr(ids, np.int64), ptr(ln), C.c_int64(batch),
                                   C.c_int(nlists), C.c_int(stride), C.c_int(k), C.c_int(ordering),
                                   ptr(os_, np.float32), ptr(oi, np.int64)))
    return os_, oi


def rerank_exact_topk(queries, cand_ids, xb, k, metric=METRIC_L2, xb_sq_norms=None):
    """rerank_exact_topk_batch with the DenseArray reader (Operations/Rerank/ExactRerank.swift:698-814):
    cand_ids [nq x C] rows of xb; returns (raw scores [nq x k], ids [nq x k]) best first, padded with +-inf / -1."""
    queries, xb = as_input(queries, np.float32), as_input(xb, np.float32)
    cand = as_input(cand_ids, np.int64)
    nq, d = _shape2(queries)
    n, _ = _shape2(xb)
    c = int(cand.shape[1])
    sc = empty_like_input(queries, (nq, k), np.float32)
    ids = empty_like_input(queries, (nq, k), np.int64)
    nr = as_input(xb_sq_norms, np.float32)
    check(lib().vix_rerank_exact_topk_f32(ptr(queries, np.float32), C.c_int64(nq), C.c_int(d), C.c_int(metric), ptr(cand, np.int64),
                                          C.c_int(c), C.c_int(k), ptr(xb, np.float32), C.c_int64(n), ptr(nr), ptr(sc, np.float32),
                                          ptr(ids, np.int64)))
    return sc, ids


# ------------------------------------------------------------------------------------------------ coarse quantiser
def centroid_batch_score(queries, centroids, metric=METRIC_L2, centroid_norms=None):
    queries, centroids = as_input(queries, np.float32), as_input(centroids, np.float32)
    q, d = _shape2(queries)
    kc, _ = _shape2(centroids)
    out = empty_like_input(queries, (q, kc), np.float32)
    cn = as_input(centroid_norms, np.float32)
    check(lib().vix_centroid_batch_score_f32(ptr(queries, np.float32), C.c_int64(q), ptr(centroids, np.float32),
                                             C.c_int(kc), C.c_int(d), C.c_int(metric), ptr(cn), ptr(out, np.float32)))
    return out


def ivf_select_nprobe_batch_f32(Q, centroids, nprobe, metric=METRIC_L2, centroid_norms=None, disabled_lists=None,
                                return_scores=True):
    Q, centroids = as_input(Q, np.float32), as_input(centroids, np.float32)
    b, d = _shape2(Q)
    kc, d2 = _shape2(centroids)
    if d != d2:
        raise VectorIndexError(-1, f"dimension mismatch: query {d} vs centroids {d2}")
    ids = empty_like_input(Q, (b, nprobe), np.int32)
    sc = empty_like_input(Q, (b, nprobe), np.float32) if return_scores else None
    cn = as_input(centroid_norms, np.float32)
    mask = as_input(disabled_lists, np.uint64)
    check(lib().vix_ivf_select_nprobe_batch_f32(ptr(Q, np.float32), C.c_int64(b), C.c_int(d), ptr(centroids, np.float32),
                                                C.c_int(kc), C.c_int(metric), C.c_int(nprobe), ptr(cn), ptr(mask),
                                                ptr(ids, np.int32), ptr(sc)))
    return (ids, sc) if return_scores else ids


def ivf_assign_f32(x, centroids, return_dist=False):
    """argmin_c L2^2(x_i, C_c), tie -> lower c; bit-exact with _vi_km12_assignAOS."""
    x, centroids = as_input(x, np.float32), as_input(centroids, np.float32)
    n, d = _shape2(x)
    kc, d2 = _shape2(centroids)
    if d != d2:
        raise VectorIndexError(-1, f"dimension mismatch: x {d} vs centroids {d2}")
    a = empty_like_input(x, (n,), np.int32)
    dist = empty_like_input(x, (n,), np.float32) if return_dist else None
    check(lib().vix_ivf_assign_f32(ptr(x, np.float32), C.c_int64(n), C.c_int(d), ptr(centroids, np.float32), C.c_int(kc),
                                   ptr(a, np.int32), ptr(dist)))
    return (a, dist) if return_dist else a


def ivf_assign_metric_f32(x, centroids, metric, centroid_norms=None):
    x, centroids = as_input(x, np.float32), as_input(centroids, np.float32)
    n, d = _shape2(x)
    kc, _ = _shape2(centroids)
    a = empty_like_input(x, (n,), np.int32)
    cn = as_input(centroid_norms, np.float32)
    check(lib().vix_ivf_assign_metric_f32(ptr(x, np.float32), C.c_int64(n), C.c_int(d), ptr(centroids, np.float32),
                                          C.c_int(kc), C.c_int(metric), ptr(cn), ptr(a, np.int32)))
    return a


# ------------------------------------------------------------------------------------------------ LUT / ADC
def pq_query_subnorms_f32(queries, m):
    """``pq_query_subnorms_f32`` (PQLUT.swift:174-187) for a batch: [nq x m] squared norms of the query sub-vectors."""
    queries = as_input(queries, np.float32)
    nq, d = _shape2(queries)
    out = empty_like_input(queries, (nq, m), np.float32)
    check(lib().vix_pq_query_subnorms_f32(ptr(queries, np.float32), C.c_int64(nq), C.c_int(d), C.c_int(m), ptr(out, np.float32)))
    return out


def pq_lut_batch_l2_f32(queries, codebooks, m, ks=256, centroid_norms=None, opts: PQLutOpts | None = None):
    queries, codebooks = as_input(queries, np.float32), as_input(codebooks, np.float32)
    nq, d = _shape2(queries)
    luts = empty_like_input(queries, (nq, m, ks), np.float32)
    cn = as_input(centroid_norms, np.float32)
    check(lib().vix_pq_lut_batch_l2_f32(ptr(queries, np.float32), C.c_int64(nq), C.c_int(d), C.c_int(m), C.c_int(ks),
                                        ptr(codebooks, np.float32), ptr(luts, np.float32), ptr(cn), _opts_ptr(opts)))
    return luts


def pq_lut_l2_f32(query, codebooks, m, ks=256, centroid_norms=None, opts: PQLutOpts | None = None):
    q = as_input(query, np.float32).reshape(1, -1)
    return pq_lut_batch_l2_f32(q, codebooks, m, ks, centroid_norms, opts)[0]


def pq_lut_residual_l2_f32(queries, coarse_ids, coarse_centroids, codebooks, m, ks=256, centroid_norms=None,
                           opts: PQLutOpts | None = None):
    queries, codebooks = as_input(queries, np.float32), as_input(codebooks, np.float32)
    coarse_centroids = as_input(coarse_centroids, np.float32)
    coarse_ids = as_input(coarse_ids, np.int32)
    nq, d = _shape2(queries)
    luts = empty_like_input(queries, (nq, m, ks), np.float32)
    cn = as_input(centroid_norms, np.float32)
    check(lib().vix_pq_lut_residual_l2_f32(ptr(queries, np.float32), ptr(coarse_ids, np.int32), C.c_int64(nq), C.c_int(d),
                                           ptr(coarse_centroids, np.float32), C.c_int(m), C.c_int(ks),
                                           ptr(codebooks, np.float32), ptr(luts, np.float32), ptr(cn), _opts_ptr(opts)))
    return luts


def _adc(fn, codes, lut, m, ks, opts, n):
    codes = as_input(codes, np.uint8)
    lut = as_input(lut, np.float32)
    out = empty_like_input(codes, (n,), np.float32)
    check(fn(ptr(codes, np.uint8), C.c_int64(n), C.c_int(m), C.c_int(ks), ptr(lut, np.float32), ptr(out, np.float32),
             _opts_ptr(opts)))
    return out


def adc_scan_u8(codes, lut, m, ks=256, opts: ADCScanOpts | None = None, n=None):
    if n is None:
        n = int(codes.shape[0])
    return _adc(lib().vix_adc_scan_u8, codes, lut, m, ks, opts, n)


def adc_scan_u4(codes, lut, m, ks=16, opts: ADCScanOpts | None = None, n=None):
    if n is None:
        n = int(codes.shape[0])
    return _adc(lib().vix_adc_scan_u4, codes, lut, m, ks, opts, n)


# ------------------------------------------------------------------------------------------------ trainers
def kmeans_cfg(batch_size=1024, epochs=10, tol=1e-4, seed=0, stream_id=0, compute_assignments=False, mode=0):
    return KMeansCfg(batch_size, epochs, tol, seed, stream_id, compute_assignments, mode)


def pq_train_cfg(algorithm=0, max_iters=25, tol=1e-4, batch_size=1024, sample_n=0, seed=42, stream_id=0,
                 empty_policy=0, mode=0):
    return PQTrainCfg(algorithm, max_iters, tol, batch_size, sample_n, seed, stream_id, empty_policy, mode)


def kmeansPlusPlusSeed(data, k, seed=42, stream_id=0):
    data = as_input(data, np.float32)
    n, d = _shape2(data)
    cents = empty_like_input(data, (k, d), np.float32)
    chosen = np.empty(k, dtype=np.int64)
    check(lib().vix_kmeanspp_seed_f32(ptr(data, np.float32), C.c_int64(n), C.c_int(d), C.c_int(k), C.c_uint64(seed),
                                      C.c_uint64(stream_id), ptr(cents, np.float32), ptr(chosen, np.int64)))
    return cents, chosen


def kmeans_minibatch_f32(x, kc, init_centroids=None, cfg: KMeansCfg | None = None, compute_assignments=False):
    x = as_input(x, np.float32)
    n, d = _shape2(x)
    cents = empty_like_input(x, (kc, d), np.float32)
    asg = empty_like_input(x, (n,), np.int32) if compute_assignments else None
    ini = as_input(init_centroids, np.float32)
    st = lib().vix_kmeans_minibatch_f32(ptr(x, np.float32), C.c_int64(n), C.c_int(d), C.c_int(kc), ptr(ini),
                                        _opts_ptr(cfg), ptr(cents, np.float32), ptr(asg))
    check(st, allow=(0, 1))
    return st, cents, asg


def pq_train_f32(x, m, ks=256, coarse_centroids=None, assignments=None, cfg: PQTrainCfg | None = None):
    x = as_input(x, np.float32)
    n, d = _shape2(x)
    if n == 0:
        raise VectorIndexError(-7, "pq_train_f32: empty input")
    if m <= 0 or d % m != 0:
        raise VectorIndexError(-1, "pq_train_f32: d must be divisible by m")
    need_n = cfg.sample_n if cfg is not None and cfg.sample_n > 0 else n
    if need_n < ks:                                                   # PQTrain.swift:127-135
        raise VectorIndexError(-7, "pq_train_f32: Insufficient training data: need at least ks vectors")
    dsub = d // m
    cb = empty_like_input(x, (m, ks, dsub), np.float32)
    norms = empty_like_input(x, (m, ks), np.float32)
    co = as_input(coarse_centroids, np.float32)
    asg = as_input(assignments, np.int32)
    st = lib().vix_pq_train_f32(ptr(x, np.float32), C.c_int64(n), C.c_int(d), C.c_int(m), C.c_int(ks), ptr(co), ptr(asg),
                                _opts_ptr(cfg), ptr(cb, np.float32), ptr(norms, np.float32))
    check(st, allow=(0, 1))
    return cb, norms


def pq_train_streaming_f32(chunks, m, ks=256, cfg=None):
    """``pq_train_streaming_f32`` (Kernels/PQTrain.swift:391-706): mini-batch PQ training over row blocks handed over one
    by one (numpy arrays or CUDA tensors, [n_c x d] each); reference parity.  Returns (codebooks [m x ks x dsub], norms)."""
    chunks = [as_input(c, np.float32) for c in chunks]
    if not chunks:
        raise VectorIndexError(-7, "pq_train_streaming_f32: empty input")
    d = int(chunks[0].shape[1])
    if m <= 0 or d % m != 0:
        raise VectorIndexError(-1, "pq_train_streaming_f32: d must be divisible by m")
    ptrs = (C.c_void_p * len(chunks))(*[ptr(c, np.float32).value or 0 for c in chunks])
    cn = (C.c_int64 * len(chunks))(*[int(c.shape[0]) for c in chunks])
    cb = np.empty((m, ks, d // m), dtype=np.float32)
    norms = np.empty((m, ks), dtype=np.float32)
    check(lib().vix_pq_train_streaming_f32(ptrs, cn, C.c_int(len(chunks)), C.c_int(d), C.c_int(m), C.c_int(ks), _opts_ptr(cfg),
                                           ptr(cb, np.float32), ptr(norms, np.float32)))
    return cb, norms


def accel_rank_candidates(queries, candidates, k, metric=METRIC_L2):
    """AccelerableIndex-shaped hand-off (AccelerableIndex.swift:15-127): candidates [c x d] in,
    (indices into candidates, distances) out, per query."""
    queries, candidates = as_input(queries, np.float32), as_input(candidates, np.float32)
    nq, d = _shape2(queries)
    c, _ = _shape2(candidates)
    idx = empty_like_input(queries, (nq, k), np.int32)
    dist = empty_like_input(queries, (nq, k), np.float32)
    check(lib().vix_accel_rank_candidates_f32(ptr(queries, np.float32), C.c_int64(nq), ptr(candidates, np.float32),
                                              C.c_int64(c), C.c_int(d), C.c_int(metric), C.c_int(k), ptr(idx, np.int32),
                                              ptr(dist, np.float32)))
    return idx, dist
