"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``vectorindex_b200`` never does.

``lib()``      -> libvix_oracle.so   (our C restatement, oracle/vix_oracle_*.c)
``ref_lib()``  -> oracle/_ref/libcpq_ref[_omp].so (the reference's own C encoder, compiled
                  unmodified from /root/reference by ``make -C oracle ref``); None if absent.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = {}

METRIC_L2, METRIC_IP, METRIC_COSINE = 0, 1, 2
ORDER_MIN, ORDER_MAX = 0, 1

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])
    if ref and os.path.exists("/root/reference/Sources/CPQEncode/pq_encode.c"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libvix_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _LIB = C.CDLL(path)
        _LIB.vo_l2sqr_direct.restype = C.c_float
        _LIB.vo_norm_l2sq.restype = C.c_float
        _LIB.vo_l2sqr_dot_fused.restype = C.c_float
        _LIB.vo_ip.restype = C.c_float
        _LIB.vo_km12_l2sq.restype = C.c_float
        _LIB.vo_km11_l2sq.restype = C.c_float
        _LIB.vo_lut_l2sqr.restype = C.c_float
        _LIB.vo_lut_dot.restype = C.c_float
        _LIB.vo_pq_sqnorm.restype = C.c_float
        _LIB.vo_lcg_next.restype = C.c_uint64
    return _LIB


def set_threads(n: int = 0) -> int:
    """Host threads of the oracle's OpenMP loops (n > 0 sets them); returns the count in effect."""
    return int(lib().vo_set_threads(C.c_int(int(n))))


def ref_lib(omp: bool = False):
    key = "omp" if omp else "st"
    if key not in _REF:
        path = os.path.join(_HERE, "_ref", "libcpq_ref_omp.so" if omp else "libcpq_ref.so")
        _REF[key] = C.CDLL(path) if os.path.exists(path) else None
    return _REF[key]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


# ---------------------------------------------------------------- pair kernels
def _pair(fn, a, b):
    a, b = _f32(a), _f32(b)
    return float(fn(_p(a, f32p), _p(b, f32p), C.c_int(a.size)))


def l2sqr_direct(q, x): return _pair(lib().vo_l2sqr_direct, q, x)
def ip(q, x): return _pair(lib().vo_ip, q, x)
def km12_l2sq(a, b): return _pair(lib().vo_km12_l2sq, a, b)
def km11_l2sq(a, b): return _pair(lib().vo_km11_l2sq, a, b)
def lut_l2sqr(a, b): return _pair(lib().vo_lut_l2sqr, a, b)
def lut_dot(a, b): return _pair(lib().vo_lut_dot, a, b)


def norm_l2sq(x):
    x = _f32(x)
    return float(lib().vo_norm_l2sq(_p(x, f32p), C.c_int(x.size)))


def pq_sqnorm(x):
    x = _f32(x)
    return float(lib().vo_pq_sqnorm(_p(x, f32p), C.c_int(x.size)))


# ---------------------------------------------------------------- block scoring
def l2sqr_block(q, xb, xb_norm=None, q_norm=float("nan")):
    q, xb = _f32(q), _f32(xb)
    n, d = xb.shape
    out = np.empty(n, dtype=np.float32)
    xn = _f32(xb_norm) if xb_norm is not None else None
    lib().vo_l2sqr_block(_p(q, f32p), _p(xb, f32p), C.c_int64(n), C.c_int(d), _p(out, f32p),
                         _p(xn, f32p), C.c_float(q_norm))
    return out


def ip_block(q, xb):
    q, xb = _f32(q), _f32(xb)
    n, d = xb.shape
    out = np.empty(n, dtype=np.float32)
    lib().vo_ip_block(_p(q, f32p), _p(xb, f32p), C.c_int64(n), C.c_int(d), _p(out, f32p))
    return out


# ---------------------------------------------------------------- selection
def select_topk(scores, k, ordering=ORDER_MIN, ids=None):
    scores = _f32(scores)
    n = scores.size
    kk = max(0, min(k, n))
    os_ = np.empty(max(kk, 1), dtype=np.float32)
    oi = np.empty(max(kk, 1), dtype=np.int32)
    idp = np.ascontiguousarray(ids, dtype=np.int32) if ids is not None else None
    got = lib().vo_select_topk(_p(scores, f32p), _p(idp, i32p), C.c_int64(n), C.c_int(k),
                               C.c_int(ordering), _p(os_, f32p), _p(oi, i32p))
    return os_[:got].copy(), oi[:got].copy()


def merge_topk(lists, k, ordering=ORDER_MIN):
    """lists: sequence of (scores, ids) each sorted best->worst."""
    nl = len(lists)
    stride = max([len(s) for s, _ in lists] + [1])
    sc = np.zeros((max(nl, 1), stride), dtype=np.float32)
    idm = np.zeros((max(nl, 1), stride), dtype=np.int32)
    lens = np.zeros(max(nl, 1), dtype=np.int32)
    for l, (s, i) in enumerate(lists):
        sc[l, :len(s)] = s
        idm[l, :len(s)] = i
        lens[l] = len(s)
    os_ = np.empty(max(k, 1), dtype=np.float32)
    oi = np.empty(max(k, 1), dtype=np.int32)
    got = lib().vo_merge_topk(_p(sc, f32p), _p(idm, i32p), _p(lens, i32p), C.c_int(nl),
                              C.c_int(stride), C.c_int(k), C.c_int(ordering), _p(os_, f32p), _p(oi, i32p))
    return os_[:got].copy(), oi[:got].copy()


# ---------------------------------------------------------------- coarse quantiser
def centroid_norms(c):
    c = _f32(c)
    kc, d = c.shape
    out = np.empty(kc, dtype=np.float32)
    lib().vo_centroid_norms(_p(c, f32p), C.c_int(kc), C.c_int(d), _p(out, f32p))
    return out


def centroid_batch_score(queries, centroids, metric=METRIC_L2, cnorms=None):
    queries, centroids = _f32(queries), _f32(centroids)
    q, d = queries.shape
    kc = centroids.shape[0]
    if cnorms is None and metric in (METRIC_L2, METRIC_COSINE):
        cnorms = centroid_norms(centroids)
    cn = _f32(cnorms) if cnorms is not None else None
    out = np.empty((q, kc), dtype=np.float32)
    lib().vo_centroid_batch_score(_p(queries, f32p), C.c_int64(q), _p(centroids, f32p), C.c_int(kc),
                                  C.c_int(d), C.c_int(metric), _p(cn, f32p), _p(out, f32p))
    return out


def probe_select_batch(queries, centroids, nprobe, metric=METRIC_L2, cnorms=None):
    queries, centroids = _f32(queries), _f32(centroids)
    q, d = queries.shape
    kc = centroids.shape[0]
    if cnorms is None and metric in (METRIC_L2, METRIC_COSINE):
        cnorms = centroid_norms(centroids)
    cn = _f32(cnorms) if cnorms is not None else None
    idx = np.empty((q, nprobe), dtype=np.int32)
    sc = np.empty((q, nprobe), dtype=np.float32)
    lib().vo_probe_select_batch(_p(queries, f32p), C.c_int64(q), _p(centroids, f32p), C.c_int(kc),
                                C.c_int(d), C.c_int(metric), _p(cn, f32p), C.c_int(nprobe),
                                _p(idx, i32p), _p(sc, f32p))
    return idx, sc


def assign(x, centroids):
    x, centroids = _f32(x), _f32(centroids)
    n, d = x.shape
    a = np.empty(n, dtype=np.int32)
    dist = np.empty(n, dtype=np.float32)
    lib().vo_assign(_p(x, f32p), C.c_int64(n), _p(centroids, f32p), C.c_int(centroids.shape[0]),
                    C.c_int(d), _p(a, i32p), _p(dist, f32p))
    return a, dist


def assign_metric(x, centroids, metric, cnorms=None):
    x, centroids = _f32(x), _f32(centroids)
    n, d = x.shape
    if cnorms is None and metric in (METRIC_L2, METRIC_COSINE):
        cnorms = centroid_norms(centroids)
    cn = _f32(cnorms) if cnorms is not None else None
    a = np.empty(n, dtype=np.int32)
    lib().vo_assign_metric(_p(x, f32p), C.c_int64(n), _p(centroids, f32p), C.c_int(centroids.shape[0]),
                           C.c_int(d), C.c_int(metric), _p(cn, f32p), _p(a, i32p))
    return a


# ---------------------------------------------------------------- PQ encode
def pq_encode_u8(x, codebooks, m, ks=256, centroid_sq=None, coarse=None, assign_=None, use_dot=True):
    x, codebooks = _f32(x), _f32(codebooks)
    n, d = x.shape
    codes = np.empty((n, m), dtype=np.uint8)
    csq = _f32(centroid_sq) if centroid_sq is not None else None
    co = _f32(coarse) if coarse is not None else None
    asg = np.ascontiguousarray(assign_, dtype=np.int32) if assign_ is not None else None
    lib().vo_pq_encode_u8(_p(x, f32p), C.c_int64(n), C.c_int(d), C.c_int(m), C.c_int(ks),
                          _p(codebooks, f32p), _p(csq, f32p), _p(co, f32p), _p(asg, i32p),
                          C.c_int(1 if use_dot else 0), _p(codes, u8p))
    return codes


def pq_encode_u4(x, codebooks, m, ks=16, coarse=None, assign_=None):
    x, codebooks = _f32(x), _f32(codebooks)
    n, d = x.shape
    codes = np.empty((n, m // 2), dtype=np.uint8)
    co = _f32(coarse) if coarse is not None else None
    asg = np.ascontiguousarray(assign_, dtype=np.int32) if assign_ is not None else None
    lib().vo_pq_encode_u4(_p(x, f32p), C.c_int64(n), C.c_int(d), C.c_int(m), C.c_int(ks),
                          _p(codebooks, f32p), _p(co, f32p), _p(asg, i32p), _p(codes, u8p))
    return codes


def pq_centroid_sq(codebooks, m, ks, dsub, swift=False):
    codebooks = _f32(codebooks)
    out = np.empty(m * ks, dtype=np.float32)
    fn = lib().vo_pq_centroid_sq_swift if swift else lib().vo_pq_centroid_sq_seq
    fn(_p(codebooks, f32p), C.c_int(m), C.c_int(ks), C.c_int(dsub), _p(out, f32p))
    return out


class PQEncodeOpts(C.Structure):
    """Mirror of PQEncodeOpts, /root/reference/Sources/CPQEncode/include/cpq_encode.h:30-38."""
    _fields_ = [("layout", C.c_int), ("use_dot_trick", C.c_bool), ("precompute_x_norm2", C.c_bool),
                ("prefetch_distance", C.c_int), ("num_threads", C.c_int), ("soa_block_B", C.c_int),
                ("interleave_g", C.c_int)]


def ref_encode(fn_name, x, codebooks, m, ks, centroid_sq=None, coarse=None, assign_=None, opts=None,
               omp=False, packed_u4=False):
    """Call one of the reference's own cpq_encode_* symbols (oracle/_ref)."""
    L = ref_lib(omp)
    if L is None:
        raise RuntimeError("oracle/_ref is not built")
    x, codebooks = _f32(x), _f32(codebooks)
    n, d = x.shape
    codes = np.zeros((n, m // 2 if packed_u4 else m), dtype=np.uint8)
    args = [_p(x, f32p), C.c_int64(n), C.c_int(d), C.c_int(m), C.c_int(ks), _p(codebooks, f32p)]
    if centroid_sq is not None:
        csq = _f32(centroid_sq)
        args.append(_p(csq, f32p))
    if coarse is not None:
        co = _f32(coarse)
        asg = np.ascontiguousarray(assign_, dtype=np.int32)
        args += [_p(co, f32p), _p(asg, i32p)]
    args.append(_p(codes, u8p))
    args.append(C.byref(opts) if opts is not None else None)
    fn = getattr(L, fn_name)
    fn.restype = None
    fn(*args)
    return codes


# ---------------------------------------------------------------- LUT / ADC
def pq_lut_l2(q, codebooks, m, ks, cnorms=None, use_dot=-1, include_q=True, strict_fp=False, q_sub_norms=None):
    q, codebooks = _f32(q), _f32(codebooks)
    lut = np.empty(m * ks, dtype=np.float32)
    cn = _f32(cnorms) if cnorms is not None else None
    qs = _f32(q_sub_norms) if q_sub_norms is not None else None
    lib().vo_pq_lut_l2(_p(q, f32p), C.c_int(q.size), C.c_int(m), C.c_int(ks), _p(codebooks, f32p),
                       _p(lut, f32p), _p(cn, f32p), _p(qs, f32p), C.c_int(use_dot),
                       C.c_int(int(include_q)), C.c_int(int(strict_fp)))
    return lut.reshape(m, ks)


def pq_lut_residual_l2(q, coarse, codebooks, m, ks, cnorms=None, use_dot=-1, include_q=True, strict_fp=False):
    q, coarse, codebooks = _f32(q), _f32(coarse), _f32(codebooks)
    lut = np.empty(m * ks, dtype=np.float32)
    cn = _f32(cnorms) if cnorms is not None else None
    lib().vo_pq_lut_residual_l2(_p(q, f32p), _p(coarse, f32p), C.c_int(q.size), C.c_int(m), C.c_int(ks),
                                _p(codebooks, f32p), _p(lut, f32p), _p(cn, f32p), C.c_int(use_dot),
                                C.c_int(int(include_q)), C.c_int(int(strict_fp)))
    return lut.reshape(m, ks)


def adc_scan_u8(codes, lut, m, ks=256, stride=0, bias=0.0, strict_fp=False):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    lut = _f32(lut)
    st = stride if stride > 0 else m
    n = codes.size // st
    out = np.empty(n, dtype=np.float32)
    lib().vo_adc_scan_u8(_p(codes, u8p), C.c_int64(n), C.c_int(m), C.c_int(ks), _p(lut, f32p),
                         _p(out, f32p), C.c_int(stride), C.c_float(bias), C.c_int(int(strict_fp)))
    return out


def adc_scan_u4(codes, lut, m, ks=16, stride=0, bias=0.0, strict_fp=False):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    lut = _f32(lut)
    st = stride if stride > 0 else m // 2
    n = codes.size // st
    out = np.empty(n, dtype=np.float32)
    lib().vo_adc_scan_u4(_p(codes, u8p), C.c_int64(n), C.c_int(m), C.c_int(ks), _p(lut, f32p),
                         _p(out, f32p), C.c_int(stride), C.c_float(bias), C.c_int(int(strict_fp)))
    return out


# ---------------------------------------------------------------- composed searches
def flat_search(queries, xb, k, metric=METRIC_L2):
    queries, xb = _f32(queries), _f32(xb)
    nq, d = queries.shape
    n = xb.shape[0]
    dist = np.empty((nq, k), dtype=np.float32)
    ids = np.empty((nq, k), dtype=np.int64)
    raw = np.empty((nq, k), dtype=np.float32)
    lib().vo_flat_search(_p(queries, f32p), C.c_int64(nq), _p(xb, f32p), C.c_int64(n), C.c_int(d),
                         C.c_int(metric), C.c_int(k), _p(dist, f32p), _p(ids, i64p), _p(raw, f32p))
    return dist, ids, raw


def ivfpq_search(queries, coarse, codebooks, cb_norms, list_offsets, codes, ids, m, ks, nprobe, k,
                 metric=METRIC_L2, coarse_norms=None):
    queries, coarse, codebooks = _f32(queries), _f32(coarse), _f32(codebooks)
    nq, d = queries.shape
    kc = coarse.shape[0]
    if coarse_norms is None:
        coarse_norms = centroid_norms(coarse)
    cn = _f32(coarse_norms)
    cbn = _f32(cb_norms) if cb_norms is not None else None
    lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float32)
    oid = np.empty((nq, k), dtype=np.int64)
    probes = np.empty((nq, nprobe), dtype=np.int32)
    lib().vo_ivfpq_search(_p(queries, f32p), C.c_int64(nq), C.c_int(d), _p(coarse, f32p), C.c_int(kc),
                          _p(cn, f32p), C.c_int(m), C.c_int(ks), _p(codebooks, f32p), _p(cbn, f32p),
                          _p(lo, i64p), _p(codes, u8p), _p(ids, i64p), C.c_int(nprobe), C.c_int(k),
                          C.c_int(metric), _p(dist, f32p), _p(oid, i64p), _p(probes, i32p))
    return dist, oid, probes


def ivfflat_search(queries, coarse, list_offsets, vecs, ids, nprobe, k, metric=METRIC_L2, coarse_norms=None):
    queries, coarse, vecs = _f32(queries), _f32(coarse), _f32(vecs)
    nq, d = queries.shape
    kc = coarse.shape[0]
    if coarse_norms is None:
        coarse_norms = centroid_norms(coarse)
    cn = _f32(coarse_norms)
    lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float32)
    oid = np.empty((nq, k), dtype=np.int64)
    lib().vo_ivfflat_search(_p(queries, f32p), C.c_int64(nq), C.c_int(d), _p(coarse, f32p), C.c_int(kc),
                            _p(cn, f32p), _p(lo, i64p), _p(vecs, f32p), _p(ids, i64p), C.c_int(nprobe),
                            C.c_int(k), C.c_int(metric), _p(dist, f32p), _p(oid, i64p))
    return dist, oid


def build_lists(assign_, kc):
    """CSR inverted lists in ascending-id order: returns (offsets[kc+1], order[n]) where ``order`` is
    the row permutation (stable by list) -- the AoS list format of Kernels/IVFAppend.swift."""
    a = np.asarray(assign_, dtype=np.int64)
    order = np.argsort(a, kind="stable")
    counts = np.bincount(a, minlength=kc)
    off = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    return off, order


# ---------------------------------------------------------------- trainers
def kmeanspp_seed(data, k, seed=42, stream=0):
    data = _f32(data)
    n, d = data.shape
    cents = np.empty((k, d), dtype=np.float32)
    chosen = np.empty(k, dtype=np.int64)
    rc = lib().vo_kmeanspp_seed(_p(data, f32p), C.c_int64(n), C.c_int(d), C.c_int(k), C.c_uint64(seed),
                                C.c_uint64(stream), _p(cents, f32p), _p(chosen, i64p))
    if rc != 0:
        raise ValueError("kmeanspp_seed: invalid parameters")
    return cents, chosen


def kmeans_minibatch(x, kc, init=None, batch_size=1024, epochs=10, tol=1e-4, seed=0, stream=0,
                     compute_assignments=False):
    x = _f32(x)
    n, d = x.shape
    cents = np.empty((kc, d), dtype=np.float32)
    ini = _f32(init) if init is not None else None
    asg = np.empty(n, dtype=np.int32) if compute_assignments else None
    cap = 1 << 16
    emp = np.zeros(cap, dtype=np.int64)
    ed = C.c_int(0)
    bd = C.c_int64(0)
    rc = lib().vo_kmeans_minibatch(_p(x, f32p), C.c_int64(n), C.c_int(d), C.c_int(kc), _p(ini, f32p),
                                   C.c_int(batch_size), C.c_int(epochs), C.c_float(tol), C.c_uint64(seed),
                                   C.c_uint64(stream), _p(cents, f32p), _p(asg, i32p), _p(emp, i64p),
                                   C.c_int(cap), C.byref(ed), C.byref(bd))
    return rc, cents, asg, dict(epochs=ed.value, batches=bd.value, empties=emp[:min(cap, bd.value)].copy())


class PQTrainCfg(C.Structure):
    _fields_ = [("ks", C.c_int), ("m", C.c_int), ("algorithm", C.c_int), ("max_iters", C.c_int),
                ("batch_size", C.c_int), ("empty_policy", C.c_int), ("sample_n", C.c_int64),
                ("seed", C.c_uint64), ("stream_id", C.c_uint64), ("tol", C.c_float),
                ("precompute_x_norm2", C.c_int), ("compute_centroid_norms", C.c_int)]


def pq_train_cfg(**kw):
    cfg = PQTrainCfg()
    lib().vo_pq_train_cfg_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def pq_train(x, m, ks, coarse=None, assign_=None, **kw):
    x = _f32(x)
    n, d = x.shape
    cfg = pq_train_cfg(m=m, ks=ks, **kw)
    dsub = d // m
    cb = np.zeros(m * ks * dsub, dtype=np.float32)
    norms = np.zeros(m * ks, dtype=np.float32)
    dist = C.c_double(0)
    co = _f32(coarse) if coarse is not None else None
    asg = np.ascontiguousarray(assign_, dtype=np.int32) if assign_ is not None else None
    rc = lib().vo_pq_train(_p(x, f32p), C.c_int64(n), C.c_int(d), C.byref(cfg), _p(co, f32p), _p(asg, i32p),
                           _p(cb, f32p), _p(norms, f32p), C.byref(dist))
    return rc, cb.reshape(m, ks, dsub), norms.reshape(m, ks), dist.value


def pq_train_streaming(chunks, d, m, ks, **kw):
    chunks = [_f32(c) for c in chunks]
    cfg = pq_train_cfg(m=m, ks=ks, **kw)
    dsub = d // m
    cb = np.zeros(m * ks * dsub, dtype=np.float32)
    arr = (f32p * len(chunks))(*[_p(c, f32p) for c in chunks])
    cn = np.array([c.size // d for c in chunks], dtype=np.int64)
    rc = lib().vo_pq_train_streaming(arr, _p(cn, i64p), C.c_int(len(chunks)), C.c_int(d), C.byref(cfg),
                                     _p(cb, f32p), None)
    return rc, cb.reshape(m, ks, dsub)
