"""Index API -- the host-side mirror of the reference's ``VectorIndexProtocol``
(/root/reference/Sources/VectorIndex/IndexProtocols.swift:50-103) for the three index kinds on the hot
path, over device-resident handles of ``libvindex_b200.so``:

  FlatIndex     FlatIndex / FlatIndexOptimized  (FlatIndex.swift, FlatIndexOptimized.swift:390-477)
  IVFIndex      IVF-Flat actor                  (IVFIndex.swift:279-451 optimize, :865-985 batchSearch)
  IVFPQIndex    the IVF-PQ composition the reference only writes down as a spec
                (docs/kernel-specs/DONE_22_adc_scan.md:831-881)

Semantics kept: ``k <= 0`` returns empty results (IVFIndex.swift:787,866); dimension mismatch raises
(:788-790); results ascend by API distance (L2: sqrt of L2^2 for Flat/IVF-Flat, ADC L2^2 for IVF-PQ as in
the composition; dot product: -dot, DistanceUtils.swift:40-46); an un-optimised IVF index answers by
linear scan (:820-832); ``nlist`` is clamped to the number of training vectors (:320); ties go to the
smaller id (TopK.swift:8-31).  Ids are integers in [0, 2^32 - 1) (the reference's TopK id type is Int32).
The metadata filter of the reference is a host closure and cannot run on the device; the only filter the
kernels honour is the list-disable bitmask of ivf_select_nprobe (IVFSelect.swift:366-395).

Multi-GPU: ``ShardedIVFPQIndex`` partitions the database by vector range over the ranks of a
``torch.distributed`` process group; queries are replicated, every rank scans its shard, and the per-rank
top-k lists are merged with ONE all-gather + the mergeTopK kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (INDEX_FLAT, INDEX_IVF_FLAT, INDEX_IVF_PQ, IndexParams, KMeansCfg, METRIC_IP, METRIC_L2, PQTrainCfg,
                   SearchStats, VectorIndexError, as_input, check, empty_like_input, lib, ptr)

_METRICS = {"euclidean": METRIC_L2, "l2": METRIC_L2, METRIC_L2: METRIC_L2,
            "dotProduct": METRIC_IP, "dot": METRIC_IP, "ip": METRIC_IP, METRIC_IP: METRIC_IP}


def _metric(m):
    if m not in _METRICS:
        raise VectorIndexError(-5, f"unsupported metric {m!r}: the GPU hot path serves euclidean and dotProduct")
    return _METRICS[m]


class _Index:
    kind = INDEX_FLAT

    def __init__(self, dimension: int, metric="euclidean", nlist: int = 256, nprobe: int = 8, m: int = 16,
                 ks: int = 256):
        p = IndexParams()
        lib().vix_index_params_default(C.byref(p))
        p.kind, p.d, p.metric = self.kind, int(dimension), _metric(metric)
        p.nlist, p.nprobe, p.m, p.ks = int(nlist), int(nprobe), int(m), int(ks)
        self.params = p
        self.dimension = int(dimension)
        self.metric = p.metric
        self._h = C.c_void_p(0)
        check(lib().vix_index_create(C.byref(p), C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h:
                lib().vix_index_destroy(self._h)
                self._h = C.c_void_p(0)
        except Exception:
            pass

    close = __del__

    # ---- VectorIndexProtocol ----
    @property
    def count(self) -> int:
        return int(lib().vix_index_count(self._h))

    def _check_dim(self, a, what):
        if a.ndim != 2 or int(a.shape[1]) != self.dimension:
            raise VectorIndexError(-1, f"{what}: expected [n x {self.dimension}], got {tuple(a.shape)}")

    def batch_insert(self, vectors, ids=None):
        x = as_input(vectors, np.float32)
        self._check_dim(x, "batch_insert")
        idp = as_input(ids, np.int64)
        check(lib().vix_index_add(self._h, ptr(x, np.float32), ptr(idp), C.c_int64(int(x.shape[0]))))

    def insert(self, id_: int, vector):
        v = np.ascontiguousarray(vector, dtype=np.float32).reshape(1, -1)
        self.batch_insert(v, np.array([id_], dtype=np.int64))

    add = batch_insert

    def clear(self):
        check(lib().vix_index_clear(self._h))

    def batch_search(self, queries, k: int, nprobe: int = 0, return_probes=False, stats=False):
        q = as_input(queries, np.float32)
        self._check_dim(q, "batch_search")
        nq = int(q.shape[0])
        kk = max(int(k), 0)
        dist = empty_like_input(q, (nq, kk), np.float32)
        ids = empty_like_input(q, (nq, kk), np.int64)
        if kk == 0 or nq == 0:
            return (dist, ids)
        npb = nprobe if nprobe > 0 else self.params.nprobe
        probes = empty_like_input(q, (nq, npb), np.int32) if return_probes else None
        st = SearchStats() if stats else None
        check(lib().vix_index_search_ex(self._h, ptr(q, np.float32), C.c_int64(nq), C.c_int(k), C.c_int(nprobe),
                                        ptr(dist, np.float32), ptr(ids, np.int64), ptr(probes),
                                        C.byref(st) if st is not None else None))
        out = (dist, ids)
        if return_probes:
            out += (probes,)
        if stats:
            out += (st,)
        return out

    def search(self, query, k: int, nprobe: int = 0):
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(1, -1)
        d, i = self.batch_search(q, k, nprobe)
        valid = i[0] >= 0
        return [(int(a), float(b)) for a, b in zip(i[0][valid], d[0][valid])]


class FlatIndex(_Index):
    kind = INDEX_FLAT

    def optimize(self):
        return None


class IVFIndex(_Index):
    """IVF-Flat.  ``optimize(training_vectors)`` trains the coarse quantiser (k-means) -- unlike the
    reference actor, vectors are added AFTER training (``batch_insert``), faiss-style, because the
    device index keeps lists, not a dictionary."""
    kind = INDEX_IVF_FLAT

    def optimize(self, training_vectors, kmeans_cfg: KMeansCfg | None = None, pq_cfg: PQTrainCfg | None = None):
        x = as_input(training_vectors, np.float32)
        self._check_dim(x, "optimize")
        if kmeans_cfg is None:
            kmeans_cfg = KMeansCfg(1024, 10, 1e-4, 42, 0, False, 1)
        if pq_cfg is None:
            pq_cfg = PQTrainCfg(0, 25, 1e-4, 1024, 0, 42, 0, 0, 1)
        check(lib().vix_index_train(self._h, ptr(x, np.float32), C.c_int64(int(x.shape[0])), C.byref(kmeans_cfg),
                                    C.byref(pq_cfg)))

    train = optimize

    def set_coarse(self, centroids):
        c = as_input(centroids, np.float32)
        self._check_dim(c, "set_coarse")
        check(lib().vix_index_set_coarse(self._h, ptr(c, np.float32), C.c_int(int(c.shape[0]))))

    def get_coarse(self):
        kc = C.c_int(0)
        check(lib().vix_index_get_coarse(self._h, None, C.byref(kc)))
        out = np.empty((kc.value, self.dimension), dtype=np.float32)
        check(lib().vix_index_get_coarse(self._h, ptr(out), C.byref(kc)))
        return out

    def list_sizes(self):
        kc = C.c_int(0)
        check(lib().vix_index_get_coarse(self._h, None, C.byref(kc)))
        out = np.empty(kc.value, dtype=np.int64)
        check(lib().vix_index_list_sizes(self._h, ptr(out)))
        return out


class IVFPQIndex(IVFIndex):
    kind = INDEX_IVF_PQ

    def set_codebooks(self, codebooks, centroid_norms=None):
        cb = as_input(codebooks, np.float32)
        cn = as_input(centroid_norms, np.float32)
        check(lib().vix_index_set_codebooks(self._h, ptr(cb, np.float32), ptr(cn)))

    def get_codebooks(self):
        m, ks, dsub = self.params.m, self.params.ks, self.dimension // self.params.m
        cb = np.empty((m, ks, dsub), dtype=np.float32)
        cn = np.empty((m, ks), dtype=np.float32)
        check(lib().vix_index_get_codebooks(self._h, ptr(cb), ptr(cn)))
        return cb, cn

    def import_lists(self, list_offsets, codes, ids):
        lo = np.ascontiguousarray(list_offsets, dtype=np.int64)
        codes = as_input(codes, np.uint8)
        ids = as_input(ids, np.int64)
        check(lib().vix_index_import_lists(self._h, ptr(lo), ptr(codes, np.uint8), ptr(ids, np.int64)))

    def export_lists(self):
        """CSR lists in the reference's AoS list format (Kernels/IVFAppend.swift:220-236) + assignments."""
        n = self.count
        kc = C.c_int(0)
        check(lib().vix_index_get_coarse(self._h, None, C.byref(kc)))
        off = np.empty(kc.value + 1, dtype=np.int64)
        codes = np.empty((n, self.params.m), dtype=np.uint8)
        ids = np.empty(n, dtype=np.int64)
        asg = np.empty(n, dtype=np.int32)
        check(lib().vix_index_export_lists(self._h, ptr(off), ptr(codes), ptr(ids), ptr(asg)))
        return off, codes, ids, asg


# ------------------------------------------------------------------------------------------------
# multi-GPU
# ------------------------------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int):
    """Contiguous vector range of ``rank`` (flat rows / vector ranges shard independently)."""
    per = (n + world - 1) // world
    b = min(n, rank * per)
    return b, min(n, b + per)


def merge_shard_results(dist_all, ids_all, k, metric=METRIC_L2):
    """mergeTopK over per-rank results.  dist_all/ids_all: [world x nq x k] (numpy or torch).  API distances
    ascend for both metrics, so the merge order is always .min with ties -> smaller id (TopKMerge.swift:66-71)."""
    from .kernels import mergeTopK
    if _lib._is_torch(dist_all):
        sc = dist_all.permute(1, 0, 2).contiguous()
        idm = ids_all.permute(1, 0, 2).contiguous()
    else:
        sc = np.ascontiguousarray(np.transpose(dist_all, (1, 0, 2)))
        idm = np.ascontiguousarray(np.transpose(ids_all, (1, 0, 2)))
    return mergeTopK(sc, idm, k, 0)


class ShardedIVFPQIndex:
    """One IVFPQIndex per rank holding a vector range; search = local fused scan + all-gather + merge."""

    def __init__(self, dimension, metric="euclidean", nlist=256, nprobe=8, m=16, ks=256, group=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = IVFPQIndex(dimension, metric, nlist, nprobe, m, ks)

    def set_parameters(self, coarse, codebooks, centroid_norms=None):
        self.local.set_coarse(coarse)
        self.local.set_codebooks(codebooks, centroid_norms)

    def add_global(self, vectors, ids=None):
        """Every rank passes the SAME global array; each keeps its own range."""
        n = int(vectors.shape[0])
        b, e = shard_range(n, self.rank, self.world)
        idl = ids[b:e] if ids is not None else np.arange(b, e, dtype=np.int64)
        if _lib._is_torch(vectors) and not _lib._is_torch(idl):
            import torch
            idl = torch.as_tensor(idl, device=vectors.device)
        self.local.batch_insert(vectors[b:e], idl)

    @classmethod
    def wrap(cls, local_index, group=None):
        """Sharded view over an already built per-rank IVFPQIndex."""
        import torch.distributed as dist
        self = cls.__new__(cls)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = local_index
        return self

    def batch_search(self, queries, k, nprobe=0):
        """Replicated queries -> local fused scan -> ONE all-gather of the packed per-rank (distance, id)
        lists -> mergeTopK kernel.  Host (numpy) queries are staged to the device once and only the merged
        [nq x k] result returns to the host."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.local.batch_search(queries, k, nprobe)
        was_numpy = not _lib._is_torch(queries)
        nccl = dist.get_backend(self.group) == "nccl"
        if was_numpy and nccl:
            dev = torch.device("cuda", torch.cuda.current_device())
            queries = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(dev, non_blocking=True)
        d_loc, i_loc = self.local.batch_search(queries, k, nprobe)
        md, mi = self.gather_merge(d_loc, i_loc, k)
        if was_numpy and _lib._is_torch(md):
            md, mi = md.cpu().numpy(), mi.cpu().numpy()
        return md, mi

    def gather_merge(self, d_loc, i_loc, k):
        """all-gather + merge of per-rank [nq x k] results (torch tensors or numpy arrays)."""
        import torch
        import torch.distributed as dist
        if not _lib._is_torch(d_loc):
            d_loc, i_loc = torch.from_numpy(d_loc), torch.from_numpy(i_loc)
            if dist.get_backend(self.group) == "nccl":
                dev = torch.device("cuda", torch.cuda.current_device())
                d_loc, i_loc = d_loc.to(dev), i_loc.to(dev)
        nq = d_loc.shape[0]
        d_all = torch.empty((self.world, nq, k), dtype=d_loc.dtype, device=d_loc.device)
        i_all = torch.empty((self.world, nq, k), dtype=i_loc.dtype, device=i_loc.device)
        dist.all_gather_into_tensor(d_all, d_loc.contiguous(), group=self.group)
        dist.all_gather_into_tensor(i_all, i_loc.contiguous(), group=self.group)
        if d_all.is_cuda:
            return merge_shard_results(d_all, i_all, k)
        return merge_shard_results_host(d_all.numpy(), i_all.numpy(), k)


def merge_shard_results_host(dist_all, ids_all, k):
    """Host restatement of the shard merge (used on CPU-only ranks and by the gloo tests): k smallest of
    the union by (distance, id); NaN / id -1 entries are padding."""
    world, nq, kk = dist_all.shape
    d = np.transpose(dist_all, (1, 0, 2)).reshape(nq, world * kk)
    i = np.transpose(ids_all, (1, 0, 2)).reshape(nq, world * kk)
    out_d = np.full((nq, k), np.nan, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        valid = i[r] >= 0
        dv, iv = d[r][valid], i[r][valid]
        order = np.lexsort((iv, dv))[:k]
        out_d[r, :order.size] = dv[order]
        out_i[r, :order.size] = iv[order]
    return out_d, out_i
