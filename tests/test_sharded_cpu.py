"""World-size-2 test of the multi-GPU host logic on CPU (gloo): list-block partitioning, the all-to-all build
routing, the two merged stages of the search (global probe lists, final top-k).  The per-rank index is an
oracle-backed stand-in with the interface of ``IVFPQIndex``'s sharding pieces, so the whole distributed
pipeline runs without a GPU and its result must equal the single-process oracle search exactly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleLocalIndex:
    """probe_range / search_with_probes / encode / add_encoded on the CPU oracle (test infrastructure)."""

    def __init__(self, d, m, metric=0):
        from oracle import oracle
        self.o, self.d, self.m, self.metric = oracle, d, m, metric
        self.assign = np.empty(0, np.int32)
        self.codes = np.empty((0, m), np.uint8)
        self.ids = np.empty(0, np.int64)

    def set_coarse(self, c):
        self.coarse = np.ascontiguousarray(c, np.float32)
        self.cnorms = self.o.centroid_norms(self.coarse)

    def set_codebooks(self, cb, norms=None):
        self.cb, self.norms = np.ascontiguousarray(cb, np.float32), norms

    def encode(self, x):
        x = np.ascontiguousarray(x, np.float32)
        asg, _ = self.o.assign(x, self.coarse)
        codes = self.o.pq_encode_u8(x, self.cb, self.m, 256, centroid_sq=self.norms.reshape(-1), coarse=self.coarse, assign_=asg)
        return asg, codes

    def add_encoded(self, assign, codes, ids):
        self.assign = np.concatenate([self.assign, np.asarray(assign, np.int32)])
        self.codes = np.concatenate([self.codes, np.asarray(codes, np.uint8).reshape(-1, self.m)])
        self.ids = np.concatenate([self.ids, np.asarray(ids, np.int64)])

    def probe_range(self, q, nprobe, begin, count):
        ids, sc = self.o.probe_select_batch(np.ascontiguousarray(q, np.float32), self.coarse[begin:begin + count], nprobe,
                                            self.metric, self.cnorms[begin:begin + count])
        ids = np.where(ids >= 0, ids + begin, ids).astype(np.int32)
        return ids, sc

    def search_with_probes(self, q, k, probes):
        o = self.o
        q = np.ascontiguousarray(q, np.float32)
        probes = np.asarray(probes)
        kc = self.coarse.shape[0]
        order = np.lexsort((self.ids, self.assign))               # lists in ascending-id order
        asg, codes, ids = self.assign[order], self.codes[order], self.ids[order]
        off = np.searchsorted(asg, np.arange(kc + 1))
        out_d = np.full((q.shape[0], k), np.nan, np.float32)
        out_i = np.full((q.shape[0], k), -1, np.int64)
        for r in range(q.shape[0]):
            cand_d, cand_i = [], []
            for l in probes[r]:
                if l < 0 or off[l + 1] == off[l]:
                    continue
                lut = o.pq_lut_residual_l2(q[r], self.coarse[l], self.cb, self.m, 256, cnorms=self.norms)
                dist = o.adc_scan_u8(codes[off[l]:off[l + 1]], lut, self.m)
                cand_d.append(dist)
                cand_i.append(ids[off[l]:off[l + 1]])
            if cand_d:
                dd, ii = np.concatenate(cand_d), np.concatenate(cand_i)
                sel = np.lexsort((ii, dd))[:k]
                out_d[r, :sel.size], out_i[r, :sel.size] = dd[sel], ii[sel]
        return out_d, out_i


def _problem():
    from oracle import oracle
    rng = np.random.default_rng(7)
    n, d, m, kc, nq = 3000, 32, 8, 24, 25
    x = (rng.standard_normal((n + nq, d)) + 2 * rng.standard_normal((12, d))[rng.integers(0, 12, n + nq)]).astype(np.float32)
    xb, q = np.ascontiguousarray(x[:n]), np.ascontiguousarray(x[n:])
    coarse = np.ascontiguousarray(xb[rng.choice(n, kc, replace=False)])
    asg, _ = oracle.assign(xb, coarse)
    rc, cb, norms, _ = oracle.pq_train(xb[:1500], m, 256, coarse=coarse, assign_=asg[:1500], max_iters=3, sample_n=0)
    assert rc == 0
    return xb, q, coarse, cb, norms, asg, m, kc


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vectorindex_b200.index import ShardedIVFPQIndex, list_block
        xb, q, coarse, cb, norms, asg, m, kc = _problem()
        nprobe, k = 5, 10
        sh = ShardedIVFPQIndex(xb.shape[1], "euclidean", nlist=kc, nprobe=nprobe, m=m, local=OracleLocalIndex(xb.shape[1], m))
        sh.set_parameters(coarse, cb, norms)
        from vectorindex_b200.index import balanced_list_bounds
        sh.set_list_bounds(balanced_list_bounds(np.bincount(asg, minlength=kc), world))
        # every rank hands in a different slice of the database, in two batches
        ids = np.arange(xb.shape[0], dtype=np.int64) * 2 + 1
        mine = np.arange(rank, xb.shape[0], world)
        half = mine.size // 2
        sh.add(xb[mine[:half]], ids[mine[:half]])
        sh.add(xb[mine[half:]], ids[mine[half:]])
        b, c = list_block(kc, rank, world, sh.bounds)
        held = sh.local.assign
        assert ((held >= b) & (held < b + c)).all()                       # only owned lists arrived here
        assert held.size == int(((asg >= b) & (asg < b + c)).sum())       # and all of their rows did
        probes = sh.global_probes(q, nprobe)
        md, mi = sh.batch_search(q, k)
        if rank == 0:
            np.savez(out_path, probes=np.asarray(probes), md=np.asarray(md), mi=np.asarray(mi))
    finally:
        dist.destroy_process_group()


def _replica_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vectorindex_b200.index import ReplicatedIVFPQIndex
        xb, q, coarse, cb, norms, asg, m, kc = _problem()
        nprobe, k = 5, 10
        rp = ReplicatedIVFPQIndex(xb.shape[1], "euclidean", nlist=kc, nprobe=nprobe, m=m, local=OracleLocalIndex(xb.shape[1], m))
        rp.set_parameters(coarse, cb, norms)
        ids = np.arange(xb.shape[0], dtype=np.int64) * 2 + 1
        cuts = [0, 1100, xb.shape[0]] if world == 2 else [0, 1100, 2200, xb.shape[0]]   # ragged contributions
        mine = np.arange(cuts[rank], cuts[rank + 1])
        half = mine.size // 3
        rp.add(xb[mine[:half]], ids[mine[:half]])
        rp.add(xb[mine[half:]], ids[mine[half:]])
        rp.add(xb[:0], ids[:0])                                           # nothing from anybody
        assert rp.local.ids.size == xb.shape[0]                           # every replica holds every row ...
        assert np.array_equal(np.sort(rp.local.ids), ids)                 # ... exactly once
        md, mi = rp.batch_search(q, k)                                    # 25 queries over 2 ranks: ragged blocks (13 + 12)
        bd, bi = rp.batch_search(q, k, gather=False)
        lo, cnt, _ = rp.query_block(q.shape[0])
        assert np.array_equal(np.asarray(bi), np.asarray(mi)[lo:lo + cnt])
        sd, si = rp.batch_search(q[:2], k)                                # fewer queries than ranks at world 3: an empty block
        assert np.array_equal(np.asarray(si), np.asarray(mi)[:2])
        assert np.array_equal(np.asarray(sd).view(np.uint32), np.asarray(md)[:2].view(np.uint32))
        np.savez(out_path + f".{rank}.npz", md=np.asarray(md), mi=np.asarray(mi), order=rp.local.ids)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_list_block_partition_covers_all_lists():
    from vectorindex_b200.index import list_block, list_owner
    for kc, world in ((65536, 8), (10, 4), (7, 8), (4096, 3)):
        blocks = [list_block(kc, r, world) for r in range(world)]
        assert sum(c for _, c in blocks) == kc
        owner = list_owner(np.arange(kc), kc, world)
        for r, (b, c) in enumerate(blocks):
            assert (owner[b:b + c] == r).all()


def test_balanced_list_bounds():
    import torch
    from vectorindex_b200.index import balanced_list_bounds, list_block, list_owner
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 3000, 4096)
    for world in (2, 3, 8):
        b = balanced_list_bounds(sizes, world)
        assert b[0] == 0 and b[-1] == sizes.size and (np.diff(b) >= 0).all() and b.size == world + 1
        work = np.array([float((sizes[b[r]:b[r + 1]].astype(np.float64) ** 2).sum()) for r in range(world)])
        assert work.max() / work.mean() < 1.01                       # equal-count blocks: several per cent
        ids = np.arange(sizes.size)
        own = list_owner(ids, sizes.size, world, b)
        assert np.array_equal(own, list_owner(torch.from_numpy(ids), sizes.size, world, b).numpy())
        for r in range(world):
            lo, cnt = list_block(sizes.size, r, world, b)
            assert (own[lo:lo + cnt] == r).all() and (own == r).sum() == cnt
    assert np.array_equal(balanced_list_bounds(np.zeros(10), 4)[[0, -1]], [0, 10])


def test_query_block_partition_covers_the_batch():
    from vectorindex_b200.index import ShardedIVFPQIndex
    for nq, world in ((10000, 8), (25, 2), (2, 3), (1, 8), (0, 4), (7, 7)):
        seen = []
        for r in range(world):
            sh = ShardedIVFPQIndex.__new__(ShardedIVFPQIndex)
            sh.rank, sh.world = r, world
            lo, cnt, per = sh.query_block(nq)
            assert 0 <= cnt <= per and per * world >= nq
            seen += list(range(lo, lo + cnt))
        assert seen == list(range(nq))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_equals_single_process_oracle(tmp_path, oracle, world):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = np.load(out)
    xb, q, coarse, cb, norms, asg, m, kc = _problem()
    ids = np.arange(xb.shape[0], dtype=np.int64) * 2 + 1
    off, order = oracle.build_lists(asg, kc)
    codes = oracle.pq_encode_u8(xb, cb, m, 256, centroid_sq=norms.reshape(-1), coarse=coarse, assign_=asg)
    od, oi, op = oracle.ivfpq_search(q, coarse, cb, norms, off, codes[order], ids[order], m, 256, 5, 10, 0)
    assert np.array_equal(res["probes"], op)                              # merged probe lists == single-GPU order
    assert np.array_equal(res["mi"], oi)
    assert np.array_equal(res["md"].view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("world", [2, 3])
def test_replicated_search_equals_single_process_oracle(tmp_path, oracle, world):
    """ReplicatedIVFPQIndex at world size 2 and 3: ragged build contributions, ragged and empty query blocks; every rank ends
    with the whole answer, bit-identical to the single-process oracle search, and the replicas hold their rows in the same
    order."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "rep")
    mp.spawn(_replica_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    xb, q, coarse, cb, norms, asg, m, kc = _problem()
    ids = np.arange(xb.shape[0], dtype=np.int64) * 2 + 1
    off, order = oracle.build_lists(asg, kc)
    codes = oracle.pq_encode_u8(xb, cb, m, 256, centroid_sq=norms.reshape(-1), coarse=coarse, assign_=asg)
    od, oi, _ = oracle.ivfpq_search(q, coarse, cb, norms, off, codes[order], ids[order], m, 256, 5, 10, 0)
    res = [np.load(out + f".{r}.npz") for r in range(world)]
    for r in res:
        assert np.array_equal(r["mi"], oi)
        assert np.array_equal(r["md"].view(np.uint32), od.view(np.uint32))
        assert np.array_equal(res[0]["order"], r["order"])


def test_host_merge_matches_definition():
    from vectorindex_b200.index import merge_shard_results_host
    rng = np.random.default_rng(3)
    world, nq, k = 3, 7, 5
    d = np.sort(np.round(rng.standard_normal((world, nq, k)) * 2).astype(np.float32), axis=2)
    i = rng.permutation(world * nq * k).reshape(world, nq, k).astype(np.int64)
    d[1, :, 3:] = np.nan
    i[1, :, 3:] = -1
    md, mi = merge_shard_results_host(d, i, k)
    for r in range(nq):
        pairs = sorted((float(d[w, r, t]), int(i[w, r, t])) for w in range(world) for t in range(k) if i[w, r, t] >= 0)[:k]
        assert [p[1] for p in pairs] == mi[r].tolist()
        assert [p[0] for p in pairs] == md[r].tolist()
