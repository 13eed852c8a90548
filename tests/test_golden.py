"""Committed golden vectors (tests/golden/, written by scripts/make_golden.py from the seeded inputs of
tests/golden_inputs.py).

cpq_encode_reference.npz are outputs of the REFERENCE's own C encoder (pq_encode.c compiled unmodified): the oracle
restatement (CPU) and the CUDA library (GPU, through the reference's cpq_* symbols) must reproduce them bit for bit.
oracle_search_small.npz freezes the oracle's output of every stage of the search path (no reference test pins LUT / ADC
arithmetic); the GPU tests compare the C-ABI entry points with it."""
import os

import numpy as np
import pytest

import golden_inputs as gi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = ["u8", "u8_nodot", "u8_csq", "res", "res_nodot", "res_csq", "u4", "res_u4"]


def _enc():
    return np.load(os.path.join(GOLD, "cpq_encode_reference.npz"))


def _srch():
    return np.load(os.path.join(GOLD, "oracle_search_small.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------ CPU: oracle vs golden
@pytest.mark.parametrize("problem", ["parity", "random"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_reproduces_reference_encoder_golden(oracle, problem, variant):
    x, cb8, cb4, coarse, assign, m = gi.encoder_problems()[problem]
    gold = _enc()
    csq = gold[f"{problem}.csq"]
    assert np.array_equal(bits(oracle.pq_centroid_sq(cb8, m, 256, x.shape[1] // m, swift=True)), bits(csq))
    res = variant.startswith("res")
    kw = dict(coarse=coarse, assign_=assign) if res else {}
    if variant.endswith("u4"):
        got = oracle.pq_encode_u4(x, cb4, m, 16, **kw)
    elif variant.endswith("csq"):
        got = oracle.pq_encode_u8(x, cb8, m, 256, centroid_sq=csq, **kw)
    else:
        got = oracle.pq_encode_u8(x, cb8, m, 256, use_dot=not variant.endswith("nodot"), **kw)
    assert np.array_equal(got, gold[f"{problem}.{variant}"])


def test_first_row_of_reference_fixture_pin():
    """first row of the n=16, d=32, m=8 fixture through the reference C path."""
    assert _enc()["parity.u8"][0].tolist() == [212, 186, 160, 117, 255, 154, 186, 249]


def test_oracle_reproduces_search_golden(oracle):
    P, g = gi.search_problem(), _srch()
    xb, q, coarse, cb, m, kc, nprobe, k = P["xb"], P["q"], P["coarse"], P["cb"], P["m"], P["kc"], P["nprobe"], P["k"]
    asg, adist = oracle.assign(xb, coarse)
    assert np.array_equal(asg, g["assign"]) and np.array_equal(bits(adist), bits(g["assign_dist"]))
    codes = oracle.pq_encode_u8(xb, cb, m, 256, centroid_sq=g["cb_norms"].reshape(-1), coarse=coarse, assign_=asg)
    assert np.array_equal(codes, g["codes"])
    pid, psc = oracle.probe_select_batch(q, coarse, nprobe, 0, oracle.centroid_norms(coarse))
    assert np.array_equal(pid, g["probe_ids"]) and np.array_equal(bits(psc), bits(g["probe_scores"]))
    fd, fi, _ = oracle.flat_search(q, xb, k, 0)
    assert np.array_equal(fi, g["flat_ids"]) and np.array_equal(bits(fd), bits(g["flat_dist"]))
    cd, ci, _ = oracle.flat_search(q, xb, k, 2)
    assert np.array_equal(ci, g["cosine_ids"]) and np.array_equal(bits(cd), bits(g["cosine_dist"]))


def _cosine_coarse(P):
    cz = P["coarse"].copy()
    cz[5] = 0.0                                                        # the degenerate centroid of the golden rows
    return cz


def test_oracle_reproduces_cosine_golden(oracle):
    """the cosine rows of the golden file: guarded CentroidBatchScore block, probe lists, list assignment and IVF-Flat
    results under the cosine metric (one centroid degenerate: score exactly 1)."""
    P, g = gi.search_problem(), _srch()
    xb, q, kc, nprobe, k = P["xb"], P["q"], P["kc"], P["nprobe"], P["k"]
    cz = _cosine_coarse(P)
    sc = oracle.centroid_batch_score(q, cz, 2)
    assert np.array_equal(bits(sc), bits(g["cbs_cosine"])) and (sc[:, 5] == 1.0).all()
    assert np.array_equal(oracle.probe_select_batch(q, cz, nprobe, 2)[0], g["probe_ids_cosine"])
    asg = oracle.assign_metric(xb, cz, 2)
    assert np.array_equal(asg, g["assign_cosine"])
    off, order = oracle.build_lists(asg, kc)
    ids = np.arange(xb.shape[0], dtype=np.int64)
    dd, ii = oracle.ivfflat_search(q, cz, off, xb[order], ids[order], nprobe, k, 2)
    assert np.array_equal(ii, g["ivfflat_cosine_ids"]) and np.array_equal(bits(dd), bits(g["ivfflat_cosine_dist"]))


def test_oracle_cosine_against_float64(oracle):
    """Cosine.run two-pass restatement vs a float64 evaluation of 1 - <q, x> / (|q| |x|): same ranking away from
    near-ties, distances within fp32 rounding; a zero row has similarity 0 (distance 1)."""
    rng = np.random.default_rng(12)
    xb = (rng.standard_normal((700, 40)) * rng.uniform(0.2, 4.0, (700, 1))).astype(np.float32)
    xb[5] = 0.0
    q = rng.standard_normal((9, 40)).astype(np.float32)
    d, i, _ = oracle.flat_search(q, xb, 700, 2)
    x64, q64 = xb.astype(np.float64), q.astype(np.float64)
    nx = np.linalg.norm(x64, axis=1)
    sim = (q64 @ x64.T) / np.linalg.norm(q64, axis=1)[:, None] / np.where(nx > 0, nx, 1.0)[None]
    ref = 1.0 - sim
    for r in range(q.shape[0]):
        assert np.allclose(d[r], ref[r][i[r]], atol=3e-6)
        assert (np.diff(d[r]) >= 0).all()
        assert d[r][np.where(i[r] == 5)[0][0]] == np.float32(1.0)


def test_oracle_ivfflat_cosine_against_float64(oracle):
    """IVF-Flat under cosine (IVFIndex.swift:376-435, 905-927; DistanceUtils.swift:22-38): the candidates of the probed
    lists ranked by a float64 evaluation of 1 - <q, x> / (|q| |x|) give the same ids and distances within fp32 rounding."""
    rng = np.random.default_rng(5)
    n, d, kc, nq, k, nprobe = 3000, 48, 24, 20, 10, 6
    x = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3, (n, 1))).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    coarse = x[rng.choice(n, kc, replace=False)].copy()
    asg = oracle.assign_metric(x, coarse, 2)
    off, order = oracle.build_lists(asg, kc)
    ids = np.arange(n, dtype=np.int64)
    dd, ii = oracle.ivfflat_search(q, coarse, off, x[order], ids[order], nprobe, k, 2)
    pid, _ = oracle.probe_select_batch(q, coarse, nprobe, 2)
    for r in range(nq):
        cand = np.concatenate([order[off[l]:off[l + 1]] for l in pid[r]])
        x64, q64 = x[cand].astype(np.float64), q[r].astype(np.float64)
        dist = 1 - (x64 @ q64) / np.linalg.norm(x64, axis=1) / np.linalg.norm(q64)
        o = np.argsort(dist, kind="stable")[:k]
        assert set(cand[o].tolist()) == set(ii[r].tolist())
        assert np.allclose(dist[o], dd[r], atol=3e-6)


# ------------------------------------------------------------------------------------------ GPU: C ABI vs golden
@pytest.mark.gpu
@pytest.mark.parametrize("problem", ["parity", "random"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_cuda_encoder_reproduces_reference_encoder_golden(vk, problem, variant):
    from vectorindex_b200._lib import PQEncodeOpts
    x, cb8, cb4, coarse, assign, m = gi.encoder_problems()[problem]
    gold = _enc()
    csq = gold[f"{problem}.csq"]
    nodot = PQEncodeOpts(0, False, False, 8, 0, 0, 0)
    if variant == "u8":
        got = vk.pq_encode_u8_f32(x, cb8, m, 256)
    elif variant == "u8_nodot":
        got = vk.pq_encode_u8_f32(x, cb8, m, 256, nodot)
    elif variant == "u8_csq":
        got = vk.pq_encode_u8_f32_withCSQ(x, cb8, csq, m, 256)
    elif variant == "res":
        got = vk.pq_encode_residual_u8_f32(x, cb8, coarse, assign, m, 256)
    elif variant == "res_nodot":
        got = vk.pq_encode_residual_u8_f32(x, cb8, coarse, assign, m, 256, nodot)
    elif variant == "res_csq":
        got = vk.pq_encode_residual_u8_f32_withCSQ(x, cb8, csq, coarse, assign, m, 256)
    elif variant == "u4":
        got = vk.pq_encode_u4_f32(x, cb4, m, 16)
    else:
        got = vk.pq_encode_residual_u4_f32(x, cb4, coarse, assign, m, 16)
    assert np.array_equal(np.asarray(got).reshape(gold[f"{problem}.{variant}"].shape), gold[f"{problem}.{variant}"])


@pytest.mark.gpu
def test_cuda_search_stages_reproduce_golden(vk):
    from vectorindex_b200.index import IVFPQIndex
    P, g = gi.search_problem(), _srch()
    xb, q, coarse, cb, m, kc, nprobe, k = P["xb"], P["q"], P["coarse"], P["cb"], P["m"], P["kc"], P["nprobe"], P["k"]
    d = xb.shape[1]
    norms = g["cb_norms"]
    asg, adist = vk.ivf_assign_f32(xb, coarse, return_dist=True)
    assert np.array_equal(asg, g["assign"]) and np.array_equal(bits(adist), bits(g["assign_dist"]))
    codes = vk.pq_encode_residual_u8_f32_withCSQ(xb, cb.reshape(-1), norms.reshape(-1), coarse, asg, m, 256)
    assert np.array_equal(np.asarray(codes).reshape(-1, m), g["codes"])
    cn = vk.row_norms_f32(coarse)
    pid, psc = vk.ivf_select_nprobe_batch_f32(q, coarse, nprobe, 0, None)
    assert np.array_equal(pid, g["probe_ids"])                      # kernel-level scores use the vDSP form: ids only
    # residual LUTs of the first probe + ADC over that list: bit-exact (PQLUT.swift:266-386, ADCScan.swift:190-283)
    luts = vk.pq_lut_residual_l2_f32(q[:8], g["probe_ids"][:8, 0].astype(np.int32), coarse, cb, m, 256, norms)
    assert np.array_equal(bits(luts), bits(g["lut_first_probe"]))
    order = np.argsort(g["assign"], kind="stable")
    off = np.concatenate([[0], np.cumsum(np.bincount(g["assign"], minlength=kc))])
    pos = 0
    for r in range(8):
        l = int(g["probe_ids"][r, 0])
        rows = order[off[l]:off[l + 1]]
        n = int(g["adc_first_probe_len"][r])
        assert rows.size == n
        if n:
            out = vk.adc_scan_u8(g["codes"][rows], g["lut_first_probe"][r], m, 256)
            assert np.array_equal(bits(out), bits(g["adc_first_probe"][pos:pos + n]))
        pos += n
    fd, fi = vk.flat_search_f32(q, xb, k, 0)
    assert np.array_equal(fi, g["flat_ids"]) and np.array_equal(bits(fd), bits(g["flat_dist"]))
    cd, ci = vk.flat_search_f32(q, xb, k, 2)
    assert np.array_equal(ci, g["cosine_ids"]) and np.array_equal(bits(cd), bits(g["cosine_dist"]))
    # the index: same lists, probe lists bit-exact, fused-scan distances within 1e-5 (re-associated sum), id sets equal
    # except at ties inside that tolerance
    idx = IVFPQIndex(d, "euclidean", nlist=kc, nprobe=nprobe, m=m)
    idx.set_coarse(coarse)
    idx.set_codebooks(cb, norms)
    idx.batch_insert(xb)
    gd, gi_, gp = idx.batch_search(q, k, return_probes=True)
    assert np.array_equal(gp, g["probe_ids"])
    np.testing.assert_allclose(gd, g["ivfpq_dist"], rtol=1e-5)
    for r in range(q.shape[0]):
        extra = set(gi_[r].tolist()) ^ set(g["ivfpq_ids"][r].tolist())
        if extra:                                                    # only entries tied with the k-th distance may differ
            kth = g["ivfpq_dist"][r, -1]
            both = {int(i): float(v) for i, v in zip(g["ivfpq_ids"][r], g["ivfpq_dist"][r])}
            both.update({int(i): float(v) for i, v in zip(gi_[r], gd[r])})
            assert all(abs(both[i] - kth) <= 1e-5 * abs(kth) for i in extra)


@pytest.mark.gpu
def test_cuda_cosine_rows_reproduce_golden(vk):
    from vectorindex_b200.index import IVFIndex
    P, g = gi.search_problem(), _srch()
    xb, q, kc, nprobe, k = P["xb"], P["q"], P["kc"], P["nprobe"], P["k"]
    cz = _cosine_coarse(P)
    assert np.array_equal(bits(vk.centroid_batch_score(q, cz, 2)), bits(g["cbs_cosine"]))
    ivf = IVFIndex(xb.shape[1], "cosine", nlist=kc, nprobe=nprobe)
    ivf.set_coarse(cz)
    ivf.batch_insert(xb)
    assert np.array_equal(ivf.list_sizes(), np.bincount(g["assign_cosine"], minlength=kc))
    gd, gi_, gp = ivf.batch_search(q, k, return_probes=True)
    assert np.array_equal(gp, g["probe_ids_cosine"])
    assert np.array_equal(gi_, g["ivfflat_cosine_ids"])
    np.testing.assert_allclose(gd, g["ivfflat_cosine_dist"], rtol=1e-5, atol=1e-6)
