"""Host-side logic that needs no device: the IDFilter bitset (IDFilterBitset / idFilterPass semantics of
Operations/Filtering/IDFilter.swift:13-135) and the metric / argument mapping of the index mirror."""
import numpy as np
import pytest

from vectorindex_b200.index import IDFilter, compose_id_filters


def _pass_reference(words, capacity, deny, i):
    """idFilterPass, restated: out-of-range ids never pass; allowlist keeps set bits, denylist keeps clear bits."""
    if i < 0 or i >= capacity:
        return False
    bit = (int(words[i >> 6]) >> (i & 63)) & 1 == 1
    return (not bit) if deny else bit


@pytest.mark.parametrize("mode", ["allow", "deny"])
def test_idfilter_matches_idfilterpass(mode):
    rng = np.random.default_rng(4)
    cap = 1000                                                      # not a multiple of 64: the last word is partial
    f = IDFilter(cap, mode)
    assert f.words.dtype == np.uint64 and f.words.size == (cap + 63) // 64
    chosen = rng.choice(cap, 300, replace=False)
    f.set(chosen)
    f.set([-5, cap, cap + 70])                                       # out of range: ignored (IDFilter.swift:52-56)
    f.set(chosen[:50], value=False)
    probe = np.concatenate([np.arange(-3, cap + 130), [2 ** 40]])
    want = np.array([_pass_reference(f.words, cap, mode == "deny", int(i)) for i in probe])
    assert np.array_equal(f.test(probe), want)
    kept = set(chosen[50:].tolist())
    inside = probe[(probe >= 0) & (probe < cap)]
    assert np.array_equal(f.test(inside), np.array([(i in kept) != (mode == "deny") for i in inside]))


def test_idfilter_initial_bit_and_empty():
    f = IDFilter(130, "allow", initial_bit=True)
    assert f.test(np.arange(130)).all() and not f.test(np.array([130, 191])).any()
    g = IDFilter(0, "deny")
    assert g.words.size == 0 and not g.test(np.array([0, 1])).any()


def test_metric_names_map_to_the_abi_values():
    from vectorindex_b200.index import _metric
    assert (_metric("euclidean"), _metric("dotProduct"), _metric("cosine")) == (0, 1, 2)
    with pytest.raises(Exception):
        _metric("manhattan")


def test_compose_id_filters_matches_idfilterpassn():
    """keep = (allow0 AND ... AND allow3) AND NOT deny, out-of-range ids never pass (IDFilter.swift:140-176)."""
    rng = np.random.default_rng(9)
    cap = 777
    allows = [IDFilter(cap, "allow").set(rng.choice(cap, 500, replace=False)) for _ in range(3)]
    deny = IDFilter(cap, "deny").set(rng.choice(cap, 200, replace=False))
    f = compose_id_filters(allows, deny)
    assert f.mode == IDFilter.ALLOW and f.capacity == cap
    ids = np.arange(-2, cap + 100)
    want = np.ones(ids.size, dtype=bool)
    for a in allows:
        want &= a.test(ids)
    want &= deny.test(ids)                                           # a denylist passes ids whose bit is clear
    assert np.array_equal(f.test(ids), want)
    assert not (int(f.words[-1]) >> (cap & 63))                      # nothing set past the domain
    # no allowlist: everything in range starts allowed; nil filters are skipped
    only_deny = compose_id_filters([None], deny)
    assert np.array_equal(only_deny.test(ids), deny.test(ids))
    assert compose_id_filters(capacity=10).test(np.arange(12)).tolist() == [True] * 10 + [False] * 2
    with pytest.raises(ValueError):
        compose_id_filters([allows[0]] * 5)
    with pytest.raises(ValueError):
        compose_id_filters([IDFilter(cap + 1, "allow")], deny)


@pytest.mark.parametrize("pipes", [1, 2])
def test_scan_table_address_algebra(pipes):
    """The index arithmetic of the fused scan's look-up tables (vix_ivfpq_scan.cu: build_lut's column formula, the per-lane
    constants and LDS immediates of lookup16, the rotated code layout), restated on integers for both table layouts: every
    look-up reads the entry that was written for its (sub-quantiser, code), a warp-wide look-up touches 32 different
    banks whatever the codes are, and the two pipelines of the two-pipeline layout never share an entry."""
    k_dual_tab = 36 * 1024
    tab_abs = 128 * 1024 if pipes == 1 else k_dual_tab               # one pipeline: any 64 KB-aligned shared address
    G = 3                                                            # m = 48
    rng = np.random.default_rng(0)

    def written(pipe, j, c, rep):                                    # build_lut: byte address of T[j][c], replica rep
        t16, jj = j >> 4, j & 15
        col = ((t16 >> 1) * 16384 + (t16 & 1) * 32) if pipes == 1 else (t16 * 16384 + pipe * 32)
        return tab_abs + 4 * (col + jj + 16 * rep + c * 64)

    def looked_up(pipe, lane, T, b, code):                           # lookup16<T>: PRMT(code word, lane constant) + immediate
        const = 4 * (16 * (lane >> 4) + ((b ^ lane) & 15)) + (128 * pipe if pipes == 2 else 0)
        assert const < 256
        base = tab_abs if pipes == 1 else 0                          # bytes 2..3 of the lane constant
        imm = ((T & 1) * 128 + (T >> 1) * 65536) if pipes == 1 else (k_dual_tab + T * 65536)
        return (base | (code << 8) | const) + imm

    owners = {}
    for pipe in range(pipes):
        for j in range(16 * G):
            for c in (0, 1, 77, 255):
                for rep in (0, 1):
                    a = written(pipe, j, c, rep)
                    assert owners.setdefault(a, (pipe, j, c, rep)) == (pipe, j, c, rep)   # no two entries share an address
        for T in range(G):
            for b in range(16):
                codes = rng.integers(0, 256, 32)
                addrs = []
                for lane in range(32):
                    j = 16 * T + ((b ^ lane) & 15)                   # byte b of slot `lane` holds this sub-quantiser (rotation)
                    a = looked_up(pipe, lane, T, b, int(codes[lane]))
                    assert a == written(pipe, j, int(codes[lane]), lane >> 4)
                    addrs.append(a)
                assert len({(a >> 2) & 31 for a in addrs}) == 32     # 32 lanes, 32 banks
    top = max(owners) + 4
    assert top <= (228 * 1024 if pipes == 2 else tab_abs + 2 * 65536)
