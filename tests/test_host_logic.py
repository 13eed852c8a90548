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
