"""Seeded INPUTS of the committed golden fixtures (tests/golden/*.npz hold only the outputs; scripts/make_golden.py
wrote them).  Everything here is plain numpy with fixed seeds, so the inputs are identical wherever the tests run."""
import numpy as np

from conftest import parity_fixture


def encoder_problems():
    """name -> (x [n x d], codebooks ks=256, codebooks ks=16, coarse [kc x d], assign [n], m)"""
    out = {}
    # the reference's own fixture (PQEncodeParity_AoS_C_vs_Swift_Tests.swift:5-31), u8 and u4 codebooks by the same formula
    x, cb8, coarse, assign = parity_fixture(16, 32, 8, 256, 4)
    _, cb4, _, _ = parity_fixture(16, 32, 8, 16, 4)
    out["parity"] = (x, cb8, cb4, coarse, assign, 8)
    rng = np.random.default_rng(20261018)
    n, d, m, kc = 257, 48, 6, 5
    x = rng.standard_normal((n, d)).astype(np.float32)
    cb8 = rng.standard_normal(m * 256 * (d // m)).astype(np.float32)
    cb4 = rng.standard_normal(m * 16 * (d // m)).astype(np.float32)
    coarse = (rng.standard_normal((kc, d)) * 0.5).astype(np.float32)
    assign = rng.integers(0, kc, n).astype(np.int32)
    out["random"] = (x, cb8, cb4, coarse, assign, m)
    return out


def search_problem():
    """small SIFT-shaped IVF-PQ problem: non-negative integer-valued components (many exact ties)"""
    rng = np.random.default_rng(96)
    n, nq, d, m, kc = 3000, 24, 32, 8, 24
    centres = np.floor(np.abs(rng.standard_normal((40, d))) * 40).clip(0, 218)
    which = rng.integers(0, 40, n + nq)
    x = np.floor(np.abs(centres[which] + 12.0 * rng.standard_normal((n + nq, d)))).clip(0, 218).astype(np.float32)
    coarse = x[rng.choice(n, kc, replace=False)].copy()
    cb = (rng.standard_normal((m, 256, d // m)) * 14.0).astype(np.float32)
    return dict(xb=np.ascontiguousarray(x[:n]), q=np.ascontiguousarray(x[n:]), coarse=coarse, cb=cb, m=m, kc=kc, nprobe=5, k=10)
