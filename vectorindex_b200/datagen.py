"""Synthetic data generators (host side, numpy) used by tests and bench.py.

They restate the reference benchmark's generators so that the inputs of BASELINE.json's configs are
the reference's own recipe:

* ``lcg_stream`` / ``lcg_unit_floats``  -- ``DataGen.LCG`` of
  /root/reference/Sources/VectorIndexBenchmarks/main.swift:535-548
  (``s = 2862933555777941757*s + 3037000493 mod 2^64``; ``f = Float(s>>11)/Float(2^53)``).
* ``bench_vectors``      -- ``DataGen.generate`` (component ``2f-1``, sequential-sum L2 normalise).
* ``gaussian_lcg``       -- ``GaussianLCG`` (Box-Muller on the same LCG), main.swift:108-121.
* ``sift_like`` / ``clustered_unit``  -- the C3 "SIFT-shaped" and C4/C5 "embedding/Deep-shaped"
  mixtures of SURVEY.md section 8(d) (recipe of main.swift:129-144 for the clustered case).

The LCG is evaluated with wrap-around uint64 numpy arithmetic in jump-ahead form, so a 128M-value
stream takes seconds, and is bit-identical to stepping the recurrence one value at a time.
"""
from __future__ import annotations

import numpy as np

LCG_A = np.uint64(2862933555777941757)
LCG_C = np.uint64(3037000493)
_CHUNK = 1 << 20


def _jump_tables(n: int):
    """A[i] = a^(i+1), G[i] = 1 + a + ... + a^i  (mod 2^64) for i in [0, n)."""
    with np.errstate(over="ignore"):
        a = np.full(n, LCG_A, dtype=np.uint64)
        A = np.cumprod(a, dtype=np.uint64)
        G = np.empty(n, dtype=np.uint64)
        G[0] = 1
        if n > 1:
            G[1:] = A[:-1]
        G = np.cumsum(G, dtype=np.uint64)
    return A, G


def lcg_stream(seed: int, n: int) -> tuple[np.ndarray, int]:
    """Return (the next n LCG states after ``seed``, final state)."""
    out = np.empty(n, dtype=np.uint64)
    s = np.uint64(seed)
    A, G = _jump_tables(min(n, _CHUNK)) if n > 0 else (None, None)
    pos = 0
    with np.errstate(over="ignore"):
        while pos < n:
            m = min(_CHUNK, n - pos)
            out[pos:pos + m] = A[:m] * s + LCG_C * G[:m]
            s = out[pos + m - 1]
            pos += m
    return out, int(s)


def lcg_unit_floats(seed: int, n: int) -> tuple[np.ndarray, int]:
    """``Float(next() >> 11) / Float(1 << 53)`` for the next n draws (main.swift:537)."""
    st, s = lcg_stream(seed, n)
    f = (st >> np.uint64(11)).astype(np.float32) / np.float32(2.0 ** 53)
    return f, s


def _normalize_rows_seq(v: np.ndarray) -> np.ndarray:
    """Row-wise ``v / sqrt(sequential fp32 sum of squares)`` (main.swift:543-545)."""
    sq = v * v
    acc = np.zeros(v.shape[0], dtype=np.float32)
    for j in range(v.shape[1]):          # sequential fp32 reduction, vectorised over rows
        acc = acc + sq[:, j]
    norm = np.sqrt(acc)
    out = v.copy()
    nz = norm > 0
    out[nz] = v[nz] / norm[nz, None]
    return out


def bench_vectors(count: int, dim: int, seed: int, normalize: bool = True) -> np.ndarray:
    """``DataGen.generate(count, dim, seed)`` of the reference benchmark (seed 123 base / 321 query)."""
    f, _ = lcg_unit_floats(seed, count * dim)
    v = (f * np.float32(2) - np.float32(1)).reshape(count, dim)
    return _normalize_rows_seq(v) if normalize else v


def lcg24_floats(seed: int, n: int) -> tuple[np.ndarray, int]:
    """``Float(s >> 40) / Float(1 << 24) * 2 - 1`` fixture generator of the reference tests
    (ScoreBlockTests.swift:24-41, PQTrainTests.swift:734-743)."""
    st, s = lcg_stream(seed, n)
    f = (st >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)
    return f * np.float32(2) - np.float32(1), s


def gaussian_lcg(seed: int, n: int) -> np.ndarray:
    """n standard normals by Box-Muller on the LCG (main.swift:108-121; cos branch then sin branch)."""
    pairs = (n + 1) // 2
    f, _ = lcg_unit_floats(seed, 2 * pairs)
    u1 = np.maximum(f[0::2], np.float32(1e-12))
    u2 = f[1::2]
    r = np.sqrt(np.float32(-2) * np.log(u1)).astype(np.float32)
    ang = (np.float32(2 * np.pi) * u2).astype(np.float32)
    out = np.empty(2 * pairs, dtype=np.float32)
    out[0::2] = r * np.cos(ang)
    out[1::2] = r * np.sin(ang)
    return out[:n]


def sift_like(count: int, dim: int, n_clusters: int, seed: int) -> np.ndarray:
    """C3 "SIFT-shaped" base: non-negative integer-valued fp32 in [0, 218] with heavy ties, drawn
    as a mixture of ``n_clusters`` centres so that IVF lists are meaningful."""
    g = gaussian_lcg(seed + 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF, n_clusters * dim)
    centres = np.clip(np.floor(np.abs(g) * np.float32(40)), 0, 218).reshape(n_clusters, dim)
    st, _ = lcg_stream(seed, count)
    which = (st % np.uint64(n_clusters)).astype(np.int64)
    noise = gaussian_lcg((seed ^ 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF, count * dim).reshape(count, dim)
    v = centres[which] + np.float32(12) * noise
    return np.clip(np.floor(np.abs(v)), 0, 218).astype(np.float32)


def clustered_unit(count: int, dim: int, n_clusters: int, seed: int) -> np.ndarray:
    """C4/C5 "embedding / Deep-shaped" base: unit-norm Gaussian-cluster mixture, sigma = 0.3/sqrt(d)
    around ``n_clusters`` random unit centres (main.swift:129-144)."""
    g = gaussian_lcg((seed + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, n_clusters * dim)
    centres = _normalize_rows_seq(g.reshape(n_clusters, dim))
    st, _ = lcg_stream(seed, count)
    which = (st % np.uint64(n_clusters)).astype(np.int64)
    sigma = np.float32(0.3) / np.sqrt(np.float32(dim))
    noise = gaussian_lcg((seed ^ 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF, count * dim).reshape(count, dim)
    v = (centres[which] + sigma * noise).astype(np.float32)
    nrm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    return (v / nrm[:, None]).astype(np.float32)
