"""Test configuration.

  -m "not gpu"  : the oracle against the reference's golden vectors / fixtures, host logic, and that the
                  C-ABI library loads and exports every symbol include/*.h declares (no compute calls).
  -m gpu        : the parity tests proper -- every call goes through the C ABI of libvindex_b200.so and is
                  compared with the CPU oracle (oracle/), which is test infrastructure only.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    if not os.path.exists(os.path.join(ROOT, "oracle", "libvix_oracle.so")):
        orc.build(ref=os.path.exists("/root/reference"))
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def vk():
    """kernel-level API of the product (loads libvindex_b200.so; fails loudly if it is missing)."""
    from vectorindex_b200 import kernels
    return kernels


def lcg24(seed, n):
    """Float(s >> 40) / 2^24 * 2 - 1 fixture generator of the reference tests (ScoreBlockTests.swift:24-41)."""
    from vectorindex_b200 import datagen
    f, _ = datagen.lcg24_floats(seed, n)
    return f


def parity_fixture(n=16, d=32, m=8, ks=256, kc=4):
    """sin/cos fixture of PQEncodeParity_AoS_C_vs_Swift_Tests.swift:5-31."""
    i = np.arange(n * d, dtype=np.int64)
    x = (np.sin((i * 131 % 1024).astype(np.float32)) * np.float32(0.25)
         + np.cos((i * 17 % 997).astype(np.float32)) * np.float32(0.125)).astype(np.float32).reshape(n, d)
    dsub = d // m
    j = np.arange(m * ks * dsub, dtype=np.int64)
    cb = (np.sin((j * 313 % 2048).astype(np.float32)) * np.float32(0.2)
          + np.cos((j * 23 % 1237).astype(np.float32)) * np.float32(0.15)).astype(np.float32)
    c = np.arange(kc * d, dtype=np.int64)
    coarse = (np.cos((c * 19 % 4096).astype(np.float32)) * np.float32(0.33)).astype(np.float32).reshape(kc, d)
    assign = (np.arange(n) % kc).astype(np.int32)
    return x, cb, coarse, assign
