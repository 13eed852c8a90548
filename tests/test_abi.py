"""The C-ABI boundary: the library loads, exports every symbol include/*.h declares, PQEncodeOpts has
the reference's LP64 layout, and -- with no GPU -- every entry point fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT
from vectorindex_b200 import _lib


def _declared_symbols():
    names = set()
    for h in ("vindex_cuda.h", "cpq_encode.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for mt in re.finditer(r"\b((?:vix|cpq)_[a-z0-9_]+)\s*\(", src):
            names.add(mt.group(1))
    return names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 45
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"symbols declared in include/*.h but not exported: {missing}"
    assert set(_lib.exported_symbols()) == declared


def test_pqencodeopts_layout_matches_reference():
    """/root/reference/tools/pq_align_check.c:22-28: size 24, offsets 0/4/5/8/12/16/20."""
    o = _lib.PQEncodeOpts
    assert C.sizeof(o) == 24
    offs = [getattr(o, f).offset for f, _ in o._fields_]
    assert offs == [0, 4, 5, 8, 12, 16, 20]


def test_version_and_pack_helpers():
    L = _lib.lib()
    assert L.vix_version() >= 100
    codes = np.array([1, 15, 7, 0, 9, 3], dtype=np.uint8)
    packed = np.zeros(3, dtype=np.uint8)
    L.cpq_pack_u4_bulk(_lib.ptr(codes), C.c_int(6), _lib.ptr(packed))
    assert packed.tolist() == [0xF1, 0x07, 0x39]                    # low nibble = even subspace (pq_encode.c:441-447)
    back = np.zeros(6, dtype=np.uint8)
    L.cpq_unpack_u4_bulk(_lib.ptr(packed), C.c_int(6), _lib.ptr(back))
    assert back.tolist() == codes.tolist()


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    from vectorindex_b200 import kernels, VectorIndexError
    x = np.zeros((4, 8), dtype=np.float32)
    with pytest.raises(VectorIndexError) as e:
        kernels.flat_search_f32(x, x, 2)
    assert e.value.status == -101 and "no CPU fallback" in str(e.value)
    cb = np.zeros((2, 256, 4), dtype=np.float32)
    with pytest.raises(VectorIndexError):
        kernels.pq_encode_u8_f32(x, cb, 2)
    with pytest.raises(VectorIndexError):
        from vectorindex_b200.index import FlatIndex
        FlatIndex(8)


def test_argument_validation_mirrors_reference_preconditions():
    from vectorindex_b200 import kernels, VectorIndexError
    x = np.zeros((4, 10), dtype=np.float32)
    with pytest.raises(VectorIndexError) as e:                      # d % m != 0 (PQEncode.swift:77-78)
        kernels.pq_encode_u8_f32(x, np.zeros(1, dtype=np.float32), 3)
    assert e.value.kind == "invalidDim"
    with pytest.raises(VectorIndexError) as e:                      # PQTrain.swift:96-135 .emptyInput
        kernels.pq_train_f32(np.zeros((0, 8), dtype=np.float32), 2)
    assert e.value.kind == "emptyInput"
    with pytest.raises(VectorIndexError) as e:                      # PQTrainTests.swift:600-625: n < ks
        kernels.pq_train_f32(np.zeros((50, 128), dtype=np.float32), 4, ks=100)
    assert e.value.kind == "emptyInput" and "Insufficient training data" in str(e.value)
    with pytest.raises(VectorIndexError) as e:                      # PQTrainTests.swift:627-652: d % m != 0
        kernels.pq_train_f32(np.zeros((1000, 100), dtype=np.float32), 7, ks=64)
    assert e.value.kind == "invalidDim" and "divisible" in str(e.value)
