"""VINDEX container reader / writer (vectorindex_b200/container.py) against the layout the reference's builder and reader
define (Kernels/VIndexContainerBuilder.swift:39-266, Kernels/VIndexMmap.swift:79-156, 322-410, 602-647).  The Swift code
cannot run here, so the byte-level test restates the documented offsets independently of the writer."""
import struct
import zlib

import numpy as np
import pytest

from vectorindex_b200 import container as vc


def _lists(rng, kc, m, max_len, id_hi):
    lens = rng.integers(0, max_len, kc)
    lens[rng.integers(0, kc)] = 0                                   # at least one empty list
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(off[-1])
    return off, rng.integers(0, 256, (n, m), dtype=np.uint8), rng.permutation(id_hi)[:n].astype(np.int64)


@pytest.mark.parametrize("id_bits", [64, 32])
def test_container_round_trip(tmp_path, id_bits):
    rng = np.random.default_rng(1)
    d, m, kc = 32, 8, 13
    off, codes, ids = _lists(rng, kc, m, 40, 100000)
    coarse = rng.standard_normal((kc, d)).astype(np.float32)
    cb = rng.standard_normal((m, 256, d // m)).astype(np.float32)
    norms = rng.standard_normal((m, 256)).astype(np.float32)
    p = str(tmp_path / "a.vindex")
    vc.write_container(p, off, codes, ids, d, m, 256, coarse, cb, norms, id_bits=id_bits)
    r = vc.read_container(p)
    assert (r["d"], r["m"], r["ks"], r["kc"], r["n"], r["id_bits"], r["group"]) == (d, m, 256, kc, int(off[-1]), id_bits, 4)
    assert np.array_equal(r["list_offsets"], off) and np.array_equal(r["codes"], codes) and np.array_equal(r["ids"], ids)
    for k_, a in (("coarse", coarse), ("codebooks", cb), ("centroid_norms", norms)):
        assert np.array_equal(r[k_].view(np.uint32), a.view(np.uint32))


def test_container_bytes_follow_the_reference_layout(tmp_path):
    rng = np.random.default_rng(2)
    d, m, kc = 16, 4, 3
    off = np.array([0, 2, 2, 5], dtype=np.int64)
    codes = rng.integers(0, 256, (5, m), dtype=np.uint8)
    ids = np.array([7, 9, 1, 2, 3], dtype=np.int64)
    p = str(tmp_path / "b.vindex")
    vc.write_container(p, off, codes, ids, d, m)
    buf = open(p, "rb").read()
    assert len(buf) >= 4096 and len(buf) % 4096 == 0                                  # VIndexMmap.swift:343
    # magic: the constant 0x00585845444E4956 of VIndexMmap.swift:80 / VIndexContainerBuilder.swift:239 stored little-endian
    # (its bytes read "VINDEXX\0"; the comment beside it says "VINDEX\0\0" -- writer and reader both use the constant)
    assert buf[:8] == bytes.fromhex("56494e4445585800")
    assert struct.unpack_from("<HHBB", buf, 8) == (1, 0, 1, 0)                         # version 1.0, little-endian
    assert struct.unpack_from("<I", buf, 20)[0] == d and struct.unpack_from("<HH", buf, 24) == (m, 256)
    assert struct.unpack_from("<I", buf, 28)[0] == kc and struct.unpack_from("<BB", buf, 32) == (64, 4)
    n_total, gen, toc_off, ntoc, hcrc = struct.unpack_from("<QQQII", buf, 40)
    assert (n_total, gen, toc_off, ntoc) == (5, 0, 256, 3)
    h = bytearray(buf[:256]); h[68:72] = b"\0\0\0\0"
    assert zlib.crc32(bytes(h)) & 0xFFFFFFFF == hcrc                                  # CRC over 256 bytes, field zeroed
    assert buf[72:256] == bytes(184)
    toc = [struct.unpack_from("<IQQIIII", buf, toc_off + 36 * i) for i in range(ntoc)]  # packed 36-byte entries
    assert [t[0] for t in toc] == [4, 5, 6]                                           # listsDesc, ids, codes
    for ty, o, sz, al, fl, crc, rs in toc:
        assert o % al == 0 and fl == 0 and rs == 0 and zlib.crc32(buf[o:o + sz]) & 0xFFFFFFFF == crc
    assert toc[2][3] == 4096                                                          # codes section page aligned
    do, io, co = toc[0][1], toc[1][1], toc[2][1]
    for l, (b, e) in enumerate(zip(off[:-1], off[1:])):
        rec = buf[do + 64 * l: do + 64 * l + 64]
        assert rec[0] == 2 and rec[1] == 4 and rec[2] == 64                           # pq8, group, id_bits
        length, cap = struct.unpack_from("<II", rec, 4)
        ids_o, codes_o, vecs_o = struct.unpack_from("<QQQ", rec, 16)
        assert struct.unpack_from("<III", rec, 40) == (8, m, 0)                       # strides
        assert length == e - b and cap >= length and ids_o % 64 == 0 and codes_o % 64 == 0 and vecs_o == 0
        assert np.array_equal(np.frombuffer(buf, "<u8", length, io + ids_o).astype(np.int64), ids[b:e])
        assert np.array_equal(np.frombuffer(buf, np.uint8, length * m, co + codes_o).reshape(-1, m), codes[b:e])


def test_container_rejects_corruption(tmp_path):
    rng = np.random.default_rng(3)
    off, codes, ids = _lists(rng, 5, 4, 20, 1000)
    p = str(tmp_path / "c.vindex")
    vc.write_container(p, off, codes, ids, 8, 4)
    raw = bytearray(open(p, "rb").read())
    bad = bytearray(raw); bad[0] ^= 1
    open(p, "wb").write(bad)
    with pytest.raises(vc.ContainerError, match="magic"):
        vc.read_container(p)
    bad = bytearray(raw); bad[21] ^= 1                               # header field without fixing the CRC
    open(p, "wb").write(bad)
    with pytest.raises(vc.ContainerError, match="header CRC"):
        vc.read_container(p)
    bad = bytearray(raw); bad[-4096] ^= 0x55                          # first byte of the codes section
    open(p, "wb").write(bad)
    with pytest.raises(vc.ContainerError, match="CRC mismatch"):
        vc.read_container(p)
    assert vc.read_container(p, verify_crcs=False)["n"] == int(off[-1])
    with pytest.raises(vc.ContainerError):
        vc.write_container(p, off, codes, ids + (1 << 33), 8, 4, id_bits=32)


@pytest.mark.gpu
def test_index_survives_the_container(tmp_path, oracle):
    """GPU index -> export_lists -> container -> import_lists into a fresh index: identical answers."""
    from test_gpu_parity import _make_ivfpq_problem, bits
    from vectorindex_b200.index import IVFPQIndex
    n, d, m, kc, nq, k, nprobe = 6000, 64, 16, 24, 32, 10, 5
    xb, q, coarse, cb, norms = _make_ivfpq_problem(oracle, n, d, m, kc, nq, seed=77)
    idx = IVFPQIndex(d, "euclidean", nlist=kc, nprobe=nprobe, m=m)
    idx.set_coarse(coarse); idx.set_codebooks(cb, norms)
    idx.batch_insert(xb, np.arange(n, dtype=np.int64) * 2)
    gd, gi = idx.batch_search(q, k)
    off, codes, lids, _ = idx.export_lists()
    p = str(tmp_path / "idx.vindex")
    vc.write_container(p, off, codes, lids, d, m, 256, idx.get_coarse(), *idx.get_codebooks())
    r = vc.read_container(p)
    idx2 = IVFPQIndex(r["d"], "euclidean", nlist=r["kc"], nprobe=nprobe, m=r["m"])
    idx2.set_coarse(r["coarse"]); idx2.set_codebooks(r["codebooks"], r["centroid_norms"])
    idx2.import_lists(r["list_offsets"], r["codes"], r["ids"])
    gd2, gi2 = idx2.batch_search(q, k)
    assert np.array_equal(gi, gi2) and np.array_equal(bits(gd), bits(gd2))
