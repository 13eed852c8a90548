"""The reference's on-disk index container ("VINDEX", S1 mmap sections) -- host-side reader / writer, so that an index
built on the GPU can be opened by the reference (`IndexMmap.open`) and a reference-built container can be loaded onto
the GPU (`IVFPQIndex.import_lists`).  Plain file IO: nothing here touches the device.

Layout, restated from the reference's writer and reader (little-endian throughout):

* header, 256 bytes (`VIndexHeader`, Kernels/VIndexMmap.swift:131-156; written by
  Kernels/VIndexContainerBuilder.swift:238-266): magic 0x00585845444E4956 u64 @0 (bytes "VINDEXX\\0"), version 1.0
  u16 @8 / @10, endianness (1 = little) u8 @12, arch u8 @13, flags u32 @16, d u32 @20, m u16 @24, ks u16 @26, kc u32 @28, id_bits u8 @32,
  code_group_g u8 @33, N_total u64 @40, generation u64 @48, toc_offset u64 @56, toc_entries u32 @64, header_crc32 u32
  @68 (CRC-32 of the 256 bytes with this field zero, VIndexMmap.swift:158-175); the rest zero;
* TOC at toc_offset: packed 36-byte entries -- type u32 @0, offset u64 @4, size u64 @12, align u32 @20, flags u32 @24,
  crc32 u32 @28 (CRC-32 of the section bytes), reserved u32 @32 (VIndexContainerBuilder.swift:204-215,
  VIndexMmap.swift:602-617); section types: centroids 1, codebooks 2, centroidNorms 3, listsDesc 4, ids 5, codes 6
  (VIndexMmap.swift:83-87);
* listsDesc: one packed 64-byte record per list -- format u8 @0 (2 = pq8), group u8 @1, id_bits u8 @2, length u32 @4,
  capacity u32 @8, ids_offset u64 @16 and codes_offset u64 @24 (relative to their sections, 64-byte aligned),
  ids_stride u32 @40, codes_stride u32 @44, vecs_stride u32 @48 (VIndexContainerBuilder.swift:176-201);
* ids: u64 (or u32) per vector; codes: AoS rows of m bytes, list order = append order (Kernels/IVFAppend.swift:735-737).

Sections start at multiples of their `align` (64; the codes section page-aligned, as the builder places it), the file
is at least 4096 bytes (VIndexMmap.swift:343).  The reference's CRC-32 (table of VIndexMmap.swift:49-56: reflected
0xEDB88320, initial / final 0xFFFFFFFF) is zlib's."""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAGIC = 0x00585845444E4956
SEC_CENTROIDS, SEC_CODEBOOKS, SEC_CENTROID_NORMS, SEC_LISTS_DESC, SEC_IDS, SEC_CODES = 1, 2, 3, 4, 5, 6
FORMAT_PQ8 = 2
PAGE = 4096


class ContainerError(ValueError):
    pass


def _align(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def _header(d, m, ks, kc, id_bits, group, n_total, toc_offset, toc_entries) -> bytes:
    h = bytearray(256)
    struct.pack_into("<QHHBB", h, 0, MAGIC, 1, 0, 1, 0)
    struct.pack_into("<IIHHI", h, 16, 0, d, m, ks, kc)
    struct.pack_into("<BB", h, 32, id_bits, group)
    struct.pack_into("<QQQII", h, 40, n_total, 0, toc_offset, toc_entries, 0)
    struct.pack_into("<I", h, 68, zlib.crc32(bytes(h)) & 0xFFFFFFFF)
    return bytes(h)


def write_container(path, list_offsets, codes, ids, d: int, m: int, ks: int = 256, coarse=None, codebooks=None,
                    centroid_norms=None, id_bits: int = 64, group: int = 4) -> None:
    """CSR lists (``list_offsets`` [kc + 1], ``codes`` [n x m] u8 and ``ids`` [n] in list order -- what
    ``IVFPQIndex.export_lists`` returns) -> a pq8 container.  ``coarse`` / ``codebooks`` / ``centroid_norms`` (the PQ
    centroid norms [m x ks]) are stored in their sections when given.  List capacity = list length."""
    off = np.ascontiguousarray(list_offsets, dtype=np.int64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1, m)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    kc, n = off.size - 1, int(off[-1])
    if kc <= 0 or codes.shape[0] != n or ids.size != n or (np.diff(off) < 0).any():
        raise ContainerError("inconsistent CSR lists")
    if id_bits not in (32, 64) or group not in (4, 8) or m % group:
        raise ContainerError("id_bits must be 32 or 64, group 4 or 8 and a divisor of m")
    if id_bits == 32 and n and (ids.min() < 0 or ids.max() > 0xFFFFFFFF):
        raise ContainerError("id exceeds 32 bits")
    id_stride = id_bits // 8
    lens = np.diff(off)

    # per-list inner offsets, 64-byte aligned inside their sections (VIndexContainerBuilder.swift:93-109)
    ids_off = np.zeros(kc, dtype=np.int64)
    codes_off = np.zeros(kc, dtype=np.int64)
    it = ct = 0
    for l in range(kc):
        it = _align(it, 64); ids_off[l] = it; it += int(lens[l]) * id_stride
        ct = _align(ct, 64); codes_off[l] = ct; ct += int(lens[l]) * m
    ids_size, codes_size = _align(max(it, 1), 64), _align(max(ct, 1), PAGE)

    extra = []                                                       # float sections in front, as the reader maps them
    for ty, arr in ((SEC_CENTROIDS, coarse), (SEC_CODEBOOKS, codebooks), (SEC_CENTROID_NORMS, centroid_norms)):
        if arr is not None:
            extra.append((ty, np.ascontiguousarray(arr, dtype="<f4").tobytes()))
    ntoc = 3 + len(extra)
    toc_offset = 256
    pos = _align(toc_offset + 36 * ntoc, 64)
    sections = []                                                    # (type, offset, size, align, bytes)
    for ty, raw in extra:
        sections.append((ty, pos, len(raw), 64, raw))
        pos = _align(pos + len(raw), 64)

    desc = bytearray(64 * kc)
    for l in range(kc):
        struct.pack_into("<BBBBII", desc, 64 * l, FORMAT_PQ8, group, id_bits, 0, int(lens[l]), int(lens[l]))
        struct.pack_into("<QQQ", desc, 64 * l + 16, int(ids_off[l]), int(codes_off[l]), 0)
        struct.pack_into("<IIII", desc, 64 * l + 40, id_stride, m, 0, 0)
    sections.append((SEC_LISTS_DESC, pos, len(desc), 64, bytes(desc)))
    pos = _align(pos + len(desc), 64)

    ids_raw = bytearray(ids_size)
    codes_raw = bytearray(codes_size)
    idt = "<u4" if id_bits == 32 else "<u8"
    for l in range(kc):
        b, e = int(off[l]), int(off[l + 1])
        if e > b:
            ids_raw[int(ids_off[l]):int(ids_off[l]) + (e - b) * id_stride] = ids[b:e].astype(idt).tobytes()
            codes_raw[int(codes_off[l]):int(codes_off[l]) + (e - b) * m] = codes[b:e].tobytes()
    sections.append((SEC_IDS, pos, ids_size, 64, bytes(ids_raw)))
    pos = _align(pos + ids_size, PAGE)
    sections.append((SEC_CODES, pos, codes_size, PAGE, bytes(codes_raw)))
    pos = _align(pos + codes_size, PAGE)
    file_size = max(pos, PAGE)

    buf = bytearray(file_size)
    buf[0:256] = _header(d, m, ks, kc, id_bits, group, n, toc_offset, ntoc)
    for i, (ty, o, sz, al, raw) in enumerate(sections):
        struct.pack_into("<IQQIIII", buf, toc_offset + 36 * i, ty, o, sz, al, 0, zlib.crc32(raw) & 0xFFFFFFFF, 0)
        buf[o:o + sz] = raw
    with open(path, "wb") as f:
        f.write(buf)


def read_container(path, verify_crcs: bool = True) -> dict:
    """A pq8 container -> dict(d, m, ks, kc, n, id_bits, group, list_offsets [kc + 1], codes [n x m], ids [n] int64, and
    coarse / codebooks / centroid_norms when their sections exist).  Mirrors the checks of `IndexMmap.open` /
    `indexInit` (VIndexMmap.swift:343-410, 602-647): size, magic, endianness, major version, header CRC, section
    alignment and CRCs."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < PAGE:
        raise ContainerError("file smaller than 4096 bytes")
    magic, vmaj, _vmin, endian, _arch = struct.unpack_from("<QHHBB", buf, 0)
    if endian != 1:
        raise ContainerError("only little-endian containers are supported")
    if magic != MAGIC:
        raise ContainerError("invalid magic number in index header")
    if vmaj != 1:
        raise ContainerError(f"unsupported index file version {vmaj}")
    _flags, d, m, ks, kc = struct.unpack_from("<IIHHI", buf, 16)
    id_bits, group = struct.unpack_from("<BB", buf, 32)
    n_total, _gen, toc_offset, ntoc, hcrc = struct.unpack_from("<QQQII", buf, 40)
    h = bytearray(buf[:256])
    struct.pack_into("<I", h, 68, 0)
    if verify_crcs and (zlib.crc32(bytes(h)) & 0xFFFFFFFF) != hcrc:
        raise ContainerError("header CRC mismatch")
    sec = {}
    for i in range(ntoc):
        ty, o, sz, al, _fl, crc, _r = struct.unpack_from("<IQQIIII", buf, toc_offset + 36 * i)
        if al and o % al:
            raise ContainerError(f"section {ty} misaligned in index file")
        if o + sz > len(buf):
            raise ContainerError(f"section {ty} exceeds the file")
        if verify_crcs and sz and (zlib.crc32(buf[o:o + sz]) & 0xFFFFFFFF) != crc:
            raise ContainerError(f"section {ty} CRC mismatch")
        sec[ty] = (o, sz)
    for need in (SEC_LISTS_DESC, SEC_IDS, SEC_CODES):
        if need not in sec:
            raise ContainerError(f"section {need} missing")
    do, dsz = sec[SEC_LISTS_DESC]
    if dsz < 64 * kc:
        raise ContainerError("listsDesc section shorter than kc records")
    io, _ = sec[SEC_IDS]
    co, _ = sec[SEC_CODES]
    lens = np.zeros(kc, dtype=np.int64)
    parts_i, parts_c = [], []
    for l in range(kc):
        fmt, _g, ib, _r0, length, cap = struct.unpack_from("<BBBBII", buf, do + 64 * l)
        ids_o, codes_o, _v = struct.unpack_from("<QQQ", buf, do + 64 * l + 16)
        ids_st, codes_st, _vs, _r1 = struct.unpack_from("<IIII", buf, do + 64 * l + 40)
        if fmt != FORMAT_PQ8:
            raise ContainerError(f"list {l}: format {fmt} is not pq8")
        if length > cap or ids_st != ib // 8 or codes_st < m:
            raise ContainerError(f"list {l}: inconsistent descriptor")
        lens[l] = length
        if length:
            raw = np.frombuffer(buf, dtype="<u4" if ib == 32 else "<u8", count=length, offset=io + ids_o)
            parts_i.append(raw.astype(np.int64))
            rows = np.frombuffer(buf, dtype=np.uint8, count=length * codes_st, offset=co + codes_o).reshape(length, codes_st)
            parts_c.append(np.ascontiguousarray(rows[:, :m]))
    out = dict(d=d, m=m, ks=ks, kc=kc, n=int(lens.sum()), n_total=n_total, id_bits=id_bits, group=group,
               list_offsets=np.concatenate([[0], np.cumsum(lens)]).astype(np.int64),
               codes=np.concatenate(parts_c) if parts_c else np.zeros((0, m), np.uint8),
               ids=np.concatenate(parts_i) if parts_i else np.zeros(0, np.int64))
    dsub = d // m if m else 0
    for key, ty, shape in (("coarse", SEC_CENTROIDS, (kc, d)), ("codebooks", SEC_CODEBOOKS, (m, ks, dsub)),
                           ("centroid_norms", SEC_CENTROID_NORMS, (m, ks))):
        if ty in sec:
            o, sz = sec[ty]
            a = np.frombuffer(buf, dtype="<f4", count=sz // 4, offset=o)
            out[key] = a.reshape(shape).copy() if a.size == int(np.prod(shape)) else a.copy()
    return out
