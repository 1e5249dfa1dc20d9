"""Chunk wire codec: tg_chunk_encode / tg_chunk_decode / tg_chunk_decode_into against the Python
restatement of pkg/util/chunk/codec.go (oracle/chunk_codec.py), hand-assembled golden bytes, and the reference's own
TestCodec round trip (codec_test.go:27-78).  Host code only — runs without a GPU."""
import ctypes as C
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import chunk_codec as W  # noqa: E402

from tidb_b200 import abi, build  # noqa: E402

TYPE_VARCHAR, TYPE_JSON = 0x0F, 0xF5


@pytest.fixture(scope="module")
def lib():
    build.build()
    return abi.load_lib()


def _col(length, bitmap, offsets, data, elem_len, keep):
    """tg_column over numpy/bytes buffers (kept alive in `keep`)"""
    c = abi.TgColumn()
    c.length = length
    c.elem_len = elem_len
    if bitmap is not None:
        b = np.frombuffer(bitmap, dtype=np.uint8).copy(); keep.append(b); c.null_bitmap = b.ctypes.data
    if offsets is not None:
        o = np.asarray(offsets, dtype=np.int64).copy(); keep.append(o); c.offsets = o.ctypes.data
    d = np.frombuffer(data, dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    keep.append(d); c.data = d.ctypes.data
    return c


def _chunk(cols, keep):
    arr = (abi.TgColumn * len(cols))(*cols)
    keep.append(arr)
    ck = abi.TgChunk(); ck.ncols = len(cols); ck.cols = arr
    return ck


def _encode(lib, ck):
    need = C.c_size_t(0)
    abi.check(lib.tg_chunk_wire_size(C.byref(ck), C.byref(need)))
    buf = (C.c_uint8 * max(need.value, 1))()
    wrote = C.c_size_t(0)
    abi.check(lib.tg_chunk_encode(C.byref(ck), buf, C.c_size_t(need.value), C.byref(wrote)))
    assert wrote.value == need.value
    return bytes(buf[:need.value])


def test_golden_bytes_fixed_and_varlen(lib):
    # hand-assembled from codec.go:49-76: 3 rows; col0 int64 [7, NULL, -1] (bitmap 0b101, nullCount 1);
    # col1 varchar ["ab", "", "xyz"] (no NULLs => no bitmap on the wire), offsets [0,2,2,5]
    keep = []
    c0 = _col(3, bytes([0b101]), None, struct.pack("<qqq", 7, 0, -1), 8, keep)
    c1 = _col(3, None, [0, 2, 2, 5], b"abxyz", -1, keep)
    wire = _encode(lib, _chunk([c0, c1], keep))
    golden = (struct.pack("<II", 3, 1) + bytes([0b101]) + struct.pack("<qqq", 7, 0, -1)
              + struct.pack("<II", 3, 0) + struct.pack("<qqqq", 0, 2, 2, 5) + b"abxyz")
    assert wire == golden
    assert wire == W.encode([W.WireColumn(3, bytes([0b101]), None, struct.pack("<qqq", 7, 0, -1)),
                             W.WireColumn(3, None, np.array([0, 2, 2, 5]), b"abxyz")])
    # a bitmap with every bit set has nullCount 0 and must NOT be written (codec.go:61)
    c2 = _col(3, bytes([0b111]), None, struct.pack("<qqq", 1, 2, 3), 8, keep)
    assert _encode(lib, _chunk([c2], keep)) == struct.pack("<II", 3, 0) + struct.pack("<qqq", 1, 2, 3)


def _decode(lib, wire, types):
    buf = np.frombuffer(wire, dtype=np.uint8).copy()
    # 8-byte aligned base so that var-len offsets can be aliased when they happen to fall on a multiple of 8
    cols = (abi.TgColumn * len(types))()
    used = C.c_size_t(0)
    tarr = (C.c_int32 * len(types))(*types)
    rc = lib.tg_chunk_decode(C.c_void_p(buf.ctypes.data), C.c_size_t(len(wire)), len(types), tarr, cols, C.byref(used))
    return rc, cols, used.value, buf


def test_reference_testcodec_round_trip(lib):
    # codec_test.go:27-78: 10 rows; col0 int64 all NULL, col1 int64 = i, col2/col3 varchar "%d.12345", col4 40-byte decimal
    # (opaque bytes here), col5 JSON (var-len opaque)
    n = 10
    keep = []
    strs = [f"{i}.12345".encode() for i in range(n)]
    offs = np.concatenate([[0], np.cumsum([len(s) for s in strs])]).astype(np.int64)
    dec = bytes((i * 7 + j) & 0xFF for i in range(n) for j in range(40))
    js = [b"\x0c" + s for s in strs]
    joffs = np.concatenate([[0], np.cumsum([len(s) for s in js])]).astype(np.int64)
    cols = [_col(n, bytes([0, 0]), None, bytes(8 * n), 8, keep),
            _col(n, None, None, np.arange(n, dtype=np.int64).tobytes(), 8, keep),
            _col(n, None, offs, b"".join(strs), -1, keep), _col(n, None, offs, b"".join(strs), -1, keep),
            _col(n, None, None, dec, 40, keep), _col(n, None, joffs, b"".join(js), -1, keep)]
    types = [abi.TYPE_LONGLONG, abi.TYPE_LONGLONG, TYPE_VARCHAR, TYPE_VARCHAR, abi.TYPE_NEWDECIMAL, TYPE_JSON]
    wire = _encode(lib, _chunk(cols, keep))
    # the restatement decodes what the library encoded ...
    wcols, rest = W.decode(wire, types)
    assert rest == b"" and [c.length for c in wcols] == [n] * 6
    assert wcols[0].null_count() == n and all(c.null_count() == 0 for c in wcols[1:])
    assert np.frombuffer(wcols[1].data, dtype="<i8").tolist() == list(range(n))
    assert [wcols[2].data[wcols[2].offsets[i]:wcols[2].offsets[i + 1]] for i in range(n)] == strs
    assert wcols[4].data == dec
    # ... and re-encodes to the same bytes
    assert W.encode([W.WireColumn(c.length, c.bitmap, c.offsets, c.data) for c in wcols]) == wire
    # the library decodes its own output into views of the buffer
    rc, dcols, used, buf = _decode(lib, wire, types)
    assert rc == 0 and used == len(wire)
    base = buf.ctypes.data
    assert dcols[0].null_bitmap is not None and dcols[1].null_bitmap is None
    got1 = np.frombuffer(wire, dtype="<i8", count=n, offset=dcols[1].data - base)
    assert got1.tolist() == list(range(n))
    assert dcols[4].elem_len == 40 and dcols[2].elem_len == -1
    o2 = np.frombuffer(wire, dtype="<i8", count=n + 1, offset=dcols[2].offsets - base)   # unaligned view, as in the Go decoder
    assert np.array_equal(o2, offs)


def test_decode_into_pinned_style_buffers_and_errors(lib):
    rng = np.random.default_rng(3)
    n = 1000
    a = rng.integers(-1 << 62, 1 << 62, n, dtype=np.int64)
    b = rng.random(n)
    notnull = rng.random(n) > 0.3
    bitmap = np.packbits(notnull, bitorder="little").tobytes()
    wire = W.encode([W.WireColumn(n, bitmap, None, a.tobytes()), W.WireColumn(n, None, None, b.tobytes())])
    types = (C.c_int32 * 2)(abi.TYPE_LONGLONG, abi.TYPE_DOUBLE)
    oa, ob = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.float64)
    na, nb_ = np.zeros((n + 7) // 8, dtype=np.uint8), np.zeros((n + 7) // 8, dtype=np.uint8)
    mc = (abi.TgMutColumn * 2)()
    mc[0].data, mc[0].null_bitmap, mc[0].elem_len = oa.ctypes.data, na.ctypes.data, 8
    mc[1].data, mc[1].null_bitmap, mc[1].elem_len = ob.ctypes.data, nb_.ctypes.data, 8
    out = abi.TgMutChunk(); out.ncols = 2; out.cols = mc; out.capacity_rows = n
    wb = np.frombuffer(wire, dtype=np.uint8).copy()
    rows, used = C.c_int64(0), C.c_size_t(0)
    abi.check(lib.tg_chunk_decode_into(C.c_void_p(wb.ctypes.data), C.c_size_t(len(wire)), 2, types, C.byref(out), C.byref(rows), C.byref(used)))
    assert rows.value == n and used.value == len(wire)
    assert np.array_equal(oa, a) and np.array_equal(ob, b)
    assert na.tobytes() == bitmap and set(nb_[:n // 8].tolist()) == {0xFF}     # no NULLs on the wire -> all-ones bitmap (setAllNotNull)
    # truncated input and too-small output are errors, not overruns
    assert lib.tg_chunk_decode_into(C.c_void_p(wb.ctypes.data), C.c_size_t(len(wire) - 5), 2, types, C.byref(out), C.byref(rows), C.byref(used)) == abi.TG_ERR_INVALID
    out.capacity_rows = n - 1
    assert lib.tg_chunk_decode_into(C.c_void_p(wb.ctypes.data), C.c_size_t(len(wire)), 2, types, C.byref(out), C.byref(rows), C.byref(used)) == abi.TG_ERR_CAPACITY
    # empty chunk: 8 header bytes per column
    empty = W.encode([W.WireColumn(0, None, None, b""), W.WireColumn(0, None, None, b"")])
    assert empty == struct.pack("<IIII", 0, 0, 0, 0)
    out.capacity_rows = n
    eb = np.frombuffer(empty, dtype=np.uint8).copy()
    abi.check(lib.tg_chunk_decode_into(C.c_void_p(eb.ctypes.data), C.c_size_t(len(empty)), 2, types, C.byref(out), C.byref(rows), C.byref(used)))
    assert rows.value == 0 and used.value == 16
