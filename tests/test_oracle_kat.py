"""Known-answer tests transcribed by hand from the reference's own unit tests: they pin the
CPU oracle (oracle/*.cpp) to pingcap/tidb @ 999e8f4c.  Each test cites the Go test it restates."""
import struct

import numpy as np
import pytest

import oracle_lib as O
from tidb_b200 import abi

L = O.lib()


def test_fnv1_64_go_stdlib_golden():
    # Go stdlib hash/fnv/fnv_test.go golden64 (FNV-1): used by the join through fnv.New64()
    # (pkg/executor/join/row_table_builder.go:103, base_join_probe.go:862)
    golden = {b"": 0xCBF29CE484222325, b"a": 0xAF63BD4C8601B7BE, b"ab": 0x08326707B4EB37B8,
              b"abc": 0xD8DCCA186BAFADCB}
    for k, v in golden.items():
        assert L.orc_fnv1_64(k, len(k)) == v
    # and the definition itself: h = h*prime ^ byte from offset basis
    h = 14695981039346656037
    for b in struct.pack("<q", -277544960):
        h = (h * 1099511628211) & ((1 << 64) - 1)
        h ^= b
    assert L.orc_fnv1_64(struct.pack("<q", -277544960), 8) == h


def test_hash_table_size():
    # pkg/executor/join/hash_table_v2_test.go:93-110 TestHashTableSize
    for rows, slots in ((10, 32), (32, 64), (33, 64), (64, 128), (65, 128)):
        assert L.orc_hash_table_length(rows) == slots
    # nextPowerOfTwo returns a power of two strictly greater than the input (hash_table_v2.go:55)
    for v, e in ((0, 2), (1, 2), (2, 4), (3, 4), (4, 8), (1023, 1024), (1024, 2048)):
        assert L.orc_next_power_of_two(v) == e


def test_setup_partition_info():
    # pkg/executor/join/row_table_builder_test.go:601-637 TestSetupPartitionInfo
    table = [(1, 1, 64), (2, 2, 63), (3, 4, 62), (4, 4, 62), (5, 8, 61), (6, 8, 61), (7, 8, 61), (8, 8, 61),
             (9, 16, 60), (10, 16, 60), (11, 16, 60), (12, 16, 60), (13, 16, 60), (14, 16, 60), (15, 16, 60),
             (16, 16, 60), (17, 16, 60), (18, 16, 60), (100, 16, 60)]
    for conc, pn, off in table:
        assert L.orc_partition_number(conc) == pn
        assert L.orc_partition_mask_offset(pn) == off


def test_tagged_bits():
    # pkg/executor/join/tagged_ptr_test.go:24-31 TestTaggedBits
    p = 0
    for i in range(65):
        assert L.orc_tagged_bits(p) == min(64 - i, 24)
        p = ((p << 1) + 1) & ((1 << 64) - 1)


def test_tag_helper_init():
    # tagged_ptr_test.go:33-41 TestTagHelperInit: mask = ^maxTaggedMask << (24 - taggedBits)
    mask = (~0xFFFFFFFFFF) & ((1 << 64) - 1)
    for bits in range(24, -1, -1):
        assert L.orc_tagged_mask(bits) == mask
        mask = (mask << 1) & ((1 << 64) - 1)


# ---- joinTableMeta KATs (pkg/executor/join/join_table_meta_test.go) -----------------------------------
T = dict(tiny=(1, 0, 0), int=(8, 0, 0), uint=(8, abi.FLAG_UNSIGNED, 0), year=(13, 0, 0), duration=(11, 0, 0),
         enum=(0xF7, 0, 0), enumint=(0xF7, 1 << 21, 0), set=(0xF8, 0, 0), bit=(16, 0, 0), json=(0xF5, 0, 0),
         float=(4, 0, 0), double=(5, 0, 0), string=(0xFD, 0, 0), binstring=(0xFC, 0, 1), date=(12, 0, 0),
         decimal=(0xF6, 0, 0))
ONE_INT64, FIXED, VARIABLE = 0, 1, 2
NORMAL, NEED_SIGN, KEEP_VAR = 0, 1, 2


def meta(key_idx, build, build_keys, probe_keys, other=None, output=(), used_flag=False):
    import ctypes as C
    def arrs(names):
        tp = (C.c_int32 * max(len(names), 1))(*[T[n][0] for n in names])
        fl = (C.c_uint32 * max(len(names), 1))(*[T[n][1] for n in names])
        bn = (C.c_int32 * max(len(names), 1))(*[T[n][2] for n in names])
        return tp, fl, bn
    ki = (C.c_int32 * len(key_idx))(*key_idx)
    b = arrs(build); bk = arrs(build_keys); pk = arrs(probe_keys)
    oth = (C.c_int32 * max(len(other or []), 1))(*(other or []))
    out = (C.c_int32 * max(len(output or []), 1))(*(output or []))
    m = O.OrcTableMeta()
    rc = L.orc_new_table_meta(len(key_idx), ki, len(build), b[0], b[1], b[2], bk[0], bk[1], bk[2],
                              pk[0], pk[1], pk[2], -1 if other is None else len(other), oth,
                              -1 if output is None else len(output), out, int(used_flag), C.byref(m))
    assert rc == 0
    return m


def test_join_table_meta_key_mode():
    # join_table_meta_test.go:27-87 TestJoinTableMetaKeyMode
    cases = [
        ([0], ["tiny"], ["tiny"], ["tiny"], ONE_INT64), ([0], ["year"], ["year"], ["year"], ONE_INT64),
        ([0], ["duration"], ["duration"], ["duration"], ONE_INT64), ([0], ["bit"], ["bit"], ["bit"], ONE_INT64),
        ([0], ["int"], ["int"], ["int"], ONE_INT64), ([0], ["uint"], ["uint"], ["uint"], ONE_INT64),
        ([0], ["date"], ["date"], ["date"], ONE_INT64), ([0], ["enumint"], ["enumint"], ["enumint"], ONE_INT64),
        ([0], ["int"], ["int"], ["uint"], FIXED), ([0], ["uint"], ["uint"], ["int"], FIXED),
        ([0], ["float"], ["float"], ["float"], FIXED), ([0], ["double"], ["double"], ["double"], FIXED),
        ([0, 1], ["date", "int"], ["date", "int"], ["date", "int"], FIXED),
        ([0, 1], ["int", "int"], ["int", "int"], ["int", "int"], FIXED),
        ([0], ["decimal"], ["decimal"], ["decimal"], VARIABLE), ([0], ["enum"], ["enum"], ["enum"], VARIABLE),
        ([0], ["set"], ["set"], ["set"], VARIABLE), ([0], ["json"], ["json"], ["json"], VARIABLE),
        ([0], ["string"], ["string"], ["string"], VARIABLE),
        ([0, 1], ["int", "string"], ["int", "string"], ["int", "string"], VARIABLE),
    ]
    for i, (ki, b, bk, pk, mode) in enumerate(cases):
        assert meta(ki, b, bk, pk).key_mode == mode, f"case {i}"


def test_join_table_meta_key_inlined_and_fixed():
    # join_table_meta_test.go:89-151 TestJoinTableMetaKeyInlinedAndFixed
    cases = [
        ([0], ["tiny"], ["tiny"], True, True, 8), ([0], ["int"], ["int"], True, True, 8),
        ([0], ["uint"], ["uint"], True, True, 8), ([0], ["year"], ["year"], True, True, 8),
        ([0], ["duration"], ["duration"], True, True, 8),
        ([0, 1], ["int", "duration"], ["int", "duration"], True, True, 16),
        ([0], ["binstring"], ["binstring"], True, False, -1),
        ([0, 1], ["binstring", "int"], ["binstring", "int"], True, False, -1),
        ([0], ["uint"], ["int"], False, True, 9), ([0], ["enumint"], ["enumint"], False, True, 8),
        ([0], ["double"], ["double"], False, True, 8), ([0], ["float"], ["float"], False, True, 8),
        ([0], ["date"], ["date"], False, True, 8), ([0], ["bit"], ["bit"], False, True, 8),
        ([0, 1], ["bit", "int"], ["bit", "int"], False, True, 16),
        ([0], ["decimal"], ["decimal"], False, False, -1), ([0], ["enum"], ["enum"], False, False, -1),
        ([0], ["set"], ["set"], False, False, -1), ([0], ["string"], ["string"], False, False, -1),
        ([0], ["json"], ["json"], False, False, -1),
        ([0, 1], ["decimal", "int"], ["decimal", "int"], False, False, -1),
        ([0, 1], ["enum", "int"], ["enum", "int"], False, False, -1),
        ([0, 1], ["enum", "decimal"], ["enum", "decimal"], False, False, -1),
    ]
    for i, (ki, b, pk, inl, fixed, klen) in enumerate(cases):
        m = meta(ki, b, b, pk)
        assert (bool(m.is_keys_inlined), bool(m.is_keys_fixed_length), m.join_keys_length) == (inl, fixed, klen), f"case {i}"


def test_join_table_meta_serialized_mode():
    # join_table_meta_test.go:167-215 TestJoinTableMetaSerializedMode
    cases = [
        ([0, 1], ["decimal", "int"], ["decimal", "int"], [NORMAL, NORMAL]),
        ([0, 1], ["uint", "int"], ["int", "int"], [NEED_SIGN, NORMAL]),
        ([0], ["uint"], ["int"], [NEED_SIGN]),
        ([0, 1], ["int", "binstring"], ["int", "binstring"], [NORMAL, KEEP_VAR]),
        ([0], ["binstring"], ["binstring"], [KEEP_VAR]),
        ([0, 1], ["int", "binstring"], ["uint", "binstring"], [NEED_SIGN, NORMAL]),
        ([0, 1], ["string", "binstring"], ["string", "binstring"], [KEEP_VAR, KEEP_VAR]),
        ([0, 1], ["string", "decimal"], ["string", "decimal"], [KEEP_VAR, KEEP_VAR]),
        ([0, 1], ["jsonx"], None, None),
    ]
    cases = cases[:-1] + [
        ([0, 1], ["json", "decimal"], ["json", "decimal"], [KEEP_VAR, KEEP_VAR]),
        ([0, 1], ["set", "enum"], ["set", "enum"], [KEEP_VAR, KEEP_VAR]),
        ([0, 1], ["enumint", "enum"], ["enumint", "enum"], [NORMAL, NORMAL]),
        ([0, 1], ["set", "enumint"], ["set", "enumint"], [NORMAL, NORMAL]),
    ]
    for i, (ki, b, pk, modes) in enumerate(cases):
        m = meta(ki, b, b, pk)
        assert list(m.serialize_modes[: m.n_serialize_modes]) == modes, f"case {i}"


def test_join_table_meta_row_columns_order():
    # join_table_meta_test.go:217-258 TestJoinTableMetaRowColumnsOrder
    cases = [
        ([0], ["string", "int"], ["string"], None, [], []),
        ([1], ["int", "int"], ["int"], None, [], [1]),
        ([2], ["int", "int", "int"], ["int"], None, [0, 1, 2], [2, 0, 1]),
        ([0], ["string", "string", "date", "decimal"], ["string"], [2, 3], [0, 1, 2, 3], [2, 3, 0, 1]),
        ([0], ["string", "string", "date", "decimal"], ["string"], [3, 2], [0, 1, 2, 3], [3, 2, 0, 1]),
        ([0], ["string", "string", "date", "decimal"], ["string"], [3, 2], [], [3, 2]),
        ([4], ["string", "string", "date", "decimal", "int"], ["int"], [2, 0], [0, 1, 2, 3, 4], [4, 2, 0, 1, 3]),
        ([0], ["string", "string", "date", "decimal", "int"], ["string"], None, [4, 1, 0, 2, 3], [4, 1, 0, 2, 3]),
        ([0], ["string", "string", "date", "decimal", "int"], ["string"], None, None, [0, 1, 2, 3, 4]),
    ]
    for i, (ki, b, keys, other, output, order) in enumerate(cases):
        m = meta(ki, b, keys, keys, other, output)
        assert list(m.row_columns_order[: m.n_row_columns]) == order, f"case {i}"


def test_join_table_meta_null_map_length():
    # join_table_meta_test.go:260-274 TestJoinTableMetaNullMapLength + newTableMeta :232-241:
    # without usedFlag ceil(cols/8); with usedFlag 4-byte aligned ((cols+1+31)/32)*4
    for ncols in (1, 7, 8, 9, 31, 32, 33):
        b = ["int"] * ncols
        assert meta([0], b, ["int"], ["int"], None, None, False).null_map_length == (ncols + 7) // 8
        assert meta([0], b, ["int"], ["int"], None, None, True).null_map_length == ((ncols + 1 + 31) // 32) * 4


def test_config2_build_row_is_32_bytes():
    # int64 key inlined + int64 payload, 2 cols → next_ptr 8 + null_map 1 + 8 + 8 → pad to 32
    m = meta([0], ["int", "int"], ["int"], ["int"], None, None, False)
    assert m.key_mode == ONE_INT64 and m.null_map_length == 1 and m.row_length == 16 and m.row_data_offset == 9


def test_group_key_codec():
    # codec.HashGroupKey (pkg/util/codec/codec.go:1761): NULL → NilFlag(0); int → varintFlag(8) +
    # binary.PutVarint (zig-zag); real → floatFlag(5) + EncodeFloat (cmp-uint64 big endian)
    import ctypes as C
    buf = C.create_string_buffer(16)
    n = L.orc_group_key_int(0, 1, buf); assert buf.raw[:n] == b"\x00"
    n = L.orc_group_key_int(0, 0, buf); assert buf.raw[:n] == b"\x08\x00"
    n = L.orc_group_key_int(1, 0, buf); assert buf.raw[:n] == b"\x08\x02"
    n = L.orc_group_key_int(-1, 0, buf); assert buf.raw[:n] == b"\x08\x01"
    n = L.orc_group_key_int(64, 0, buf); assert buf.raw[:n] == b"\x08\x80\x01"
    n = L.orc_group_key_int(-(1 << 63), 0, buf); assert buf.raw[:n] == b"\x08" + b"\xff" * 9 + b"\x01"
    n = L.orc_group_key_real(0.0, 0, buf); assert buf.raw[:n] == b"\x05\x80" + b"\x00" * 7
    n = L.orc_group_key_real(-0.0, 0, buf); assert buf.raw[:n] == b"\x05\x80" + b"\x00" * 7   # -0 groups with +0
    n = L.orc_group_key_real(1.0, 0, buf); assert buf.raw[:n] == b"\x05\xbf\xf0" + b"\x00" * 6
    n = L.orc_group_key_real(-1.0, 0, buf); assert buf.raw[:n] == b"\x05\x40\x0f" + b"\xff" * 6
