"""Test-side loader for oracle/liboracle.so (TEST INFRASTRUCTURE — see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

from tidb_b200 import abi
from tidb_b200.chunk import Chunk, Column, MutChunk, chunk_array
from tidb_b200.plan import AggPlan, FilterItem, JoinPlan, filter_array

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORC_DIR = os.path.join(_ROOT, "oracle")
_ORC_SO = os.path.join(_ORC_DIR, "liboracle.so")
_lib = None


class OrcTableMeta(C.Structure):
    _fields_ = [("key_mode", C.c_int32), ("is_keys_inlined", C.c_int32),
                ("is_keys_fixed_length", C.c_int32), ("join_keys_length", C.c_int32),
                ("null_map_length", C.c_int32), ("row_length", C.c_int32),
                ("is_fixed_length", C.c_int32), ("row_data_offset", C.c_int32),
                ("n_row_columns", C.c_int32), ("row_columns_order", C.c_int32 * 64),
                ("n_serialize_modes", C.c_int32), ("serialize_modes", C.c_int32 * 16),
                ("column_count_needed_for_other_condition", C.c_int32)]


def build_oracle() -> str:
    srcs = [os.path.join(_ORC_DIR, f) for f in ("join.cpp", "agg.cpp", "vec.cpp", "common.hpp", "oracle.h")]
    srcs.append(os.path.join(_ROOT, "include", "tidbgpu.h"))
    if (not os.path.exists(_ORC_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_ORC_SO) for s in srcs):
        subprocess.run(["make", "-C", _ORC_DIR, "liboracle.so"], check=True, capture_output=True)
    return _ORC_SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_fnv1_64.restype = C.c_uint64
        _lib.orc_fnv1_64.argtypes = [C.c_char_p, C.c_size_t]
        _lib.orc_next_power_of_two.restype = C.c_uint64
        _lib.orc_next_power_of_two.argtypes = [C.c_uint64]
        _lib.orc_hash_table_length.restype = C.c_uint64
        _lib.orc_hash_table_length.argtypes = [C.c_uint64]
        _lib.orc_partition_number.restype = C.c_uint32
        _lib.orc_partition_number.argtypes = [C.c_uint32]
        _lib.orc_partition_mask_offset.restype = C.c_int32
        _lib.orc_partition_mask_offset.argtypes = [C.c_uint32]
        _lib.orc_tagged_bits.restype = C.c_uint8
        _lib.orc_tagged_bits.argtypes = [C.c_uint64]
        _lib.orc_tagged_mask.restype = C.c_uint64
        _lib.orc_tagged_mask.argtypes = [C.c_uint8]
        _lib.orc_group_key_int.argtypes = [C.c_int64, C.c_int, C.c_char_p]
        _lib.orc_group_key_real.argtypes = [C.c_double, C.c_int, C.c_char_p]
        for f in ("orc_join_result_rows", "orc_join_row_count", "orc_join_total_row_bytes",
                  "orc_join_hash_table_slots", "orc_agg_result_rows"):
            getattr(_lib, f).restype = C.c_int64
        for f in ("orc_join_build_seconds", "orc_join_probe_seconds", "orc_agg_seconds"):
            getattr(_lib, f).restype = C.c_double
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_vec_compare_int.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p]
        _lib.orc_vec_compare_real.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        _lib.orc_vec_arith_int.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_void_p]
        _lib.orc_vec_arith_real.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    return _lib


def _err() -> str:
    m = lib().orc_last_error()
    return m.decode() if m else ""


class OracleJoin:
    """Runs the HashJoinV2Exec restatement on lists of chunks; returns columns as (values, nulls)."""

    def __init__(self, plan: JoinPlan, concurrency: int = 5):
        self.plan = plan
        self._desc, self._keep = plan.to_struct()
        self._h = C.c_void_p()
        rc = lib().orc_join_open(C.byref(self._desc), C.c_int32(concurrency), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"oracle join open failed ({rc}): {_err()}")

    def run(self, build: Sequence[Chunk], probe: Sequence[Chunk]):
        ba, pa = chunk_array(build), chunk_array(probe)
        rc = lib().orc_join_run(self._h, ba, C.c_int64(len(build)), pa, C.c_int64(len(probe)))
        if rc != 0:
            raise RuntimeError(f"oracle join run failed ({rc}): {_err()}")
        n = lib().orc_join_result_rows(self._h)
        schema = self.plan.out_schema()
        dts = [_np_dtype(t) for t in schema]
        out = MutChunk([np.dtype(d).itemsize for d in dts], n, dts)
        rc = lib().orc_join_result_fetch(self._h, C.byref(out.struct))
        if rc != 0:
            raise RuntimeError(f"oracle join fetch failed ({rc}): {_err()}")
        return n, out.columns(n)

    def build(self, build: Sequence[Chunk]) -> None:
        ba = chunk_array(build)
        rc = lib().orc_join_build(self._h, ba, C.c_int64(len(build)))
        if rc != 0:
            raise RuntimeError(f"oracle join build failed ({rc}): {_err()}")

    def probe(self, probe_arr, n_chunks: int) -> int:
        """probe_arr: a prepared chunk_array (so that marshalling stays outside timed regions)"""
        rc = lib().orc_join_probe(self._h, probe_arr, C.c_int64(n_chunks))
        if rc != 0:
            raise RuntimeError(f"oracle join probe failed ({rc}): {_err()}")
        return lib().orc_join_result_rows(self._h)

    def stat(self, name: str):
        return getattr(lib(), "orc_join_" + name)(self._h)

    def close(self):
        if self._h:
            lib().orc_join_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleAgg:
    def __init__(self, plan: AggPlan, partial_concurrency: int = 5, final_concurrency: int = 5):
        self.plan = plan
        self._desc, self._keep = plan.to_struct()
        self._h = C.c_void_p()
        rc = lib().orc_agg_open(C.byref(self._desc), C.c_int32(partial_concurrency), C.c_int32(final_concurrency),
                                C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"oracle agg open failed ({rc}): {_err()}")

    def run(self, chunks: Sequence[Chunk]):
        ca = chunk_array(chunks)
        rc = lib().orc_agg_run(self._h, ca, C.c_int64(len(chunks)))
        if rc != 0:
            raise RuntimeError(f"oracle agg run failed ({rc}): {_err()}")
        n = lib().orc_agg_result_rows(self._h)
        dts = [agg_out_dtype(f) for f in self.plan.funcs]
        out = MutChunk([8] * len(dts), n, dts)
        rc = lib().orc_agg_result_fetch(self._h, C.byref(out.struct))
        if rc != 0:
            raise RuntimeError(f"oracle agg fetch failed ({rc}): {_err()}")
        return n, out.columns(n)

    def seconds(self) -> float:
        return lib().orc_agg_seconds(self._h)

    def close(self):
        if self._h:
            lib().orc_agg_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def agg_out_dtype(f) -> type:
    if f.name == abi.AGG_COUNT:
        return np.int64
    if f.name in (abi.AGG_SUM, abi.AGG_AVG):
        return np.float64
    return np.float64 if f.arg_type in (abi.TYPE_DOUBLE, abi.TYPE_FLOAT) else np.int64


def _np_dtype(t) -> type:
    if t.tp == abi.TYPE_DOUBLE:
        return np.float64
    if t.tp == abi.TYPE_FLOAT:
        return np.float32
    return np.int64


# ---- VecEval wrappers ------------------------------------------------------------------------------
def _colptr(col: Optional[Column]):
    if col is None:
        return None, None
    s = col.to_struct()
    return C.addressof(s), s


def vec_compare_int(op, a: Column, b: Optional[Column], b_const=0, a_unsigned=False, b_unsigned=False):
    n = a.length
    res = np.zeros(n, dtype=np.int64); nulls = np.zeros((n + 7) // 8, dtype=np.uint8)
    pa, ka = _colptr(a); pb, kb = _colptr(b)
    rc = lib().orc_vec_compare_int(op, int(a_unsigned), int(b_unsigned), pa, pb, C.c_int64(b_const),
                                   C.c_void_p(res.ctypes.data), C.c_void_p(nulls.ctypes.data))
    assert rc == 0, _err()
    return res, np.unpackbits(nulls, bitorder="little")[:n] == 0


def vec_compare_real(op, a: Column, b: Optional[Column], b_const=0.0):
    n = a.length
    res = np.zeros(n, dtype=np.int64); nulls = np.zeros((n + 7) // 8, dtype=np.uint8)
    pa, ka = _colptr(a); pb, kb = _colptr(b)
    rc = lib().orc_vec_compare_real(op, pa, pb, C.c_double(b_const), C.c_void_p(res.ctypes.data), C.c_void_p(nulls.ctypes.data))
    assert rc == 0, _err()
    return res, np.unpackbits(nulls, bitorder="little")[:n] == 0


def vec_arith_int(op, a: Column, b: Optional[Column], b_const=0, a_unsigned=False, b_unsigned=False):
    n = a.length
    res = np.zeros(n, dtype=np.int64); nulls = np.zeros((n + 7) // 8, dtype=np.uint8)
    pa, ka = _colptr(a); pb, kb = _colptr(b)
    rc = lib().orc_vec_arith_int(op, int(a_unsigned), int(b_unsigned), pa, pb, C.c_int64(b_const),
                                 C.c_void_p(res.ctypes.data), C.c_void_p(nulls.ctypes.data))
    return rc, res, np.unpackbits(nulls, bitorder="little")[:n] == 0


def vec_arith_real(op, a: Column, b: Optional[Column], b_const=0.0):
    n = a.length
    res = np.zeros(n, dtype=np.float64); nulls = np.zeros((n + 7) // 8, dtype=np.uint8)
    pa, ka = _colptr(a); pb, kb = _colptr(b)
    rc = lib().orc_vec_arith_real(op, pa, pb, C.c_double(b_const), C.c_void_p(res.ctypes.data), C.c_void_p(nulls.ctypes.data))
    return rc, res, np.unpackbits(nulls, bitorder="little")[:n] == 0


def vec_filter(chk: Chunk, items: Sequence[FilterItem]):
    phys = chk.columns[0].length
    sel = np.zeros(phys, dtype=np.uint8)
    cnt = C.c_int64(0)
    cs = chk.to_struct()
    fa = filter_array(items)
    rc = lib().orc_vec_filter(C.byref(cs), fa, C.c_int32(len(items)), C.c_void_p(sel.ctypes.data), C.byref(cnt))
    assert rc == 0, _err()
    return sel.astype(bool), cnt.value
