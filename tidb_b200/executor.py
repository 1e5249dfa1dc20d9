"""Host-side mirror of the reference's executor interface for the GPU operators.

The reference's contract is exec.Executor (pkg/executor/internal/exec/executor.go:51-77):
Open(ctx) / Next(ctx, req *chunk.Chunk) / Close(), children pulled with exec.Next(ctx, child, chk),
zero rows = EOF.  A cgo shim implementing that interface forwards to the C-ABI exactly like the
classes below do (INTEGRATION.md shows the Go source); they exist in Python only because this image
has no Go toolchain.  Names follow the reference: MockDataSource (internal/testutil/testutil.go:63),
HashJoinV2Exec (join/hash_join_v2.go:608), HashAggExec (aggregate/agg_hash_executor.go:93).

There is no CPU fallback: every operator fails if libtidbgpu.so is missing or no CUDA device exists.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import abi
from .chunk import Chunk, Column, MutChunk
from .plan import AggPlan, FieldType, JoinPlan

MAX_CHUNK_SIZE = 1024  # tidb_max_chunk_size default (vardef/tidb_vars.go:1464)


def np_dtype_of(t: FieldType):
    if t.tp == abi.TYPE_DOUBLE:
        return np.float64
    if t.tp == abi.TYPE_FLOAT:
        return np.float32
    return np.int64   # unsigned columns keep their bit pattern (chunk.Column stores raw 8 bytes)


class Executor:
    """exec.Executor: open → next* → close.  next() returns a Chunk; 0 rows means EOF."""

    def __init__(self, schema: Sequence[FieldType], children: Sequence["Executor"] = ()):
        self.schema = list(schema)
        self.children = list(children)

    def open(self) -> None:            # BaseExecutorV2.Open executor.go:301: open the children first
        for c in self.children:
            c.open()

    def next(self, required_rows: int = MAX_CHUNK_SIZE) -> Chunk:
        raise NotImplementedError

    def close(self) -> None:
        for c in self.children:
            c.close()

    def empty_chunk(self) -> Chunk:
        return Chunk([Column(np.zeros(0, dtype=np_dtype_of(t))) for t in self.schema])


class MockDataSource(Executor):
    """Replays pre-generated chunks: the fake backend of every operator test / benchmark
    (pkg/executor/internal/testutil/testutil.go:63 MockDataSource, :269 BuildMockDataSource)."""

    def __init__(self, schema: Sequence[FieldType], chunks: Sequence[Chunk]):
        super().__init__(schema)
        self.chunks = list(chunks)
        self._pos = 0

    def open(self) -> None:
        self._pos = 0

    def next(self, required_rows: int = MAX_CHUNK_SIZE) -> Chunk:
        if self._pos >= len(self.chunks):
            return self.empty_chunk()
        c = self.chunks[self._pos]
        self._pos += 1
        return c


def _out_chunk(schema: Sequence[FieldType], capacity: int) -> MutChunk:
    dts = [np_dtype_of(t) for t in schema]
    return MutChunk([np.dtype(d).itemsize for d in dts], capacity, dts)


class HashJoinExec(Executor):
    """GPU replacement of join.HashJoinV2Exec behind the same Open/Next/Close surface.

    Build side is a pipeline breaker (fetchAndBuildHashTable hash_join_v2.go:1266): the first Next drains
    the build child into tg_join_build_push and calls tg_join_build_finish.  The probe side streams:
    Next pulls probe chunks (fetchProbeSideChunks hash_join_base.go:161) into tg_join_probe_push until
    tg_join_next has rows to hand out (hash_join_v2.go:1176-1186)."""

    def __init__(self, plan: JoinPlan, left: Executor, right: Executor):
        super().__init__(plan.out_schema(), [left, right])
        self.plan = plan
        self.build_child, self.probe_child = (right, left) if plan.build_is_right else (left, right)
        self._h = C.c_void_p()
        self._lib = None
        self._prepared = False
        self._probe_done = False
        self._out: Optional[MutChunk] = None

    def open(self) -> None:
        super().open()
        self._lib = abi.load_lib()
        desc, self._keep = self.plan.to_struct()
        self._h = C.c_void_p()
        abi.check(self._lib.tg_join_open(C.byref(desc), C.byref(self._h)))
        self._prepared = False
        self._probe_done = False

    def _build(self) -> None:
        while True:
            chk = self.build_child.next(MAX_CHUNK_SIZE)
            if chk.num_rows() == 0:
                break
            cs = chk.to_struct()
            abi.check(self._lib.tg_join_build_push(self._h, C.byref(cs)))
        abi.check(self._lib.tg_join_build_finish(self._h))
        self._prepared = True

    def next(self, required_rows: int = MAX_CHUNK_SIZE) -> Chunk:
        if not self._h:
            raise RuntimeError("next before open")
        if not self._prepared:
            self._build()
        if self._out is None or self._out.capacity < required_rows:
            self._out = _out_chunk(self.schema, max(required_rows, 8))
        n = C.c_int64(0)
        while True:
            abi.check(self._lib.tg_join_next(self._h, C.byref(self._out.struct), C.c_int64(required_rows), C.byref(n)))
            if n.value > 0 or self._probe_done:
                break
            chk = self.probe_child.next(MAX_CHUNK_SIZE)
            if chk.num_rows() == 0:
                abi.check(self._lib.tg_join_probe_finish(self._h))
                self._probe_done = True
            else:
                cs = chk.to_struct()
                abi.check(self._lib.tg_join_probe_push(self._h, C.byref(cs)))
        cols = self._out.columns(n.value)
        return Chunk([Column(v, nl if nl.any() else None) for v, nl in cols])

    def stats(self) -> abi.TgJoinStats:
        s = abi.TgJoinStats()
        abi.check(self._lib.tg_join_get_stats(self._h, C.byref(s)))
        return s

    def close(self) -> None:
        if self._h:
            self._lib.tg_join_close(self._h)
            self._h = C.c_void_p()
        super().close()


class HashAggExec(Executor):
    """GPU replacement of aggregate.HashAggExec.  Output schema: one column per aggregate function in
    descriptor order (group columns are emitted through firstrow() funcs, as the reference's plans do)."""

    def __init__(self, plan: AggPlan, child: Executor, out_schema: Optional[Sequence[FieldType]] = None):
        super().__init__(out_schema or [self._ret_type(plan, f) for f in plan.funcs], [child])
        self.plan = plan
        self._h = C.c_void_p()
        self._lib = None
        self._prepared = False
        self._out: Optional[MutChunk] = None

    @staticmethod
    def _ret_type(plan: AggPlan, f) -> FieldType:
        if f.name == abi.AGG_COUNT:
            return FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)
        if f.name in (abi.AGG_SUM, abi.AGG_AVG):
            return FieldType(abi.TYPE_DOUBLE, 0)
        return FieldType(plan.col_types[f.arg_col].tp, plan.col_types[f.arg_col].flag & ~abi.FLAG_NOT_NULL)

    def open(self) -> None:
        super().open()
        self._lib = abi.load_lib()
        desc, self._keep = self.plan.to_struct()
        self._h = C.c_void_p()
        abi.check(self._lib.tg_agg_open(C.byref(desc), C.byref(self._h)))
        self._prepared = False

    def next(self, required_rows: int = MAX_CHUNK_SIZE) -> Chunk:
        if not self._prepared:
            child = self.children[0]
            while True:   # fetchChildData agg_hash_executor.go:449
                chk = child.next(MAX_CHUNK_SIZE)
                if chk.num_rows() == 0:
                    break
                cs = chk.to_struct()
                abi.check(self._lib.tg_agg_push(self._h, C.byref(cs)))
            abi.check(self._lib.tg_agg_finish(self._h))
            self._prepared = True
        if self._out is None or self._out.capacity < required_rows:
            self._out = _out_chunk(self.schema, max(required_rows, 8))
        n = C.c_int64(0)
        abi.check(self._lib.tg_agg_next(self._h, C.byref(self._out.struct), C.c_int64(required_rows), C.byref(n)))
        cols = self._out.columns(n.value)
        return Chunk([Column(v, nl if nl.any() else None) for v, nl in cols])

    def stats(self) -> abi.TgAggStats:
        s = abi.TgAggStats()
        abi.check(self._lib.tg_agg_get_stats(self._h, C.byref(s)))
        return s

    def close(self) -> None:
        if self._h:
            self._lib.tg_agg_close(self._h)
            self._h = C.c_void_p()
        super().close()


def drain(e: Executor, required_rows: int = MAX_CHUNK_SIZE) -> List[Chunk]:
    """open → next until EOF → close, like the reference's test helpers."""
    e.open()
    out = []
    try:
        while True:
            c = e.next(required_rows)
            if c.num_rows() == 0:
                break
            out.append(c)
    finally:
        e.close()
    return out
