"""chunk.Column / chunk.Chunk in the reference's exact memory layout, on numpy buffers.

Mirrors pkg/util/chunk/column.go:74-82 (Column{length, nullBitmap, offsets, data}) and
pkg/util/chunk/chunk.go:35-54 (Chunk{columns, sel, ...}): fixed-width little-endian `data`,
LSB-first `nullBitmap` where bit 1 means NOT NULL, optional selection vector `sel`.
Only fixed-width columns are modelled (the GPU path's scope).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import abi


def pack_not_null_bitmap(nulls: np.ndarray) -> np.ndarray:
    """bool array (True = NULL) -> Column.nullBitmap bytes (bit 1 = NOT NULL, LSB first)."""
    return np.packbits(~np.asarray(nulls, dtype=bool), bitorder="little")


def unpack_nulls(bitmap: np.ndarray, n: int) -> np.ndarray:
    """Column.nullBitmap bytes -> bool array (True = NULL) of length n."""
    bits = np.unpackbits(np.asarray(bitmap, dtype=np.uint8), bitorder="little")[:n]
    return bits == 0


class Column:
    """One fixed-width chunk.Column.  `data` is a 1-D numpy array of int64/uint64/float64/float32."""

    def __init__(self, data: np.ndarray, nulls: Optional[np.ndarray] = None):
        data = np.ascontiguousarray(data)
        if data.dtype.itemsize not in (4, 8):
            raise ValueError("only 4/8-byte fixed-width columns are modelled")
        self.data = data
        self.length = int(data.shape[0])
        self.elem_len = int(data.dtype.itemsize)
        if nulls is not None:
            nulls = np.asarray(nulls, dtype=bool)
            if nulls.shape[0] != self.length:
                raise ValueError("nulls length mismatch")
            self.null_bitmap: Optional[np.ndarray] = pack_not_null_bitmap(nulls)
        else:
            self.null_bitmap = None

    # Column.IsNull column.go:225
    def is_null(self, i: int) -> bool:
        if self.null_bitmap is None:
            return False
        return ((int(self.null_bitmap[i >> 3]) >> (i & 7)) & 1) == 0

    def nulls(self) -> np.ndarray:
        if self.null_bitmap is None:
            return np.zeros(self.length, dtype=bool)
        return unpack_nulls(self.null_bitmap, self.length)

    def to_struct(self) -> abi.TgColumn:
        s = abi.TgColumn()
        s.length = self.length
        s.null_bitmap = self.null_bitmap.ctypes.data if self.null_bitmap is not None else None
        s.offsets = None
        s.data = self.data.ctypes.data if self.length else None
        s.elem_len = self.elem_len
        return s

    def slice(self, lo: int, hi: int) -> "Column":
        nl = self.nulls()[lo:hi] if self.null_bitmap is not None else None
        return Column(self.data[lo:hi].copy(), nl)


class Chunk:
    """chunk.Chunk: columns + optional sel (logical row i -> physical row sel[i])."""

    def __init__(self, columns: Sequence[Column], sel: Optional[np.ndarray] = None):
        self.columns: List[Column] = list(columns)
        self.sel = None if sel is None else np.ascontiguousarray(sel, dtype=np.int64)
        self._keep = None

    # Chunk.NumRows chunk.go:384
    def num_rows(self) -> int:
        if self.sel is not None:
            return int(self.sel.shape[0])
        return self.columns[0].length if self.columns else 0

    def num_cols(self) -> int:
        return len(self.columns)

    def to_struct(self) -> abi.TgChunk:
        n = len(self.columns)
        arr = (abi.TgColumn * max(n, 1))()
        for i, c in enumerate(self.columns):
            arr[i] = c.to_struct()
        s = abi.TgChunk()
        s.ncols = n
        s.cols = C.cast(arr, C.POINTER(abi.TgColumn))
        s.sel = self.sel.ctypes.data if self.sel is not None else None
        s.nsel = self.num_rows() if self.sel is not None else 0
        self._keep = arr
        return s

    def split(self, max_rows: int) -> List["Chunk"]:
        """Cut into chunks of at most max_rows physical rows (tidb_max_chunk_size = 1024)."""
        assert self.sel is None
        n = self.num_rows()
        out = []
        for lo in range(0, n, max_rows):
            hi = min(n, lo + max_rows)
            out.append(Chunk([c.slice(lo, hi) for c in self.columns]))
        return out


def chunk_array(chunks: Sequence[Chunk]):
    """C array of tg_chunk for a list of chunks (keeps the backing structs alive on the result)."""
    arr = (abi.TgChunk * max(len(chunks), 1))()
    for i, c in enumerate(chunks):
        arr[i] = c.to_struct()
    arr._chunks = list(chunks)  # keep alive
    return arr


class MutChunk:
    """Caller-owned output chunk (tg_mut_chunk): numpy buffers the library fills."""

    def __init__(self, elem_lens: Sequence[int], capacity_rows: int, dtypes: Optional[Sequence] = None):
        self.capacity = int(capacity_rows)
        self.data = []
        self.bitmaps = []
        for i, el in enumerate(elem_lens):
            dt = dtypes[i] if dtypes is not None else (np.int64 if el == 8 else np.float32)
            self.data.append(np.zeros(max(self.capacity, 1), dtype=dt))
            self.bitmaps.append(np.zeros((max(self.capacity, 1) + 7) // 8, dtype=np.uint8))
        self._cols = (abi.TgMutColumn * max(len(elem_lens), 1))()
        for i, el in enumerate(elem_lens):
            self._cols[i].null_bitmap = self.bitmaps[i].ctypes.data
            self._cols[i].data = self.data[i].ctypes.data
            self._cols[i].elem_len = el
        self.struct = abi.TgMutChunk()
        self.struct.ncols = len(elem_lens)
        self.struct.cols = C.cast(self._cols, C.POINTER(abi.TgMutColumn))
        self.struct.capacity_rows = self.capacity

    def columns(self, nrows: int):
        """-> list of (values ndarray[:nrows], nulls bool ndarray[:nrows])"""
        return [(d[:nrows].copy(), unpack_nulls(b, nrows)) for d, b in zip(self.data, self.bitmaps)]
