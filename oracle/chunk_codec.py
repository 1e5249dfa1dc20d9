"""TEST INFRASTRUCTURE — CPU restatement of TiDB's chunk wire codec, the checker for tg_chunk_encode / tg_chunk_decode.

Follows pkg/util/chunk/codec.go: Encode :41, encodeColumn :49-76, DecodeToChunk :93, decodeColumn :101-140,
setAllNotNull :145-152, getFixedLen :165-179.  Pure Python/numpy, small inputs only.  Pinned by: the round trip of the
reference's TestCodec (codec_test.go:27-78: all-NULL int64 column, int64, two var-len columns, 40-byte decimal) and the
hand-assembled byte strings in tests/test_chunk_codec.py.  Parity pinned by KAT, not by the running Go binary.
"""
import struct
from typing import List, Optional, Tuple

import numpy as np

# mysql type codes (parser/mysql/type.go:17-48) with a fixed chunk width (codec.go:165-179)
_FIXED = {4: 4, 1: 8, 2: 8, 9: 8, 3: 8, 8: 8, 5: 8, 13: 8, 11: 8, 10: 8, 12: 8, 7: 8, 0xF6: 40}


def fixed_len(mysql_type: int) -> int:
    return _FIXED.get(mysql_type, -1)


class WireColumn:
    """length, NOT-NULL bitmap (bytes, bit 1 = not null, LSB first) or None, offsets (int64[length+1]) or None, data bytes"""

    def __init__(self, length: int, bitmap: Optional[bytes], offsets: Optional[np.ndarray], data: bytes):
        self.length, self.bitmap, self.offsets, self.data = length, bitmap, offsets, data

    def null_count(self) -> int:   # Column.nullCount, column.go:243
        if self.bitmap is None:
            return 0
        bits = np.unpackbits(np.frombuffer(self.bitmap, dtype=np.uint8), bitorder="little")[:self.length]
        return int(self.length - bits.sum())


def encode_column(col: WireColumn) -> bytes:   # codec.go:49-76
    out = struct.pack("<I", col.length) + struct.pack("<I", col.null_count())
    if col.null_count() > 0:
        out += col.bitmap[:(col.length + 7) // 8]
    if col.offsets is not None:
        out += np.asarray(col.offsets, dtype="<i8")[:col.length + 1].tobytes()
    return out + col.data


def encode(cols: List[WireColumn]) -> bytes:   # codec.go:41-47
    return b"".join(encode_column(c) for c in cols)


def decode(buf: bytes, mysql_types: List[int]) -> Tuple[List[WireColumn], bytes]:   # codec.go:93-140
    cols = []
    for tp in mysql_types:
        length, nulls = struct.unpack_from("<II", buf, 0)
        buf = buf[8:]
        nb = (length + 7) // 8
        if nulls > 0:
            bitmap, buf = buf[:nb], buf[nb:]
        else:
            bitmap = b"\xff" * nb              # setAllNotNull (:145): an all-ones bitmap is materialised
        fl = fixed_len(tp)
        offsets = None
        if fl == -1:
            offsets = np.frombuffer(buf[:(length + 1) * 8], dtype="<i8").copy()
            buf = buf[(length + 1) * 8:]
            nbytes = int(offsets[length])
        else:
            nbytes = fl * length
        cols.append(WireColumn(length, bitmap, offsets, buf[:nbytes]))
        buf = buf[nbytes:]
    return cols, buf
