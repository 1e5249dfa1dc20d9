"""TEST INFRASTRUCTURE (oracle): CPU restatement of TopNExec's row order, used only by tests/.

Follows pkg/executor/sortexec/topn.go: rows are compared item by item with the column type's CompareFunc
(pkg/util/chunk/compare.go:64 cmpNull — NULL before every value; :74 cmpInt64, :82 cmpUint64, :112 cmpFloat64 = Go
cmp.Compare, NaN before everything) and DESC negates the comparison (topn.go:157 greaterRow); the result is rows
[offset, offset + count) of that order (topn.go:285 heap of offset+count rows, :346 final sort).  Ties keep no
particular order in the reference (heap), so callers compare the ORDER BY columns of tied rows, not their identity.
Parity pinning: no golden vectors exist in the reference for TopN beyond SQL results on tiny tables
(tests/integrationtest/r/executor/sort.result-style); pinned by this restatement + the hand cases in tests/test_topn.py.
"""
from __future__ import annotations

import functools
import math
from typing import List, Sequence, Tuple


def _cmp_value(a, b, kind: str) -> int:
    if kind == "real":          # Go cmp.Compare on float64: NaN < everything, NaN == NaN
        an, bn = isinstance(a, float) and math.isnan(a), isinstance(b, float) and math.isnan(b)
        if an or bn:
            return 0 if (an and bn) else (-1 if an else 1)
    return -1 if a < b else (1 if a > b else 0)


def topn_rows(rows: Sequence[Tuple], kinds: Sequence[str], by_items: Sequence[Tuple[int, bool]], offset: int, count: int) -> List[Tuple]:
    """rows: tuples with None for NULL; kinds[c] in {"int", "uint", "real"}; by_items: (column, desc)"""
    def cmp_rows(x, y):
        for col, desc in by_items:
            a, b = x[col], y[col]
            if a is None or b is None:
                c = 0 if (a is None and b is None) else (-1 if a is None else 1)     # cmpNull
            else:
                if kinds[col] == "uint":
                    a, b = a % (1 << 64), b % (1 << 64)
                c = _cmp_value(a, b, kinds[col])
            if desc:
                c = -c
            if c:
                return c
        return 0
    ordered = sorted(rows, key=functools.cmp_to_key(cmp_rows))
    return ordered[offset:offset + count]
