"""Nested-loop join generators: an independent restatement of the reference's own test oracles.

  genInnerJoinResult            pkg/executor/join/inner_join_probe_test.go:80-125
  genLeftOuterJoinResult        left_outer_join_probe_test.go:33-105   (right outer: mirrored)
  genAntiSemiJoinResult         anti_semi_join_probe_test.go:35-99
  genLeftOuterSemiOrSemiJoin... left_outer_semi_join_probe_test.go:47-165
  checkChunksEqual              inner_join_probe_test.go:137-193  (sorted row multisets)

Rows are Python tuples with None for NULL.  Used to cross-check oracle/join.cpp (hash-table
restatement) at small sizes; never used by the product.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from tidb_b200 import abi
from tidb_b200.chunk import Chunk
from tidb_b200.plan import FilterItem, JoinPlan

Row = Tuple


def chunk_rows(chk: Chunk) -> List[Row]:
    """Logical rows of a chunk as tuples (None = NULL), honouring sel."""
    n = chk.num_rows()
    cols = []
    for c in chk.columns:
        nl = c.nulls()
        vals = c.data
        cols.append((vals, nl))
    rows = []
    for l in range(n):
        p = int(chk.sel[l]) if chk.sel is not None else l
        rows.append(tuple(None if nl[p] else vals[p].item() for vals, nl in cols))
    return rows


def _cmp(op: int, a, b) -> bool:
    return {abi.CMP_LT: a < b, abi.CMP_LE: a <= b, abi.CMP_GT: a > b, abi.CMP_GE: a >= b,
            abi.CMP_EQ: a == b, abi.CMP_NE: a != b}[op]


def filter_ok(row: Row, items: Sequence[FilterItem]) -> bool:
    for it in items:
        a = row[it.lhs_col]
        b = row[it.rhs_col] if it.rhs_col >= 0 else (it.const_f64 if it.is_real else it.const_i64)
        if a is None or b is None or not _cmp(it.op, a, b):
            return False
    return True


def _key(row: Row, idx: Sequence[int], types, float_norm=True):
    k = []
    for i in idx:
        v = row[i]
        if v is None:
            return None
        if isinstance(v, float) and v == 0:
            v = 0.0
        if types[i].tp in (abi.TYPE_DATE, abi.TYPE_DATETIME, abi.TYPE_TIMESTAMP):
            # date-time keys compare by their calendar fields (codec.go:697-707 serializes Time.ToPackedUint); the low 4 bits
            # of the CoreTime word are type / fsp (types/time.go:243-251)
            w = int(v) & ((1 << 64) - 1)
            v = ("time", w >> 50, (w >> 46) & 0xf, (w >> 41) & 0x1f, (w >> 36) & 0x1f, (w >> 30) & 0x3f, (w >> 24) & 0x3f, (w >> 4) & 0xfffff)
        k.append(v)
    return tuple(k)


def nested_loop_join(plan: JoinPlan, left: Sequence[Chunk], right: Sequence[Chunk]) -> List[Row]:
    lrows = [r for c in left for r in chunk_rows(c)]
    rrows = [r for c in right for r in chunk_rows(c)]
    lu = plan.lused if plan.lused is not None else list(range(len(plan.left_types)))
    ru = plan.rused if plan.rused is not None else list(range(len(plan.right_types)))
    # filters live on the build / probe child
    lfilter = plan.probe_filter if plan.build_is_right else plan.build_filter
    rfilter = plan.build_filter if plan.build_is_right else plan.probe_filter
    lkeys = [(_key(r, plan.left_keys, plan.left_types) if filter_ok(r, lfilter) else None) for r in lrows]
    rkeys = [(_key(r, plan.right_keys, plan.right_types) if filter_ok(r, rfilter) else None) for r in rrows]
    lpass = [filter_ok(r, lfilter) for r in lrows]
    rpass = [filter_ok(r, rfilter) for r in rrows]
    # mixed signed/unsigned keys compare by value: a negative signed never equals an unsigned
    def norm(k, types, idx):
        if k is None:
            return None
        out = []
        for v, i in zip(k, idx):
            t = types[i]
            if t.tp != abi.TYPE_DOUBLE and t.tp != abi.TYPE_FLOAT and t.unsigned and v < 0:
                v += 1 << 64
            out.append(v)
        return tuple(out)
    lkeys = [norm(k, plan.left_types, plan.left_keys) for k in lkeys]
    rkeys = [norm(k, plan.right_types, plan.right_keys) for k in rkeys]

    def out_row(l: Optional[Row], r: Optional[Row]) -> Row:
        a = tuple(l[i] for i in lu) if l is not None else tuple(None for _ in lu)
        b = tuple(r[i] for i in ru) if r is not None else tuple(None for _ in ru)
        return a + b

    def pair_ok(l: Row, r: Row) -> bool:
        """OtherCondition on one candidate pair: every CNF item non-NULL true (inner_join_probe_test.go builds its expected
        rows the same way: key equality AND the other condition on the joined row)"""
        for it in getattr(plan, "other_cond", []) or []:
            a = (l if it.lhs_side == 0 else r)[it.lhs_col]
            b = ((l if it.rhs_side == 0 else r)[it.rhs_col]) if it.rhs_side >= 0 else (it.const_f64 if it.is_real else it.const_i64)
            if a is None or b is None or not _cmp(it.op, a, b):
                return False
        return True

    jt = plan.join_type
    res: List[Row] = []
    rindex = {}
    for j, k in enumerate(rkeys):
        if k is not None:
            rindex.setdefault(k, []).append(j)
    if jt == abi.JOIN_INNER:
        for i, l in enumerate(lrows):
            for j in rindex.get(lkeys[i], []) if lkeys[i] is not None else []:
                if pair_ok(l, rrows[j]):
                    res.append(out_row(l, rrows[j]))
    elif jt == abi.JOIN_LEFT_OUTER:
        for i, l in enumerate(lrows):
            ms = rindex.get(lkeys[i], []) if lkeys[i] is not None else []
            ms = [j for j in ms if pair_ok(l, rrows[j])]
            for j in ms:
                res.append(out_row(l, rrows[j]))
            if not ms:
                res.append(out_row(l, None))
    elif jt == abi.JOIN_RIGHT_OUTER:
        lindex = {}
        for i, k in enumerate(lkeys):
            if k is not None:
                lindex.setdefault(k, []).append(i)
        for j, r in enumerate(rrows):
            ms = lindex.get(rkeys[j], []) if rkeys[j] is not None else []
            ms = [i for i in ms if pair_ok(lrows[i], r)]
            for i in ms:
                res.append(out_row(lrows[i], r))
            if not ms:
                res.append(out_row(None, r))
    elif jt in (abi.JOIN_SEMI, abi.JOIN_ANTI_SEMI):
        for i, l in enumerate(lrows):
            m = lkeys[i] is not None and any(pair_ok(l, rrows[j]) for j in rindex.get(lkeys[i], []))
            if (jt == abi.JOIN_SEMI) == m:
                # semi: left rows removed by the left filter never match; anti: they are results
                res.append(tuple(l[c] for c in lu))
    elif jt in (abi.JOIN_LEFT_OUTER_SEMI, abi.JOIN_ANTI_LEFT_OUTER_SEMI):
        anti = jt == abi.JOIN_ANTI_LEFT_OUTER_SEMI
        for i, l in enumerate(lrows):
            m = lkeys[i] is not None and lkeys[i] in rindex
            flag = (0 if m else 1) if anti else (1 if m else 0)
            res.append(tuple(l[c] for c in lu) + (flag,))
    else:
        raise ValueError("join type")
    return res


def columns_to_rows(ncols_vals_nulls) -> List[Row]:
    """[(values, nulls)] per column -> list of row tuples with None for NULL"""
    if not ncols_vals_nulls:
        return []
    n = len(ncols_vals_nulls[0][0])
    cols = [[None if nl[i] else v[i].item() for i in range(n)] for v, nl in ncols_vals_nulls]
    return list(zip(*cols)) if cols else []


def sort_rows(rows: Sequence[Row]) -> List[Row]:
    """checkChunksEqual ordering: NULL first, then by value, column by column."""
    def key(r):
        return tuple((0, 0) if v is None else (1, v) for v in r)
    return sorted(rows, key=key)


def assert_rows_equal(expected: Sequence[Row], got: Sequence[Row]) -> None:
    assert len(expected) == len(got), f"row count {len(got)} != expected {len(expected)}"
    e, g = sort_rows(expected), sort_rows(got)
    for i, (a, b) in enumerate(zip(e, g)):
        assert a == b, f"row {i}: got {b}, expected {a}"


def columns_sorted(cols) -> np.ndarray:
    """Fast multiset canonical form for large all-int64 results: lexsorted 2-D array.

    NULLs are mapped through a parallel flag column so that NULL != 0."""
    if not cols:
        return np.zeros((0, 0), dtype=np.int64)
    mats = []
    for v, nl in cols:
        vv = np.asarray(v).view(np.int64) if np.asarray(v).dtype.itemsize == 8 else np.asarray(v).astype(np.int64)
        vv = np.where(nl, 0, vv)
        mats.append(nl.astype(np.int64))
        mats.append(vv)
    m = np.stack(mats, axis=1)
    order = np.lexsort(m.T[::-1])
    return m[order]
