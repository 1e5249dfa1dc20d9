"""TopNExec on the device (tg_topn, csrc/topn.cu) against the oracle restatement of sortexec.TopNExec's order
(oracle/topn.py), through the executor shim's Open / Next / Close.  Ties are unordered in the reference (heap), so rows
are compared on their ORDER BY columns position by position and as a multiset only when the boundary has no tie."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import topn as OT   # noqa: E402
from tidb_b200 import abi   # noqa: E402
from tidb_b200.chunk import Chunk, Column   # noqa: E402
from tidb_b200.executor import MockDataSource, TopNExec, drain   # noqa: E402
from tidb_b200.plan import FieldType   # noqa: E402

pytestmark = pytest.mark.gpu
INT = FieldType(abi.TYPE_LONGLONG, 0)
UINT = FieldType(abi.TYPE_LONGLONG, abi.FLAG_UNSIGNED)
DBL = FieldType(abi.TYPE_DOUBLE, 0)


def _rows(chunks):
    out = []
    for c in chunks:
        cols = [(col.data, col.nulls()) for col in c.columns]
        for i in range(c.num_rows()):
            out.append(tuple(None if nl[i] else v[i].item() for v, nl in cols))
    return out


def _check(tbl, schema, kinds, by, offset, count, required_rows=1024):
    got = _rows(drain(TopNExec(MockDataSource(schema, tbl.split(1024)), by, offset, count), required_rows))
    exp = OT.topn_rows(_rows([tbl]), kinds, by, offset, count)
    assert len(got) == len(exp)
    key = lambda r: tuple(r[c] for c, _ in by)
    same = lambda a, b: all((x is None and y is None) or (x is not None and y is not None and (x == y or (x != x and y != y))) for x, y in zip(a, b))
    for g, e in zip(got, exp):
        assert same(key(g), key(e)), (g, e)      # same ORDER BY values at every output position
    return got, exp


def test_topn_hand_cases():
    # ORDER BY a DESC, b LIMIT 2,3 on a tiny table with NULLs: NULL sorts first ascending, last descending
    a = np.array([5, 3, 5, 1, 9, 7, 5], dtype=np.int64); an = np.array([False, False, False, True, False, False, False])
    b = np.array([1.5, 2.0, 0.5, 9.0, -1.0, 4.0, 0.25]); bn = np.array([False] * 6 + [True])
    tbl = Chunk([Column(a, an), Column(b, bn)])
    got, exp = _check(tbl, [INT, DBL], ["int", "real"], [(0, True), (1, False)], 2, 3)
    assert got == exp == [(5, None), (5, 0.5), (5, 1.5)]
    got, exp = _check(tbl, [INT, DBL], ["int", "real"], [(0, False)], 0, 2)
    assert [g[0] for g in got] == [None, 3]
    assert _rows(drain(TopNExec(MockDataSource([INT, DBL], [tbl]), [(0, False)], 100, 5))) == []     # offset beyond the input
    assert len(_rows(drain(TopNExec(MockDataSource([INT, DBL], [tbl]), [(1, True)], 0, 100)))) == 7  # count beyond the input


@pytest.mark.parametrize("by,offset,count", [([(2, True), (0, False)], 0, 10), ([(0, False)], 1000, 50), ([(1, True), (2, False), (0, True)], 5, 2000),
                                             ([(3, False)], 0, 17)])
def test_topn_random_vs_oracle(by, offset, count):
    rng = np.random.default_rng(50 + offset + count)
    n = 200_000
    a = rng.integers(-1000, 1000, n).astype(np.int64); an = rng.random(n) < 0.01                  # heavy ties + NULLs
    u = rng.integers(0, 1 << 63, n).astype(np.int64) * 2 + rng.integers(0, 2, n); un = rng.random(n) < 0.01   # full uint64 range
    x = np.floor(rng.random(n) * 1e9) / 1000 - 3e5; xn = rng.random(n) < 0.02
    x[:4] = [np.nan, -0.0, np.inf, -np.inf]
    y = rng.integers(0, 5, n).astype(np.int64)                                                    # 5 distinct values: the threshold ties massively
    tbl = Chunk([Column(a, an), Column(u, un), Column(x, xn), Column(y)])
    _check(tbl, [INT, UINT, DBL, INT], ["int", "uint", "real", "int"], by, offset, count, required_rows=333)


def test_topn_q3_shape_result():
    # ORDER BY revenue DESC, o_orderdate LIMIT 10 (TPC-H Q3's TopN, tpch_suite_out.json:102): distinct revenues -> exact rows
    rng = np.random.default_rng(9)
    n = 1_000_000
    rev = rng.permutation(n).astype(np.float64) * 0.37
    tbl = Chunk([Column(np.arange(n, dtype=np.int64)), Column(rev), Column(rng.integers(0, 2406, n).astype(np.int64))])
    got, exp = _check(tbl, [INT, DBL, INT], ["int", "real", "int"], [(1, True), (2, False)], 0, 10)
    assert got == exp
