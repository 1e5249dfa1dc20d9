"""SelectionExec / ProjectionExec shims (tidb_b200/executor.py) against the oracle's row-at-a-time evaluators
(oracle/vec.cpp restating expression.VectorizedFilter chunk_executor.go:413 and the builtin*Sig.vecEval* functions),
through the Open / Next / Close contract of exec.Executor (executor.go:51-77): at most RequiredRows rows per Next,
0 rows = EOF, Open again after Close re-executes, an overflow on a non-NULL row fails the call (types.ErrOverflow)."""
import numpy as np
import pytest

import oracle_lib as O
from tidb_b200 import abi
from tidb_b200.chunk import Chunk, Column
from tidb_b200.executor import MockDataSource, ProjectionExec, SelectionExec, drain
from tidb_b200.plan import ColRef, Const, FieldType, FilterItem, ScalarFunc

pytestmark = pytest.mark.gpu

INT = FieldType(abi.TYPE_LONGLONG, 0)
DBL = FieldType(abi.TYPE_DOUBLE, 0)


def _table(rng, n):
    a = rng.integers(-50, 50, n).astype(np.int64); an = rng.random(n) < 0.1
    b = rng.integers(-50, 50, n).astype(np.int64); bn = rng.random(n) < 0.1
    x = np.floor(rng.random(n) * 1000) / 8; xn = rng.random(n) < 0.1
    y = np.floor(rng.random(n) * 10) / 100; yn = rng.random(n) < 0.05
    return Chunk([Column(a, an), Column(b, bn), Column(x, xn), Column(y, yn)]), [INT, INT, DBL, DBL]


def _rows(chunks):
    out = []
    for c in chunks:
        cols = [(col.data, col.nulls()) for col in c.columns]
        for i in range(c.num_rows()):
            out.append(tuple(None if nl[i] else v[i].item() for v, nl in cols))
    return out


@pytest.mark.parametrize("required_rows", [1024, 100, 7])
def test_selection_exec_vs_oracle(required_rows):
    rng = np.random.default_rng(3)
    tbl, schema = _table(rng, 20_000)
    filters = [FilterItem(abi.CMP_GT, 0, const_i64=-20), FilterItem(abi.CMP_NE, 0, rhs_col=1), FilterItem(abi.CMP_LT, 2, is_real=True, const_f64=100.0)]
    e = SelectionExec(MockDataSource(schema, tbl.split(1024)), filters, batch_rows=5000)
    chunks = drain(e, required_rows)
    assert all(0 < c.num_rows() <= required_rows for c in chunks)          # Next honours RequiredRows; the last Next returned 0 rows
    sel, nsel = O.vec_filter(tbl, filters)
    exp = _rows([Chunk([Column(c.data[sel], c.nulls()[sel]) for c in tbl.columns])])
    assert _rows(chunks) == exp                                            # Selection keeps the child's row order (select.go:765)
    assert e.launches < len(tbl.split(1024))                               # child chunks are batched into few launches
    # Open again after Close: the executor re-executes (Apply re-execution contract, executor.go:301)
    assert _rows(drain(e, 1024)) == exp
    # an always-false filter: EOF at once
    e2 = SelectionExec(MockDataSource(schema, tbl.split(1024)), [FilterItem(abi.CMP_GT, 0, const_i64=1000)])
    assert drain(e2) == []


def test_projection_exec_vs_oracle_and_overflow():
    rng = np.random.default_rng(4)
    tbl, schema = _table(rng, 10_000)
    # l_extendedprice * (1 - l_discount) as the planner writes it once the constant sits on the right: x * ((y * -1) + 1);
    # a + b; a < b; and a plain column reference (swapped through, not copied)
    one_minus_y = ScalarFunc("arith", abi.ARITH_PLUS, (ScalarFunc("arith", abi.ARITH_MUL, (ColRef(3), Const(-1.0, True)), is_real=True), Const(1.0, True)), is_real=True)
    exprs = [ColRef(1), ScalarFunc("arith", abi.ARITH_MUL, (ColRef(2), one_minus_y), is_real=True),
             ScalarFunc("arith", abi.ARITH_PLUS, (ColRef(0), ColRef(1))), ScalarFunc("cmp", abi.CMP_LT, (ColRef(0), ColRef(1))),
             ScalarFunc("arith", abi.ARITH_MINUS, (ColRef(0), Const(7)))]
    e = ProjectionExec(MockDataSource(schema, tbl.split(1024)), exprs, batch_rows=4096)
    chunks = drain(e, 333)
    assert all(0 < c.num_rows() <= 333 for c in chunks)
    a, b, x, y = tbl.columns
    _, m1, m1n = O.vec_arith_real(abi.ARITH_MUL, y, None, -1.0)
    _, om, omn = O.vec_arith_real(abi.ARITH_PLUS, Column(m1, m1n), None, 1.0)
    _, rev, revn = O.vec_arith_real(abi.ARITH_MUL, x, Column(om, omn))
    _, s_, sn = O.vec_arith_int(abi.ARITH_PLUS, a, b)
    lt, ltn = O.vec_compare_int(abi.CMP_LT, a, b)
    _, mi, min_ = O.vec_arith_int(abi.ARITH_MINUS, a, None, 7)
    exp = _rows([Chunk([b, Column(rev, revn), Column(s_, sn), Column(lt, ltn), Column(mi, min_)])])
    assert _rows(chunks) == exp
    # overflow on a non-NULL row fails the whole Next (builtin_arithmetic_vec.go:957 -> types.ErrOverflow)
    big = Chunk([Column(np.array([1, (1 << 63) - 1, 5], dtype=np.int64)), Column(np.array([1, 1, 1], dtype=np.int64))])
    e3 = ProjectionExec(MockDataSource([INT, INT], [big]), [ScalarFunc("arith", abi.ARITH_PLUS, (ColRef(0), ColRef(1)))])
    e3.open()
    with pytest.raises(abi.TgError) as ei:
        e3.next()
    assert ei.value.code == abi.TG_ERR_OVERFLOW
    e3.close()
    # ... but not when the overflowing row is NULL
    bign = Chunk([Column(np.array([1, (1 << 63) - 1, 5], dtype=np.int64), np.array([False, True, False])), Column(np.array([1, 1, 1], dtype=np.int64))])
    out = drain(ProjectionExec(MockDataSource([INT, INT], [bign]), [ScalarFunc("arith", abi.ARITH_PLUS, (ColRef(0), ColRef(1)))]))
    assert _rows(out) == [(2,), (None,), (6,)]
