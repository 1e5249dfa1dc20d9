"""oracle/agg.cpp and oracle/vec.cpp against the reference's known answers and numpy re-derivations."""
import numpy as np
import pytest

import oracle_lib as O
from nested_loop import columns_to_rows
from tidb_b200 import abi
from tidb_b200.chunk import Chunk, Column
from tidb_b200.plan import AggFunc, AggPlan, FieldType, FilterItem

INT = FieldType(abi.TYPE_LONGLONG, 0)
DBL = FieldType(abi.TYPE_DOUBLE, 0)


def run_agg(plan, chunks, pc=5, fc=5):
    a = O.OracleAgg(plan, pc, fc)
    n, cols = a.run(chunks)
    a.close()
    return columns_to_rows(cols) if n else []


# ---- aggfuncs known answers ----------------------------------------------------------------------------
def test_aggfunc_kats_rows_0_to_4():
    # aggfuncs/func_sum_test.go:39-47 TestSum (DOUBLE, 5 rows 0..4 → 10.0; empty → NULL),
    # func_count_test.go TestCount (→ 5; empty → 0), func_avg_test.go:38-47 TestAvg (→ 2.0; empty → NULL)
    x = Column(np.arange(5, dtype=np.float64))
    plan = AggPlan([DBL], [], [AggFunc(abi.AGG_SUM, 0, abi.TYPE_DOUBLE), AggFunc(abi.AGG_COUNT, 0, abi.TYPE_DOUBLE),
                               AggFunc(abi.AGG_AVG, 0, abi.TYPE_DOUBLE)])
    assert run_agg(plan, [Chunk([x])]) == [(10.0, 5, 2.0)]
    empty = Chunk([Column(np.zeros(0, dtype=np.float64))])
    assert run_agg(plan, [empty]) == [(None, 0, None)]
    assert run_agg(plan, []) == [(None, 0, None)]


def test_aggfunc_merge_partial_kats():
    # func_sum_test.go:28-37 TestMergePartialResult4Sum: partial(0..4)=10, partial(2..4)=9, merged 19;
    # func_count_test.go:43 (5, 3, 8); func_avg_test.go:25-33 (2.0, 3.0, 2.375).
    # Two partial workers each see one chunk; the final worker merges (agg_hash_final_worker.go:73).
    a = Chunk([Column(np.arange(5, dtype=np.float64))])
    b = Chunk([Column(np.arange(2, 5, dtype=np.float64))])
    plan = AggPlan([DBL], [], [AggFunc(abi.AGG_SUM, 0, abi.TYPE_DOUBLE), AggFunc(abi.AGG_COUNT, 0, abi.TYPE_DOUBLE),
                               AggFunc(abi.AGG_AVG, 0, abi.TYPE_DOUBLE)])
    for pc, fc in ((2, 1), (2, 3), (1, 1)):
        assert run_agg(plan, [a, b], pc, fc) == [(19.0, 8, 2.375)]


def test_sql_aggregate_result_goldens():
    # tests/integrationtest/r/executor/aggregate.result:11-14: count(c) over NULL → 0, over one value → 1
    t = Chunk([Column(np.array([1, 2], dtype=np.int64)),
               Column(np.array([0, 1], dtype=np.int64), np.array([True, False]))])
    plan = AggPlan([INT, INT], [0], [AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_COUNT, 1)])
    assert sorted(run_agg(plan, [t])) == [(1, 0), (2, 1)]
    # :18-21: sum(b/c) group by id → 0.3333333333333333, 0.16666666666666666 (the division is a
    # projection below the agg; float32 inputs widened to double)
    q = Column(np.array([np.float32(1) / np.float32(3), np.float32(1) / np.float32(6)], dtype=np.float64))
    q = Column(np.array([1.0 / 3.0, 1.0 / 6.0]))
    plan = AggPlan([INT, DBL], [0], [AggFunc(abi.AGG_SUM, 1, abi.TYPE_DOUBLE)])
    got = sorted(run_agg(plan, [Chunk([Column(np.array([1, 2], dtype=np.int64)), q])]))
    assert [repr(r[0]) for r in got] == ["0.16666666666666666", "0.3333333333333333"]
    # :53-58: count(a) on an empty table → one row 0; with GROUP BY → empty set
    empty = Chunk([Column(np.zeros(0, dtype=np.int64))])
    assert run_agg(AggPlan([INT], [], [AggFunc(abi.AGG_COUNT, 0)]), [empty]) == [(0,)]
    assert run_agg(AggPlan([INT], [0], [AggFunc(abi.AGG_COUNT, 0)]), [empty]) == []


def test_parallel_hash_agg_20_groups():
    # pkg/executor/test/aggregate/aggregate_test.go:385-420 TestParallelHashAgg: 20 groups × 20 rows
    # of v=1 → every SUM is 20 (groups are strings there; int group ids here), max_chunk_size 32
    g = np.tile(np.arange(20, dtype=np.int64), 20)
    v = np.ones(400, dtype=np.float64)
    chunks = Chunk([Column(g), Column(v)]).split(32)
    plan = AggPlan([INT, DBL], [0], [AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_SUM, 1, abi.TYPE_DOUBLE)])
    assert sorted(run_agg(plan, chunks)) == [(i, 20.0) for i in range(20)]


def test_agg_random_vs_numpy():
    rng = np.random.default_rng(7)
    n = 5000
    g = rng.integers(-20, 20, n).astype(np.int64)
    gn = rng.random(n) < 0.05          # NULL group keys form ONE group (codec.go:1766-1771)
    x = rng.random(n) * 1e7
    xn = rng.random(n) < 0.1
    y = rng.integers(-1000, 1000, n).astype(np.int64)
    chunks = Chunk([Column(g, gn), Column(x, xn), Column(y)]).split(1024)
    plan = AggPlan([INT, DBL, INT], [0], [
        AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_SUM, 1, abi.TYPE_DOUBLE), AggFunc(abi.AGG_COUNT, 1, abi.TYPE_DOUBLE),
        AggFunc(abi.AGG_AVG, 1, abi.TYPE_DOUBLE), AggFunc(abi.AGG_COUNT, -1), AggFunc(abi.AGG_MIN, 2), AggFunc(abi.AGG_MAX, 2),
        AggFunc(abi.AGG_MIN, 1, abi.TYPE_DOUBLE), AggFunc(abi.AGG_MAX, 1, abi.TYPE_DOUBLE)])
    got = {r[0]: r for r in run_agg(plan, chunks)}
    keys = set(None if nl else int(v) for v, nl in zip(g, gn))
    assert set(got.keys()) == keys
    for k in keys:
        m = gn if k is None else (~gn & (g == k))
        xs = x[m & ~xn]
        r = got[k]
        assert r[2] == len(xs) and r[4] == int(m.sum())
        if len(xs):
            assert r[1] == pytest.approx(xs.sum(), rel=1e-9) and r[3] == pytest.approx(xs.mean(), rel=1e-9)
            assert r[7] == xs.min() and r[8] == xs.max()
        else:
            assert r[1] is None and r[3] is None and r[7] is None
        assert r[5] == y[m].min() and r[6] == y[m].max()


@pytest.mark.parametrize("ngroup,pc,fc", [(2, 1, 1), (3, 5, 5), (4, 3, 2)])
def test_agg_multi_column_group_by_and_fused_argument_vs_python(ngroup, pc, fc):
    # GetGroupKey over several columns (agg_util.go:106: the encoded columns concatenated, NULL = NilFlag, so NULL is a key value
    # of its own in every position) and the aggregate argument l * (1 - d) evaluated below the aggregate (func_sum.go:90
    # args[0].EvalReal): the oracle against a plain Python dictionary
    rng = np.random.default_rng(300 + ngroup)
    n = 6000
    gcols = [Column(rng.integers(-3, 3, n).astype(np.int64), rng.random(n) < 0.08) for _ in range(ngroup)]
    a = np.floor(rng.random(n) * 1000) / 4
    b = np.floor(rng.random(n) * 10) / 100
    an = rng.random(n) < 0.05
    cols = gcols + [Column(a, an), Column(b)]
    ia, ib = ngroup, ngroup + 1
    plan = AggPlan([INT] * ngroup + [DBL, DBL], list(range(ngroup)),
                   [AggFunc(abi.AGG_FIRSTROW, c) for c in range(ngroup)] +
                   [AggFunc(abi.AGG_SUM, ia, abi.TYPE_DOUBLE, arg_col2=ib, arg_expr=abi.ARGEXPR_MUL_CSUB, arg_const=1.0),
                    AggFunc(abi.AGG_AVG, ia, abi.TYPE_DOUBLE, arg_col2=ib, arg_expr=abi.ARGEXPR_MUL),
                    AggFunc(abi.AGG_COUNT, ia, abi.TYPE_DOUBLE), AggFunc(abi.AGG_COUNT, -1)])
    got = {tuple(r[:ngroup]): r[ngroup:] for r in run_agg(plan, Chunk(cols).split(777), pc, fc)}
    exp = {}
    gv = [(c.data, c.nulls()) for c in gcols]
    for i in range(n):
        k = tuple(None if nl[i] else int(v[i]) for v, nl in gv)
        e = exp.setdefault(k, [0.0, 0.0, 0, 0])
        e[3] += 1
        if not an[i]:
            e[0] += a[i] * (1.0 - b[i]); e[1] += a[i] * b[i]; e[2] += 1
    assert set(got) == set(exp)
    for k, (s1, s2, cnt, rows) in exp.items():
        r = got[k]
        assert r[2] == cnt and r[3] == rows
        if cnt:
            assert r[0] == pytest.approx(s1, rel=1e-9) and r[1] == pytest.approx(s2 / cnt, rel=1e-9)
        else:
            assert r[0] is None and r[1] is None


# ---- VecEval -----------------------------------------------------------------------------------------
def test_vec_compare_int_and_nulls():
    a = Column(np.array([1, 2, 3, -4, 5], dtype=np.int64), np.array([0, 0, 1, 0, 0], dtype=bool))
    b = Column(np.array([2, 2, 2, 2, 2], dtype=np.int64), np.array([0, 1, 0, 0, 0], dtype=bool))
    for op, fn in ((abi.CMP_LT, np.less), (abi.CMP_LE, np.less_equal), (abi.CMP_GT, np.greater),
                   (abi.CMP_GE, np.greater_equal), (abi.CMP_EQ, np.equal), (abi.CMP_NE, np.not_equal)):
        res, nulls = O.vec_compare_int(op, a, b)
        assert list(nulls) == [False, True, True, False, False]   # MergeNulls column.go:906
        exp = fn(a.data, b.data).astype(np.int64)
        assert all(res[i] == exp[i] for i in range(5) if not nulls[i])
        res, nulls = O.vec_compare_int(op, a, None, 2)
        assert list(nulls) == [False, False, True, False, False]
    # types.CompareInt mixed signedness (types/compare.go:86): unsigned 2^64-1 > signed -1
    u = Column(np.array([-1], dtype=np.int64))
    res, _ = O.vec_compare_int(abi.CMP_GT, u, None, -1, a_unsigned=True, b_unsigned=False)
    assert res[0] == 1
    res, _ = O.vec_compare_int(abi.CMP_EQ, u, None, -1, a_unsigned=True, b_unsigned=False)
    assert res[0] == 0
    res, _ = O.vec_compare_int(abi.CMP_LT, u, None, -1, a_unsigned=False, b_unsigned=True)
    assert res[0] == 1


def test_vec_compare_real_nan_ordering():
    # Go cmp.Compare: NaN < everything, NaN == NaN (builtin_compare_vec_generated.go:54)
    a = Column(np.array([np.nan, 1.0, np.nan, -0.0]))
    b = Column(np.array([1.0, np.nan, np.nan, 0.0]))
    res, _ = O.vec_compare_real(abi.CMP_LT, a, b); assert list(res) == [1, 0, 0, 0]
    res, _ = O.vec_compare_real(abi.CMP_EQ, a, b); assert list(res) == [0, 0, 1, 1]
    res, _ = O.vec_compare_real(abi.CMP_GT, a, b); assert list(res) == [0, 1, 0, 0]


def test_vec_arith_int_overflow_rules():
    mx, mn = (1 << 63) - 1, -(1 << 63)
    one = Column(np.array([1], dtype=np.int64))
    big = Column(np.array([mx], dtype=np.int64))
    # plusSS overflow (builtin_arithmetic_vec.go:957)
    rc, _, _ = O.vec_arith_int(abi.ARITH_PLUS, big, one); assert rc == abi.TG_ERR_OVERFLOW
    rc, r, _ = O.vec_arith_int(abi.ARITH_PLUS, big, None, -1); assert rc == 0 and r[0] == mx - 1
    # a NULL row never raises (the `if result.IsNull(i) continue` guard)
    nb = Column(np.array([mx], dtype=np.int64), np.array([True]))
    rc, _, nulls = O.vec_arith_int(abi.ARITH_PLUS, nb, one); assert rc == 0 and nulls[0]
    # unsigned + unsigned: 2^64-1 + 1 overflows, 2^63 + 2^63-1 does not
    rc, _, _ = O.vec_arith_int(abi.ARITH_PLUS, Column(np.array([-1], dtype=np.int64)), one, a_unsigned=True, b_unsigned=True)
    assert rc == abi.TG_ERR_OVERFLOW
    rc, r, _ = O.vec_arith_int(abi.ARITH_PLUS, Column(np.array([mn], dtype=np.int64)), big, a_unsigned=True, b_unsigned=True)
    assert rc == 0 and r[0] == -1
    # minus: MinInt64 - 1 overflows signed; unsigned 0 - 1 overflows; 5 - 3 fine
    rc, _, _ = O.vec_arith_int(abi.ARITH_MINUS, Column(np.array([mn], dtype=np.int64)), one); assert rc == abi.TG_ERR_OVERFLOW
    rc, _, _ = O.vec_arith_int(abi.ARITH_MINUS, Column(np.array([0], dtype=np.int64)), one, a_unsigned=True, b_unsigned=True)
    assert rc == abi.TG_ERR_OVERFLOW
    rc, r, _ = O.vec_arith_int(abi.ARITH_MINUS, Column(np.array([5], dtype=np.int64)), None, 3); assert rc == 0 and r[0] == 2
    # multiply: MinInt64 * -1 overflows (builtin_arithmetic_vec.go:667), 2^32 * 2^31 overflows, small ok
    rc, _, _ = O.vec_arith_int(abi.ARITH_MUL, Column(np.array([-1], dtype=np.int64)), None, mn); assert rc == abi.TG_ERR_OVERFLOW
    rc, _, _ = O.vec_arith_int(abi.ARITH_MUL, Column(np.array([1 << 32], dtype=np.int64)), None, 1 << 31); assert rc == abi.TG_ERR_OVERFLOW
    rc, r, _ = O.vec_arith_int(abi.ARITH_MUL, Column(np.array([-7], dtype=np.int64)), None, 6); assert rc == 0 and r[0] == -42


def test_vec_arith_real_overflow_rules():
    big = Column(np.array([1.7e308]))
    rc, _, _ = O.vec_arith_real(abi.ARITH_PLUS, big, big); assert rc == abi.TG_ERR_OVERFLOW     # +Inf
    rc, _, _ = O.vec_arith_real(abi.ARITH_MUL, big, None, 10.0); assert rc == abi.TG_ERR_OVERFLOW
    # NaN: plus/minus raise (!IsFinite), multiply does not (only IsInf is checked, :51-58)
    nan = Column(np.array([np.nan]))
    rc, _, _ = O.vec_arith_real(abi.ARITH_PLUS, nan, None, 1.0); assert rc == abi.TG_ERR_OVERFLOW
    rc, r, _ = O.vec_arith_real(abi.ARITH_MUL, nan, None, 1.0); assert rc == 0 and np.isnan(r[0])
    # Q3 projection l_price * (1 - l_disc)
    p = Column(np.array([100.0, 50.0])); d = Column(np.array([0.1, 0.0]))
    rc, one_minus, _ = O.vec_arith_real(abi.ARITH_MINUS, Column(np.array([1.0, 1.0])), d); assert rc == 0
    rc, r, _ = O.vec_arith_real(abi.ARITH_MUL, p, Column(one_minus)); assert rc == 0 and list(r) == [90.0, 50.0]


def test_vec_filter_sel_and_nulls():
    a = Column(np.array([5, 1, 7, 9, 3], dtype=np.int64), np.array([0, 0, 1, 0, 0], dtype=bool))
    b = Column(np.array([1.0, 2.0, 3.0, 4.0, 5.0]))
    chk = Chunk([a, b])
    items = [FilterItem(abi.CMP_GT, 0, const_i64=2), FilterItem(abi.CMP_LT, 1, is_real=True, const_f64=4.5)]
    sel, n = O.vec_filter(chk, items)
    assert list(sel) == [True, False, False, True, False] and n == 2
    # rows outside sel are never selected (chunk_executor.go:441-457)
    chk.sel = np.array([1, 3, 4], dtype=np.int64)
    sel, n = O.vec_filter(chk, items)
    assert list(sel) == [False, False, False, True, False] and n == 1
