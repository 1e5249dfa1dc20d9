"""oracle/join.cpp (hash-table restatement of HashJoinV2Exec) against
  (a) the SQL known answers the reference keeps for hash joins, and
  (b) the independent nested-loop restatement of the reference's own test generators
      (tests/nested_loop.py), in the style of testJoinProbe (inner_join_probe_test.go:228):
      random chunks, 1/3 of the build rows copied into the probe side, optional sel, NULL keys,
      duplicate keys, both build sides, every join type; comparison = sorted row multisets
      (checkChunksEqual inner_join_probe_test.go:137)."""
import itertools

import numpy as np
import pytest

import oracle_lib as O
from nested_loop import assert_rows_equal, columns_to_rows, nested_loop_join
from tidb_b200 import abi
from tidb_b200.chunk import Chunk, Column
from tidb_b200.plan import FieldType, FilterItem, JoinPlan

INT = FieldType(abi.TYPE_LONGLONG, 0)
INT_NN = FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)
UINT_NN = FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL | abi.FLAG_UNSIGNED)
DBL = FieldType(abi.TYPE_DOUBLE, 0)


def col(vals, nulls=None, dtype=np.int64):
    return Column(np.array(vals, dtype=dtype), None if nulls is None else np.array(nulls, dtype=bool))


def run_oracle(plan, left, right, conc=5):
    build, probe = (right, left) if plan.build_is_right else (left, right)
    j = O.OracleJoin(plan, conc)
    n, cols = j.run(build, probe)
    j.close()
    return columns_to_rows(cols) if n else []


# ---- (a) SQL known answers ----------------------------------------------------------------------
def test_sql_hash_join_result_t128_s2():
    # tests/integrationtest/r/executor/jointest/hash_join.result:1-22: t(1..128) join s(1,128)
    t = Chunk([col(list(range(1, 129)), [False] * 128)])
    s = Chunk([col([1, 128], [False, False])])
    plan = JoinPlan(abi.JOIN_INNER, [INT], [INT], [0], [0], build_is_right=True)
    for chunks in (t.split(32), [t]):   # tidb_max_chunk_size=32 in the golden file
        got = run_oracle(plan, chunks, [s])
        assert sorted(got) == [(1, 1), (128, 128)]


def test_sql_hash_join_result_negative_duplicates():
    # hash_join.result:36-60: t.a ∈ {148307968, -1327693824, -277544960}; s.a has -277544960 twice
    t = [Chunk([col([148307968, -1327693824, -277544960], [False] * 3)])]
    s = [Chunk([col([-277544960, 2, 2, -277544960, 2, 6], [False] * 6)])]
    inner = JoinPlan(abi.JOIN_INNER, [INT], [INT], [0], [0], lused=[0], rused=[])
    assert sorted(run_oracle(inner, t, s)) == [(-277544960,), (-277544960,)]
    left = JoinPlan(abi.JOIN_LEFT_OUTER, [INT], [INT], [0], [0], lused=[0], rused=[])
    assert sorted(run_oracle(left, t, s)) == [(-1327693824,), (-277544960,), (-277544960,), (148307968,)]


def test_sql_outer_join_nil_rows():
    # pkg/executor/test/jointest/join_test.go:66-80: t(1,1),(2,2); t1(2,3),(4,4)
    t = [Chunk([col([1, 2], [False, False]), col([1, 2], [False, False])])]
    t1 = [Chunk([col([2, 4], [False, False]), col([3, 4], [False, False])])]
    lo = JoinPlan(abi.JOIN_LEFT_OUTER, [INT, INT], [INT, INT], [0], [0])
    assert sorted(run_oracle(lo, t, t1), key=str) == sorted([(1, 1, None, None), (2, 2, 2, 3)], key=str)
    # "t left outer join t1 on t.c1 = t1.c1 and t.c1 != 1": left condition → probe filter; the
    # filtered outer row is still emitted NULL-padded
    lo_f = JoinPlan(abi.JOIN_LEFT_OUTER, [INT, INT], [INT, INT], [0], [0],
                    probe_filter=[FilterItem(abi.CMP_NE, 0, const_i64=1)])
    assert sorted(run_oracle(lo_f, t, t1), key=str) == sorted([(1, 1, None, None), (2, 2, 2, 3)], key=str)
    # "t1 right outer join t": (nil nil 1 1) appears
    ro = JoinPlan(abi.JOIN_RIGHT_OUTER, [INT, INT], [INT, INT], [0], [0], build_is_right=False)
    assert sorted(run_oracle(ro, t1, t), key=str) == sorted([(None, None, 1, 1), (2, 3, 2, 2)], key=str)


# ---- (b) randomized, oracle vs nested loop ------------------------------------------------------------
def gen_side(rng, rows, ncols, key_col, null_frac, key_range, dup, key_dtype=np.int64):
    cols = []
    for c in range(ncols):
        if c == key_col:
            if key_dtype == np.float64:
                v = rng.integers(-key_range, key_range, rows).astype(np.float64) / 2
            else:
                v = rng.integers(-key_range, key_range, rows).astype(np.int64)
            if dup:
                v = v[rng.integers(0, max(rows // 4, 1), rows)] if rows else v
        else:
            v = rng.integers(-1 << 40, 1 << 40, rows).astype(np.int64)
        nulls = rng.random(rows) < null_frac if null_frac > 0 else None
        cols.append(Column(v, nulls))
    return cols


def make_case(rng, n_build, n_probe, null_frac, dup, with_sel, key_dtype=np.int64, ncols=3):
    kt = DBL if key_dtype == np.float64 else INT
    ltypes = [kt if c == 1 else INT for c in range(ncols)]
    rtypes = [kt if c == 0 else INT for c in range(ncols)]
    right = gen_side(rng, n_build, ncols, 0, null_frac, 50, dup, key_dtype)
    left = gen_side(rng, n_probe, ncols, 1, null_frac, 50, dup, key_dtype)
    # copy 1/3 of the right keys into the left side to force matches (testJoinProbe :262-275)
    if n_build and n_probe:
        k = max(1, n_probe // 3)
        src = rng.integers(0, n_build, k)
        dst = rng.choice(n_probe, k, replace=False)
        lk = left[1].data.copy()
        lk[dst] = right[0].data[src]
        left[1] = Column(lk, left[1].nulls() if left[1].null_bitmap is not None else None)
    lchunk, rchunk = Chunk(left), Chunk(right)
    lchunks, rchunks = lchunk.split(37) or [], rchunk.split(29) or []
    if with_sel:
        for lst in (lchunks, rchunks):
            for ch in lst:
                n = ch.columns[0].length
                keep = np.sort(rng.choice(n, max(1, n * 2 // 3), replace=False))
                ch.sel = keep.astype(np.int64)
    return ltypes, rtypes, lchunks, rchunks


JOIN_TYPES = [abi.JOIN_INNER, abi.JOIN_LEFT_OUTER, abi.JOIN_RIGHT_OUTER, abi.JOIN_SEMI, abi.JOIN_ANTI_SEMI,
              abi.JOIN_LEFT_OUTER_SEMI, abi.JOIN_ANTI_LEFT_OUTER_SEMI]


@pytest.mark.parametrize("jt", JOIN_TYPES)
@pytest.mark.parametrize("build_is_right", [True, False])
@pytest.mark.parametrize("nulls,dup,with_sel", [(0.0, False, False), (0.15, True, False), (0.1, True, True)])
def test_oracle_vs_nested_loop(jt, build_is_right, nulls, dup, with_sel):
    if jt in (abi.JOIN_LEFT_OUTER_SEMI, abi.JOIN_ANTI_LEFT_OUTER_SEMI) and not build_is_right:
        pytest.skip("NewJoinProbe panics: left outer semi needs right build (base_join_probe.go:913)")
    rng = np.random.default_rng(1234 + jt * 7 + int(build_is_right))
    ltypes, rtypes, l, r = make_case(rng, 300, 400, nulls, dup, with_sel)
    semi = jt >= abi.JOIN_SEMI
    plan = JoinPlan(jt, ltypes, rtypes, [1], [0], build_is_right=build_is_right,
                    lused=[0, 1, 2], rused=[] if semi else [2, 0])
    for conc in (1, 5):
        got = run_oracle(plan, l, r, conc)
        assert_rows_equal(nested_loop_join(plan, l, r), got)


@pytest.mark.parametrize("jt", [abi.JOIN_INNER, abi.JOIN_LEFT_OUTER, abi.JOIN_RIGHT_OUTER, abi.JOIN_SEMI,
                                abi.JOIN_ANTI_SEMI])
def test_oracle_filters(jt):
    rng = np.random.default_rng(99 + jt)
    ltypes, rtypes, l, r = make_case(rng, 200, 300, 0.1, True, False)
    semi = jt >= abi.JOIN_SEMI
    for build_is_right in (True, False):
        lf = [FilterItem(abi.CMP_GT, 0, const_i64=0)]
        rf = [FilterItem(abi.CMP_LT, 1, const_i64=1 << 39), FilterItem(abi.CMP_NE, 2, rhs_col=1)]
        plan = JoinPlan(jt, ltypes, rtypes, [1], [0], build_is_right=build_is_right,
                        lused=None, rused=[] if semi else None,
                        build_filter=rf if build_is_right else lf, probe_filter=lf if build_is_right else rf)
        # the reference only attaches a filter to the OUTER side of an outer join / any side of inner
        if jt == abi.JOIN_LEFT_OUTER:
            plan.build_filter, plan.probe_filter = ([], lf) if build_is_right else (lf, [])
        if jt == abi.JOIN_RIGHT_OUTER:
            plan.build_filter, plan.probe_filter = (rf, []) if build_is_right else ([], rf)
        if semi:
            plan.build_filter, plan.probe_filter = ([], lf) if build_is_right else (lf, [])
            if jt == abi.JOIN_ANTI_SEMI:
                continue  # anti semi + left filter is planned as other-condition, out of scope here
        got = run_oracle(plan, l, r)
        assert_rows_equal(nested_loop_join(plan, l, r), got)


def test_oracle_double_and_mixed_sign_keys():
    rng = np.random.default_rng(5)
    ltypes, rtypes, l, r = make_case(rng, 150, 200, 0.1, True, False, key_dtype=np.float64)
    # -0.0 must join with +0.0 (codec.go:676-682)
    l[0].columns[1].data[0] = -0.0
    r[0].columns[0].data[0] = 0.0
    l[0].columns[1] = Column(l[0].columns[1].data, None)
    r[0].columns[0] = Column(r[0].columns[0].data, None)
    plan = JoinPlan(abi.JOIN_INNER, ltypes, rtypes, [1], [0])
    got = run_oracle(plan, l, r)
    exp = nested_loop_join(plan, l, r)
    assert_rows_equal(exp, got)
    assert any(row[1] == 0 for row in got)
    # signed vs unsigned keys: FixedSerializedKey with sign flag (join_table_meta.go:296-303):
    # -1 (signed) must NOT match 2^64-1 (unsigned) although the raw bytes are equal
    lt, rt = [INT_NN], [UINT_NN]
    lc = [Chunk([col([-1, 5, 7])])]
    rc = [Chunk([col([-1, 5, 9])])]   # -1 here is 0xFFFF... unsigned
    plan = JoinPlan(abi.JOIN_INNER, lt, rt, [0], [0])
    assert sorted(run_oracle(plan, lc, rc)) == [(5, 5)]


def test_oracle_multi_key():
    rng = np.random.default_rng(6)
    n = 300
    a = rng.integers(0, 6, n); b = rng.integers(0, 6, n)
    c = rng.integers(0, 6, n); d = rng.integers(0, 6, n)
    l = [Chunk([Column(a), Column(b, rng.random(n) < 0.1)])]
    r = [Chunk([Column(c), Column(d, rng.random(n) < 0.1)])]
    plan = JoinPlan(abi.JOIN_INNER, [INT_NN, INT], [INT_NN, INT], [0, 1], [0, 1])
    assert_rows_equal(nested_loop_join(plan, l, r), run_oracle(plan, l, r))


def test_oracle_empty_sides():
    e = [Chunk([Column(np.zeros(0, dtype=np.int64))])]
    f = [Chunk([col([1, 2, 3])])]
    for jt in (abi.JOIN_INNER, abi.JOIN_LEFT_OUTER, abi.JOIN_SEMI, abi.JOIN_ANTI_SEMI):
        semi = jt >= abi.JOIN_SEMI
        plan = JoinPlan(jt, [INT], [INT], [0], [0], rused=[] if semi else None)
        assert_rows_equal(nested_loop_join(plan, f, e), run_oracle(plan, f, e))
        assert_rows_equal(nested_loop_join(plan, e, f), run_oracle(plan, e, f))


def test_row_layout_alignment_and_table_size():
    # row_table_builder_test.go:72-83: every row is 8-byte aligned; config-2 shaped row is 32 bytes
    # (SURVEY §8 a6); hash_table_v2.go:67: slots = max(nextPow2(valid keys), 32) per partition
    n = 1000
    keys = np.arange(n, dtype=np.int64) * 7919
    plan = JoinPlan(abi.JOIN_INNER, [INT_NN, INT_NN], [INT_NN, INT_NN], [0], [0])
    j = O.OracleJoin(plan, 1)   # concurrency 1 → 1 partition
    j.run([Chunk([Column(keys), Column(keys * 3)])], [Chunk([Column(keys[:10]), Column(keys[:10])])])
    assert j.stat("row_count") == n
    assert j.stat("total_row_bytes") == 32 * n
    assert j.stat("partitions") == 1
    assert j.stat("hash_table_slots") == 1024   # nextPowerOfTwo(1000) = 1024
    j.close()
    j = O.OracleJoin(plan, 5)   # default concurrency 5 → 8 partitions (TestSetupPartitionInfo)
    j.run([Chunk([Column(keys), Column(keys * 3)])], [Chunk([Column(keys[:10]), Column(keys[:10])])])
    assert j.stat("partitions") == 8
    j.close()


@pytest.mark.parametrize("jt,build_is_right", [(abi.JOIN_INNER, True), (abi.JOIN_INNER, False), (abi.JOIN_LEFT_OUTER, True),
                                               (abi.JOIN_RIGHT_OUTER, False), (abi.JOIN_SEMI, True), (abi.JOIN_ANTI_SEMI, True)])
def test_oracle_other_condition_vs_nested_loop(jt, build_is_right):
    # OtherCondition: a candidate pair (equal keys) is a match only if every residual item is non-NULL true
    # (inner_join_probe.go:72-79, base_join_probe.go:758); the oracle's restatement against the independent nested loop
    from tidb_b200.plan import OtherCond
    rng = np.random.default_rng(900 + jt)
    ltypes, rtypes, l, r = make_case(rng, 600, 800, 0.1, True, False)
    semi = jt >= abi.JOIN_SEMI
    other = [OtherCond(abi.CMP_LT, 0, 0, 1, 1), OtherCond(abi.CMP_NE, 1, 2, -1, -1, const_i64=3)]
    plan = JoinPlan(jt, ltypes, rtypes, [1], [0], build_is_right=build_is_right, lused=[0, 1, 2], rused=[] if semi else [2, 0], other_cond=other)
    assert_rows_equal(nested_loop_join(plan, l, r), run_oracle(plan, l, r))


def make_multikey_case(rng, n_build, n_probe, null_frac, nkeys=2, key_range=6, unsigned_second=False):
    """two sides of 4 columns each; the first `nkeys` columns of the right side and columns 1..nkeys of the left side are
    the equal-condition keys, drawn from a small range so that pairs agreeing on SOME but not ALL key columns are common"""
    def side(rows, first_key):
        cols = []
        for c in range(4):
            is_key = first_key <= c < first_key + nkeys
            v = rng.integers(-key_range, key_range, rows).astype(np.int64) if is_key else rng.integers(-1 << 40, 1 << 40, rows).astype(np.int64)
            nulls = rng.random(rows) < null_frac if null_frac > 0 else None
            cols.append(Column(v, nulls))
        return cols
    right, left = side(n_build, 0), side(n_probe, 1)
    u = FieldType(abi.TYPE_LONGLONG, abi.FLAG_UNSIGNED)
    rtypes = [INT] * 4
    ltypes = [INT] * 4
    if unsigned_second:
        ltypes[2] = u      # left key 2 unsigned vs right key 2 signed: equal bits of a negative value must not match
    return ltypes, rtypes, Chunk(left).split(41), Chunk(right).split(23)


@pytest.mark.parametrize("jt,build_is_right", [(abi.JOIN_INNER, True), (abi.JOIN_INNER, False), (abi.JOIN_LEFT_OUTER, True),
                                               (abi.JOIN_RIGHT_OUTER, False), (abi.JOIN_SEMI, True), (abi.JOIN_ANTI_SEMI, True),
                                               (abi.JOIN_LEFT_OUTER, False), (abi.JOIN_SEMI, False)])
@pytest.mark.parametrize("nkeys,unsigned_second", [(2, False), (3, False), (2, True)])
def test_oracle_multi_column_keys_vs_nested_loop(jt, build_is_right, nkeys, unsigned_second):
    # several equal conditions: FixedSerializedKey (join_table_meta.go:174-178, codec.go:822 SerializeKeys): a pair matches
    # iff EVERY key column is non-NULL and equal by value
    rng = np.random.default_rng(4100 + jt * 11 + nkeys)
    ltypes, rtypes, l, r = make_multikey_case(rng, 500, 700, 0.08, nkeys, unsigned_second=unsigned_second)
    semi = jt >= abi.JOIN_SEMI
    plan = JoinPlan(jt, ltypes, rtypes, list(range(1, 1 + nkeys)), list(range(nkeys)), build_is_right=build_is_right,
                    lused=[0, 1, 2, 3], rused=[] if semi else [3, 0, 1])
    want = nested_loop_join(plan, l, r)
    assert len(want) > 0
    assert_rows_equal(want, run_oracle(plan, l, r))


def core_time(year, month, day, hour=0, minute=0, second=0, micro=0, fsp_tt=0):
    """types.CoreTime bit fields + the fspTt nibble (types/time.go:235-251): year 14 bits at 50, month 4 at 46, day 5 at 41,
    hour 5 at 36, minute 6 at 30, second 6 at 24, microsecond 20 at 4; fspTt = 0b1110 for DATE, fsp << 1 | is_timestamp else"""
    v = (year << 50) | (month << 46) | (day << 41) | (hour << 36) | (minute << 30) | (second << 24) | (micro << 4) | fsp_tt
    return v - (1 << 64) if v >= (1 << 63) else v


DATE_TT, DATETIME0_TT, DATETIME6_TT, TIMESTAMP0_TT = 0b1110, 0, 6 << 1, 1


def make_time_case(rng, n_build, n_probe, null_frac):
    """left: (payload, DATETIME(6) key, payload), right: (DATE key, payload): days from a small window so that dates repeat;
    a third of the left keys sit at midnight (they can match a DATE), the rest carry a time of day (they cannot)"""
    def days(n):
        d = rng.integers(0, 40, n)
        return [(2023 + int(x) // 28 // 12, 1 + (int(x) // 28) % 12, 1 + int(x) % 28) for x in d]
    rk = [core_time(y, m, d, fsp_tt=DATE_TT) for (y, m, d) in days(n_build)]
    lk = []
    for (y, m, d) in days(n_probe):
        if rng.random() < 0.34:
            lk.append(core_time(y, m, d, fsp_tt=DATETIME6_TT))
        else:
            lk.append(core_time(y, m, d, int(rng.integers(0, 24)), int(rng.integers(0, 60)), int(rng.integers(0, 60)), int(rng.integers(0, 1000000)), DATETIME6_TT))
    nl = lambda n: (rng.random(n) < null_frac) if null_frac > 0 else None
    left = [Column(rng.integers(-1 << 40, 1 << 40, n_probe).astype(np.int64), nl(n_probe)), Column(np.array(lk, dtype=np.int64), nl(n_probe)),
            Column(rng.integers(-1 << 40, 1 << 40, n_probe).astype(np.int64), nl(n_probe))]
    right = [Column(np.array(rk, dtype=np.int64), nl(n_build)), Column(rng.integers(-1 << 40, 1 << 40, n_build).astype(np.int64), nl(n_build))]
    ltypes = [INT, FieldType(abi.TYPE_DATETIME, 0), INT]
    rtypes = [FieldType(abi.TYPE_DATE, 0), INT]
    return ltypes, rtypes, Chunk(left).split(61), Chunk(right).split(47)


def test_oracle_time_key_known_pairs():
    # DATE '2024-02-29' joins DATETIME '2024-02-29 00:00:00' whatever the fsp, and the zero date joins the zero datetime
    # (Time.ToPackedUint, types/time.go:646-657: IsZero → 0); a time of day or another day never does
    d = [core_time(2024, 2, 29, fsp_tt=DATE_TT), core_time(0, 0, 0, fsp_tt=DATE_TT), core_time(9999, 12, 31, fsp_tt=DATE_TT)]
    t = [core_time(2024, 2, 29, fsp_tt=DATETIME0_TT), core_time(2024, 2, 29, fsp_tt=DATETIME6_TT), core_time(2024, 2, 29, 0, 0, 0, 1, DATETIME6_TT),
         core_time(0, 0, 0, fsp_tt=DATETIME0_TT), core_time(2024, 2, 28, fsp_tt=DATETIME0_TT), core_time(9999, 12, 31, fsp_tt=TIMESTAMP0_TT)]
    plan = JoinPlan(abi.JOIN_INNER, [FieldType(abi.TYPE_DATETIME, 0)], [FieldType(abi.TYPE_DATE, 0)], [0], [0])
    got = run_oracle(plan, [Chunk([col(t)])], [Chunk([col(d)])])
    assert sorted(got) == sorted([(t[0], d[0]), (t[1], d[0]), (t[3], d[1]), (t[5], d[2])])


@pytest.mark.parametrize("jt,build_is_right", [(abi.JOIN_INNER, True), (abi.JOIN_INNER, False), (abi.JOIN_LEFT_OUTER, True),
                                               (abi.JOIN_LEFT_OUTER, False), (abi.JOIN_SEMI, True), (abi.JOIN_ANTI_SEMI, False)])
def test_oracle_time_keys_vs_nested_loop(jt, build_is_right):
    rng = np.random.default_rng(7300 + jt)
    ltypes, rtypes, l, r = make_time_case(rng, 300, 500, 0.07)
    semi = jt >= abi.JOIN_SEMI
    plan = JoinPlan(jt, ltypes, rtypes, [1], [0], build_is_right=build_is_right, lused=[0, 1, 2], rused=[] if semi else [1, 0])
    want = nested_loop_join(plan, l, r)
    assert len(want) > 0
    assert_rows_equal(want, run_oracle(plan, l, r))


@pytest.mark.parametrize("seed", range(24))
def test_oracle_fuzz_keys_filters_other_condition_vs_nested_loop(seed):
    """randomized differential test of the oracle against the independent nested loop over the shapes the device accepts:
    1-3 key columns (integer family, mixed signedness), optional OtherCondition items (column vs column across sides, column
    vs constant), optional filter on the probe side, NULLs everywhere, duplicate keys, ragged chunks"""
    from tidb_b200.plan import OtherCond
    rng = np.random.default_rng(88000 + seed)
    nkeys = int(rng.integers(1, 4))
    jt, brt = [(abi.JOIN_INNER, True), (abi.JOIN_INNER, False), (abi.JOIN_LEFT_OUTER, True), (abi.JOIN_RIGHT_OUTER, False),
               (abi.JOIN_SEMI, True), (abi.JOIN_ANTI_SEMI, True)][int(rng.integers(0, 6))]
    ltypes, rtypes, l, r = make_multikey_case(rng, int(rng.integers(1, 400)), int(rng.integers(1, 600)), float(rng.choice([0.0, 0.1])), nkeys,
                                              key_range=int(rng.integers(2, 9)), unsigned_second=bool(rng.integers(0, 2)) and nkeys >= 2)
    other = []
    if rng.random() < 0.6:
        other.append(OtherCond(int(rng.integers(0, 6)), 0, 0, 1, 3))                                  # left.col0 OP right.col3
    if rng.random() < 0.4:
        other.append(OtherCond(int(rng.integers(0, 6)), 1, 3, -1, -1, const_i64=int(rng.integers(-1 << 39, 1 << 39))))   # right.col3 OP const
    semi = jt >= abi.JOIN_SEMI
    # a filter on the probe side (the outer side of the outer joins): filtered outer rows are still emitted NULL-padded
    pf = [FilterItem(abi.CMP_GT, 0 if brt else 3, const_i64=int(rng.integers(-1 << 39, 1 << 39)))] if rng.random() < 0.5 else []
    plan = JoinPlan(jt, ltypes, rtypes, list(range(1, 1 + nkeys)), list(range(nkeys)), build_is_right=brt,
                    lused=[0, 1, 2, 3], rused=[] if semi else [3, 0, 1], other_cond=other, probe_filter=pf)
    assert_rows_equal(nested_loop_join(plan, l, r), run_oracle(plan, l, r, int(rng.integers(1, 6))))
