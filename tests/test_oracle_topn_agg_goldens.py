"""SQL known answers of the reference for ORDER BY ... LIMIT (TopNExec, sortexec/topn.go) and for a multi-column GROUP BY
(HashAggExec, agg_util.go:106 GetGroupKey), against oracle/topn.py and oracle/agg.cpp.  CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import topn as OT   # noqa: E402
from test_oracle_agg_vec import INT, run_agg   # noqa: E402
from tidb_b200 import abi   # noqa: E402
from tidb_b200.chunk import Chunk, Column   # noqa: E402
from tidb_b200.plan import AggFunc, AggPlan   # noqa: E402


def test_sql_limit_offset_goldens():
    # tests/integrationtest/r/executor/executor.result:388-407: select * from t order by a limit 1, k
    t = [(1, 1), (2, 2), (3, 30), (4, 40), (5, 5), (6, 6)]
    shuffled = [t[i] for i in (4, 0, 5, 2, 1, 3)]
    for k, want in ((1, [(2, 2)]), (2, [(2, 2), (3, 30)]), (3, [(2, 2), (3, 30), (4, 40)]), (4, [(2, 2), (3, 30), (4, 40), (5, 5)])):
        assert OT.topn_rows(shuffled, ["int", "int"], [(0, False)], 1, k) == want
    # executor.result:1394-1400: order by a desc limit 1 → (3, 1, 3)
    assert OT.topn_rows([(1, 3, 1), (2, 2, 2), (3, 1, 3)], ["int"] * 3, [(0, True)], 0, 1) == [(3, 1, 3)]
    # NULL ordering (util/chunk/compare.go:64 cmpNull; MySQL: NULL first ascending, last descending), executor.result:384-385 data
    rows = [(1, 1), (2, 2), (None, None)]
    assert OT.topn_rows(rows, ["int", "int"], [(0, False)], 0, 1) == [(None, None)]
    assert OT.topn_rows(rows, ["int", "int"], [(0, True)], 0, 2) == [(2, 2), (1, 1)]


def test_sql_multi_column_group_by_golden():
    # tests/integrationtest/r/executor/aggregate.result:59-81: t(a, b, c) = (0,0,0),(1,1,1),(3,3,6),(3,2,5),(2,1,4),(1,1,3),(1,1,2);
    # select count(a) from t where b > 0 group by a, b → {1, 1, 1, 3}; ... order by a limit 1 → 3
    t = [(0, 0, 0), (1, 1, 1), (3, 3, 6), (3, 2, 5), (2, 1, 4), (1, 1, 3), (1, 1, 2)]
    kept = [r for r in t if r[1] > 0]      # the Selection below the aggregate
    cols = [Column(np.array([r[c] for r in kept], dtype=np.int64)) for c in range(3)]
    plan = AggPlan([INT, INT, INT], [0, 1], [AggFunc(abi.AGG_COUNT, 0), AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_FIRSTROW, 1)])
    for pc, fc in ((1, 1), (5, 5)):
        got = run_agg(plan, Chunk(cols).split(2), pc, fc)
        assert sorted(r[0] for r in got) == [1, 1, 1, 3]
        assert sorted(got) == [(1, 2, 1), (1, 3, 2), (1, 3, 3), (3, 1, 1)]
        assert OT.topn_rows(got, ["int"] * 3, [(1, False)], 0, 1)[0][0] == 3
