"""SQL known answers the reference keeps for joins with several equal conditions and with a non-equi residual
(tests/integrationtest/r/executor/jointest/join.result), transcribed as (plan, left chunks, right chunks, expected rows).
SQL results do not depend on the join algorithm, so they pin the semantics of OtherCondition (inner_join_probe.go:72-79) and
of multi-column keys (join_table_meta.go:174-178) for the oracle (tests/test_oracle_sql_goldens.py) and for the CUDA path
(tests/test_gpu_join.py) alike."""
import numpy as np

from tidb_b200 import abi
from tidb_b200.chunk import Chunk, Column
from tidb_b200.plan import FieldType, JoinPlan, OtherCond

INT = FieldType(abi.TYPE_LONGLONG, 0)


def _table(rows, ncols):
    cols = []
    for c in range(ncols):
        vals = [r[c] for r in rows]
        cols.append(Column(np.array([0 if v is None else v for v in vals], dtype=np.int64), np.array([v is None for v in vals], dtype=bool)))
    return [Chunk(cols)]


def cases():
    out = []
    # join.result:202-212: k left join t on k.a = t.a and k.pk > t.pk → count(*) = 33
    k = [(0, 8), (0, 23), (1, 21), (1, 33), (1, 52), (2, 17), (2, 34), (2, 39), (2, 40), (2, 66), (2, 67), (3, 9), (3, 25), (3, 41), (3, 48),
         (4, 4), (4, 11), (4, 15), (4, 26), (4, 27), (4, 31), (4, 35), (4, 45), (4, 47), (4, 49)]
    t = [(3, 4), (3, 5), (3, 27), (3, 29), (3, 57), (3, 58), (3, 79), (3, 84), (3, 92), (3, 95)]
    plan = JoinPlan(abi.JOIN_LEFT_OUTER, [INT, INT], [INT, INT], [0], [0], build_is_right=True, other_cond=[OtherCond(abi.CMP_GT, 0, 1, 1, 1)])
    out.append(("left join, k.a = t.a and k.pk > t.pk (join.result:202-212)", plan, _table(k, 2), _table(t, 2), 33, None))
    # join.result:236-244: t t1 join t t2 on t1.b = t2.b and t1.a = t2.a
    t = [(1, 1), (1, 2), (2, 1), (2, 2)]
    plan = JoinPlan(abi.JOIN_INNER, [INT, INT], [INT, INT], [1, 0], [1, 0])
    out.append(("self join on (b, a) (join.result:236-244)", plan, _table(t, 2), _table(t, 2), 4, [(1, 1, 1, 1), (1, 2, 1, 2), (2, 1, 2, 1), (2, 2, 2, 2)]))
    # join.result:541-559: t join s on two equal conditions, three ways
    t = [(1, 1), (1, 2), (2, 2)]
    s = [(1, 1), (2, 2), (2, 1)]
    plan = JoinPlan(abi.JOIN_INNER, [INT, INT], [INT, INT], [0, 1], [0, 1])
    out.append(("t.a = s.a and t.b = s.b (join.result:547-550)", plan, _table(t, 2), _table(s, 2), 2, [(1, 1, 1, 1), (2, 2, 2, 2)]))
    plan = JoinPlan(abi.JOIN_INNER, [INT, INT], [INT, INT], [0, 1], [0, 0])
    out.append(("t.a = s.a and t.b = s.a (join.result:551-555)", plan, _table(t, 2), _table(s, 2), 3, [(1, 1, 1, 1), (2, 2, 2, 2), (2, 2, 2, 1)]))
    plan = JoinPlan(abi.JOIN_INNER, [INT, INT], [INT, INT], [0, 0], [0, 1])
    out.append(("t.a = s.a and t.a = s.b (join.result:556-560)", plan, _table(t, 2), _table(s, 2), 3, [(1, 1, 1, 1), (1, 2, 1, 1), (2, 2, 2, 2)]))
    # join.result:1258-1262: (t.c, t.d) = any (select * from t) → semi join on two key columns, t = (1,1),(2,2),(3,4)
    t = [(1, 1), (2, 2), (3, 4)]
    plan = JoinPlan(abi.JOIN_SEMI, [INT, INT], [INT, INT], [0, 1], [0, 1], build_is_right=True, lused=[0], rused=[])
    out.append(("(t.c, t.d) = any (select * from t) (join.result:1258-1262)", plan, _table(t, 2), _table(t, 2), 3, [(1,), (2,), (3,)]))
    # join.result:586-598: tt1(ts timestamp) = '2001-01-01 00:00:00', tt3(ts datetime) the same instant (session time zone UTC):
    # select * from tt1 where ts in (select ts from tt3) → the row; date-time keys compare by calendar fields, not by type bits
    def core_time(y, mo, d, h=0, mi=0, sec=0, us=0, fsp_tt=0):      # types/time.go:235-251
        v = (y << 50) | (mo << 46) | (d << 41) | (h << 36) | (mi << 30) | (sec << 24) | (us << 4) | fsp_tt
        return v - (1 << 64) if v >= (1 << 63) else v
    ts, dt = core_time(2001, 1, 1, fsp_tt=1), core_time(2001, 1, 1, fsp_tt=0)      # TIMESTAMP(0): fspTt = 0b0001, DATETIME(0): 0
    plan = JoinPlan(abi.JOIN_SEMI, [FieldType(abi.TYPE_TIMESTAMP, 0)], [FieldType(abi.TYPE_DATETIME, 0)], [0], [0], build_is_right=True, lused=[0], rused=[])
    out.append(("timestamp in (select datetime) (join.result:586-598)", plan, _table([(ts,)], 1), _table([(dt,), (core_time(2001, 1, 2),)], 1), 1, [(ts,)]))
    # join.result:1276-1277: A join B on A.c = B.c and A.c > 100 → empty (the one-side condition is a probe filter)
    return out
