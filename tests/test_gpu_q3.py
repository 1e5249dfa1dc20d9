"""Q3-shape pipeline (BASELINE configs[3]) at reduced scale (SF = 0.1): the device pipeline of tidb_b200/q3.py — two joins with
fused Selections, HashAgg on three GROUP BY columns with the projection fused into its update, TopN — against the SAME plan
run through the ORACLE operators (oracle/join.cpp, oracle/agg.cpp, oracle/topn.py) on host copies of the same columns.
Keys / dates / priorities and the group count bit-exact, SUM within 1e-6 relative, TopN rows equal (distinct revenues)."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle_lib as O
from tidb_b200 import abi, q3
from tidb_b200.chunk import Chunk, Column
from tidb_b200.plan import AggFunc, AggPlan, FilterItem, JoinPlan

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import topn as OT   # noqa: E402

pytestmark = pytest.mark.gpu
INT, DBL = q3.INT, q3.DBL


def oracle_q3(h):
    """h: dict of host numpy columns.  The reference plan, operator by operator (tpch_suite_out.json:99-123)."""
    j1 = JoinPlan(abi.JOIN_INNER, [INT] * 4, [INT] * 2, [1], [0], build_is_right=True, lused=[0, 2, 3], rused=[],
                  build_filter=[FilterItem(abi.CMP_EQ, 1, const_i64=q3.SEGMENT)], probe_filter=[FilterItem(abi.CMP_LT, 2, const_i64=q3.DATE)])
    n1, c1 = O.OracleJoin(j1, 4).run(Chunk([Column(h["c_custkey"]), Column(h["c_seg"])]).split(4096),
                                     Chunk([Column(h["o_orderkey"]), Column(h["o_custkey"]), Column(h["o_date"]), Column(h["o_prio"])]).split(4096))
    j2 = JoinPlan(abi.JOIN_INNER, [INT, DBL, DBL, INT], [INT] * 3, [0], [0], build_is_right=True, lused=[0, 1, 2], rused=[1, 2],
                  probe_filter=[FilterItem(abi.CMP_GT, 3, const_i64=q3.DATE)])
    n2, c2 = O.OracleJoin(j2, 4).run(Chunk([Column(v.copy()) for v, _ in c1]).split(4096),
                                     Chunk([Column(h["l_orderkey"]), Column(h["l_price"]), Column(h["l_disc"]), Column(h["l_ship"])]).split(4096))
    ap = AggPlan([INT, DBL, DBL, INT, INT], [0, 3, 4],
                 [AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_SUM, 1, abi.TYPE_DOUBLE, arg_col2=2, arg_expr=abi.ARGEXPR_MUL_CSUB, arg_const=1.0),
                  AggFunc(abi.AGG_FIRSTROW, 3), AggFunc(abi.AGG_FIRSTROW, 4)])
    a = O.OracleAgg(ap, 4, 4)
    ng, ca = a.run(Chunk([Column(v.copy()) for v, _ in c2]).split(4096))
    a.close()
    return n1, n2, ng, [v.copy() for v, _ in ca]


def test_q3_shape_vs_oracle_operators():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        d = q3.gen(dev, 15_000, 150_000, 600_000)
        t = {}
        got = q3.run(d, dev, stream, topn=10, timings=t)
        stream.synchronize()
    h = {k: v.cpu().numpy() for k, v in d.__dict__.items()}
    n1, n2, ng, (ok, rev, od, op) = oracle_q3(h)
    assert t["rows"] == {"j1_out": n1, "j2_out": n2, "groups": ng}              # row counts of every operator bit-exact
    g_ok = got["orderkey"].cpu().numpy(); order = np.argsort(g_ok); eo = np.argsort(ok)
    assert np.array_equal(g_ok[order], ok[eo])
    assert np.array_equal(got["o_date"].cpu().numpy()[order], od[eo]) and np.array_equal(got["o_prio"].cpu().numpy()[order], op[eo])
    assert np.allclose(got["revenue"].cpu().numpy()[order], rev[eo], rtol=1e-6, atol=0)
    # TopN 10: ORDER BY revenue DESC, o_orderdate
    rows = list(zip(ok.tolist(), rev.tolist(), od.tolist(), op.tolist()))
    exp_top = OT.topn_rows(rows, ["int", "real", "int", "int"], [(1, True), (2, False)], 0, 10)
    top = got["top"]
    assert len(top[0]) == len(exp_top) == 10
    for i, e in enumerate(exp_top):
        assert (int(top[0][i]), int(top[2][i]), int(top[3][i])) == (e[0], e[2], e[3])
        assert top[1][i] == pytest.approx(e[1], rel=1e-6)
    # and the plain torch rendering still agrees (kept as a second, independent check)
    ref = q3.reference(d)
    assert torch.equal(got["orderkey"][torch.argsort(got["orderkey"])], ref["orderkey"])
