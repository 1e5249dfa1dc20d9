"""TPC-H Q3-shape pipeline on device-resident synthetic columns (BASELINE.json configs[3]).

Plan shape follows the reference's TiFlash MPP plan for Q3 (pkg/planner/core/casetest/tpch/testdata/tpch_suite_out.json:99-123):

    HashAgg(group by l_orderkey, o_orderdate, o_shippriority; sum(l_extendedprice * (1 - l_discount)))
      HashJoin(lineitem.l_orderkey = orders.o_orderkey)            probe: Selection(l_shipdate > D) on lineitem
        HashJoin(orders.o_custkey = customer.c_custkey)            probe: Selection(o_orderdate < D) on orders
          Selection(c_mktsegment = S) on customer                  (build side)

Everything runs through the C-ABI operators: the three Selections are fused into the joins as build / probe
filters (tg_filter_item), and the
aggregation is tg_agg with all three GROUP BY columns (tag-claimed multi-word slots) and the projection
l_extendedprice * (1 - l_discount) fused into its update kernel (tg_agg_func.arg_expr), followed by tg_topn.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict

import torch

from . import abi
from .device import DeviceAgg, DeviceJoin
from .plan import AggFunc, AggPlan, FieldType, FilterItem, JoinPlan

INT = FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)
DBL = FieldType(abi.TYPE_DOUBLE, abi.FLAG_NOT_NULL)
SEGMENT, DATE = 1, 1168   # c_mktsegment = 1, o_orderdate < 1168 < l_shipdate  


@dataclass
class Q3Data:
    c_custkey: torch.Tensor; c_seg: torch.Tensor
    o_orderkey: torch.Tensor; o_custkey: torch.Tensor; o_date: torch.Tensor; o_prio: torch.Tensor
    l_orderkey: torch.Tensor; l_price: torch.Tensor; l_disc: torch.Tensor; l_ship: torch.Tensor

    def scanned_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.__dict__.values())


def gen(dev, n_cust: int, n_orders: int, n_line: int, seed: int = 42) -> Q3Data:
    g = torch.Generator(device=dev); g.manual_seed(seed)
    ri = lambda lo, hi, n: torch.randint(lo, hi, (n,), device=dev, generator=g, dtype=torch.int64)
    return Q3Data(
        c_custkey=torch.randperm(n_cust, device=dev, generator=g, dtype=torch.int64), c_seg=ri(0, 5, n_cust),
        o_orderkey=torch.randperm(n_orders, device=dev, generator=g, dtype=torch.int64) * 4 + 1, o_custkey=ri(0, n_cust, n_orders),
        o_date=ri(0, 2406, n_orders), o_prio=ri(0, 5, n_orders),
        l_orderkey=ri(0, n_orders, n_line) * 4 + 1,
        l_price=torch.floor(torch.rand(n_line, device=dev, generator=g, dtype=torch.float64) * 100000) / 100,
        l_disc=torch.floor(torch.rand(n_line, device=dev, generator=g, dtype=torch.float64) * 11) / 100,
        l_ship=ri(0, 2406, n_line))


def _view(ptr: int, n: int, dev, dt="<i8"):
    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (n,), "typestr": dt, "data": (ptr, False), "version": 3}
    return torch.as_tensor(a, device=dev)


def _col_struct(ptr: int, n: int, nulls: int = 0):
    c = abi.TgColumn(); c.length = n; c.data = ptr or None; c.elem_len = 8; c.null_bitmap = nulls or None; c.offsets = None
    return c


def run(d: Q3Data, dev, stream, topn: int = 10, timings: Dict[str, float] = None, keep_groups: bool = True) -> Dict[str, torch.Tensor]:
    """The whole query on the device.  -> {orderkey, revenue, o_date, o_prio} of every group (unordered, device tensors) and,
    under "top", the TopN rows (host numpy arrays, ORDER BY revenue DESC, o_orderdate LIMIT topn).
    Operators: J1 = orders JOIN customer (Selections fused as build/probe filters), J2 = lineitem JOIN J1 (J1's device-resident
    output is the build side, borrowed, no copy), HashAgg GROUP BY (l_orderkey, o_orderdate, o_shippriority) with
    SUM(l_extendedprice * (1 - l_discount)) evaluated INSIDE the update kernel (tg_agg_func.arg_expr: no projected column,
    the constant is a scalar), TopN (tg_topn)."""
    import numpy as np
    lib = abi.load_lib()
    st = stream.cuda_stream
    di = dev.index or 0
    marks = []

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append((name, e))
    mark("start")
    # J1: orders (probe, filter o_date < D) JOIN customer (build, filter c_seg = S); keep o_orderkey, o_date, o_prio
    j1 = DeviceJoin(JoinPlan(abi.JOIN_INNER, [INT] * 4, [INT] * 2, [1], [0], build_is_right=True, lused=[0, 2, 3], rused=[],
                             build_filter=[FilterItem(abi.CMP_EQ, 1, const_i64=SEGMENT)], probe_filter=[FilterItem(abi.CMP_LT, 2, const_i64=DATE)],
                             device=di, stream=st))
    j1.build([d.c_custkey, d.c_seg])
    mark("J1 build (customer, c_mktsegment filter fused)")
    n1, c1, _ = j1.probe([d.o_orderkey, d.o_custkey, d.o_date, d.o_prio])
    mark("J1 probe (orders, o_orderdate filter fused)")
    # J2: lineitem (probe, filter l_ship > D) JOIN J1 (build on o_orderkey); keep l_orderkey, l_price, l_disc, o_date, o_prio
    j2 = DeviceJoin(JoinPlan(abi.JOIN_INNER, [INT, DBL, DBL, INT], [INT] * 3, [0], [0], build_is_right=True, lused=[0, 1, 2], rused=[1, 2],
                             probe_filter=[FilterItem(abi.CMP_GT, 3, const_i64=DATE)], device=di, stream=st))
    j2.build([_view(p, n1, dev) for p in c1])     # borrowed until build_finish: J1's result buffers are read in place
    mark("J2 build (J1 output, in place)")
    j1.close()
    n2, c2, _ = j2.probe([d.l_orderkey, d.l_price, d.l_disc, d.l_ship])
    mark("J2 probe (lineitem, l_shipdate filter fused)")
    lk = _view(c2[0], n2, dev); price = _view(c2[1], n2, dev, "<f8"); disc = _view(c2[2], n2, dev, "<f8")
    jd = _view(c2[3], n2, dev); jp = _view(c2[4], n2, dev)
    # HashAgg: GROUP BY l_orderkey, o_orderdate, o_shippriority; SUM(l_extendedprice * (1 - l_discount)) fused
    agg = DeviceAgg(AggPlan([INT, DBL, DBL, INT, INT], [0, 3, 4],
                            [AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_SUM, 1, abi.TYPE_DOUBLE, arg_col2=2, arg_expr=abi.ARGEXPR_MUL_CSUB, arg_const=1.0),
                             AggFunc(abi.AGG_FIRSTROW, 3), AggFunc(abi.AGG_FIRSTROW, 4)],
                            device=di, stream=st, expected_groups=max(n1, 1)))
    agg.push([lk, price, disc, jd, jp])
    ng, ca, na = agg.finish()
    mark("HashAgg (3 GROUP BY columns, projection fused)")
    out = {"groups": ng}
    if keep_groups:   # copies of the whole aggregate result (verification); the query's own result is the TopN below
        out.update({"orderkey": _view(ca[0], ng, dev).clone(), "revenue": _view(ca[1], ng, dev, "<f8").clone(),
                    "o_date": _view(ca[2], ng, dev).clone(), "o_prio": _view(ca[3], ng, dev).clone()})
    # TopN: ORDER BY revenue DESC, o_orderdate LIMIT topn  (tpch_suite_out.json:102)
    if topn > 0:
        cols = (abi.TgColumn * 4)(*[_col_struct(ca[k], ng, na[k]) for k in range(4)])
        ck = abi.TgChunk(); ck.ncols = 4; ck.cols = C.cast(cols, C.POINTER(abi.TgColumn)); ck.sel = None; ck.nsel = 0
        tps = (C.c_int32 * 4)(abi.TYPE_LONGLONG, abi.TYPE_DOUBLE, abi.TYPE_LONGLONG, abi.TYPE_LONGLONG)
        fls = (C.c_uint32 * 4)(0, 0, 0, 0)
        items = (abi.TgSortItem * 2)(abi.TgSortItem(1, 1), abi.TgSortItem(2, 0))
        from .chunk import MutChunk
        oc = MutChunk([8, 8, 8, 8], max(topn, 8), [np.int64, np.float64, np.int64, np.int64])
        got = C.c_int64(0)
        abi.check(lib.tg_topn(di, 1, C.byref(ck), tps, fls, items, 2, C.c_int64(0), C.c_int64(topn), C.byref(oc.struct), C.byref(got), C.c_void_p(st)))
        out["top"] = [v.copy() for v, _ in oc.columns(got.value)]
        mark("TopN")
    agg.close(); j2.close()
    if timings is not None:
        stream.synchronize()
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            timings[name] = timings.get(name, 0.0) + a.elapsed_time(b)
        timings["rows"] = {"j1_out": n1, "j2_out": n2, "groups": ng}
    return out


def reference(d: Q3Data) -> Dict[str, torch.Tensor]:
    """the same query with plain torch ops (verification only)"""
    cust_ok = torch.zeros(int(d.c_custkey.max().item()) + 1, dtype=torch.bool, device=d.c_custkey.device)
    cust_ok[d.c_custkey[d.c_seg == SEGMENT]] = True
    om = (d.o_date < DATE) & cust_ok[d.o_custkey]
    n_ok = int(d.o_orderkey.max().item()) + 1
    order_row = torch.full((n_ok,), -1, dtype=torch.int64, device=om.device)
    idx = torch.nonzero(om).flatten()
    order_row[d.o_orderkey[idx]] = idx
    lm = d.l_ship > DATE
    lk = d.l_orderkey[lm]
    orow = order_row[lk]
    keep = orow >= 0
    lk, orow = lk[keep], orow[keep]
    rev = (d.l_price[lm][keep] * (1 - d.l_disc[lm][keep]))
    keys, inv = torch.unique(lk, return_inverse=True)
    s = torch.zeros(keys.numel(), dtype=torch.float64, device=lk.device).scatter_add_(0, inv, rev)
    first = torch.zeros(keys.numel(), dtype=torch.int64, device=lk.device).scatter_(0, inv, orow)
    return {"orderkey": keys, "revenue": s, "o_date": d.o_date[first], "o_prio": d.o_prio[first]}
