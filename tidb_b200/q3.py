"""TPC-H Q3-shape pipeline on device-resident synthetic columns (BASELINE.json configs[3]).

Plan shape follows the reference's TiFlash MPP plan for Q3 (pkg/planner/core/casetest/tpch/testdata/tpch_suite_out.json:99-123):

    HashAgg(group by l_orderkey, o_orderdate, o_shippriority; sum(l_extendedprice * (1 - l_discount)))
      HashJoin(lineitem.l_orderkey = orders.o_orderkey)            probe: Selection(l_shipdate > D) on lineitem
        HashJoin(orders.o_custkey = customer.c_custkey)            probe: Selection(o_orderdate < D) on orders
          Selection(c_mktsegment = S) on customer                  (build side)

Everything runs through the C-ABI operators: the three Selections are fused into the joins as build / probe
filters (tg_filter_item), the projection l_extendedprice * (1 - l_discount) is two VecEval kernels, and the
aggregation is tg_agg.  o_orderdate and o_shippriority are functionally dependent on the order key, so the
aggregate groups by l_orderkey and carries them as MAX() — the multi-column GROUP BY itself is not offloaded yet
(tg_agg_supported declines it).
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


def run(d: Q3Data, dev, stream) -> Dict[str, torch.Tensor]:
    """-> {orderkey, revenue, o_date, o_prio} (unordered), all device tensors"""
    lib = abi.load_lib()
    st = stream.cuda_stream
    di = dev.index or 0
    # J1: orders (probe, filter o_date < D) ⋈ customer (build, filter c_seg = S); keep o_orderkey, o_date, o_prio
    j1 = DeviceJoin(JoinPlan(abi.JOIN_INNER, [INT] * 4, [INT] * 2, [1], [0], build_is_right=True, lused=[0, 2, 3], rused=[],
                             build_filter=[FilterItem(abi.CMP_EQ, 1, const_i64=SEGMENT)], probe_filter=[FilterItem(abi.CMP_LT, 2, const_i64=DATE)],
                             device=di, stream=st))
    j1.build([d.c_custkey, d.c_seg])
    n1, c1, _ = j1.probe([d.o_orderkey, d.o_custkey, d.o_date, d.o_prio])
    ok, od, op = (_view(p, n1, dev).clone() for p in c1)
    j1.close()
    # J2: lineitem (probe, filter l_ship > D) ⋈ J1 (build on o_orderkey); keep l_orderkey, l_price, l_disc, o_date, o_prio
    j2 = DeviceJoin(JoinPlan(abi.JOIN_INNER, [INT, DBL, DBL, INT], [INT] * 3, [0], [0], build_is_right=True, lused=[0, 1, 2], rused=[1, 2],
                             probe_filter=[FilterItem(abi.CMP_GT, 3, const_i64=DATE)], device=di, stream=st))
    j2.build([ok, od, op])
    n2, c2, _ = j2.probe([d.l_orderkey, d.l_price, d.l_disc, d.l_ship])
    lk = _view(c2[0], n2, dev); price = _view(c2[1], n2, dev, "<f8"); disc = _view(c2[2], n2, dev, "<f8")
    jd = _view(c2[3], n2, dev); jp = _view(c2[4], n2, dev)
    # projection: l_extendedprice * (1 - l_discount)   (builtinArithmeticMinusRealSig / MultiplyRealSig)
    one_minus = torch.empty(n2, dtype=torch.float64, device=dev); rev = torch.empty(n2, dtype=torch.float64, device=dev)
    nb = torch.empty((n2 + 7) // 8 + 8, dtype=torch.uint8, device=dev)

    def col(t):
        c = abi.TgColumn(); c.length = t.numel(); c.data = t.data_ptr() if t.numel() else None; c.elem_len = 8; c.null_bitmap = None
        return c
    ones = torch.ones(n2, dtype=torch.float64, device=dev)
    c_ones, c_disc, c_price = col(ones), col(disc), col(price)
    abi.check(lib.tg_vec_arith_real(di, 1, abi.ARITH_MINUS, C.byref(c_ones), C.byref(c_disc), C.c_double(0), C.c_void_p(one_minus.data_ptr()),
                                    C.c_void_p(nb.data_ptr()), C.c_void_p(st)))
    c_om = col(one_minus)
    abi.check(lib.tg_vec_arith_real(di, 1, abi.ARITH_MUL, C.byref(c_price), C.byref(c_om), C.c_double(0), C.c_void_p(rev.data_ptr()),
                                    C.c_void_p(nb.data_ptr()), C.c_void_p(st)))
    # aggregate
    agg = DeviceAgg(AggPlan([INT, DBL, INT, INT], [0], [AggFunc(abi.AGG_FIRSTROW, 0), AggFunc(abi.AGG_SUM, 1, abi.TYPE_DOUBLE),
                                                        AggFunc(abi.AGG_MAX, 2), AggFunc(abi.AGG_MAX, 3)],
                            device=di, stream=st, expected_groups=max(n1, 1)))
    agg.push([lk, rev, jd, jp])
    ng, ca, _ = agg.finish()
    out = {"orderkey": _view(ca[0], ng, dev).clone(), "revenue": _view(ca[1], ng, dev, "<f8").clone(),
           "o_date": _view(ca[2], ng, dev).clone(), "o_prio": _view(ca[3], ng, dev).clone()}
    agg.close(); j2.close()
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
