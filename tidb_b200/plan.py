"""Descriptor builders: the information executorBuilder hands to the operators.

JoinPlan mirrors what buildHashJoinV2FromChildExecs passes to HashJoinV2Exec
(pkg/executor/builder.go:1771-1931: child schemas, key column indices, LUsed/RUsed, JoinType,
RightAsBuildSide, Build/ProbeFilter); AggPlan mirrors buildHashAggFromChildExec
(builder.go:2106-2181: GroupByItems, AggFuncDescs).  Both render to the C-ABI structs of
include/tidbgpu.h.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from . import abi


@dataclass
class FieldType:
    """types.FieldType reduced to what the path needs: MySQL type code + flag bits."""
    tp: int = abi.TYPE_LONGLONG
    flag: int = 0

    @property
    def not_null(self) -> bool:
        return bool(self.flag & abi.FLAG_NOT_NULL)

    @property
    def unsigned(self) -> bool:
        return bool(self.flag & abi.FLAG_UNSIGNED)


def _i32(vals: Sequence[int]):
    return (C.c_int32 * max(len(vals), 1))(*vals)


def _u32(vals: Sequence[int]):
    return (C.c_uint32 * max(len(vals), 1))(*vals)


@dataclass
class FilterItem:
    """One CNF item `col OP const` / `col OP col` (tg_filter_item)."""
    op: int
    lhs_col: int
    rhs_col: int = -1
    is_real: bool = False
    lhs_unsigned: bool = False
    const_i64: int = 0
    const_f64: float = 0.0
    rhs_unsigned: bool = False

    def to_struct(self) -> abi.TgFilterItem:
        s = abi.TgFilterItem()
        s.op, s.lhs_col, s.rhs_col = self.op, self.lhs_col, self.rhs_col
        s.is_real, s.lhs_unsigned = int(self.is_real), int(self.lhs_unsigned)
        s.rhs_unsigned = int(self.rhs_unsigned)
        s.const_i64, s.const_f64 = self.const_i64, self.const_f64
        return s


def filter_array(items: Sequence[FilterItem]):
    arr = (abi.TgFilterItem * max(len(items), 1))()
    for i, it in enumerate(items):
        arr[i] = it.to_struct()
    return arr


@dataclass
class OtherCond:
    """One CNF item of HashJoinV2Exec.OtherCondition: `side.col OP side.col` (or a constant with rhs_side = -1) over the
    joined row; side 0 = left child, 1 = right child (tg_other_item)."""
    op: int
    lhs_side: int
    lhs_col: int
    rhs_side: int = -1
    rhs_col: int = -1
    is_real: bool = False
    lhs_unsigned: bool = False
    rhs_unsigned: bool = False
    const_i64: int = 0
    const_f64: float = 0.0

    def to_struct(self) -> abi.TgOtherItem:
        s = abi.TgOtherItem()
        s.op, s.is_real = self.op, int(self.is_real)
        s.lhs_side, s.lhs_col, s.rhs_side, s.rhs_col = self.lhs_side, self.lhs_col, self.rhs_side, self.rhs_col
        s.lhs_unsigned, s.rhs_unsigned = int(self.lhs_unsigned), int(self.rhs_unsigned)
        s.const_i64, s.const_f64 = self.const_i64, self.const_f64
        return s


@dataclass
class JoinPlan:
    join_type: int
    left_types: List[FieldType]
    right_types: List[FieldType]
    left_keys: List[int]
    right_keys: List[int]
    build_is_right: bool = True
    lused: Optional[List[int]] = None      # None = all columns (Go nil)
    rused: Optional[List[int]] = None
    build_filter: List[FilterItem] = field(default_factory=list)
    probe_filter: List[FilterItem] = field(default_factory=list)
    device: int = 0
    stream: int = 0
    load_factor: float = 0.0
    other_cond: List[OtherCond] = field(default_factory=list)

    def out_schema(self) -> List[FieldType]:
        lu = self.lused if self.lused is not None else list(range(len(self.left_types)))
        ru = self.rused if self.rused is not None else list(range(len(self.right_types)))
        out = [self.left_types[i] for i in lu] + [self.right_types[i] for i in ru]
        if self.join_type in (abi.JOIN_LEFT_OUTER_SEMI, abi.JOIN_ANTI_LEFT_OUTER_SEMI):
            out.append(FieldType(abi.TYPE_LONGLONG, 0))
        return out

    def to_struct(self) -> Tuple[abi.TgJoinDesc, list]:
        keep = []
        d = abi.TgJoinDesc()
        d.join_type = self.join_type
        d.build_is_right = int(self.build_is_right)
        d.n_left_cols, d.n_right_cols = len(self.left_types), len(self.right_types)
        for name, vals, mk in (("left_types", [t.tp for t in self.left_types], _i32),
                               ("left_flags", [t.flag for t in self.left_types], _u32),
                               ("right_types", [t.tp for t in self.right_types], _i32),
                               ("right_flags", [t.flag for t in self.right_types], _u32),
                               ("left_key_idx", self.left_keys, _i32),
                               ("right_key_idx", self.right_keys, _i32)):
            arr = mk(vals)
            keep.append(arr)
            setattr(d, name, arr)
        d.nkeys = len(self.left_keys)
        if self.lused is None:
            d.n_lused = -1
        else:
            arr = _i32(self.lused); keep.append(arr); d.lused = arr; d.n_lused = len(self.lused)
        if self.rused is None:
            d.n_rused = -1
        else:
            arr = _i32(self.rused); keep.append(arr); d.rused = arr; d.n_rused = len(self.rused)
        d.n_build_filter, d.n_probe_filter = len(self.build_filter), len(self.probe_filter)
        if self.build_filter:
            arr = filter_array(self.build_filter); keep.append(arr); d.build_filter = arr
        if self.probe_filter:
            arr = filter_array(self.probe_filter); keep.append(arr); d.probe_filter = arr
        d.device = self.device
        d.stream = self.stream or None
        d.load_factor = self.load_factor
        d.n_other_cond = len(self.other_cond)
        if self.other_cond:
            arr = (abi.TgOtherItem * len(self.other_cond))(*[o.to_struct() for o in self.other_cond]); keep.append(arr); d.other_cond = arr
        return d, keep


@dataclass
class AggFunc:
    name: int
    arg_col: int = -1
    arg_type: int = abi.TYPE_LONGLONG
    arg_flag: int = 0
    mode: int = abi.AGGMODE_COMPLETE
    arg_col2: int = -1
    arg_expr: int = 0          # abi.ARGEXPR_*: the argument as arg_col * arg_col2 / arg_col * (arg_const - arg_col2)
    arg_const: float = 0.0


@dataclass
class AggPlan:
    col_types: List[FieldType]
    group_by: List[int]
    funcs: List[AggFunc]
    device: int = 0
    stream: int = 0
    expected_groups: int = 0

    def to_struct(self) -> Tuple[abi.TgAggDesc, list]:
        keep = []
        d = abi.TgAggDesc()
        d.n_cols, d.n_group_by = len(self.col_types), len(self.group_by)
        a = _i32([t.tp for t in self.col_types]); keep.append(a); d.col_types = a
        a = _u32([t.flag for t in self.col_types]); keep.append(a); d.col_flags = a
        a = _i32(self.group_by); keep.append(a); d.group_by_cols = a
        fa = (abi.TgAggFunc * max(len(self.funcs), 1))()
        for i, f in enumerate(self.funcs):
            fa[i].name, fa[i].mode, fa[i].arg_col = f.name, f.mode, f.arg_col
            fa[i].arg_type, fa[i].arg_flag, fa[i].arg_col2 = f.arg_type, f.arg_flag, f.arg_col2
            fa[i].arg_expr, fa[i].arg_const = f.arg_expr, f.arg_const
        keep.append(fa)
        d.funcs = fa
        d.n_funcs = len(self.funcs)
        d.device = self.device
        d.stream = self.stream or None
        d.expected_groups = self.expected_groups
        return d, keep


# ---------------------------------------------------------------------------------------------------------------
# Scalar expressions for ProjectionExec (expression.Expression reduced to what the VecEval kernels offload)
# ---------------------------------------------------------------------------------------------------------------
class Expr:
    def ret_type(self, schema: Sequence[FieldType]) -> FieldType:
        raise NotImplementedError


@dataclass
class ColRef(Expr):
    """expression.Column: passed through (EvaluatorSuite swaps plain column references, evaluator.go:128)"""
    idx: int

    def ret_type(self, schema):
        return schema[self.idx]


@dataclass
class Const(Expr):
    """expression.Constant: handed to the kernels as a scalar (the reference materialises a column, vectorized.go:23)"""
    value: float
    is_real: bool = False

    def ret_type(self, schema):
        return FieldType(abi.TYPE_DOUBLE if self.is_real else abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)


@dataclass
class ScalarFunc(Expr):
    """builtinArithmetic{Plus,Minus,Multiply}{Int,Real}Sig / builtin{LT,LE,GT,GE,EQ,NE}{Int,Real}Sig over two arguments
    (kind "arith": op = abi.ARITH_*, kind "cmp": op = abi.CMP_*); the right argument may be a Const"""
    kind: str
    op: int
    args: Tuple[Expr, Expr]
    is_real: bool = False
    a_unsigned: bool = False
    b_unsigned: bool = False

    def ret_type(self, schema):
        if self.kind == "arith":
            return FieldType(abi.TYPE_DOUBLE if self.is_real else abi.TYPE_LONGLONG, 0)
        return FieldType(abi.TYPE_LONGLONG, 0)
