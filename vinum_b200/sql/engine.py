"""Query execution on the device operators.

Host control flow only: it decides WHICH kernels run over WHICH columns, the way the
reference's planner chains its operators (vinum/planner/planner.py:330-507):

    scan(used columns -> HBM) -> WHERE -> [pre-aggregate projection] -> GROUP BY / aggregate
        -> HAVING -> ORDER BY -> SELECT projection (+ column names) -> LIMIT/OFFSET -> host table

Numeric / temporal / boolean columns live in HBM as `DeviceColumn`s and every operator on
them is a CUDA kernel (vinum_b200.ops, vinum_b200.aggregate).  An aggregate query never
materialises its WHERE: a `column <op> literal` predicate is fused into the aggregate kernel,
anything else is passed to it as a byte mask.

String columns stay on the HOST ("string predicate falls back to host"): their
predicates run the reference's own NumPy / pyarrow.compute calls and only the resulting mask
travels to the device; string GROUP BY keys are dictionary-encoded on the host and grouped
on the device by their int32 codes.  Scalar functions (`to_int`, `sqrt`, `np.*` ...) are
NumPy callables in the reference (vinum/core/functions.py:341-367) and are applied on the
host here as well (functions.py in this package).
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from .. import _lib as L
from .. import ops
from ..aggregate import Aggregator
from ..strings import StringMinMax
from ..device import DeviceBatch, DeviceColumn, Stream, default_stream, vk_dtype_of
from .ast import (NUMPY_AGG_MAPPING, Column, Expression, Literal, Node, Op, Query, SortOrder,
                  contains_aggregate, is_aggregate_call, walk)
from .functions import call_host_function
from .parser import ParserError, parse_sql

_TOPK_MIN_ROWS = 1 << 20   # below this a full radix sort costs less than the select's round trips


class OperatorError(Exception):
    """vinum/errors/__init__.py: OperatorError."""


class HostColumn:
    """A column the device path does not handle (strings): a pyarrow array on the host."""

    def __init__(self, arr):
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
        self.arr = arr

    @property
    def length(self) -> int:
        return len(self.arr)

    def to_numpy(self) -> np.ndarray:
        return self.arr.to_numpy(zero_copy_only=False)


Value = Union[DeviceColumn, HostColumn, int, float, bool, str, None, list]

_AGG_CODES = {"count_star": L.AGG_COUNT_STAR, "count": L.AGG_COUNT, "min": L.AGG_MIN, "max": L.AGG_MAX,
              "sum": L.AGG_SUM, "avg": L.AGG_AVG}
_CMP_FLIP = {"==": "==", "!=": "!=", ">": "<", ">=": "<=", "<": ">", "<=": ">="}
_ARITH = {Op.ADDITION: "+", Op.SUBTRACTION: "-", Op.MULTIPLICATION: "*", Op.DIVISION: "/", Op.MODULUS: "%",
          Op.BINARY_AND: "&", Op.BINARY_OR: "|", Op.BINARY_XOR: "#"}
_CMP = {Op.EQUALS: "==", Op.NOT_EQUALS: "!=", Op.GREATER_THAN: ">", Op.GREATER_THAN_OR_EQUAL: ">=",
        Op.LESS_THAN: "<", Op.LESS_THAN_OR_EQUAL: "<="}
_NP_ARITH = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "%": np.mod, "&": np.bitwise_and,
             "|": np.bitwise_or, "#": np.bitwise_xor}
_NP_CMP = {"==": np.equal, "!=": np.not_equal, ">": np.greater, ">=": np.greater_equal, "<": np.less,
           "<=": np.less_equal}


def _is_col(v) -> bool:
    return isinstance(v, (DeviceColumn, HostColumn))


def _is_temporal_value(v) -> bool:
    if isinstance(v, (np.datetime64, np.timedelta64)):
        return True
    return isinstance(v, DeviceColumn) and v.arrow_type is not None and pa.types.is_temporal(v.arrow_type)


def _temporal_numpy_dtype(t: pa.DataType) -> Optional[str]:
    """NumPy dtype a temporal Arrow column is viewed as on the host (what to_numpy() of the
    reference's NumPy path yields)."""
    if pa.types.is_timestamp(t):
        return f"datetime64[{t.unit}]"
    if pa.types.is_date32(t):
        return "datetime64[D]"
    if pa.types.is_date64(t):
        return "datetime64[ms]"
    if pa.types.is_duration(t):
        return f"timedelta64[{t.unit}]"
    return None


def _is_num(v) -> bool:
    return isinstance(v, (int, float, bool, np.integer, np.floating, np.bool_))


class Frame:
    """Named columns of equal length: the batch flowing between operators."""

    def __init__(self, n: int, cols: Optional[Dict[str, Value]] = None):
        self.n = n
        self.cols: Dict[str, Value] = dict(cols or {})
        self.virtual = False   # no table column behind it (literal-only SELECT)


class Engine:
    def __init__(self, table: pa.Table, stream: Optional[Stream] = None, exchange=None):
        self.table = table
        self._st = stream
        # sharded execution (vinum_b200.sharded): `table` is this rank's row range; aggregate results go
        # through `exchange(aggregator, stream) -> (aggregator, raw groups)` before they are finalised
        self.exchange = exchange
        self.stats = {"kernels_before": int(L.lib.vk_launch_count())}

    @property
    def st(self) -> Stream:
        """The CUDA stream every operator of this query runs on (created on first use: binding and
        planning need no device)."""
        if self._st is None:
            self._st = default_stream()
        return self._st

    # ================================================================ binding
    def _bind(self, q: Query) -> Query:
        """Alias substitution, column validation, aggregate detection (Binder.bind,
        vinum/planner/binder.py:41-89)."""
        aliases = {e.alias: e for e in q.select if e.alias}

        def subst(node):
            if node is None:
                return None
            if isinstance(node, Column) and node.name in aliases:
                return _copy(aliases[node.name])
            if isinstance(node, Expression):
                return Expression(node.op, tuple(subst(a) for a in node.args), node.function_name, node.alias)
            return node

        where = subst(q.where)
        group_by = tuple(subst(g) for g in q.group_by)
        having = subst(q.having)
        order_by = tuple(subst(o) for o in q.order_by)
        names = set(self.table.schema.names)
        for clause in (q.select, (where,), group_by, (having,), order_by):
            for root in clause:
                if root is None:
                    continue
                for node in walk(root):
                    if isinstance(node, Column) and node.name not in names:
                        raise ParserError(f"Column '{node.name}' is not found.")  # binder.py:141-146
        bound = Query(q.select, q.distinct, where, group_by, having, order_by, q.sort_order, q.limit, q.offset,
                      q.has_group_clause)
        bound.is_aggregate = (q.distinct or q.has_group_clause or any(contains_aggregate(e) for e in q.select))
        if (q.distinct or q.has_group_clause) and group_by:
            self._check_group_by(q.select, group_by)
        return bound

    @staticmethod
    def _check_group_by(select, group_by) -> None:
        """binder.py:209-265: a SELECT item is a GROUP BY column / expression or contains an aggregate."""
        usage = 'Only aggregate functions and columns present in the "GROUP BY" clause are allowed.'
        gb_keys = {g.key() for g in group_by}
        for e in select:
            if isinstance(e, Column):
                if e.key() not in gb_keys:
                    raise ParserError(f'Column "{e.name}" is not part of the "GROUP BY" clause. {usage}.')
            elif isinstance(e, Expression):
                if e.key() not in gb_keys and not contains_aggregate(e):
                    opn = e.function_name or e.op.name
                    raise ParserError(f'Operator "{opn}" is neither aggregate function nor part of the '
                                      f'"GROUP BY" clause. {usage}.')
            else:
                raise ParserError(f'Literal "{e.value}" is not allowed in the aggregate query. {usage}.')

    # ============================================================== execution
    def execute(self, q: Query) -> pa.Table:
        q = self._bind(q)
        used = _used_columns(q)
        streamed = self._stream_aggregate(q) if q.is_aggregate else None
        if streamed is not None:
            frame, resolved = streamed
        elif q.is_aggregate:
            frame = self._scan(used)
            frame, resolved = self._aggregate(q, frame)
        else:
            frame = self._scan(used)
            resolved = {}
            if not used:
                # no column is referenced at all: the reference skips the table and projects over
                # ONE empty batch (EmptyTableReaderOperator, planner.py:345-362, algebra.py:282-287);
                # the projected arrays / broadcast literals define the row count
                frame.n = 1
                frame.virtual = True
            if q.where is not None:
                frame = self._filter(frame, self._predicate(q.where, frame, resolved))
        return self._tail(q, frame, resolved)

    # -------------------------------------------------------------------- scan
    def _scan(self, used: Sequence[str]) -> Frame:
        """TableReaderOperator + the used-columns projection (planner.py:345-375): only the
        referenced columns are copied to the device."""
        t = self.table
        frame = Frame(t.num_rows)
        for name in used:
            col = t.column(name)
            if vk_dtype_of(col.type) is not None:
                frame.cols[name] = DeviceColumn.from_arrow(col, self.st)
            else:
                frame.cols[name] = HostColumn(col)
        return frame

    # ------------------------------------------------------------- expressions
    def _eval(self, node: Node, frame: Frame, resolved: Dict) -> Value:
        k = node.key() if not isinstance(node, Literal) else None
        if k is not None and k in resolved:
            return frame.cols[resolved[k]]
        if isinstance(node, Literal):
            return node.value
        if isinstance(node, Column):
            if node.name not in frame.cols:
                raise ParserError(f'Column "{node.name}" is not part of the "GROUP BY" clause. Only aggregate '
                                  f'functions and columns present in the "GROUP BY" clause are allowed..')
            return frame.cols[node.name]
        op = node.op
        if op in _ARITH:
            ch = self._chain(node, frame, resolved)
            if ch is not None and len(ch) >= 3:
                # two or more operations over plain columns: one pass, no intermediate array (row a3)
                self.stats["fused_expr"] = self.stats.get("fused_expr", 0) + 1
                return ops.eval_chain(ch, self.st)
            return self._fold(_ARITH[op], [self._eval(a, frame, resolved) for a in node.args], self._arith2)
        if op == Op.NEGATION or op == Op.BINARY_NOT:
            x = self._eval(node.args[0], frame, resolved)
            sym = "neg" if op == Op.NEGATION else "~"
            if isinstance(x, DeviceColumn):
                return ops.arith(sym, x, None, self.st)
            if isinstance(x, HostColumn):
                raise OperatorError(f"operator {sym} is not defined for column type {x.arr.type}")
            return (np.negative if op == Op.NEGATION else np.invert)(x).item()
        if op in _CMP:
            fused = self._chain_compare(node, frame, resolved)
            if fused is not None:
                self.stats["fused_expr"] = self.stats.get("fused_expr", 0) + 1
                return ops.compare_chains(fused[0], fused[1], fused[2], self.st)
            a, b = (self._eval(x, frame, resolved) for x in node.args)
            return self._compare(_CMP[op], a, b)
        if op in (Op.AND, Op.OR):
            masks = [self._as_mask(self._eval(a, frame, resolved), frame.n) for a in node.args]
            out = masks[0]
            for m in masks[1:]:  # folded left like BINARY_EXPRESSIONS (core/base.py:145-151)
                out = ops.mask_and(out, m, self.st) if op == Op.AND else ops.mask_or(out, m, self.st)
            return out
        if op == Op.NOT:
            return ops.mask_not(self._as_mask(self._eval(node.args[0], frame, resolved), frame.n), self.st)
        if op in (Op.IS_NULL, Op.IS_NOT_NULL):
            x = self._eval(node.args[0], frame, resolved)
            if isinstance(x, DeviceColumn):
                return ops.is_null(x, self.st) if op == Op.IS_NULL else ops.is_valid(x, self.st)
            if isinstance(x, HostColumn):
                m = pc.is_null(x.arr) if op == Op.IS_NULL else pc.is_valid(x.arr)
                return self._upload_mask(m.to_numpy(zero_copy_only=False))
            return (x is None) == (op == Op.IS_NULL)
        if op in (Op.IN, Op.NOT_IN):
            x = self._eval(node.args[0], frame, resolved)
            values = node.args[1].value
            neg = op == Op.NOT_IN
            if isinstance(x, DeviceColumn) and all(_is_num(v) for v in values):
                return ops.isin(x, list(values), neg, self.st)
            if isinstance(x, HostColumn):
                m = np.isin(x.to_numpy(), values, invert=neg)  # expressions.py:39-40
                return self._upload_mask(m)
            if isinstance(x, DeviceColumn):
                raise OperatorError("IN over a numeric column needs numeric literals")
            return (x in values) != neg
        if op in (Op.BETWEEN, Op.NOT_BETWEEN):
            x, lo, hi = (self._eval(a, frame, resolved) for a in node.args)
            neg = op == Op.NOT_BETWEEN
            if isinstance(x, DeviceColumn) and _is_num(lo) and _is_num(hi):
                return ops.between(x, lo, hi, neg, self.st)
            if neg:   # expressions.py:46-48
                return ops.mask_or(self._as_mask(self._compare("<", x, lo), frame.n),
                                   self._as_mask(self._compare(">", x, hi), frame.n), self.st)
            return ops.mask_and(self._as_mask(self._compare(">=", x, lo), frame.n),
                                self._as_mask(self._compare("<=", x, hi), frame.n), self.st)
        if op in (Op.LIKE, Op.NOT_LIKE):
            x, pattern = (self._eval(a, frame, resolved) for a in node.args)
            if not isinstance(x, HostColumn) or not isinstance(pattern, str):
                raise OperatorError("LIKE needs a string column and a string pattern")
            rx = re.compile("^" + pattern.replace("_", ".").replace("%", ".*") + "$")  # functions.py:299-304
            inv = op == Op.NOT_LIKE
            m = np.fromiter((bool(rx.match(v)) != inv for v in x.to_numpy()), dtype=bool, count=x.length)
            return self._upload_mask(m)
        if op == Op.CONCAT:
            return self._host_call("concat", [self._eval(a, frame, resolved) for a in node.args], frame.n)
        if op == Op.FUNCTION:
            if is_aggregate_call(node):
                raise OperatorError(f"aggregate function {node.function_name}() is not allowed here")
            return self._host_call(node.function_name, [self._eval(a, frame, resolved) for a in node.args], frame.n)
        raise OperatorError(f"unsupported expression {op}")

    # ---------------------------------------------------------- expression fusion
    def _chain(self, node: Node, frame: Frame, resolved: Dict):
        """`node` as a left-deep chain [(None, t0), (op1, t1), ...] over null-free int64 / float64 device
        columns and numeric literals (vk_expr.cuh), or None.  a + (b * c) is rewritten (b * c) + a for
        the commutative operators: the same IEEE / wrapping operations in the same tree order."""
        if os.environ.get("VINUM_B200_FUSE_EXPR", "1") == "0":
            return None
        if isinstance(node, Literal):
            return [(None, node.value)] if ops.chain_term_ok(node.value) else None
        k = node.key()
        if k in resolved:
            v = frame.cols[resolved[k]]
            return [(None, v)] if isinstance(v, DeviceColumn) and ops.chain_term_ok(v) else None
        if isinstance(node, Column):
            v = frame.cols.get(node.name)
            return [(None, v)] if isinstance(v, DeviceColumn) and ops.chain_term_ok(v) else None
        if not isinstance(node, Expression) or node.op not in _ARITH or len(node.args) < 2:
            return None
        sym = _ARITH[node.op]
        out = self._chain(node.args[0], frame, resolved)
        for arg in node.args[1:]:   # n-ary nodes fold left (BINARY_EXPRESSIONS, core/base.py:145-151)
            rhs = self._chain(arg, frame, resolved)
            if out is None or rhs is None:
                return None
            if len(rhs) == 1:
                out = out + [(sym, rhs[0][1])]
            elif len(out) == 1 and sym in ("+", "*", "&", "|", "#"):
                out = rhs + [(sym, out[0][1])]
            else:
                return None
            if len(out) > ops.MAX_CHAIN_TERMS:
                return None
        if not any(isinstance(t, DeviceColumn) for _, t in out):
            return None   # constants only: folded on the host
        floats = any((isinstance(t, DeviceColumn) and t.dtype == L.F64) or isinstance(t, (float, np.floating)) for _, t in out)
        if floats and any(o in ("&", "|", "#") for o, _ in out[1:]):
            return None   # NumPy raises for bitwise operators on floats: keep that path
        return out

    def _chain_compare(self, node: Node, frame: Frame, resolved: Dict):
        """(lhs chain, op, rhs chain) when a comparison's two sides are chains and fusing saves a pass
        (`a * 10 > b`); None for plain `column <op> column / literal`, which has its own kernels."""
        if not (isinstance(node, Expression) and node.op in _CMP and len(node.args) == 2):
            return None
        lc = self._chain(node.args[0], frame, resolved)
        rc = self._chain(node.args[1], frame, resolved) if lc is not None else None
        if lc is None or rc is None or (len(lc) < 2 and len(rc) < 2):
            return None
        return lc, _CMP[node.op], rc

    def _predicate(self, w: Node, frame: Frame, resolved: Dict) -> "ops.Predicate":
        """WHERE / HAVING as the predicate the consumer kernel evaluates itself: `column <op> literal`
        (vectorised in registers), a fused expression chain, or -- anything else -- a byte mask."""
        if isinstance(w, Expression) and w.op in _CMP and len(w.args) == 2:
            a, b = w.args
            sym = _CMP[w.op]
            if isinstance(b, Column) and isinstance(a, Literal):
                a, b, sym = b, a, _CMP_FLIP[sym]
            if isinstance(a, Column) and isinstance(b, Literal) and _is_num(b.value) and not isinstance(b.value, bool):
                col = frame.cols.get(a.name)
                if isinstance(col, DeviceColumn) and col.dtype != L.BOOL8:
                    return ops.Predicate.compare(col, sym, b.value)
            fused = self._chain_compare(w, frame, resolved)
            if fused is not None:
                self.stats["fused_expr"] = self.stats.get("fused_expr", 0) + 1
                return ops.Predicate.expr(*fused)
        return ops.Predicate.from_mask(self._as_mask(self._eval(w, frame, resolved), frame.n))

    @staticmethod
    def _fold(sym, args, fn):
        out = args[0]
        for a in args[1:]:
            out = fn(sym, out, a)
        return out

    def _arith2(self, sym: str, a: Value, b: Value) -> Value:
        if isinstance(a, HostColumn) or isinstance(b, HostColumn) or isinstance(a, str) or isinstance(b, str):
            raise OperatorError(f"operator {sym} is not defined for string operands")
        if _is_temporal_value(a) or _is_temporal_value(b):
            # datetime / timedelta arithmetic is NumPy's in the reference (np.add / np.subtract on
            # datetime64 arrays): a host step here as well
            x = self._device_to_numpy(a) if isinstance(a, DeviceColumn) else a
            y = self._device_to_numpy(b) if isinstance(b, DeviceColumn) else b
            return self._from_host(_NP_ARITH[sym](x, y))
        if isinstance(a, DeviceColumn) or isinstance(b, DeviceColumn):
            return ops.arith(sym, a, b, self.st)
        with np.errstate(all="ignore"):
            return _NP_ARITH[sym](a, b).item()   # constant folding with the same NumPy ufunc

    def _compare(self, sym: str, a: Value, b: Value) -> Value:
        if isinstance(a, DeviceColumn) and (isinstance(b, DeviceColumn) or _is_num(b)):
            return ops.compare(a, sym, b, self.st)
        if isinstance(b, DeviceColumn) and _is_num(a):
            return ops.compare(b, _CMP_FLIP[sym], a, self.st)
        if _is_col(a) or _is_col(b):
            # a string operand: the reference's NumPy lambda on object arrays (expressions.py:30-36)
            x = a.to_numpy() if isinstance(a, HostColumn) else (a.to_numpy(self.st) if isinstance(a, DeviceColumn) else a)
            y = b.to_numpy() if isinstance(b, HostColumn) else (b.to_numpy(self.st) if isinstance(b, DeviceColumn) else b)
            m = _NP_CMP[sym](x, y)
            return self._upload_mask(np.asarray(m, dtype=bool))
        return bool(_NP_CMP[sym](a, b))

    def _host_call(self, name: str, args: List[Value], n: int) -> Value:
        host_args = []
        for a in args:
            if isinstance(a, DeviceColumn):
                host_args.append(self._device_to_numpy(a))
            elif isinstance(a, HostColumn):
                host_args.append(a.arr)
            else:
                host_args.append(a)
        res = call_host_function(name, host_args)
        return self._from_host(res)

    def _device_to_numpy(self, a: DeviceColumn) -> np.ndarray:
        """The NumPy view the reference computes on (RecordBatch.get_np_column, record_batch.py:
        100-125): NULL -> NaN for numbers, NaT for temporal types."""
        arr = a.to_numpy(self.st)
        valid = a.validity_to_numpy(self.st)
        tdt = _temporal_numpy_dtype(a.arrow_type) if a.arrow_type is not None else None
        if tdt is not None:
            arr = arr.view(tdt) if arr.dtype.itemsize == 8 else arr.astype(np.int64).astype(tdt)   # date32: int32 days
            if valid is not None:
                arr = arr.copy()
                arr[~valid] = np.datetime64("NaT") if tdt.startswith("datetime") else np.timedelta64("NaT")
            return arr
        if valid is not None:
            arr = arr.astype(np.float64)
            arr[~valid] = np.nan
        return arr

    def _from_host(self, res) -> Value:
        if isinstance(res, (pa.Array, pa.ChunkedArray)):
            if vk_dtype_of(res.type) is not None:
                return DeviceColumn.from_arrow(res, self.st)
            return HostColumn(res)
        if isinstance(res, np.ndarray) and res.shape != () and res.dtype.kind in "Mm":
            return self._from_host(pa.array(res))      # NaT -> NULL; datetime64[D] -> date32, [s] -> timestamp[s]
        if isinstance(res, (np.datetime64, np.timedelta64)):
            return res
        if isinstance(res, np.ndarray) and res.shape != ():
            if res.dtype.kind in "iufb" and res.dtype.itemsize <= 8 and res.dtype != np.float16:
                col = DeviceColumn.from_numpy(res if res.dtype != np.bool_ else res.astype(np.bool_), self.st)
                col._host_ref = res
                return col
            return HostColumn(pa.array(res))
        if isinstance(res, np.generic):
            return res.item()
        if isinstance(res, np.ndarray):
            return res.item()
        return res

    def _upload_mask(self, m: np.ndarray) -> DeviceColumn:
        m = np.ascontiguousarray(m, dtype=np.bool_)
        col = DeviceColumn.from_numpy(m, self.st)
        col._host_ref = m
        return col

    def _as_mask(self, v: Value, n: int) -> DeviceColumn:
        if isinstance(v, DeviceColumn):
            if v.dtype != L.BOOL8:
                return ops.compare(v, "!=", 0, self.st)
            if v.has_nulls:
                # a NULL condition drops the row (RecordBatch.filter with a nullable mask keeps only
                # true slots on this path, record_batch.py:85-90): the raw byte under a NULL is garbage
                plain = DeviceColumn(v.data, None, v.offset, v.length, L.BOOL8, pa.bool_(), 0, v.data_ptr)
                return ops.mask_and(plain, ops.is_valid(v, self.st), self.st)
            return v
        if isinstance(v, HostColumn):
            raise OperatorError("a string expression cannot be used as a condition")
        return self._upload_mask(np.full(n, bool(v), dtype=np.bool_))

    # ------------------------------------------------------- filter / sort
    def _filter(self, frame: Frame, mask) -> Frame:
        """FilterOperator (algebra.py:108-123): one compaction kernel over every device column; `mask` is
        a byte-mask column or an `ops.Predicate` the kernel evaluates itself (no mask is written)."""
        names = [k for k, v in frame.cols.items() if isinstance(v, DeviceColumn)]
        host = [k for k, v in frame.cols.items() if isinstance(v, HostColumn)]
        pred = mask if isinstance(mask, ops.Predicate) else ops.Predicate.from_mask(mask)
        if (host or not names) and pred.mask is None:
            # host (string) columns are filtered on the host: they need the mask itself
            if pred.chains is not None:
                mask = ops.compare_chains(pred.chains[0], pred.op, pred.chains[1], self.st)
            else:
                mask = ops.compare(pred.column, pred.op, pred.scalar, self.st)
            pred = ops.Predicate.from_mask(mask)
        else:
            mask = pred.mask
        out = Frame(0)
        if names:
            batch = DeviceBatch([frame.cols[k] for k in names], names, frame.n)
            res = ops.filter_batch(batch, pred, self.st)
            out.n = res.num_rows
            for k, c in zip(names, res.columns):
                out.cols[k] = c
        if host or not names:
            hm = mask.to_numpy(self.st).astype(bool)
            out.n = int(hm.sum())
            for k in host:
                out.cols[k] = HostColumn(frame.cols[k].arr.filter(pa.array(hm)))
        return out

    def _sort(self, frame: Frame, keys: List[Value], orders: Sequence[SortOrder], top: Optional[int] = None) -> Frame:
        """SortOperator (algebra.py:126-201) -> device radix sort + gather.  With a LIMIT only the
        first `top` = offset + limit rows of the permutation are gathered (SliceOperator,
        algebra.py:204-247, pulled below the gather)."""
        dev_keys, dev_orders = [], []
        for k, o in zip(keys, orders):
            if not _is_col(k):
                continue  # a constant key does not order anything
            if isinstance(k, HostColumn):
                k = self._string_ranks(k)
            if k.dtype == L.BOOL8:
                raise OperatorError("Sorting by boolean column is not supported yet.")  # algebra.py:191-201
            dev_keys.append(k)
            dev_orders.append(L.DESC if o == SortOrder.DESC else L.ASC)
        if not dev_keys or frame.n == 0:
            return frame
        m = frame.n if top is None else min(top, frame.n)
        if 0 < m < frame.n and frame.n >= _TOPK_MIN_ROWS and os.environ.get("VINUM_B200_TOPK", "1") != "0":
            # radix select of the LIMIT's rows, then a sort of those only (1e8 rows, k <= 1e5: 1.7 ms vs 8.6 ms)
            idx = ops.sort_top(dev_keys, dev_orders, m, self.st)
            sorted0 = None
            self.stats["sort_topk"] = True
        else:
            # the first key's own column comes out of the last radix pass already sorted (no gather)
            idx, sorted0 = ops.sort_indices_keys(dev_keys, dev_orders, self.st)
            if m < frame.n:
                idx = idx.slice(0, m)
                sorted0 = sorted0.slice(0, m) if sorted0 is not None else None
        out = Frame(m)
        hidx = None
        for name, v in frame.cols.items():
            if sorted0 is not None and v is dev_keys[0]:
                out.cols[name] = sorted0
            elif isinstance(v, DeviceColumn):
                out.cols[name] = ops.take(v, idx, self.st)
            else:
                if hidx is None:
                    hidx = pa.array(idx.to_numpy(self.st))
                out.cols[name] = HostColumn(v.arr.take(hidx))
        return out

    def _string_ranks(self, col: HostColumn) -> DeviceColumn:
        """Order-preserving int32 codes of a string column (NULLs stay NULL -> sorted last)."""
        arr = col.arr
        uniq = pc.unique(arr).drop_null()
        order = pc.sort_indices(uniq)
        sorted_uniq = uniq.take(order)
        codes = pc.index_in(arr, value_set=sorted_uniq)
        return DeviceColumn.from_arrow(codes, self.st)

    # ------------------------------------------------------------- aggregate
    def _streamable(self, q: Query) -> bool:
        """Eligibility test of the streaming fast path without running it."""
        return self._stream_aggregate(q, dry_run=True) is not None

    def _stream_aggregate(self, q: Query, dry_run: bool = False) -> Optional[Tuple[Frame, Dict]]:
        """The streaming fast path (vinum_b200.executor.filter_aggregate): when every group key and
        aggregate argument is a plain null-free numeric column and the WHERE is `column <op>
        literal`, the table is never resident as a whole -- 2^24-row chunks of just the
        referenced columns are copied host -> device on a copy stream (true DMA from pinned
        Arrow buffers) while the fused filter -> aggregate kernel consumes the previous chunk."""
        from ..executor import filter_aggregate
        if q.distinct or not q.group_by:
            return None
        schema = self.table.schema

        def plain(node) -> Optional[str]:
            if not isinstance(node, Column):
                return None
            col = self.table.column(node.name)
            dt = vk_dtype_of(col.type)
            if dt is None or dt == L.BOOL8 or col.null_count:
                return None
            return node.name

        keys = [plain(g) for g in q.group_by]
        if any(k is None for k in keys) or len(set(keys)) != len(keys):
            return None
        calls: List[Expression] = []
        seen = set()
        for root in list(q.select) + [q.having] + list(q.order_by):
            if root is None:
                continue
            for node in walk(root):
                if is_aggregate_call(node) and node.key() not in seen:
                    seen.add(node.key())
                    calls.append(node)
        if not calls:
            return None
        funcs = []
        for i, call in enumerate(calls):
            fname = NUMPY_AGG_MAPPING.get(call.function_name.lower(), call.function_name.lower())
            if fname == "count_star":
                funcs.append(("COUNT_STAR", "", f"__agg{i}"))
                continue
            if len(call.args) != 1 or plain(call.args[0]) is None:
                return None
            funcs.append((fname.upper(), call.args[0].name, f"__agg{i}"))
        where = None
        if q.where is not None:
            w = q.where
            if not (isinstance(w, Expression) and w.op in _CMP and len(w.args) == 2):
                return None
            a, b = w.args
            sym = _CMP[w.op]
            if isinstance(b, Column) and isinstance(a, Literal):
                a, b, sym = b, a, _CMP_FLIP[sym]
            if not (isinstance(b, Literal) and _is_num(b.value) and not isinstance(b.value, bool)) or plain(a) is None:
                return None
            where = (a.name, sym, b.value)
        if dry_run:
            return Frame(0), {}
        stats: dict = {}
        rb = filter_aggregate(self.table, keys, funcs, where, stats=stats, exchange=self.exchange)
        self.stats.update({"agg_path": stats.get("agg_path"), "streamed": True, "h2d_bytes": stats.get("h2d_bytes"),
                           "d2h_bytes": stats.get("d2h_bytes")})
        out = Frame(rb.num_rows)
        resolved: Dict = {}
        for i, (g, name) in enumerate(zip(q.group_by, keys)):
            col = self._from_host(rb.column(i))
            out.cols[f"__key{i}"] = col
            out.cols[name] = col
            resolved[g.key()] = f"__key{i}"
        for i, call in enumerate(calls):
            out.cols[f"__agg{i}"] = self._from_host(rb.column(len(keys) + i))
            resolved[call.key()] = f"__agg{i}"
        return out, resolved

    def _aggregate(self, q: Query, frame: Frame) -> Tuple[Frame, Dict]:
        """WHERE + pre-aggregate projection + AggregateOperator (planner.py:383-470,
        core/aggregate.py:32-124) on the fused / masked device aggregate, one batch."""
        state = _AggState(q)
        try:
            self._agg_update(q, frame, state)
            return self._agg_finish(state)
        finally:
            state.close()

    def _agg_update(self, q: Query, frame: Frame, state: "_AggState") -> None:
        """One batch of the stream into the aggregate state (AggregateOperator.next,
        aggregate.py:114-121): predicate, key and argument expressions are evaluated on this
        batch's columns, then ONE update of the device aggregate."""
        st = self.st
        # ---- predicate: fused `column <op> literal` or a byte mask; never a compaction ----
        pred = self._predicate(q.where, frame, {}) if q.where is not None else None

        first = state.agg is None
        key_vals: List[DeviceColumn] = []
        keep = []
        for k, g in enumerate(state.group_exprs):
            v = self._eval(g, frame, {})
            if not _is_col(v):
                v = self._from_host(np.full(frame.n, v))
            if isinstance(v, HostColumn):
                if first:
                    state.key_kind.append("dict")
                    state.dicts.append(({}, [], v.arr.type))
                codes = _encode_with_dictionary(v.arr, state.dicts[k][0], state.dicts[k][1])
                keep.append(codes)
                key_vals.append(DeviceColumn.from_arrow(codes, st))
            elif v.dtype == L.BOOL8:
                if first:
                    state.key_kind.append("bool")
                    state.dicts.append(None)
                key_vals.append(DeviceColumn(v.data, v.validity, v.offset, v.length, L.U8, pa.uint8(), v.null_count,
                                             v.data_ptr))
            else:
                if first:
                    state.key_kind.append(None)
                    state.dicts.append(None)
                key_vals.append(v)

        specs, agg_vals, str_calls = [], [], []
        for call in state.agg_calls:
            fname = NUMPY_AGG_MAPPING.get(call.function_name.lower(), call.function_name.lower())
            code = _AGG_CODES[fname]
            if code == L.AGG_COUNT_STAR:
                specs.append((code, None))
                agg_vals.append(None)
                continue
            if len(call.args) != 1:
                raise OperatorError(f"{fname}() takes exactly one argument")
            v = self._eval(call.args[0], frame, {})
            if not _is_col(v):
                v = self._from_host(np.full(frame.n, v))
            if isinstance(v, HostColumn):
                if code in (L.AGG_MIN, L.AGG_MAX):
                    # rank codes of this batch reduced on the device, winners merged at finish (vinum_b200.strings);
                    # the slot in the main aggregate is a validity COUNT that _agg_finish replaces
                    str_calls.append((len(specs), code == L.AGG_MIN, v))
                    code = L.AGG_COUNT
                elif code != L.AGG_COUNT:
                    raise OperatorError(f"{fname}() over a string column is not on the device path")
                # COUNT(str): only the validity matters
                valid = pc.is_valid(v.arr)
                dummy = pa.array(np.zeros(v.length, dtype=np.uint8), mask=~valid.to_numpy(zero_copy_only=False))
                keep.append(dummy)
                v = DeviceColumn.from_arrow(dummy, st)
            specs.append((code, v.arrow_type))
            agg_vals.append(v)

        if first:
            state.agg = Aggregator([k.arrow_type for k in key_vals], specs)
        if key_vals or any(v is not None for v in agg_vals):
            state.agg.update(key_vals, agg_vals, pred, st)
        else:
            state.agg.update_count_rows(frame.n, pred, st)
        for i, is_min, v in str_calls:
            if i not in state.str_minmax:
                state.str_minmax[i] = StringMinMax([k.arrow_type for k in key_vals], is_min, v.arr.type)
            state.str_minmax[i].update(key_vals, v.arr, pred, st)
        state.batches += 1
        if keep:
            st.sync()   # the async copies read host arrays that die with this batch

    def _agg_finish(self, state: "_AggState") -> Tuple[Frame, Dict]:
        """AggregateOperator result (aggregate.py:122): finalise, read the groups back, give
        dictionary-coded / boolean keys their user types again."""
        if self.exchange is not None:
            if state.str_minmax:
                raise OperatorError("sharded execution: MIN / MAX over a string column is not implemented")
            if any(kind is not None for kind in state.key_kind):
                raise OperatorError("sharded execution: GROUP BY over string / boolean keys needs a shared "
                                    "dictionary across ranks, which is not implemented")
            local_path = state.agg.last_path
            state.agg, raw = self.exchange(state.agg, self.st)
            keys_out, aggs_out = state.agg.result_arrays(self.st, raw=raw)
            self.stats["agg_path"] = local_path
        else:
            keys_out, aggs_out = state.agg.result_arrays(self.st)
            self.stats["agg_path"] = state.agg.last_path
        self.stats["agg_batches"] = state.batches
        aggs_out = list(aggs_out)
        for i, smm in state.str_minmax.items():
            aggs_out[i] = smm.result(keys_out, len(aggs_out[i]))
        out = Frame(len(aggs_out[0]) if aggs_out else (len(keys_out[0]) if keys_out else 1))
        resolved: Dict = {}
        for i, (g, arr) in enumerate(zip(state.group_exprs, keys_out)):
            kind = state.key_kind[i]
            if kind == "dict":
                _table, values, t = state.dicts[i]
                dictionary = pa.array(values, type=t)
                arr = dictionary.take(arr) if len(dictionary) else pa.nulls(len(arr), t)
            elif kind == "bool":
                arr = arr.cast(pa.bool_())
            name = f"__key{i}"
            out.cols[name] = self._from_host(arr)
            resolved[g.key()] = name
            if isinstance(g, Column):
                out.cols.setdefault(g.name, out.cols[name])
        for i, (call, arr) in enumerate(zip(state.agg_calls, aggs_out)):
            name = f"__agg{i}"
            out.cols[name] = self._from_host(arr)
            resolved[call.key()] = name
        return out, resolved

    # --------------------------------------------------------------- streams
    def execute_stream(self, q: Query, batches, schema: pa.Schema) -> pa.Table:
        """The query over a STREAM of record batches that need not fit in memory together
        (StreamReader.sql, vinum/api/stream_reader.py:25-94; FileReaderOperator,
        algebra.py:268-279).  An aggregate query folds every batch into the device aggregate
        and only the groups survive; any other query keeps just the rows that pass its WHERE
        (compacted on the device, batch by batch) and finishes on those."""
        self.table = schema.empty_table()
        q = self._bind(q)
        used = _used_columns(q)
        if q.is_aggregate:
            state = _AggState(q)
            try:
                for tbl in batches:
                    self.table = tbl
                    self._agg_update(q, self._scan(used), state)
                if state.agg is None:
                    self.table = schema.empty_table()
                    self._agg_update(q, self._scan(used), state)
                frame, resolved = self._agg_finish(state)
            finally:
                state.close()
            return self._tail(q, frame, resolved)
        # non-aggregate: WHERE per batch, the survivors are concatenated on the host
        kept = []
        rows = 0
        need = None if (q.order_by or q.limit is None) else q.offset + q.limit
        for tbl in batches:
            self.table = tbl
            frame = self._scan(used)
            if q.where is not None:
                frame = self._filter(frame, self._predicate(q.where, frame, {}))
            kept.append(pa.table({name: (v.to_arrow(self.st) if isinstance(v, DeviceColumn) else v.arr)
                                  for name, v in frame.cols.items()}) if frame.cols else pa.table({"__n": pa.nulls(frame.n)}))
            rows += frame.n
            if need is not None and rows >= need:
                break
        self.table = pa.concat_tables(kept) if kept else pa.table({n: pa.array([], type=schema.field(n).type) for n in used})
        frame = self._scan(used)
        if not used:
            frame.n = rows
        return self._tail(q, frame, {})

    def _tail(self, q: Query, frame: Frame, resolved: Dict) -> pa.Table:
        if q.having is not None:
            frame = self._filter(frame, self._as_mask(self._eval(q.having, frame, resolved), frame.n))
        if q.order_by:
            top = None if q.limit is None else q.offset + q.limit
            frame = self._sort(frame, [self._eval(o, frame, resolved) for o in q.order_by], q.sort_order, top)
        values = [self._eval(e, frame, resolved) for e in q.select]
        n = frame.n
        if frame.virtual:
            lengths = [v.length for v in values if _is_col(v)]
            n = max(lengths) if lengths else 1
            if any(length != n for length in lengths):
                raise OperatorError("SELECT expressions produce arrays of different sizes")   # algebra.py:89-105
        out = self._materialize(values, output_names(q.select), n, q.limit, q.offset)
        self.stats["kernels"] = int(L.lib.vk_launch_count()) - self.stats["kernels_before"]
        return out

    # ------------------------------------------------------------ materialise
    def _materialize(self, values: List[Value], names: List[str], n: int, limit: Optional[int], offset: int) -> pa.Table:
        """ProjectOperator(col_names) + SliceOperator + MaterializeTableOperator
        (planner.py:484-505, algebra.py:204-247,290-295)."""
        lo, hi = 0, n
        if limit is not None:
            lo = min(offset, n)
            hi = min(lo + limit, n)
        arrays = []
        for v in values:
            if isinstance(v, DeviceColumn):
                arrays.append(v.slice(lo, hi - lo).to_arrow(self.st))
            elif isinstance(v, HostColumn):
                arrays.append(v.arr.slice(lo, hi - lo))
            else:
                arrays.append(pa.array(np.repeat(np.array([v]), hi - lo)) if v is not None
                              else pa.nulls(hi - lo))   # algebra.py:76-87
        return pa.Table.from_arrays(arrays, names=names)


class _AggState:
    """Streaming state of one aggregate operator: the device aggregate, the aggregate calls
    and group expressions of the query, and one host dictionary per string key that lives as
    long as the stream (so codes mean the same thing in every batch)."""

    def __init__(self, q: Query):
        self.agg: Optional[Aggregator] = None
        self.batches = 0
        self.str_minmax: Dict[int, StringMinMax] = {}    # aggregate call index -> MIN / MAX over a string column
        self.key_kind: List[Optional[str]] = []
        self.dicts: List = []
        self.group_exprs: List[Node] = []
        seen = set()
        for g in list(q.group_by) + (list(q.select) if q.distinct else []):
            if isinstance(g, Literal):
                continue
            if g.key() not in seen:
                seen.add(g.key())
                self.group_exprs.append(g)
        self.agg_calls: List[Expression] = []
        seen_a = set()
        for root in list(q.select) + [q.having] + list(q.order_by):
            if root is None:
                continue
            for node in walk(root):
                if is_aggregate_call(node) and node.key() not in seen_a:
                    seen_a.add(node.key())
                    self.agg_calls.append(node)

    def close(self):
        if self.agg is not None:
            self.agg.close()
            self.agg = None


def _encode_with_dictionary(arr, table: dict, values: list) -> pa.Array:
    """String array -> int32 codes against a dictionary that persists across batches."""
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    enc = pc.dictionary_encode(arr)
    local = enc.dictionary.to_pylist()
    mapping = np.empty(max(len(local), 1), dtype=np.int32)
    for j, v in enumerate(local):
        code = table.get(v)
        if code is None:
            code = len(values)
            table[v] = code
            values.append(v)
        mapping[j] = code
    idx = enc.indices
    if idx.null_count:
        null_mask = idx.is_null().to_numpy(zero_copy_only=False)
        local_codes = idx.fill_null(0).to_numpy(zero_copy_only=False)
    else:
        null_mask = None
        local_codes = idx.to_numpy()
    codes = mapping[local_codes] if len(local) else np.zeros(len(arr), dtype=np.int32)
    return pa.array(codes, type=pa.int32(), mask=null_mask)


def _used_columns(q: Query) -> List[str]:
    used: List[str] = []
    for clause in (q.select, (q.where,), q.group_by, (q.having,), q.order_by):
        for root in clause:
            if root is not None:
                for node in walk(root):
                    if isinstance(node, Column) and node.name not in used:
                        used.append(node.name)
    return used


def _copy(node):
    if isinstance(node, Expression):
        return Expression(node.op, node.args, node.function_name, node.alias)
    if isinstance(node, Column):
        return Column(node.name, node.alias)
    return Literal(node.value, node.alias)


def output_names(select: Sequence[Node]) -> List[str]:
    """QueryPlanner._column_names (planner.py:290-323): alias / column name / function name,
    else col_N; repeated names get _1, _2, ... suffixes."""
    out, index, unnamed = [], {}, 0
    for e in select:
        name = e.output_name()
        if not name:
            name = f"col_{unnamed}"
            unnamed += 1
        if name in index:
            index[name] += 1
            name = f"{name}_{index[name]}"
        else:
            index[name] = 0
        out.append(name)
    return out


CHUNK_ROWS = 1 << 26   # tables above this are executed as a stream of zero-copy slices


def execute_sql(sql: str, table: pa.Table, stream: Optional[Stream] = None, stats: Optional[dict] = None,
                chunk_rows: int = CHUNK_ROWS, exchange=None) -> pa.Table:
    """Parse + plan + run one SELECT over a host pyarrow.Table; the result is a host table."""
    q = parse_sql(sql, table.schema.names)
    eng = Engine(table, stream, exchange)
    if table.num_rows > chunk_rows and exchange is None:
        bound = eng._bind(q)
        if not (bound.is_aggregate and eng._streamable(bound)):
            # bound the device footprint: the table goes through in row slices
            slices = (table.slice(o, chunk_rows) for o in range(0, table.num_rows, chunk_rows))
            out = eng.execute_stream(q, slices, table.schema)
            if stats is not None:
                stats.update(eng.stats)
            return out
    out = eng.execute(q)
    if stats is not None:
        stats.update(eng.stats)
    return out


def execute_sql_stream(sql: str, reader, stream: Optional[Stream] = None, stats: Optional[dict] = None,
                       batch_rows: int = 1 << 22) -> pa.Table:
    """One SELECT over a `pyarrow.RecordBatchReader`-like stream (`.schema`, iteration yields
    RecordBatches): small reader batches are coalesced into device-sized ones."""
    schema = reader.schema
    q = parse_sql(sql, schema.names)
    eng = Engine(schema.empty_table(), stream)

    def coalesced():
        pending, rows = [], 0
        for b in reader:
            if b.num_rows == 0:
                continue
            pending.append(b)
            rows += b.num_rows
            if rows >= batch_rows:
                yield pa.Table.from_batches(pending, schema=schema).combine_chunks()
                pending, rows = [], 0
        if pending:
            yield pa.Table.from_batches(pending, schema=schema).combine_chunks()

    out = eng.execute_stream(q, coalesced(), schema)
    if stats is not None:
        stats.update(eng.stats)
    return out
