"""Query lowering: ``Task`` tree -> logical GPU plan -> expression programs for the scan kernel.

Input is the reference's plan language: a ``Task`` chain whose payloads are ``Col`` trees (built
either by this package's mirror classes or by the reference's own ``mini_spark`` classes -- nodes
are recognised by class name and attribute names, never by identity).  Output is

* a small logical plan (:class:`LTable`, :class:`LSelect`, :class:`LAggregate`, :class:`LJoin`)
  in which consecutive filters/projections are fused, AVG is expanded to SUM/COUNT
  (reference ``plan.py:190-203``, ``sql.py:436-446``) and expressions are typed, and
* per fused operator, a postfix *expression program* (:func:`compile_program`) for
  ``msc_scan_aggregate`` / ``msc_scan_project``.

It replaces the reference's physical planner + Zig code generation
(``src/mini_spark/plan.py:135-235``, ``codegen.py``, ``templates/plan.zig``): there is no compiler
at query time, the program is interpreted in registers by one pre-compiled kernel.

Nothing in this module touches the GPU, so it is unit-tested on CPU.
"""

from __future__ import annotations

import operator as _op
import os
import struct
from dataclasses import dataclass, field
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Iterable, Optional, Protocol, Sequence

from .constants import ColumnType, Schema
from .io import BlockFile, datetime_to_timestamp
from .native import K, OP, P_F32, P_F64, P_I32, P_I64, P_U8, P_U16, P_U32, RV

# value types of the IR
INT, FLOAT, TS, STR, BOOL = "I", "F", "T", "S", "B"
_FROM_COLUMN_TYPE = {ColumnType.INTEGER: INT, ColumnType.FLOAT: FLOAT, ColumnType.TIMESTAMP: TS, ColumnType.STRING: STR}
_OPNAME = {
    _op.add: "add", _op.sub: "sub", _op.mul: "mul", _op.truediv: "truediv", _op.floordiv: "floordiv",
    _op.mod: "mod", _op.eq: "eq", _op.ne: "ne", _op.lt: "lt", _op.le: "le", _op.gt: "gt", _op.ge: "ge",
    _op.and_: "and", _op.or_: "or",
}
_ARITH = {"add", "sub", "mul", "truediv", "floordiv", "mod"}
_CMP = {"eq", "ne", "lt", "le", "gt", "ge"}


class LoweringError(Exception):
    """The query uses something the GPU engine does not implement."""


# ------------------------------------------------------------------------------------------------
# typed expression IR (hashable -> common sub-expression detection by equality)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Expr:
    type: str


@dataclass(frozen=True)
class EInput(Expr):
    """Column ``index`` of the operator's input relation."""
    index: int = 0


@dataclass(frozen=True)
class EConst(Expr):
    value: Any = None  # int | float | str (STR) | int microseconds (TS)


@dataclass(frozen=True)
class EBin(Expr):
    op: str = ""
    left: Expr = None  # type: ignore[assignment]
    right: Expr = None  # type: ignore[assignment]


@dataclass(frozen=True)
class ECast(Expr):
    """INT -> FLOAT coercion (reference sql.py:277-290)."""
    child: Expr = None  # type: ignore[assignment]


@dataclass(frozen=True)
class ELike(Expr):
    child: Expr = None  # type: ignore[assignment]
    pattern: str = ""


@dataclass(frozen=True)
class EConcat(Expr):
    """STRING + STRING, flattened: parts are STR inputs or STR constants."""
    parts: tuple = ()


@dataclass(frozen=True)
class ECode(Expr):
    """The integer dictionary code of a STR expression, typed INT (hash-join keys are 64-bit ints)."""
    child: Expr = None  # type: ignore[assignment]


@dataclass(frozen=True)
class ETranslate(Expr):
    """STR expression re-coded into another column's dictionary (join keys); token picks the target."""
    child: Expr = None  # type: ignore[assignment]
    token: str = ""


def expr_inputs(e: Expr) -> set[int]:
    if isinstance(e, EInput):
        return {e.index}
    out: set[int] = set()
    for child in expr_children(e):
        out |= expr_inputs(child)
    return out


def expr_children(e: Expr) -> tuple[Expr, ...]:
    if isinstance(e, EBin):
        return (e.left, e.right)
    if isinstance(e, (ECast, ELike, ETranslate, ECode)):
        return (e.child,)
    if isinstance(e, EConcat):
        return tuple(e.parts)
    return ()


def substitute(e: Expr, inputs: Sequence[Expr]) -> Expr:
    """Replace every ``EInput(i)`` by ``inputs[i]`` (fuses a projection into its consumer)."""
    if isinstance(e, EInput):
        return inputs[e.index]
    if isinstance(e, EBin):
        return EBin(e.type, e.op, substitute(e.left, inputs), substitute(e.right, inputs))
    if isinstance(e, ECast):
        return ECast(e.type, substitute(e.child, inputs))
    if isinstance(e, ELike):
        return ELike(e.type, substitute(e.child, inputs), e.pattern)
    if isinstance(e, ETranslate):
        return ETranslate(e.type, substitute(e.child, inputs), e.token)
    if isinstance(e, ECode):
        return ECode(e.type, substitute(e.child, inputs))
    if isinstance(e, EConcat):
        return _concat(*[substitute(p, inputs) for p in e.parts])
    return e


def remap(e: Expr, mapping: dict[int, int]) -> Expr:
    return substitute(e, _Remap(mapping, e))


class _Remap:
    def __init__(self, mapping: dict[int, int], root: Expr) -> None:
        self.mapping = mapping
        self.types = _input_types(root)

    def __getitem__(self, i: int) -> Expr:
        return EInput(self.types[i], self.mapping[i])


def _input_types(e: Expr) -> dict[int, str]:
    if isinstance(e, EInput):
        return {e.index: e.type}
    out: dict[int, str] = {}
    for c in expr_children(e):
        out.update(_input_types(c))
    return out


def _concat(*parts: Expr) -> Expr:
    flat: list[Expr] = []
    for p in parts:
        flat.extend(p.parts if isinstance(p, EConcat) else [p])
    return EConcat(STR, tuple(flat))


def show(e: Expr) -> str:
    if isinstance(e, EInput):
        return f"${e.index}"
    if isinstance(e, EConst):
        return repr(e.value)
    if isinstance(e, EBin):
        return f"({show(e.left)} {e.op} {show(e.right)})"
    if isinstance(e, ECast):
        return f"float({show(e.child)})"
    if isinstance(e, ELike):
        return f"like({show(e.child)}, {e.pattern!r})"
    if isinstance(e, ETranslate):
        return f"recode({show(e.child)})"
    if isinstance(e, ECode):
        return f"code({show(e.child)})"
    if isinstance(e, EConcat):
        return "concat(" + ", ".join(show(p) for p in e.parts) + ")"
    return "?"


# ------------------------------------------------------------------------------------------------
# Col tree -> Expr   (duck-typed on the reference's class / attribute names, sql.py:16-468)
# ------------------------------------------------------------------------------------------------
def _kind(obj: Any) -> str:
    return type(obj).__name__


def lower_col(col: Any, schema: Schema) -> Expr:
    """Type and lower one ``Col`` tree against the schema of the operator input."""
    kind = _kind(col)
    if kind == "Lit":
        return _lower_literal(col.value)
    if kind in ("AliasColumn", "AggCol"):
        return lower_col(col.original_col, schema)
    if kind == "LikeColumn":
        child = lower_col(col.original_col, schema)
        if child.type != STR:
            raise AssertionError("LIKE operator can only be applied to string columns")
        return ELike(BOOL, child, col.pattern)
    if kind == "BinaryOperatorColumn":
        return _lower_binary(col, schema)
    if kind in ("Col", "SchemaCol"):
        for i, (name, ctype) in enumerate(schema):
            if name == col.name:
                return EInput(_FROM_COLUMN_TYPE[ctype], i)
        raise ValueError(f'Column "{col.name}" not found in schema {schema}')
    raise LoweringError(f"unsupported column expression {kind}")


def _lower_literal(value: Any) -> Expr:
    if type(value) is bool:
        return EConst(BOOL, int(value))
    if type(value) is int:
        return EConst(INT, value)
    if type(value) is float:
        return EConst(FLOAT, value)
    if type(value) is str:
        return EConst(STR, value)
    if type(value) is datetime:
        return EConst(TS, datetime_to_timestamp(value))
    raise LoweringError(f"unsupported literal {value!r}")


def _lower_binary(col: Any, schema: Schema) -> Expr:
    op = _OPNAME.get(col.operator)
    if op is None:
        raise LoweringError(f"unsupported operator {col.operator}")
    left = lower_col(col.left_side, schema)
    right = lower_col(col.right_side, schema)
    if op in ("and", "or"):
        for side in (left, right):
            if side.type not in (BOOL, INT):
                raise TypeError(f"Type mismatch in binary operation: {op} needs boolean operands")
        return EBin(BOOL, op, left, right)
    # an ISO string literal compared with a TIMESTAMP is parsed (reference sql.py:291-298)
    if left.type == STR and right.type == TS and isinstance(left, EConst):
        left = EConst(TS, datetime_to_timestamp(datetime.fromisoformat(left.value)))
    if right.type == STR and left.type == TS and isinstance(right, EConst):
        right = EConst(TS, datetime_to_timestamp(datetime.fromisoformat(right.value)))
    lt, rt = left.type, right.type
    lt = INT if lt == BOOL else lt
    rt = INT if rt == BOOL else rt
    if op == "truediv":
        if lt not in (INT, FLOAT) or rt not in (INT, FLOAT):
            raise TypeError(f"Type mismatch in binary operation: {lt} / {rt}")
        left = ECast(FLOAT, left) if lt == INT else left
        right = ECast(FLOAT, right) if rt == INT else right
        return EBin(FLOAT, op, left, right)
    if {lt, rt} == {INT, FLOAT}:
        left = ECast(FLOAT, left) if lt == INT else left
        right = ECast(FLOAT, right) if rt == INT else right
        lt = rt = FLOAT
    if lt != rt:
        raise TypeError(f"Type mismatch in binary operation: {lt} {op} {rt}")
    if op in _CMP:
        return EBin(BOOL, op, left, right)  # (STRING <, <=, >, >= compare like Python strings: through the dictionaries' order)
    if lt == STR:
        if op != "add":
            raise TypeError(f"unsupported operand type(s) for {op}: 'str' and 'str'")
        return _concat(left, right)
    if lt == TS:
        raise LoweringError("arithmetic on TIMESTAMP is not implemented on the GPU engine")
    return EBin(lt, op, left, right)


# ------------------------------------------------------------------------------------------------
# logical plan
# ------------------------------------------------------------------------------------------------
@dataclass
class LNode:
    schema: Schema

    def describe(self, indent: int = 0) -> str:
        raise NotImplementedError


@dataclass
class LTable(LNode):
    path: Path = Path()
    alias: str = ""

    def describe(self, indent: int = 0) -> str:
        return " " * indent + f"Table({self.path}) {[n for n, _ in self.schema]}"


@dataclass
class LSelect(LNode):
    """Fused filter + projection: rows of ``child`` passing all ``filters``, columns = ``outputs``."""
    child: LNode = None  # type: ignore[assignment]
    filters: list[Expr] = field(default_factory=list)
    outputs: list[Expr] = field(default_factory=list)

    def describe(self, indent: int = 0) -> str:
        pad = " " * indent
        lines = [pad + "Select " + ", ".join(f"{n}={show(e)}" for (n, _), e in zip(self.schema, self.outputs))]
        if self.filters:
            lines.append(pad + "  where " + " and ".join(show(f) for f in self.filters))
        lines.append(self.child.describe(indent + 2))
        return "\n".join(lines)


@dataclass
class LAggregate(LNode):
    """GROUP BY ``group`` with aggregates ``aggs`` = [(kind, expr)], kind in sum/min/max/count."""
    child: LNode = None  # type: ignore[assignment]
    group: Expr = None  # type: ignore[assignment]
    aggs: list[tuple[str, Expr]] = field(default_factory=list)
    groups_seen: int = field(default=0, compare=False)  # groups this node produced the last time it ran (a cached plan sizes its next hash table with it)

    def describe(self, indent: int = 0) -> str:
        pad = " " * indent
        aggs = ", ".join(f"{k}({show(e)})" for k, e in self.aggs)
        return pad + f"Aggregate by {show(self.group)}: {aggs}\n" + self.child.describe(indent + 2)


@dataclass
class LJoin(LNode):
    left: LNode = None  # type: ignore[assignment]
    right: LNode = None  # type: ignore[assignment]
    left_key: Expr = None  # type: ignore[assignment]
    right_key: Expr = None  # type: ignore[assignment]

    def describe(self, indent: int = 0) -> str:
        pad = " " * indent
        return (pad + f"HashJoin {show(self.left_key)} = {show(self.right_key)}\n" + self.left.describe(indent + 2) + "\n"
                + self.right.describe(indent + 2))


def identity_select(node: LNode) -> LSelect:
    return LSelect(list(node.schema), node, [], [EInput(_FROM_COLUMN_TYPE[t], i) for i, (_, t) in enumerate(node.schema)])


_MAY_RAISE = {"truediv", "floordiv", "mod"}  # ZeroDivisionError in the reference (sql.py:262-266)


def _cannot_raise(e: Expr) -> bool:
    if isinstance(e, EBin) and e.op in _MAY_RAISE:
        return False
    return all(_cannot_raise(c) for c in expr_children(e))


def push_filters_below_join(filters: Sequence[Expr], join: "LJoin") -> tuple[list[Expr], "LJoin"]:
    """Predicate pushdown through an (inner) join: a conjunct that reads columns of one side only filters that side
    BEFORE the join.  The reference filters after it (`FilterTask` above `BroadcastHashJoinTask`), which gives the same
    rows for an inner join; here it decides whether sf10 builds on 15 M orders and probes 60 M lineitems or on the 2.3 M
    and 17 M that pass `o_orderdate BETWEEN ..` / `l_shipmode LIKE ..`.  Conjuncts that could raise (division, modulo)
    stay above the join, so a row the join would have dropped can never report a division by zero."""
    nl = len(join.left.schema)
    left_f: list[Expr] = []
    right_f: list[Expr] = []
    keep: list[Expr] = []
    for conjunct in split_conjunctions(filters):
        ins = expr_inputs(conjunct)
        if ins and _cannot_raise(conjunct) and all(i < nl for i in ins):
            left_f.append(conjunct)
        elif ins and _cannot_raise(conjunct) and all(i >= nl for i in ins):
            right_f.append(remap(conjunct, {i: i - nl for i in ins}))
        else:
            keep.append(conjunct)
    if not left_f and not right_f:
        return list(filters), join

    def filtered(side: LNode, extra: list[Expr]) -> LNode:
        if not extra:
            return side
        sel = identity_select(side)
        sel.filters = extra
        return fuse_selects(sel)

    return keep, LJoin(join.schema, filtered(join.left, left_f), filtered(join.right, right_f), join.left_key, join.right_key)


def fuse_selects(node: LNode) -> LNode:
    """Merge stacked LSelects bottom-up by inlining the inner projection into the outer one; filters above a join move
    below it where they can (push_filters_below_join)."""
    if isinstance(node, LSelect):
        child = fuse_selects(node.child)
        if isinstance(child, LSelect):
            filters = list(child.filters) + [substitute(f, child.outputs) for f in node.filters]
            outputs = [substitute(o, child.outputs) for o in node.outputs]
            node = LSelect(node.schema, child.child, filters, outputs)
            child = node.child
        if isinstance(child, LJoin) and node.filters and os.environ.get("MINISPARK_JOIN_PUSHDOWN", "1") != "0":
            filters, child = push_filters_below_join(node.filters, child)
            return LSelect(node.schema, child, filters, node.outputs)
        return LSelect(node.schema, child, node.filters, node.outputs)
    if isinstance(node, LAggregate):
        return LAggregate(node.schema, fuse_selects(node.child), node.group, node.aggs)
    if isinstance(node, LJoin):
        return LJoin(node.schema, fuse_selects(node.left), fuse_selects(node.right), node.left_key, node.right_key)
    return node


def lower_task(task: Any) -> LNode:
    """Lower a *validated* task tree (``task.validate_schema()`` has run) to a fused logical plan.

    Mirrors what ``PhysicalPlan.generate_physical_plan`` does for the reference engines
    (plan.py:224-235): AVG expansion and the trailing projection (plan.py:190-203), output-name
    cleanup of ``alias.`` prefixes (plan.py:206-222).
    """
    node = _lower(task)
    names = [n for n, _ in node.schema]
    if any("." in n for n in names):
        sel = identity_select(node)
        sel.schema = [(n.split(".")[-1], t) for n, t in node.schema]
        node = sel
    return fuse_selects(node)


def _table_schema(task: Any) -> Schema:
    file_schema = BlockFile(task.file_path).file_schema
    if getattr(task, "alias", ""):
        return [(f"{task.alias}.{n}", t) for n, t in file_schema]
    return list(file_schema)


def _lower(task: Any) -> LNode:
    kind = _kind(task)
    if kind == "LoadTableBlockTask":
        return LTable(_table_schema(task), Path(task.file_path), getattr(task, "alias", "") or "")
    if kind == "ProjectTask":
        child = _lower(task.parent_task)
        cols: list[Any] = []
        for col in task.columns:  # '*' (re-)expansion, tasks.py:88-93
            if _kind(col) == "Col" and col.name == "*":
                cols.extend(_NameRef(n) for n, _ in child.schema)
            else:
                cols.append(col)
        outputs = [lower_col(c, child.schema) for c in cols]
        schema = [(c.name, _column_type(e, c, child.schema)) for c, e in zip(cols, outputs)]
        return LSelect(schema, child, [], outputs)
    if kind == "FilterTask":
        child = _lower(task.parent_task)
        cond = lower_col(task.condition, child.schema)
        if cond.type not in (BOOL, INT):
            raise TypeError("filter condition must be boolean")
        sel = identity_select(child)
        sel.filters = [cond]
        return sel
    if kind == "AggregateTask":
        return _lower_aggregate(task)
    if kind == "BroadcastHashJoinTask":
        left = _lower(task.parent_task)
        right = _lower(task.right_side_task)
        cond = task.join_condition
        if _kind(cond) != "BinaryOperatorColumn":
            raise AssertionError("Only equi-join is supported")
        left_col, right_col = cond.extract_left_right_key(left.schema, right.schema)
        lkey = lower_col(left_col, left.schema)
        rkey = lower_col(right_col, right.schema)
        if lkey.type != rkey.type:
            raise TypeError(f"Type mismatch in join keys: {lkey.type} vs {rkey.type}")
        return LJoin(list(left.schema) + list(right.schema), left, right, lkey, rkey)
    if kind in ("WriteToLocalFileTask", "WriteToShufflePartitions", "LoadShuffleFilesTask"):
        return _lower(task.parent_task)  # planner artefacts of the reference engines: transparent here
    raise LoweringError(f"unsupported task {kind}")


class _NameRef:
    """Stand-in for ``Col(name)`` produced by '*' expansion."""

    def __init__(self, name: str) -> None:
        self.name = name


_NameRef.__name__ = "Col"


def _column_type(e: Expr, col: Any, schema: Schema) -> ColumnType:
    # the reference types comparisons after their left operand (sql.py:303); mirror it when available
    infer = getattr(col, "infer_type", None)
    if infer is not None:
        try:
            return infer(schema)
        except Exception:  # noqa: BLE001 - fall through to the IR type
            pass
    return {INT: ColumnType.INTEGER, BOOL: ColumnType.INTEGER, FLOAT: ColumnType.FLOAT, TS: ColumnType.TIMESTAMP,
            STR: ColumnType.STRING}[e.type]


def _lower_aggregate(task: Any) -> LNode:
    child = _lower(task.parent_task)
    group = lower_col(task.group_by_column, child.schema)
    group_type = _column_type(group, task.group_by_column, child.schema)
    aggs: list[tuple[str, Expr]] = []
    agg_schema: Schema = [(task.group_by_column.name, group_type)]
    post: list[tuple[str, ColumnType, Expr]] = [(task.group_by_column.name, group_type, EInput(group.type, 0))]

    def add(kind: str, e: Expr, name: str) -> EInput:
        aggs.append((kind, e))
        ctype = ColumnType.FLOAT if e.type == FLOAT else ColumnType.INTEGER
        agg_schema.append((name, ctype))
        return EInput(FLOAT if e.type == FLOAT else INT, len(aggs))

    for agg in task.agg_columns:
        e = lower_col(agg.original_col, child.schema)
        if e.type == BOOL:
            e = EBin(INT, "add", e, EConst(INT, 0))
        if e.type not in (INT, FLOAT):
            raise AssertionError(f"aggregate over non-numeric column {agg.name}")
        if agg.type == "avg":  # AVG = SUM / COUNT, projected after the aggregate (sql.py:436-446)
            s = add("sum", e, f"{agg.name}_sum")
            c = add("count", EConst(INT, 1), f"{agg.name}_count")
            num = s if s.type == FLOAT else ECast(FLOAT, s)
            post.append((agg.name, ColumnType.FLOAT, EBin(FLOAT, "truediv", num, ECast(FLOAT, c))))
        elif agg.type in ("sum", "min", "max"):
            is_count = agg.type == "sum" and isinstance(e, EConst) and e.type == INT and e.value == 1
            ref = add("count" if is_count else agg.type, e, agg.name)
            post.append((agg.name, ColumnType.FLOAT if e.type == FLOAT else ColumnType.INTEGER, ref))
        else:
            raise LoweringError(f"unsupported aggregate {agg.type}")
    node = LAggregate(agg_schema, child, group, aggs)
    return LSelect([(n, t) for n, t, _ in post], node, [], [e for _, _, e in post])


# ------------------------------------------------------------------------------------------------
# expression program compiler
# ------------------------------------------------------------------------------------------------
@dataclass
class Binding:
    """How the scan reads one input column."""
    phys: int                    # MSC_P_*
    staged: Optional[int] = None  # slot in desc.staged (direct read)
    gather: Optional[int] = None  # slot in desc.gather ...
    index: Optional[int] = None   # ... through this staged index-vector slot
    dict_id: Any = None           # dictionary handle for STR columns
    probe: bool = False           # gather column of the build side of a fused join: read through the program's PROBE result


@dataclass
class ProbeSpec:
    """The probe half of a hash join carried by a scan's row program (MSC_OP_PROBE): after the scan's own `pre_filters`,
    every row looks its key up in the join table (LUT slot `lut`); rows without a match are filtered out, and the
    build-side columns (bindings with probe=True) are read through the matched build row."""
    key: "Expr"
    lut: int
    compact: bool = False  # the table has 8-byte slots (32-bit keys): lets a specialised kernel compile that format only


class Resolver(Protocol):
    """Run-time services the compiler needs for STRING operands (implemented by the engine)."""

    def binding(self, index: int) -> Binding: ...
    def literal_code(self, dict_id: Any, text: str) -> int: ...           # -1 when absent
    def like_lut(self, dict_id: Any, pattern: str) -> int: ...            # -> LUT slot
    def translate_lut(self, dict_id: Any, token: str) -> tuple[int, Any]: ...  # -> (LUT slot, target dict)
    def same_dict(self, a: Any, b: Any) -> bool: ...
    def recode_lut(self, src: Any, dst: Any) -> int: ...                  # LUT slot: codes of src -> codes of dst
    def order_lut(self, dict_id: Any, op: str, text: str) -> int: ...     # LUT slot (u8): entry <op> text, op in lt/le/gt/ge
    def rank_luts(self, a: Any, b: Any) -> tuple[int, int]: ...           # LUT slots (u32): ranks of both dictionaries' entries in their common order


MAX_TEMPS = K["MSC_VM_MAX_TEMPS"]
FAST_SHAPES = os.environ.get("MSC_SCAN_FAST", "1") != "0"  # 0: force the generic interpreter path (A/B testing)
_F_OPS = {"add": "ADD_F", "sub": "SUB_F", "mul": "MUL_F", "truediv": "DIV_F", "floordiv": "FLOORDIV_F", "mod": "MOD_F",
          "lt": "LT_F", "le": "LE_F", "gt": "GT_F", "ge": "GE_F", "eq": "EQ_F", "ne": "NE_F"}
_I_OPS = {"add": "ADD_I", "sub": "SUB_I", "mul": "MUL_I", "floordiv": "FLOORDIV_I", "mod": "MOD_I",
          "lt": "LT_I", "le": "LE_I", "gt": "GT_I", "ge": "GE_I", "eq": "EQ_I", "ne": "NE_I", "and": "AND", "or": "OR"}
SRC_TEMP, SRC_STAGED, SRC_CONST, SRC_GATHER, SRC_LUT = (K[f"MSC_SRC_{n}"] for n in ("TEMP", "STAGED", "CONST", "GATHER", "LUT"))
SRC_GATHER_T = K["MSC_SRC_GATHER_T"]
SRC_I2F = K["MSC_SRC_I2F"]
DST_TEMP, DST_FILTER, DST_GROUP, DST_AGG, DST_OUT, DST_NONE = (K[f"MSC_DST_{n}"] for n in ("TEMP", "FILTER", "GROUP", "AGG", "OUT", "NONE"))
_DST_NAME = {DST_TEMP: "t", DST_FILTER: "filter", DST_GROUP: "group", DST_AGG: "agg", DST_OUT: "out", DST_NONE: "none"}
_SRC_NAME = {SRC_TEMP: "t", SRC_STAGED: "col", SRC_CONST: "const", SRC_GATHER: "gather", SRC_LUT: "lut", SRC_GATHER_T: "gather", 0: "-"}


def f64_bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", float(x)))[0]


def src(kind: int, index: int, i2f: bool = False) -> int:
    if index < 0 or index > 0xFFF:
        raise LoweringError("operand index out of range")
    return index | (kind << 12) | ((SRC_I2F << 12) if i2f else 0)


@dataclass
class Program:
    code: list[int] = field(default_factory=list)    # u32 words, two per instruction
    consts: list[int] = field(default_factory=list)
    ntemps: int = 0
    text: list[str] = field(default_factory=list)    # disassembly, for explain/tests
    regvm: list[int] = field(default_factory=list)   # same query for the register-resident interpreter (may be empty)
    regvm_count_slot: int = -1                       # accumulator slot of COUNT (the regvm kernel's group-presence counter)
    regvm_text: list[str] = field(default_factory=list)

    def words(self) -> list[int]:
        return list(self.code)


def _fmt_operand(o: int) -> str:
    kind, idx = (o >> 12) & 7, o & 0xFFF
    if kind == 0:
        return "-"
    if kind == SRC_GATHER_T:
        body = f"gather{idx & 63}[t{idx >> 6}]"
    else:
        body = f"{_SRC_NAME[kind]}{idx & 63}[ix{idx >> 6}]" if kind == SRC_GATHER else f"{_SRC_NAME[kind]}{idx}"
    return f"f64({body})" if o & 0x8000 else body


class ProgramBuilder:
    """Emits three-address instructions; intermediates live in numbered temporaries (shared memory
    slots of the scan kernel), repeated sub-expressions are computed once and kept in a temporary."""

    def __init__(self, resolver: Resolver) -> None:
        self.r = resolver
        self.p = Program()
        self.free: list[int] = list(range(MAX_TEMPS))
        self.cse: dict[Expr, int] = {}
        self.uses_left: dict[Expr, int] = {}
        self.candidates: dict[Expr, int] = {}
        self.staged_phys: dict[int, int] = {}
        self.probe_temp: Optional[int] = None  # temporary that holds the build row a PROBE matched (kept until END)

    # -- low level ------------------------------------------------------------------------------
    def alloc(self) -> int:
        if not self.free:
            raise LoweringError(f"expression needs more than {MAX_TEMPS} temporaries")
        t = self.free.pop(0)
        self.p.ntemps = max(self.p.ntemps, t + 1)
        return t

    def release(self, temps: Iterable[int]) -> None:
        for t in temps:
            if t not in self.free:
                self.free.append(t)
        self.free.sort()

    def const(self, bits: int) -> int:
        bits &= 0xFFFFFFFFFFFFFFFF
        if bits >= 1 << 63:
            bits -= 1 << 64
        if bits in self.p.consts:
            return self.p.consts.index(bits)
        if len(self.p.consts) >= K["MSC_VM_MAX_CONSTS"]:
            raise LoweringError("too many constants in one program")
        self.p.consts.append(bits)
        return len(self.p.consts) - 1

    def emit(self, opname: str, a: int = 0, b: int = 0, dkind: int = DST_NONE, didx: int = 0, tee: int = 0,
             agg_kind: Optional[int] = None, out_u32: bool = False) -> None:
        if len(self.p.code) + 2 >= K["MSC_VM_MAX_CODE"]:
            raise LoweringError("expression program too long")
        if didx < 0 or didx > 0x7F:
            raise LoweringError("destination index out of range")
        fast = self.fast_id(opname, a, b, dkind, tee, agg_kind, out_u32) if FAST_SHAPES else 0
        self.p.code.append(OP[opname] | (dkind << 6) | (tee << 9) | (didx << 13) | (fast << 20))
        self.p.code.append(a | (b << 16))
        dst = _DST_NAME[dkind] + (str(didx) if dkind in (DST_TEMP, DST_AGG, DST_OUT) else "")
        if tee:
            dst += f",t{tee - 1}"
        self.p.text.append(f"{dst} <- {opname}({_fmt_operand(a)}, {_fmt_operand(b)})" + (f"  [fast {fast}]" if fast else ""))

    # -- fast shapes (include/minispark_cuda.h MSC_FAST_*) ----------------------------------------
    def _fk(self, operand: int) -> Optional[int]:
        kind, idx, i2f = (operand >> 12) & 7, operand & 0xFFF, bool(operand & 0x8000)
        if kind == SRC_TEMP and not i2f:
            return K["MSC_FK_TEMP"]
        if kind == SRC_CONST and not i2f:
            return K["MSC_FK_CONST"]
        if kind != SRC_STAGED:
            return None
        phys = self.staged_phys.get(idx)
        if phys == P_I32:
            return K["MSC_FK_I32F"] if i2f else K["MSC_FK_I32"]
        if i2f:
            return None
        return {P_F32: K["MSC_FK_F32"], P_F64: K["MSC_FK_F64"], P_I64: K["MSC_FK_I64"], P_U8: K["MSC_FK_U8"],
                P_U16: K["MSC_FK_U16"], P_U32: K["MSC_FK_U32"]}.get(phys)

    def fast_id(self, opname: str, a: int, b: int, dkind: int, tee: int, agg_kind: Optional[int], out_u32: bool) -> int:
        """Shape id of a pre-compiled specialisation of this instruction, or 0 (generic path)."""
        fa, fb = self._fk(a), self._fk(b)
        float_kinds = (K["MSC_FK_F32"], K["MSC_FK_F64"], K["MSC_FK_I32F"])
        if opname in ("ADD_F", "SUB_F", "MUL_F") and fa is not None and fb is not None and fa <= 4 and fb <= 4:
            if dkind == DST_TEMP and not tee:
                dk = 0
            elif dkind == DST_AGG and agg_kind == K["MSC_AGG_SUM_F"]:
                dk = 2 if tee else 1
            else:
                return 0
            opi = ("ADD_F", "SUB_F", "MUL_F").index(opname)
            return K["MSC_FAST_ARITH"] + ((opi * 5 + fa) * 5 + fb) * 3 + dk
        if tee or fa is None:
            return 0
        if opname == "MOV" and dkind == DST_AGG and agg_kind is not None:
            return K["MSC_FAST_AGGMOV"] + agg_kind * 10 + fa
        if dkind == DST_FILTER and fb == K["MSC_FK_CONST"] and opname[:2] in ("LT", "LE", "GT", "GE", "EQ", "NE") and len(opname) == 4:
            is_f = opname.endswith("_F")
            if fa in (K["MSC_FK_TEMP"], K["MSC_FK_CONST"]) or (fa in float_kinds) != is_f:
                return 0
            return K["MSC_FAST_CMP"] + ("LT", "LE", "GT", "GE", "EQ", "NE").index(opname[:2]) * 10 + fa
        if opname == "MOV" and dkind == DST_GROUP and fa not in float_kinds and fa != K["MSC_FK_CONST"]:
            return K["MSC_FAST_GROUP"] + fa
        if opname == "MOV" and dkind == DST_OUT and fa != K["MSC_FK_CONST"]:
            code_kinds = (K["MSC_FK_U8"], K["MSC_FK_U16"], K["MSC_FK_U32"])
            # (the handlers the kernel compiles, scan_kernel.cuh out_valid: a code goes to a U32 column, numbers to 64-bit ones)
            if (out_u32 and fa not in (*code_kinds, K["MSC_FK_TEMP"])) or (not out_u32 and fa in code_kinds):
                return 0
            return K["MSC_FAST_OUT"] + fa * 2 + int(out_u32)
        return 0

    # -- CSE ------------------------------------------------------------------------------------
    def plan_cse(self, roots: Iterable[Expr]) -> None:
        counts: dict[Expr, int] = {}

        def walk(e: Expr) -> None:
            if isinstance(e, (EInput, EConst)) or (isinstance(e, ECast) and isinstance(e.child, (EInput, EConst))):
                return
            counts[e] = counts.get(e, 0) + 1
            if counts[e] == 1:
                for c in expr_children(e):
                    walk(c)

        for root in roots:
            walk(root)
        self.candidates = {e: n for e, n in counts.items() if n > 1}

    def _take(self, e: Expr) -> tuple[int, list[int]]:
        t = self.cse[e]
        self.uses_left[e] -= 1
        if self.uses_left[e] <= 0:
            del self.cse[e]
            return src(SRC_TEMP, t), [t]
        return src(SRC_TEMP, t), []

    # -- operands --------------------------------------------------------------------------------
    def _leaf(self, e: Expr) -> Optional[int]:
        if isinstance(e, EInput):
            b = self.r.binding(e.index)
            if b.staged is not None:
                self.staged_phys[b.staged] = b.phys
                return src(SRC_STAGED, b.staged)
            if b.probe:
                if self.probe_temp is None:
                    raise LoweringError("a build-side column is read before the join's probe")
                return src(SRC_GATHER_T, b.gather | (self.probe_temp << 6))
            return src(SRC_GATHER, b.gather | (b.index << 6))
        if isinstance(e, EConst):
            if e.type == STR:
                raise LoweringError("a string literal is only supported as an operand of =, != or +")
            return src(SRC_CONST, self.const(f64_bits(e.value) if e.type == FLOAT else int(e.value)))
        if isinstance(e, ECast):
            if isinstance(e.child, EConst):
                return src(SRC_CONST, self.const(f64_bits(float(e.child.value))))
            if isinstance(e.child, EInput):
                return self._leaf(e.child) | (SRC_I2F << 12)
        return None

    def operand(self, e: Expr) -> tuple[int, list[int]]:
        """Encoded operand holding ``e`` and the temporaries to release once it has been consumed."""
        leaf = self._leaf(e)
        if leaf is not None:
            return leaf, []
        if e in self.cse:
            return self._take(e)
        if isinstance(e, ECast):
            enc, frees = self.operand(e.child)
            return enc | (SRC_I2F << 12), frees
        t = self.alloc()
        self.materialize(e, DST_TEMP, t)
        if e in self.cse:  # materialize registered it as a shared value
            return self._take(e)
        return src(SRC_TEMP, t), [t]

    # -- strings ---------------------------------------------------------------------------------
    def string_operand(self, e: Expr) -> tuple[int, list[int], Any]:
        """(operand holding the dictionary code of a STR expression, temps to free, its dictionary)."""
        if isinstance(e, ECode):
            return self.string_operand(e.child)
        if isinstance(e, EInput):
            return self._leaf(e), [], self.r.binding(e.index).dict_id
        if isinstance(e, ETranslate):
            enc, frees, d = self.string_operand(e.child)
            slot, target = self.r.translate_lut(d, e.token)
            if slot < 0:
                return enc, frees, target
            t = self.alloc()
            self.emit("LUT32", enc, src(SRC_LUT, slot), DST_TEMP, t)
            self.release(frees)
            return src(SRC_TEMP, t), [t], target
        raise LoweringError(f"string expression {show(e)} must be materialised first")

    def _string_compare(self, e: EBin) -> tuple[str, int, int, list[int]]:
        left, right = e.left, e.right
        if e.op in ("lt", "le", "gt", "ge"):
            return self._string_order(e)
        if isinstance(left, EConst) and not isinstance(right, EConst):
            left, right = right, left
        opname = "EQ_I" if e.op == "eq" else "NE_I"
        if isinstance(left, EConst):  # literal vs literal folds to a constant
            truth = (left.value == right.value) == (e.op == "eq")
            return "MOV", src(SRC_CONST, self.const(int(truth))), 0, []
        a, frees, d = self.string_operand(left)
        if isinstance(right, EConst):
            return opname, a, src(SRC_CONST, self.const(self.r.literal_code(d, right.value))), frees  # -1 matches no code
        b, frees_b, d2 = self.string_operand(right)
        if not self.r.same_dict(d, d2):
            t = self.alloc()
            self.emit("LUT32", b, src(SRC_LUT, self.r.recode_lut(d2, d)), DST_TEMP, t)
            self.release(frees_b)
            b, frees_b = src(SRC_TEMP, t), [t]
        return opname, a, b, frees + frees_b

    def _string_order(self, e: EBin) -> tuple[str, int, int, list[int]]:
        """STRING <, <=, >, >= (Python compares code points, sql.py:262-266; the format is ASCII, io.py:101): against a
        literal a u8 lookup table over the column's dictionary entries, between two columns the entries' ranks in the
        common order of both dictionaries."""
        import operator as _op

        left, right, op = e.left, e.right, e.op
        if isinstance(left, EConst) and isinstance(right, EConst):
            return "MOV", src(SRC_CONST, self.const(int(getattr(_op, op)(left.value, right.value)))), 0, []
        if isinstance(left, EConst):  # literal <op> column  ==  column <flipped op> literal
            left, right, op = right, left, {"lt": "gt", "le": "ge", "gt": "lt", "ge": "le"}[op]
        a, frees, d = self.string_operand(left)
        if isinstance(right, EConst):
            return "LUT8", a, src(SRC_LUT, self.r.order_lut(d, op, right.value)), frees
        b, frees_b, d2 = self.string_operand(right)
        slot_a, slot_b = self.r.rank_luts(d, d2)
        ta, tb = self.alloc(), self.alloc()
        self.emit("LUT32", a, src(SRC_LUT, slot_a), DST_TEMP, ta)
        self.emit("LUT32", b, src(SRC_LUT, slot_b), DST_TEMP, tb)
        self.release(frees + frees_b)
        return _I_OPS[op], src(SRC_TEMP, ta), src(SRC_TEMP, tb), [ta, tb]

    # -- one expression root into a destination -------------------------------------------------
    def materialize(self, e: Expr, dkind: int, didx: int = 0, agg_kind: Optional[int] = None, out_u32: bool = False) -> Any:
        """Evaluate ``e`` into the destination; returns the dictionary for STR-valued roots."""
        if e in self.cse:
            enc, frees = self._take(e)
            self.emit("MOV", enc, 0, dkind, didx, agg_kind=agg_kind, out_u32=out_u32)
            self.release(frees)
            return None
        dict_id = None
        frees: list[int] = []
        if e.type == STR or isinstance(e, ECode):
            a, frees, dict_id = self.string_operand(e)
            opname, b = "MOV", 0
        elif isinstance(e, EBin) and (e.left.type == STR or e.right.type == STR):
            opname, a, b, frees = self._string_compare(e)
        elif isinstance(e, ELike):
            a, frees, d = self.string_operand(e.child)
            opname, b = "LUT8", src(SRC_LUT, self.r.like_lut(d, e.pattern))
        elif isinstance(e, EBin):
            a, fa = self.operand(e.left)
            b, fb = self.operand(e.right)
            frees = fa + fb
            is_float = (e.left.type == FLOAT) if e.op in _CMP else (e.type == FLOAT)
            table = _F_OPS if is_float else _I_OPS
            if e.op not in table:
                raise LoweringError(f"operator {e.op} not available for type {e.type}")
            opname = table[e.op]
        elif isinstance(e, EConcat):
            raise LoweringError("internal: concat must be materialised before compilation")
        else:  # leaf or cast
            a, frees = self.operand(e)
            opname, b = "MOV", 0
        tee = 0
        shared = self.candidates.get(e, 0)
        if shared > 1 and e not in self.cse:
            if dkind == DST_TEMP:
                keep = didx
            else:
                keep = self.alloc()
                tee = keep + 1
            self.cse[e] = keep
            self.uses_left[e] = shared  # every occurrence (including this one, if it is an operand) calls _take
            if dkind != DST_TEMP:
                self.uses_left[e] -= 1  # this occurrence went straight to its destination
                if self.uses_left[e] <= 0:
                    del self.cse[e]
                    self.release([keep])
                    tee = 0
        self.emit(opname, a, b, dkind, didx, tee, agg_kind=agg_kind, out_u32=out_u32)
        self.release(frees)
        return dict_id

    def emit_probe(self, spec: ProbeSpec) -> None:
        """t <- PROBE(key, table); FILTER <- t >= 0.  The temporary stays allocated: every build-side column reads through it."""
        if spec.key.type == STR or isinstance(spec.key, ECode):
            a, frees, _ = self.string_operand(spec.key)
        else:
            a, frees = self.operand(spec.key)
        t = self.alloc()
        self.emit("PROBE", a, src(SRC_LUT, spec.lut | (K["MSC_PROBE_COMPACT"] if spec.compact else 0)), DST_TEMP, t)
        self.release(frees)
        self.emit("GE_I", src(SRC_TEMP, t), src(SRC_CONST, self.const(0)), DST_FILTER)
        self.probe_temp = t

    def end(self) -> Program:
        self.p.code.extend([OP["END"], 0])
        self.p.text.append("END")
        return self.p


def split_conjunctions(filters: Sequence[Expr]) -> list[Expr]:
    """`a AND b` as a filter is the same as filtering by a, then by b (each becomes one FILTER instruction)."""
    out: list[Expr] = []

    def walk(e: Expr) -> None:
        if isinstance(e, EBin) and e.op == "and" and e.left.type == BOOL and e.right.type == BOOL:
            walk(e.left)
            walk(e.right)
        else:
            out.append(e)

    for f in filters:
        walk(f)
    return out


_AGG_KINDS = {("sum", FLOAT): K["MSC_AGG_SUM_F"], ("sum", INT): K["MSC_AGG_SUM_I"], ("min", FLOAT): K["MSC_AGG_MIN_F"],
              ("max", FLOAT): K["MSC_AGG_MAX_F"], ("min", INT): K["MSC_AGG_MIN_I"], ("max", INT): K["MSC_AGG_MAX_I"]}


@dataclass
class AggregateProgram:
    program: Program
    agg_kinds: list[int]       # MSC_AGG_* per unique accumulator slot
    slot_of: list[int]         # slot of each requested aggregate (duplicates share a slot)
    group_dict: Any            # dictionary of a STR group key (None otherwise)


def compile_aggregate(resolver: Resolver, filters: Sequence[Expr], group: Expr, aggs: Sequence[tuple[str, Expr]],
                      probe: Optional[ProbeSpec] = None, pre_filters: Sequence[Expr] = ()) -> AggregateProgram:
    """``probe``: the scan also carries the probe half of a join -- `pre_filters` (the probe side's own) run first, then the
    probe, then `filters` (which may read build-side columns)."""
    b = ProgramBuilder(resolver)
    norm = [(k, EConst(INT, 1) if k == "count" else (EBin(INT, "add", e, EConst(INT, 0)) if e.type == BOOL else e)) for k, e in aggs]
    # accumulator slots: plain SUM(column) aggregates first and next to each other, so that the register interpreter
    # can take runs of them in one instruction (AGGCOL<n>, gen_regvm.py); the slot order is internal (slot_of maps back)
    def plain_sum(key: tuple[str, Expr]) -> bool:
        kind, e = key
        return kind == "sum" and e.type == FLOAT and (isinstance(e, EInput) or (isinstance(e, ECast) and isinstance(e.child, EInput)))

    unique: dict[tuple[str, Expr], int] = {}
    for key in sorted(dict.fromkeys(norm), key=lambda k: 0 if plain_sum(k) else 1):
        if len(unique) >= K["MSC_VM_MAX_AGGS"]:
            raise LoweringError("too many aggregates in one GROUP BY")
        unique[key] = len(unique)
    slot_of = [unique[key] for key in norm]
    count_key = ("count", EConst(INT, 1))
    if REGVM_ENABLED and count_key not in unique and len(unique) < K["MSC_VM_MAX_AGGS"]:
        unique[count_key] = len(unique)  # the regvm kernel tells present groups by their row count
    filters = split_conjunctions(filters)
    pre_filters = split_conjunctions(pre_filters)
    b.plan_cse([*pre_filters, *([probe.key] if probe else []), *filters, group, *[e for (k, e) in unique if k != "count"]])
    for f in pre_filters:
        b.materialize(f, DST_FILTER)
    if probe is not None:
        b.emit_probe(probe)
    for f in filters:
        b.materialize(f, DST_FILTER)
    group_dict = b.materialize(group, DST_GROUP)
    for (kind, e) in unique:  # stage the plain-SUM columns back to back (a run needs adjacent stage slots)
        if plain_sum((kind, e)):
            leaf = e if isinstance(e, EInput) else e.child
            b.r.binding(leaf.index)
    kinds: list[int] = []
    for (kind, e), slot in unique.items():
        if kind == "count":
            kinds.append(K["MSC_AGG_SUM_I"])
            b.emit("MOV", src(SRC_CONST, b.const(1)), 0, DST_AGG, slot, agg_kind=K["MSC_AGG_SUM_I"])
            continue
        kinds.append(_AGG_KINDS[(kind, FLOAT if e.type == FLOAT else INT)])
        b.materialize(e, DST_AGG, slot, agg_kind=kinds[-1])
    program = b.end()
    if REGVM_ENABLED and probe is None:
        try:
            program.regvm, program.regvm_text = compile_regvm(b, filters, group, unique)
            program.regvm_count_slot = unique.get(count_key, -1)
        except _RegvmUnsupported:
            program.regvm, program.regvm_text = [], []
    return AggregateProgram(program, kinds, slot_of, group_dict)


# ------------------------------------------------------------------------------------------------
# second encoding for the register-resident interpreter (csrc/gen_regvm.py, scan_regvm.cu)
# ------------------------------------------------------------------------------------------------
REGVM_ENABLED = os.environ.get("MSC_SCAN_REGVM", "1") != "0"
RV_MAX_DEPTH = RV["MAX_DEPTH"]
RV_MAX_TEMPS = RV["MAX_TEMPS"]


class _RegvmUnsupported(Exception):
    """The query uses something outside the regvm op set; the C++ interpreter runs it instead."""


def _float_const(e: Expr) -> Optional[float]:
    if isinstance(e, EConst) and e.type == FLOAT:
        return float(e.value)
    if isinstance(e, ECast) and isinstance(e.child, EConst) and e.child.type in (INT, BOOL):
        return float(e.child.value)
    return None


def compile_regvm(b: "ProgramBuilder", filters: Sequence[Expr], group: Expr, unique: dict) -> tuple[list[int], list[str]]:
    """Postfix stack code (one u32 per instruction) for a dense aggregate scan, or raise _RegvmUnsupported.

    Column operands are emitted as staged slots (the library resolves them to shared-memory offsets and
    checks their physical types against the handler); constants share the main program's pool."""
    words: list[int] = []
    text: list[str] = []

    def emit(name: str, a1: int = 0, a2: int = 0) -> None:
        if name not in RV:
            raise _RegvmUnsupported(name)
        if not (0 <= a1 < 256 and 0 <= a2 < 256):
            raise _RegvmUnsupported("operand out of range")
        words.append(RV[name] | (a1 << 8) | (a2 << 16))
        text.append(f"{name} {a1} {a2}")

    def column(e: Expr) -> Binding:
        if not isinstance(e, EInput):
            raise _RegvmUnsupported("not a column")
        bd = b.r.binding(e.index)
        if bd.staged is None:
            raise _RegvmUnsupported("gathered column")
        return bd

    flip = {"lt": "gt", "le": "ge", "gt": "lt", "ge": "le", "eq": "eq", "ne": "ne"}
    for f in filters:
        if not (isinstance(f, EBin) and f.op in _CMP):
            raise _RegvmUnsupported("filter shape")
        left, right, op = f.left, f.right, f.op
        if isinstance(left, EConst) and not isinstance(right, EConst):
            left, right, op = right, left, flip[op]
        if not isinstance(right, EConst) or right.type == STR or left.type == STR:
            raise _RegvmUnsupported("filter operand")
        bd = column(left)
        ty = {P_I64: "I64", P_I32: "I32", P_F32: "F32", P_F64: "F64"}.get(bd.phys)
        if ty is None or (ty[0] == "F") != (left.type == FLOAT) or (left.type == FLOAT) != (right.type == FLOAT):
            raise _RegvmUnsupported("filter type")
        bits = f64_bits(right.value) if ty[0] == "F" else int(right.value)
        emit(f"CMPCOL_{op.upper()}_{ty}", bd.staged, b.const(bits))

    if not (isinstance(group, EInput) and group.type == STR):
        raise _RegvmUnsupported("group key")
    gb = column(group)
    emit({P_U8: "GROUP_U8", P_U16: "GROUP_U16", P_U32: "GROUP_U32"}.get(gb.phys, "?"), gb.staged)

    counts: dict[Expr, int] = {}

    def walk(e: Expr) -> None:
        if isinstance(e, (EInput, EConst)) or _float_const(e) is not None or (isinstance(e, ECast) and isinstance(e.child, EInput)):
            return
        counts[e] = counts.get(e, 0) + 1
        if counts[e] == 1:
            for c in expr_children(e):
                walk(c)

    for (kind, e) in unique:
        if kind != "count":
            walk(e)
    temps: dict[Expr, int] = {}

    def float_leaf(e: Expr) -> Optional[tuple[str, int]]:
        """(column kind, staged slot) of a FLOAT-valued column leaf: f32 / f64 column or an i32 column cast to float."""
        if isinstance(e, EInput) and e.type == FLOAT:
            bd = column(e)
            kind = {P_F32: "F32", P_F64: "F64"}.get(bd.phys)
        elif isinstance(e, ECast) and isinstance(e.child, EInput) and e.child.type in (INT, BOOL):
            bd = column(e.child)
            kind = {P_I32: "I32F"}.get(bd.phys)
        else:
            return None
        if kind is None:
            raise _RegvmUnsupported("column type")
        return kind, bd.staged

    def load(e: Expr, d: int) -> bool:
        """Push a column leaf at depth d; False when e is not a loadable leaf."""
        leaf = float_leaf(e)
        if leaf is not None:
            emit(f"LD_{leaf[0]}_D{d}", leaf[1])
            return True
        if isinstance(e, EInput) and e.type in (INT, TS):
            bd = column(e)
            name = {P_I32: "LD_I32", P_I64: "LD_I64"}.get(bd.phys)
            if name is None:
                raise _RegvmUnsupported("column type")
            emit(f"{name}_D{d}", bd.staged)
            return True
        return False

    def gen(e: Expr, d: int, tee: bool = True) -> None:
        """Evaluate e into stack slot d (depth d -> d + 1), preferring the fused instruction forms."""
        if d >= RV_MAX_DEPTH:
            raise _RegvmUnsupported("expression too deep")
        if e in temps:
            emit(f"GET{temps[e]}_D{d}")
            return
        if load(e, d):
            return
        c = _float_const(e)
        if c is not None:
            emit(f"CONST_D{d}", b.const(f64_bits(c)))
            return
        if not (isinstance(e, EBin) and e.type == FLOAT and e.op in ("add", "sub", "mul")):
            raise _RegvmUnsupported(f"operator {getattr(e, 'op', type(e).__name__)}")
        left, right, op = e.left, e.right, e.op
        cl, cr = _float_const(left), _float_const(right)
        if (cl is None) != (cr is None):  # one side is a constant
            other, cval = (left, cr) if cr is not None else (right, cl)
            if cr is not None:   # x op c
                name, cval = {"add": ("ADDC", cval), "sub": ("ADDC", -cval), "mul": ("MULC", cval)}[op]
            else:                # c op x
                name = {"add": "ADDC", "sub": "RSUBC", "mul": "MULC"}[op]
            leaf = float_leaf(other) if other not in temps else None
            if leaf is not None:
                emit(f"LD{name}_{leaf[0]}_D{d}", leaf[1], b.const(f64_bits(cval)))
            else:
                gen(other, d)
                emit(f"{name}_D{d + 1}", b.const(f64_bits(cval)))
        else:
            def operand_form(x: Expr) -> Optional[tuple[str, int]]:
                if x in temps:
                    return f"T{temps[x]}", 0
                leaf = float_leaf(x)
                return (f"COL_{leaf[0]}", leaf[1]) if leaf is not None else None

            rform, lform = operand_form(right), operand_form(left)
            if rform is not None:      # top = left op right-operand
                gen(left, d)
                emit({"add": "ADD", "sub": "SUB", "mul": "MUL"}[op] + f"{rform[0]}_D{d + 1}", rform[1])
            elif lform is not None:    # top = left-operand op right  (reversed subtraction)
                gen(right, d)
                emit({"add": "ADD", "sub": "RSUB", "mul": "MUL"}[op] + f"{lform[0]}_D{d + 1}", lform[1])
            else:
                gen(left, d)
                gen(right, d + 1)
                emit({"add": "ADDF", "sub": "SUBF", "mul": "MULF"}[op] + f"_D{d + 2}")
        if tee and counts.get(e, 0) > 1 and len(temps) < RV_MAX_TEMPS:
            temps[e] = len(temps)
            emit(f"TEE{temps[e]}_D{d + 1}")

    def plain_sum_leaf(kind: str, e: Expr) -> Optional[tuple[str, int]]:
        if kind != "sum" or e.type != FLOAT or e in temps:
            return None
        if isinstance(e, EInput) or (isinstance(e, ECast) and isinstance(e.child, EInput)):
            return float_leaf(e)
        return None

    items = list(unique.items())
    skip = 0
    for pos, ((kind, e), slot) in enumerate(items):
        if skip:
            skip -= 1
            continue
        if kind == "count":
            emit("COUNT", slot)
            continue
        is_float = e.type == FLOAT
        first = plain_sum_leaf(kind, e)
        if first is not None:  # a run of SUM(column) over adjacent stage slots into adjacent accumulators: one instruction
            run = 1
            while run < 4 and pos + run < len(items):
                (k2, e2), s2 = items[pos + run]
                nxt = plain_sum_leaf(k2, e2)
                if nxt is None or nxt[0] != first[0] or nxt[1] != first[1] + run or s2 != slot + run:
                    break
                run += 1
            if run > 1 and f"AGGCOL{run}_{first[0]}" in RV:
                emit(f"AGGCOL{run}_{first[0]}", first[1], slot)
                skip = run - 1
                continue
        if kind == "sum" and is_float and e not in temps:
            if isinstance(e, EInput):
                bd = column(e)
                fused = {P_F32: "AGGCOL_F32", P_F64: "AGGCOL_F64"}.get(bd.phys)
                if fused:
                    emit(fused, bd.staged, slot)
                    continue
            if isinstance(e, ECast) and isinstance(e.child, EInput) and column(e.child).phys == P_I32:
                emit("AGGCOL_I32F", column(e.child).staged, slot)
                continue
        if is_float:
            shared = kind == "sum" and e not in temps and counts.get(e, 0) > 1 and len(temps) < RV_MAX_TEMPS
            gen(e, 0, tee=not shared)
            if shared:  # aggregate and keep the value for a later expression in one instruction
                temps[e] = len(temps)
                emit(f"AGGT{temps[e]}_SUMF_D1", slot)
                continue
        elif not load(e, 0):
            raise _RegvmUnsupported("integer expression")
        emit(f"AGG_{kind.upper()}{'F' if is_float else 'I'}_D1", slot)
    emit("END")
    if len(words) > K["MSC_VM_MAX_CODE2"]:
        raise _RegvmUnsupported("program too long")
    return words, text


def compile_build(resolver: Resolver, filters: Sequence[Expr], key: Expr) -> tuple[Program, Any]:
    """The build half of a join as a scan (msc_scan_join_build): the side's filters, RANK, GROUP <- key.  Returns the program
    and the dictionary of a STR key (joined on its code)."""
    b = ProgramBuilder(resolver)
    filters = split_conjunctions(filters)
    b.plan_cse([*filters, key])
    for f in filters:
        b.materialize(f, DST_FILTER)
    if filters:
        b.emit("RANK")
    key_dict = b.materialize(key, DST_GROUP)
    return b.end(), key_dict


@dataclass
class ProjectProgram:
    program: Program
    out_phys: list[int]
    out_dicts: list[Any]


def compile_project(resolver: Resolver, filters: Sequence[Expr], outputs: Sequence[Expr], probe: Optional[ProbeSpec] = None,
                    pre_filters: Sequence[Expr] = ()) -> ProjectProgram:
    b = ProgramBuilder(resolver)
    filters = split_conjunctions(filters)
    pre_filters = split_conjunctions(pre_filters)
    b.plan_cse([*pre_filters, *([probe.key] if probe else []), *filters, *outputs])
    for f in pre_filters:
        b.materialize(f, DST_FILTER)
    if probe is not None:
        b.emit_probe(probe)
    for f in filters:
        b.materialize(f, DST_FILTER)
    if filters or pre_filters or probe is not None:
        b.emit("RANK")
    if len(outputs) > K["MSC_VM_MAX_OUT"]:
        raise LoweringError("too many output columns in one projection")
    out_phys: list[int] = []
    out_dicts: list[Any] = []
    for i, e in enumerate(outputs):
        out_phys.append(P_U32 if e.type == STR else (P_F64 if e.type == FLOAT else P_I64))
        out_dicts.append(b.materialize(e, DST_OUT, i, out_u32=out_phys[-1] == P_U32))
    return ProjectProgram(b.end(), out_phys, out_dicts)
