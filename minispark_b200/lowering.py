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
import struct
from dataclasses import dataclass, field
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Iterable, Optional, Protocol, Sequence

from .constants import ColumnType, Schema
from .io import BlockFile, datetime_to_timestamp
from .native import K, OP, P_F32, P_F64, P_I32, P_I64, P_U8, P_U16, P_U32

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
        if lt == STR and op not in ("eq", "ne"):
            raise LoweringError("ordering comparisons on STRING are not implemented on the GPU engine")
        return EBin(BOOL, op, left, right)
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


def fuse_selects(node: LNode) -> LNode:
    """Merge stacked LSelects bottom-up by inlining the inner projection into the outer one."""
    if isinstance(node, LSelect):
        child = fuse_selects(node.child)
        if isinstance(child, LSelect):
            filters = list(child.filters) + [substitute(f, child.outputs) for f in node.filters]
            outputs = [substitute(o, child.outputs) for o in node.outputs]
            return LSelect(node.schema, child.child, filters, outputs)
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


class Resolver(Protocol):
    """Run-time services the compiler needs for STRING operands (implemented by the engine)."""

    def binding(self, index: int) -> Binding: ...
    def literal_code(self, dict_id: Any, text: str) -> int: ...           # -1 when absent
    def like_lut(self, dict_id: Any, pattern: str) -> int: ...            # -> LUT slot
    def translate_lut(self, dict_id: Any, token: str) -> tuple[int, Any]: ...  # -> (LUT slot, target dict)
    def same_dict(self, a: Any, b: Any) -> bool: ...
    def recode_lut(self, src: Any, dst: Any) -> int: ...                  # LUT slot: codes of src -> codes of dst


_LOAD = {P_U8: "LOAD_U8", P_U16: "LOAD_U16", P_U32: "LOAD_U32", P_I32: "LOAD_I32", P_I64: "LOAD_I64", P_F32: "LOAD_F32",
         P_F64: "LOAD_F64"}
_F_OPS = {"add": "ADD_F", "sub": "SUB_F", "mul": "MUL_F", "truediv": "DIV_F", "floordiv": "FLOORDIV_F", "mod": "MOD_F",
          "lt": "LT_F", "le": "LE_F", "gt": "GT_F", "ge": "GE_F", "eq": "EQ_F", "ne": "NE_F"}
_I_OPS = {"add": "ADD_I", "sub": "SUB_I", "mul": "MUL_I", "floordiv": "FLOORDIV_I", "mod": "MOD_I",
          "lt": "LT_I", "le": "LE_I", "gt": "GT_I", "ge": "GE_I", "eq": "EQ_I", "ne": "NE_I", "and": "AND", "or": "OR"}
MAX_DEPTH = K["MSC_VM_MAX_DEPTH"]
MAX_TEMPS = K["MSC_VM_MAX_TEMPS"]


def f64_bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", float(x)))[0]


@dataclass
class Program:
    code: list[int] = field(default_factory=list)
    consts: list[int] = field(default_factory=list)
    text: list[str] = field(default_factory=list)  # disassembly, for explain/tests

    def words(self) -> list[int]:
        return list(self.code)


class ProgramBuilder:
    """Emits the postfix program, tracking the static stack depth of every instruction."""

    def __init__(self, resolver: Resolver) -> None:
        self.r = resolver
        self.p = Program()
        self.depth = 0
        self.temps: dict[Expr, int] = {}
        self.temp_candidates: set[Expr] = set()
        self.str_dict: dict[Expr, Any] = {}

    # -- emission helpers -----------------------------------------------------------------------
    def emit(self, name: str, arg: int = 0, delta: int = 0) -> None:
        op = OP[name]
        if self.depth > MAX_DEPTH or (delta > 0 and self.depth + delta > MAX_DEPTH):
            raise LoweringError(f"expression needs more than {MAX_DEPTH} stack slots")
        if arg < 0 or arg > 0xFFFF:
            raise LoweringError("instruction argument out of range")
        self.p.code.append(op | (self.depth << 8) | (arg << 16))
        self.p.text.append(f"{name} {arg}" if arg or name in ("CONST", "TEE", "GET", "LUT8", "LUT32") or name.startswith(("LOAD", "AGG", "STORE")) else name)
        self.depth += delta

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

    # -- CSE ------------------------------------------------------------------------------------
    def plan_temps(self, roots: Iterable[Expr]) -> None:
        """Pick up to MAX_TEMPS repeated non-trivial sub-expressions to keep in TEE/GET temporaries."""
        counts: dict[Expr, int] = {}

        def walk(e: Expr) -> None:
            if isinstance(e, (EInput, EConst)):
                return
            counts[e] = counts.get(e, 0) + 1
            if counts[e] == 1:
                for c in expr_children(e):
                    walk(c)

        for root in roots:
            walk(root)

        def cost(e: Expr) -> int:
            return 1 + sum(cost(c) for c in expr_children(e))

        repeated = sorted((e for e, n in counts.items() if n > 1 and e.type != STR and not isinstance(e, ECast)),
                          key=lambda e: -cost(e) * (counts[e] - 1))
        self.temp_candidates = set(repeated[:MAX_TEMPS])

    # -- expression evaluation: leaves the value on the stack -------------------------------------
    def value(self, e: Expr) -> None:
        if e in self.temps:
            self.emit("GET", self.temps[e], +1)
            return
        self._value(e)
        if e in self.temp_candidates and e not in self.temps and len(self.temps) < MAX_TEMPS:
            slot = len(self.temps)
            self.temps[e] = slot
            self.emit("TEE", slot, 0)

    def _value(self, e: Expr) -> None:
        if isinstance(e, EInput):
            b = self.r.binding(e.index)
            if b.staged is not None:
                self.emit(_LOAD[b.phys], b.staged, +1)
            else:
                self.emit(_LOAD[b.phys].replace("LOAD", "LOADG"), b.index | (b.gather << 8), +1)
            return
        if isinstance(e, EConst):
            if e.type == STR:
                raise LoweringError("a string literal is only supported as an operand of =, !=, + or as a selected column")
            self.emit("CONST", self.const(f64_bits(e.value) if e.type == FLOAT else int(e.value)), +1)
            return
        if isinstance(e, ECast):
            if isinstance(e.child, EConst):  # fold float(int literal)
                self.emit("CONST", self.const(f64_bits(float(e.child.value))), +1)
                return
            self.value(e.child)
            self.emit("I2F")
            return
        if isinstance(e, ELike):
            d = self.string(e.child)
            self.emit("LUT8", self.r.like_lut(d, e.pattern))
            return
        if isinstance(e, (ETranslate, ECode)):
            self.string(e.child if isinstance(e, ECode) else e)
            return
        if isinstance(e, EConcat):
            raise LoweringError("internal: concat must be materialised before compilation")
        if isinstance(e, EBin):
            if e.left.type == STR or e.right.type == STR:
                self._string_compare(e)
                return
            self.value(e.left)
            self.value(e.right)
            is_float = (e.left.type == FLOAT) if e.op in _CMP else (e.type == FLOAT)
            table = _F_OPS if is_float else _I_OPS
            if e.op not in table:
                raise LoweringError(f"operator {e.op} not available for type {e.type}")
            self.emit(table[e.op], 0, -1)
            return
        raise LoweringError(f"cannot compile {e}")

    def string(self, e: Expr) -> Any:
        """Push the dictionary code of a STR expression; returns the dictionary it is coded in."""
        if isinstance(e, EInput):
            b = self.r.binding(e.index)
            self._value(e)
            return b.dict_id
        if isinstance(e, ETranslate):
            src = self.string(e.child)
            slot, target = self.r.translate_lut(src, e.token)
            if slot >= 0:
                self.emit("LUT32", slot)
            return target
        raise LoweringError(f"string expression {show(e)} must be materialised first")

    def _string_compare(self, e: EBin) -> None:
        left, right = e.left, e.right
        if isinstance(left, EConst) and not isinstance(right, EConst):
            left, right = right, left
        if isinstance(left, EConst):  # literal vs literal: fold
            truth = (left.value == right.value) == (e.op == "eq")
            self.emit("CONST", self.const(int(truth)), +1)
            return
        d = self.string(left)
        if isinstance(right, EConst):
            code = self.r.literal_code(d, right.value)
            self.emit("CONST", self.const(code), +1)  # -1 never equals a code
        else:
            d2 = self.string(right)
            if not self.r.same_dict(d, d2):
                self.emit("LUT32", self.r.recode_lut(d2, d))
        self.emit("EQ_I" if e.op == "eq" else "NE_I", 0, -1)

    # -- statement-level helpers ------------------------------------------------------------------
    def filter(self, e: Expr) -> None:
        self.value(e)
        self.emit("FILTER", 0, -1)

    def end(self) -> Program:
        if self.depth != 0:
            raise LoweringError("internal: unbalanced expression stack")
        self.p.code.append(OP["END"])
        self.p.text.append("END")
        if len(self.p.code) > K["MSC_VM_MAX_CODE"]:
            raise LoweringError("expression program too long")
        return self.p


_AGG_OPS = {("sum", FLOAT): ("AGG_SUM_F", K["MSC_AGG_SUM_F"]), ("sum", INT): ("AGG_SUM_I", K["MSC_AGG_SUM_I"]),
            ("min", FLOAT): ("AGG_MIN_F", K["MSC_AGG_MIN_F"]), ("max", FLOAT): ("AGG_MAX_F", K["MSC_AGG_MAX_F"]),
            ("min", INT): ("AGG_MIN_I", K["MSC_AGG_MIN_I"]), ("max", INT): ("AGG_MAX_I", K["MSC_AGG_MAX_I"])}


@dataclass
class AggregateProgram:
    program: Program
    agg_kinds: list[int]       # MSC_AGG_* per unique accumulator slot
    slot_of: list[int]         # slot of each requested aggregate (duplicates share a slot)
    group_dict: Any            # dictionary of a STR group key (None otherwise)


def compile_aggregate(resolver: Resolver, filters: Sequence[Expr], group: Expr, aggs: Sequence[tuple[str, Expr]]) -> AggregateProgram:
    b = ProgramBuilder(resolver)
    norm = [(k, EConst(INT, 1) if k == "count" else (EBin(INT, "add", e, EConst(INT, 0)) if e.type == BOOL else e)) for k, e in aggs]
    b.plan_temps([*filters, group, *[e for k, e in norm if k != "count"]])
    for f in filters:
        b.filter(f)
    group_dict = None
    if group.type == STR:
        group_dict = b.string(group)
    else:
        b.value(group)
    b.emit("GROUP", 0, -1)
    slots: dict[tuple[str, Expr], int] = {}
    kinds: list[int] = []
    slot_of: list[int] = []
    for kind, e in norm:
        key = (kind, e)
        if key in slots:
            slot_of.append(slots[key])
            continue
        slot = len(kinds)
        if slot >= K["MSC_VM_MAX_AGGS"]:
            raise LoweringError("too many aggregates in one GROUP BY")
        slots[key] = slot
        slot_of.append(slot)
        if kind == "count":
            kinds.append(K["MSC_AGG_SUM_I"])
            b.emit("AGG_COUNT", slot, 0)
            continue
        opname, agg_kind = _AGG_OPS[(kind, FLOAT if e.type == FLOAT else INT)]
        kinds.append(agg_kind)
        b.value(e)
        b.emit(opname, slot, -1)
    return AggregateProgram(b.end(), kinds, slot_of, group_dict)


@dataclass
class ProjectProgram:
    program: Program
    out_phys: list[int]
    out_dicts: list[Any]


def compile_project(resolver: Resolver, filters: Sequence[Expr], outputs: Sequence[Expr]) -> ProjectProgram:
    b = ProgramBuilder(resolver)
    b.plan_temps([*filters, *outputs])
    for f in filters:
        b.filter(f)
    if filters:
        b.emit("RANK")
    out_phys: list[int] = []
    out_dicts: list[Any] = []
    if len(outputs) > K["MSC_VM_MAX_OUT"]:
        raise LoweringError("too many output columns in one projection")
    for i, e in enumerate(outputs):
        if e.type == STR:
            out_dicts.append(b.string(e))
            b.emit("STORE_U32", i, -1)
            out_phys.append(P_U32)
        elif isinstance(e, ECode):  # integer code of a string key; remember which dictionary it is coded in
            out_dicts.append(b.string(e.child))
            b.emit("STORE_I64", i, -1)
            out_phys.append(P_I64)
        else:
            b.value(e)
            out_dicts.append(None)
            if e.type == FLOAT:
                b.emit("STORE_F64", i, -1)
                out_phys.append(P_F64)
            else:
                b.emit("STORE_I64", i, -1)
                out_phys.append(P_I64)
    return ProjectProgram(b.end(), out_phys, out_dicts)
