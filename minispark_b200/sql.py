"""Column-expression DSL: ``Col``, ``Lit``, ``AggCol`` and friends.

Host-side mirror of the reference's ``src/mini_spark/sql.py`` *interface* (class names, attribute
names, derived column names and type rules) so queries written against the reference build the same
expression trees here.  It deliberately has no per-row interpreter (``sql.py:262-266`` is the
reference's compute path): trees built from these classes are *input* to
:mod:`minispark_b200.lowering`, which compiles them to the expression program evaluated by the
fused CUDA scan kernel.

Naming rules reproduced (they become output column names, reference ``sql.py:260,369,409,464``):
``<l>_<op>_<r>`` for binary operators, ``lit_<v>`` for literals, ``<agg>_<col>`` for aggregates,
``<col>_like_<pattern>`` for LIKE, ``count`` for ``Functions.count()``.
"""

from __future__ import annotations

import operator as _op
import re
from datetime import datetime
from typing import Any, Callable, Iterable, Iterator, Literal

from .constants import ColumnType, ColumnTypePython, Schema

AggregationType = Literal["sum", "min", "max", "avg"]

# operator -> symbol used by explain(); the key's __name__ is what ends up in column names
BINOP_SYMBOLS: dict[Callable[[Any, Any], Any], str] = {
    _op.add: "+", _op.sub: "-", _op.mul: "*", _op.truediv: "/", _op.floordiv: "//", _op.mod: "%",
    _op.pow: "**", _op.eq: "==", _op.ne: "!=", _op.lt: "<", _op.le: "<=", _op.gt: ">", _op.ge: ">=",
    _op.and_: "and", _op.or_: "or",
}
_DUNDERS = {
    "__lt__": _op.lt, "__le__": _op.le, "__gt__": _op.gt, "__ge__": _op.ge, "__eq__": _op.eq,
    "__ne__": _op.ne, "__add__": _op.add, "__sub__": _op.sub, "__mul__": _op.mul,
    "__truediv__": _op.truediv, "__floordiv__": _op.floordiv, "__mod__": _op.mod,
    "__and__": _op.and_, "__or__": _op.or_,
}


def _lookup(schema: Schema, name: str) -> tuple[int, ColumnType] | None:
    for pos, (col_name, col_type) in enumerate(schema):
        if col_name == name:
            return pos, col_type
    return None


class Col:
    """A reference to a named column; operators build :class:`BinaryOperatorColumn` trees."""

    def __init__(self, name: str) -> None:
        self.name = name

    # comparison / arithmetic / boolean dunders are attached below from _DUNDERS
    def __invert__(self) -> "Col":
        raise NotImplementedError  # NOT is unsupported, as in the reference (sql.py:44-45)

    def __hash__(self) -> int:
        return hash((type(self), self.name))

    def like(self, pattern: str) -> "Col":
        return LikeColumn(self, pattern)

    def between(self, start: "Col | ColumnTypePython", end: "Col | ColumnTypePython") -> "Col":
        # `start <= self` in the reference (sql.py:72-73): a Col start keeps that order, a literal
        # start is reflected by Python into `self >= start`.
        if isinstance(start, Col):
            lower = BinaryOperatorColumn(start, self, _op.le)
        else:
            lower = BinaryOperatorColumn(self, start, _op.ge)
        upper = BinaryOperatorColumn(self, end, _op.le)
        return BinaryOperatorColumn(lower, upper, _op.and_)

    def alias(self, name: str) -> "Col":
        return AliasColumn(self, name)

    def normalize_agg_columns(self) -> "Col":
        return self

    @property
    def all_nested_columns(self) -> Iterable["Col"]:
        yield self

    def infer_type(self, schema: Schema) -> ColumnType:
        hit = _lookup(schema, self.name)
        if hit is None:
            raise ValueError(f'Column "{self.name}" not found in schema {schema}')
        return hit[1]

    def __str__(self) -> str:
        return self.name

    __repr__ = __str__


def _attach_operators() -> None:
    def make(fn: Callable[[Any, Any], Any]) -> Callable[["Col", Any], "Col"]:
        def method(self: "Col", other: Any) -> "Col":
            return BinaryOperatorColumn(self, other, fn)
        return method
    for dunder, fn in _DUNDERS.items():
        setattr(Col, dunder, make(fn))


class SchemaCol(Col):
    """A column bound to a position (kept for interface parity; unused by the GPU lowering)."""

    def __init__(self, name: str, col_pos: int) -> None:
        super().__init__(name)
        self.col_pos = col_pos


class AliasColumn(Col):
    def __init__(self, original_col: Col, name: str) -> None:
        super().__init__(name)
        self.original_col = original_col

    def __hash__(self) -> int:
        return hash((type(self), hash(self.original_col), self.name))

    @property
    def all_nested_columns(self) -> Iterable[Col]:
        yield self
        yield from self.original_col.all_nested_columns

    def infer_type(self, schema: Schema) -> ColumnType:
        return self.original_col.infer_type(schema)

    def __str__(self) -> str:
        return f"({self.original_col}) AS {self.name}"

    __repr__ = __str__


class LikeColumn(Col):
    """SQL LIKE: ``%`` = any run, ``_`` = any one char, anchored both ends (sql.py:178-179)."""

    def __init__(self, original_col: Col, pattern: str) -> None:
        super().__init__(f"{original_col.name}_like_{pattern}")
        self.original_col = original_col
        self.pattern = pattern
        self.regex = self.generate_regex(pattern)

    @staticmethod
    def generate_regex(pattern: str) -> str:
        return "^" + re.escape(pattern).replace("%", ".*").replace("_", ".") + "$"

    def __hash__(self) -> int:
        return hash((type(self), hash(self.original_col), self.pattern))

    @property
    def all_nested_columns(self) -> Iterable[Col]:
        yield self
        yield from self.original_col.all_nested_columns

    def infer_type(self, schema: Schema) -> ColumnType:
        if self.original_col.infer_type(schema) != ColumnType.STRING:
            raise AssertionError("LIKE operator can only be applied to string columns")
        return ColumnType.STRING

    def __str__(self) -> str:
        return f"({self.original_col}) LIKE '{self.pattern}'"

    __repr__ = __str__


class BinaryOperatorColumn(Col):
    """``left <op> right``.  ``infer_type`` applies the reference's coercions (sql.py:277-303)."""

    def __init__(self, left_side: Any, right_side: Any, operator: Callable[[Any, Any], Any]) -> None:
        self.left_side: Col = left_side if isinstance(left_side, Col) else Lit(left_side)
        self.right_side: Col = right_side if isinstance(right_side, Col) else Lit(right_side)
        self.operator = operator
        self.left_type_convert_to: ColumnType | None = None
        self.right_type_convert_to: ColumnType | None = None
        super().__init__(f"{self.left_side.name}_{operator.__name__}_{self.right_side.name}")

    def __hash__(self) -> int:
        return hash((type(self), hash(self.left_side), hash(self.right_side), self.operator))

    @property
    def all_nested_columns(self) -> Iterable[Col]:
        yield self
        yield from self.left_side.all_nested_columns
        yield from self.right_side.all_nested_columns

    def infer_type(self, schema: Schema) -> ColumnType:
        lt = self.left_side.infer_type(schema)
        rt = self.right_side.infer_type(schema)
        if self.operator is _op.truediv:  # '/' is always true division -> FLOAT
            self.left_type_convert_to = None if lt == ColumnType.FLOAT else ColumnType.FLOAT
            self.right_type_convert_to = None if rt == ColumnType.FLOAT else ColumnType.FLOAT
            return ColumnType.FLOAT
        if {lt, rt} == {ColumnType.INTEGER, ColumnType.FLOAT}:  # INT (+) FLOAT -> FLOAT
            self.left_type_convert_to = ColumnType.FLOAT if lt == ColumnType.INTEGER else None
            self.right_type_convert_to = ColumnType.FLOAT if rt == ColumnType.INTEGER else None
            return ColumnType.FLOAT
        # an ISO string literal against a TIMESTAMP is parsed in place
        if lt == ColumnType.STRING and rt == ColumnType.TIMESTAMP:
            if type(self.left_side) is not Lit:
                raise AssertionError("only a string literal can be compared with a TIMESTAMP")
            self.left_side.value = datetime.fromisoformat(str(self.left_side.value))
            lt = ColumnType.TIMESTAMP
        if rt == ColumnType.STRING and lt == ColumnType.TIMESTAMP:
            if type(self.right_side) is not Lit:
                raise AssertionError("only a string literal can be compared with a TIMESTAMP")
            self.right_side.value = datetime.fromisoformat(str(self.right_side.value))
            rt = ColumnType.TIMESTAMP
        if lt != rt:
            raise TypeError(f"Type mismatch in binary operation: {lt} {self.operator} {rt}")
        return lt

    def normalize_agg_columns(self) -> Col:
        return BinaryOperatorColumn(
            self.left_side.normalize_agg_columns(), self.right_side.normalize_agg_columns(), self.operator
        )

    def extract_left_right_key(self, left_schema: Schema, right_schema: Schema) -> tuple[Col, Col]:
        """Which side of an equi-join condition belongs to which input (sql.py:343-355)."""
        if type(self.left_side) is not Col or type(self.right_side) is not Col:
            raise AssertionError("join keys must be plain columns")
        a, b = self.left_side.name, self.right_side.name
        if a == b:
            raise AssertionError("Join keys must be different columns")
        left_names = {n for n, _ in left_schema}
        right_names = {n for n, _ in right_schema}
        if a in left_names and b in right_names:
            return self.left_side, self.right_side
        if a in right_names and b in left_names:
            return self.right_side, self.left_side
        raise ValueError("Join keys must be from different tables")

    def __str__(self) -> str:
        return f"({self.left_side}) {BINOP_SYMBOLS[self.operator]} ({self.right_side})"

    __repr__ = __str__


class Lit(Col):
    def __init__(self, value: ColumnTypePython) -> None:
        self.value = value
        super().__init__(f"lit_{value}")

    def __hash__(self) -> int:
        return hash((type(self), self.value))

    @property
    def all_nested_columns(self) -> Iterator[Col]:
        return iter(())

    def infer_type(self, schema: Schema) -> ColumnType:  # noqa: ARG002
        return ColumnType.of(self.value)

    def __str__(self) -> str:
        return str(self.value)

    __repr__ = __str__


class AggCol(Col):
    """``type(original_col)`` aggregated per group; ``type`` in sum/min/max/avg (sql.py:398-446)."""

    def __init__(self, agg_type: AggregationType, original_col: Col) -> None:
        super().__init__(f"{agg_type}_{original_col.name}")
        self.original_col = original_col
        self.type = agg_type

    def infer_type(self, schema: Schema) -> ColumnType:
        return ColumnType.FLOAT if self.type == "avg" else self.original_col.infer_type(schema)

    def alias(self, name: str) -> "AggCol":  # renames in place, like the reference
        self.name = name
        return self

    @property
    def all_nested_columns(self) -> Iterable[Col]:
        yield self
        if self.original_col is not None:
            yield from self.original_col.all_nested_columns

    def normalize_agg_columns(self) -> Col:
        return Col(self.name)

    def expand_avg(self) -> Iterable["AggCol"]:
        if self.type != "avg":
            yield self
            return
        yield AggCol("sum", self.original_col).alias(f"{self.name}_sum")
        yield AggCol("sum", Lit(1)).alias(f"{self.name}_count")

    def projection(self) -> Col:
        if self.type == "avg":
            return (Col(f"{self.name}_sum") / Col(f"{self.name}_count")).alias(self.name)
        return Col(self.name)

    def __str__(self) -> str:
        return f"{self.type}({self.original_col}) AS {self.name}"

    __repr__ = __str__


class Functions:
    @staticmethod
    def min(col: Col) -> AggCol:
        return AggCol("min", col)

    @staticmethod
    def max(col: Col) -> AggCol:
        return AggCol("max", col)

    @staticmethod
    def sum(col: Col) -> AggCol:
        return AggCol("sum", col)

    @staticmethod
    def avg(col: Col) -> AggCol:
        return AggCol("avg", col)

    @staticmethod
    def count() -> AggCol:
        return AggCol("sum", Lit(1)).alias("count")


_attach_operators()
