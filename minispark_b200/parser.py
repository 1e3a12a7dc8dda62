"""SQL front end: ``parse_sql(text) -> DataFrame``.

The reference parses SQL with a ``parsimonious`` PEG grammar and a node visitor
(``src/mini_spark/parser.py:14-69`` grammar, ``:112-388`` visitor).  ``parsimonious`` is not
available here, so this is a from-scratch hand-written recursive-descent parser for the same
language that builds the same ``DataFrame`` / ``Col`` trees; it is pinned by restatements of the
reference's SQL->DataFrame cases (``tests/test_parser.py``) and SQL->rows cases (``tests/test_e2e.py``).

Language quirks kept on purpose (they are observable): table names are single-quoted; numbers are
integers; ``COUNT()`` takes no argument; every join type parses but runs as an inner join
(``parser.py:133``); exactly one GROUP BY column is executable (``dataframe.py:64-65``); BETWEEN takes
a column and two literals/columns; NOT parses but is unsupported by ``Col`` (``sql.py:44-45``); HAVING
aggregates are added as hidden ``_having_*`` aggregates and filtered after the GROUP BY
(``parser.py:153-161``).
"""

from __future__ import annotations

import operator
import re
from typing import Any, Callable, Optional

from .dataframe import DataFrame
from .sql import AggCol, Col, Lit
from .sql import Functions as F


class ParseError(Exception):
    """The text is not in the SQL subset (the reference raises parsimonious.ParseError)."""


class SemanticError(Exception):
    def __init__(self, message: str) -> None:
        super().__init__(message)


class GroupByError(SemanticError):
    pass


_WS = re.compile(r"\s+")
_COLUMN = re.compile(r"[A-Za-z_][A-Za-z0-9_\.]*")
_IDENT = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")
_NUMBER = re.compile(r"-?[0-9]+(\.[0-9]+)?")
_STRING = re.compile(r"'([^']*)'")
_TABLE = re.compile(r"'([a-zA-Z0-9_\-\./ ]+)'")
_COMPARATORS = [("<=", operator.le), (">=", operator.ge), ("!=", operator.ne), ("=", operator.eq),
                ("<", operator.lt), (">", operator.gt)]
_AGGREGATES = ("COUNT", "SUM", "AVG", "MIN", "MAX")
_JOIN_TYPES = (("JOIN",), ("LEFT", "JOIN"), ("RIGHT", "JOIN"), ("INNER", "JOIN"), ("FULL", "JOIN"))


class _Backtrack(Exception):
    pass


class _Parser:
    def __init__(self, text: str) -> None:
        self.text = text
        self.pos = 0

    # -- primitives ---------------------------------------------------------------------------------
    def fail(self, what: str) -> "_Backtrack":
        return _Backtrack(f"expected {what} at offset {self.pos}: {self.text[self.pos:self.pos + 30]!r}")

    def ws(self, required: bool = False) -> bool:
        m = _WS.match(self.text, self.pos)
        if m:
            self.pos = m.end()
            return True
        if required:
            raise self.fail("whitespace")
        return False

    def word(self, literal: str) -> None:
        if not self.text.startswith(literal, self.pos):
            raise self.fail(repr(literal))
        self.pos += len(literal)

    def regex(self, pattern: re.Pattern, what: str) -> re.Match:
        m = pattern.match(self.text, self.pos)
        if not m:
            raise self.fail(what)
        self.pos = m.end()
        return m

    def attempt(self, rule: Callable[[], Any]) -> tuple[bool, Any]:
        start = self.pos
        try:
            return True, rule()
        except _Backtrack:
            self.pos = start
            return False, None

    def first_of(self, *rules: Callable[[], Any]) -> Any:
        last: Optional[_Backtrack] = None
        for rule in rules:
            start = self.pos
            try:
                return rule()
            except _Backtrack as e:
                self.pos = start
                last = e
        raise last if last else self.fail("alternative")

    # -- query --------------------------------------------------------------------------------------
    def query(self) -> DataFrame:
        self.ws()
        self.word("SELECT")
        self.ws(True)
        select_list = self.select_list()
        self.ws(True)
        self.word("FROM")
        self.ws(True)
        df = self.table_reference()
        joins = []
        while True:
            ok, join = self.attempt(self.join_clause)
            if not ok:
                break
            joins.append(join)
        ok, where = self.attempt(self.where_clause)
        ok_group, group = self.attempt(self.group_by_clause)
        self.ws()
        self.word(";")
        self.ws()
        if self.pos != len(self.text):
            raise self.fail("end of query")

        for other, cond in joins:
            df = df.join(other, on=cond, how="inner")
        if ok:
            df = df.filter(where)
        if not ok_group:
            return df.select(*select_list)
        group_cols, having = group
        group_names = {c.name for c in group_cols}
        agg_cols = [c for c in select_list if type(c) is AggCol]
        stray = [c for c in select_list if type(c) is not AggCol and c.name not in group_names]
        if stray:
            raise GroupByError(
                "All selected columns must be aggregate functions or part of the key when using GROUP BY:\n" f"{stray}"
            )
        if having is not None:
            hidden = [c for c in having.all_nested_columns if type(c) is AggCol]
            for c in hidden:
                c.name = f"_having_{c.name}"
            agg_cols.extend(hidden)
        df = df.group_by(*group_cols).agg(*agg_cols)
        if having is not None:
            df = df.filter(having.normalize_agg_columns())
        return df.select(*[Col(c.name) for c in select_list])

    def select_list(self) -> list[Col]:
        items = [self.select_item()]
        while True:
            def more() -> Col:
                self.ws()
                self.word(",")
                self.ws()
                return self.select_item()
            ok, item = self.attempt(more)
            if not ok:
                return items
            items.append(item)

    def select_item(self) -> Col:
        return self.first_of(self.star, self.aggregate_function_call, self.expr_aliased)

    def star(self) -> Col:
        self.word("*")
        return Col("*")

    def alias(self) -> str:
        self.ws(True)
        self.word("AS")
        self.ws(True)
        return self.regex(_IDENT, "identifier").group(0)

    def aggregate_function_call(self) -> AggCol:
        for name in _AGGREGATES:
            if self.text.startswith(name + "(", self.pos):
                break
        else:
            raise self.fail("aggregate function")
        self.pos += len(name) + 1
        ok, arg = self.attempt(self.expr)
        self.word(")")
        ok_alias, alias = self.attempt(self.alias)
        if name == "COUNT":
            if ok:
                raise AssertionError("COUNT() takes no argument")
            agg = F.count()
        else:
            if not ok:
                raise AssertionError(f"{name}() needs one argument")
            agg = {"SUM": F.sum, "AVG": F.avg, "MIN": F.min, "MAX": F.max}[name](arg)
        return agg.alias(alias) if ok_alias else agg

    def expr_aliased(self) -> Col:
        col = self.expr()
        ok, alias = self.attempt(self.alias)
        return col.alias(alias) if ok else col

    def table_reference(self) -> DataFrame:
        name = self.regex(_TABLE, "quoted table name").group(1)
        df = DataFrame().table(name)
        ok, alias = self.attempt(self.alias)
        return df.alias(alias) if ok else df

    def join_clause(self) -> tuple[DataFrame, Col]:
        self.ws(True)

        def join_type(words: tuple[str, ...]) -> Callable[[], None]:
            def rule() -> None:
                for i, w in enumerate(words):
                    if i:
                        self.ws(True)
                    self.word(w)
            return rule

        self.first_of(*[join_type(words) for words in _JOIN_TYPES])
        self.ws(True)
        table = self.table_reference()
        self.ws(True)
        self.word("ON")
        self.ws(True)
        return table, self.condition()

    def where_clause(self) -> Col:
        self.ws(True)
        self.word("WHERE")
        self.ws(True)
        return self.condition()

    def group_by_clause(self) -> tuple[list[Col], Optional[Col]]:
        self.ws(True)
        self.word("GROUP")
        self.ws(True)
        self.word("BY")
        self.ws(True)
        cols = [self.column_name()]
        while True:
            def more() -> Col:
                self.ws()
                self.word(",")
                self.ws()
                return self.column_name()
            ok, col = self.attempt(more)
            if not ok:
                break
            cols.append(col)

        def having() -> Col:
            self.ws(True)
            self.word("HAVING")
            self.ws(True)
            return self.condition()

        ok, cond = self.attempt(having)
        return cols, (cond if ok else None)

    # -- conditions ---------------------------------------------------------------------------------
    def condition(self) -> Col:
        return self._chain(self.and_expr, "OR", operator.or_)

    def and_expr(self) -> Col:
        return self._chain(self.not_expr, "AND", operator.and_)

    def _chain(self, operand: Callable[[], Col], keyword: str, combine: Callable[[Any, Any], Col]) -> Col:
        left = operand()
        while True:
            def more() -> Col:
                self.ws(True)
                self.word(keyword)
                self.ws(True)
                return operand()
            ok, right = self.attempt(more)
            if not ok:
                return left
            left = combine(left, right)

    def not_expr(self) -> Col:
        def negation() -> None:
            self.word("NOT")
            self.ws(True)
        negated, _ = self.attempt(negation)
        pred = self.predicate()
        return ~pred if negated else pred  # Col.__invert__ raises NotImplementedError, as in the reference

    def predicate(self) -> Any:
        return self.first_of(self.comparison, self.parenthised_condition, self.string_literal, self.between, self.like)

    def comparison(self) -> Col:
        left = self.expr()
        self.ws()
        for symbol, fn in _COMPARATORS:
            if self.text.startswith(symbol, self.pos):
                self.pos += len(symbol)
                break
        else:
            raise self.fail("comparator")
        self.ws()
        right = self.expr()
        return fn(left, right)

    def parenthised_condition(self) -> Col:
        self.word("(")
        self.ws()
        cond = self.condition()
        self.ws()
        self.word(")")
        return cond

    def between(self) -> Col:
        col = self.column_name()
        self.ws(True)
        self.word("BETWEEN")
        self.ws(True)
        start = self.first_of(self.string_literal, self.column_name)
        self.ws(True)
        self.word("AND")
        self.ws(True)
        end = self.first_of(self.string_literal, self.column_name)
        return col.between(start, end)

    def like(self) -> Col:
        col = self.expr()
        self.ws(True)
        self.word("LIKE")
        self.ws(True)
        return col.like(self.string_literal())

    # -- arithmetic ---------------------------------------------------------------------------------
    def expr(self) -> Any:
        return self._binary(self.mul_expr, {"+": operator.add, "-": operator.sub})

    def mul_expr(self) -> Any:
        return self._binary(self.atom, {"*": operator.mul, "/": operator.truediv})

    def _binary(self, operand: Callable[[], Any], ops: dict[str, Callable[[Any, Any], Any]]) -> Any:
        left = operand()
        while True:
            def more() -> tuple[Callable[[Any, Any], Any], Any]:
                self.ws()
                if self.pos >= len(self.text) or self.text[self.pos] not in ops:
                    raise self.fail("operator")
                fn = ops[self.text[self.pos]]
                self.pos += 1
                self.ws()
                return fn, operand()
            ok, found = self.attempt(more)
            if not ok:
                return left
            fn, right = found
            left = fn(left, right)

    def atom(self) -> Any:
        return self.first_of(self.function_call, self.number, self.column_name, self.parenthised_expr, self.string_literal)

    def function_call(self) -> Col:
        name = self.regex(_IDENT, "identifier").group(0)
        self.ws()
        self.word("(")
        self.ws()
        ok, args = self.attempt(self.argument_list)
        self.ws()
        self.word(")")
        args = args if ok else []
        if name == "COUNT":
            if args:
                raise AssertionError("COUNT() takes no argument")
            return F.count()
        if name == "SUM":
            if len(args) != 1:
                raise AssertionError("SUM() needs one argument")
            return F.sum(args[0])
        raise SemanticError(f"Unsupported function: {name}")

    def argument_list(self) -> list[Any]:
        args = [self.expr()]
        while True:
            def more() -> Any:
                self.ws()
                self.word(",")
                self.ws()
                return self.expr()
            ok, arg = self.attempt(more)
            if not ok:
                return args
            args.append(arg)

    def parenthised_expr(self) -> Any:
        self.word("(")
        self.ws()
        inner = self.expr()
        self.ws()
        self.word(")")
        return inner

    def number(self) -> Lit:
        text = self.regex(_NUMBER, "number").group(0)
        return Lit(int(text))  # integers only, like the reference (parser.py:352-353)

    def column_name(self) -> Col:
        return Col(self.regex(_COLUMN, "column name").group(0))

    def string_literal(self) -> str:
        return self.regex(_STRING, "string literal").group(1)


def parse_sql(sql: str) -> DataFrame:
    parser = _Parser(sql)
    try:
        return parser.query()
    except _Backtrack as e:
        raise ParseError(str(e)) from None
