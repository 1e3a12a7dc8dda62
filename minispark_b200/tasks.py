"""Logical operator tree (``Task`` chain) that a ``DataFrame`` builds and an engine executes.

Mirror of the *shape* of the reference's ``src/mini_spark/tasks.py``: same class names, field names
and ``validate_schema`` rules/errors (``tasks.py:86-102,123-129,179-182,242-252,312-333``), so the
reference's DataFrame-level tests read unchanged.  What is intentionally absent is the reference's
per-chunk Python compute (``execute``/``generate_chunks``/``write``, ``tasks.py:79-84,117-121,
167-177,201-240,270-310,347-375``): in this framework those operators are CUDA kernels reached
through :class:`minispark_b200.execution.CudaExecutionEngine`.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Iterable, Literal, Optional

from .constants import Schema
from .io import BlockFile
from .sql import AggCol, BinaryOperatorColumn, Col, LikeColumn

JoinType = Literal["inner", "left", "right", "outer"]


def nice_schema(schema: Schema | None) -> str:
    if schema is None:
        return "None"
    return "[" + ", ".join(f"{name}:{col_type.name}" for name, col_type in schema) + "]"


def _plain_names(columns: Iterable[Col]) -> set[str]:
    return {c.name for column in columns for c in column.all_nested_columns if type(c) is Col}


def _require_known(columns: Iterable[Col], schema: Schema, what: str) -> None:
    known = {name for name, _ in schema}
    unknown = [name for name in _plain_names(columns) if name not in known]
    if unknown:
        raise ValueError(f"Unknown columns in {what}: {unknown}")


@dataclass
class Task:
    parent_task: "Task" = field(repr=False)
    inferred_schema: Optional[Schema] = None

    def validate_schema(self) -> Schema:
        return self.parent_task.validate_schema()

    @property
    def task_chain(self) -> Iterable["Task"]:
        if type(self) is VoidTask:
            return
        yield from self.parent_task.task_chain
        yield self

    def _describe(self) -> str:
        return type(self).__name__ + "()"

    def explain(self, lvl: int = 0) -> None:
        indent = "  " * lvl + ("+- " if lvl > 0 else "")
        print(f"{indent} {self._describe()}:{nice_schema(self.inferred_schema)}")  # noqa: T201
        self.parent_task.explain(lvl + 1)


@dataclass
class VoidTask(Task):
    parent_task: Optional[Task] = None  # type: ignore[assignment]

    def validate_schema(self) -> Schema:
        return []

    def explain(self, lvl: int = 0) -> None:  # noqa: ARG002
        return


@dataclass(kw_only=True)
class ProducerTask(Task):
    pass


@dataclass(kw_only=True)
class ConsumerTask(Task):
    pass


@dataclass(kw_only=True)
class WriterTask(Task):
    pass


@dataclass(kw_only=True)
class LoadTableBlockTask(ProducerTask):
    file_path: Path
    alias: str = ""

    @property
    def file_schema(self) -> Schema:
        return BlockFile(self.file_path).file_schema

    def validate_schema(self) -> Schema:
        if self.parent_task.validate_schema() != []:
            raise AssertionError("a table load must be the leaf of the task chain")
        if not self.alias:
            return self.file_schema
        return [(f"{self.alias}.{name}", col_type) for name, col_type in self.file_schema]

    def _describe(self) -> str:
        return f"LoadTableBlockTask({self.file_path})"


@dataclass(kw_only=True)
class LoadShuffleFilesTask(ProducerTask):
    def _describe(self) -> str:
        return "LoadShuffleFile()"


@dataclass(kw_only=True)
class ProjectTask(ConsumerTask):
    columns: list[Col]

    def validate_schema(self) -> Schema:
        schema = self.parent_task.validate_schema()
        expanded: list[Col] = []
        for col in self.columns:  # '*' expands to every input column
            if type(col) is Col and col.name == "*":
                expanded.extend(Col(name) for name, _ in schema)
            else:
                expanded.append(col)
        self.columns = expanded
        _require_known(self.columns, schema, "projection")
        return [(col.name, col.infer_type(schema)) for col in self.columns]

    def _describe(self) -> str:
        return f"Project({', '.join(str(c) for c in self.columns)})"


@dataclass(kw_only=True)
class FilterTask(ConsumerTask):
    condition: Col

    def __post_init__(self) -> None:
        if type(self.condition) not in (BinaryOperatorColumn, LikeColumn):
            raise AssertionError(type(self.condition))

    def validate_schema(self) -> Schema:
        schema = self.parent_task.validate_schema()
        self.condition.infer_type(schema)
        return schema

    def _describe(self) -> str:
        return f"Filter({self.condition})"


@dataclass(kw_only=True)
class BroadcastHashJoinTask(ProducerTask):
    """Equi-join of ``parent_task`` (left) with ``right_side_task``; always inner (tasks.py:230-239)."""

    right_side_task: Task
    join_condition: Col
    how: JoinType = "inner"
    left_key: Optional[Col] = None
    right_key: Optional[Col] = None
    left_schema: Optional[Schema] = None
    right_schema: Optional[Schema] = None

    def validate_schema(self) -> Schema:
        self.left_schema = self.parent_task.validate_schema()
        self.right_schema = self.right_side_task.validate_schema()
        _require_known([self.join_condition], self.left_schema + self.right_schema, "Join")
        if type(self.join_condition) is not BinaryOperatorColumn:
            raise AssertionError("Only equi-join is supported")
        self.left_key, self.right_key = self.join_condition.extract_left_right_key(
            self.left_schema, self.right_schema
        )
        return self.left_schema + self.right_schema

    def _describe(self) -> str:
        return f'Join({self.join_condition}, "{self.how}")'

    def explain(self, lvl: int = 0) -> None:
        indent = "  " * lvl + ("+- " if lvl > 0 else "")
        print(f"{indent} {self._describe()}:{nice_schema(self.inferred_schema)}")  # noqa: T201
        self.parent_task.explain(lvl + 1)
        self.right_side_task.explain(lvl + 1)


@dataclass(kw_only=True)
class AggregateTask(ConsumerTask):
    group_by_column: Col
    agg_columns: list[AggCol]
    before_shuffle: bool = True

    def validate_schema(self) -> Schema:
        schema = self.parent_task.validate_schema()
        if not self.before_shuffle:
            return schema
        _require_known([*self.agg_columns, self.group_by_column], schema, "aggregation")
        out = [(self.group_by_column.name, self.group_by_column.infer_type(schema))]
        out += [(agg.name, agg.infer_type(schema)) for agg in self.agg_columns]
        return out

    def _describe(self) -> str:
        return (f"AggregateTask(group_by: {self.group_by_column}, agg: {self.agg_columns}, "
                f"before_shuffle:{self.before_shuffle})")


@dataclass
class WriteToShufflePartitions(WriterTask):
    key_column: Optional[Col] = None

    def validate_schema(self) -> Schema:
        schema = self.parent_task.validate_schema()
        if self.key_column is None:
            return schema
        _require_known([self.key_column], schema, "GroupBy")
        if self.key_column.name in {name for name, _ in schema}:
            return schema
        return [(self.key_column.name, self.key_column.infer_type(schema)), *schema]

    def _describe(self) -> str:
        return f"WriteToShufflePartitions({self.key_column})"


@dataclass(kw_only=True)
class WriteToLocalFileTask(WriterTask):
    def _describe(self) -> str:
        return "WriteToLocalFileTask()"
