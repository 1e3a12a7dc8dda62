"""Fluent ``DataFrame`` builder over the task tree (interface of ``src/mini_spark/dataframe.py``).

``collect()``/``show()`` call exactly ``engine.execute_full_task(task)`` then
``engine.collect_results(...)`` (reference ``dataframe.py:71-79``), which is the boundary the
``CudaExecutionEngine`` plugs into.  Unlike the reference, a DataFrame built without an engine
has *no* default engine: there is no CPU execution path in this framework.
"""

from __future__ import annotations

from copy import deepcopy
from pathlib import Path
from typing import TYPE_CHECKING

from tabulate import tabulate

from .tasks import (AggregateTask, BroadcastHashJoinTask, FilterTask, JoinType, LoadTableBlockTask,
                    ProjectTask, Task, VoidTask)

if TYPE_CHECKING:
    from .constants import Row, Schema
    from .execution import ExecutionEngine
    from .sql import AggCol, Col


class GroupedData:
    def __init__(self, df: "DataFrame", column: "Col") -> None:
        self.df = df
        self.group_column = column

    def agg(self, *agg_columns: "AggCol") -> "DataFrame":
        self.df.task = AggregateTask(self.df.task, group_by_column=self.group_column,
                                     agg_columns=list(agg_columns))
        return self.df


class DataFrame:
    def __init__(self, engine: "ExecutionEngine | None" = None) -> None:
        self.engine = engine
        self.task: Task = VoidTask()

    @property
    def schema(self) -> "Schema":
        return self.task.validate_schema()

    def table(self, file_path: str) -> "DataFrame":
        self.task = LoadTableBlockTask(self.task, file_path=Path(file_path))
        return self

    def alias(self, alias_name: str) -> "DataFrame":
        if type(self.task) is not LoadTableBlockTask:
            raise AssertionError("Alias can only be applied to table")
        self.task.alias = alias_name
        return self

    def select(self, *columns: "Col") -> "DataFrame":
        self.task = ProjectTask(self.task, columns=list(columns))
        return self

    def filter(self, column: "Col") -> "DataFrame":
        self.task = FilterTask(self.task, condition=column)
        return self

    def group_by(self, column: "Col") -> GroupedData:
        return GroupedData(self, column)

    def join(self, other_df: "DataFrame", on: "Col", how: JoinType) -> "DataFrame":
        self.task = BroadcastHashJoinTask(self.task, right_side_task=other_df.task, join_condition=on, how=how)
        return self

    def _engine(self) -> "ExecutionEngine":
        if self.engine is None:
            raise RuntimeError("DataFrame has no engine: construct it as DataFrame(CudaExecutionEngine())")
        return self.engine

    def collect(self) -> "list[Row]":
        engine = self._engine()
        return list(engine.collect_results(engine.execute_full_task(self.task)))

    def show(self, n: int = 10) -> int:
        engine = self._engine()
        rows = list(engine.collect_results(engine.execute_full_task(self.task), limit=n))
        print(tabulate(rows, tablefmt="rounded_outline", headers="keys"))  # noqa: T201
        return len(rows)

    def explain(self, *, full: bool = False) -> None:
        task = deepcopy(self.task)
        print("Logical Plan")  # noqa: T201
        task.validate_schema()
        task.explain()
        if full:
            from .lowering import lower_task  # noqa: PLC0415
            print("GPU plan")  # noqa: T201
            print(lower_task(deepcopy(self.task)).describe())  # noqa: T201
