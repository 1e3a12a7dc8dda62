"""Type ordinals and sizing constants of the BlockFile / engine contract.

Mirrors the public names of the reference's ``src/mini_spark/constants.py`` (ordinals at
``constants.py:18-23`` are the on-disk contract; ``ROWS_PER_BLOCK`` at ``constants.py:7``;
``MAX_INT``/``MIN_INT`` at ``constants.py:14-15`` seed MIN/MAX aggregates, ``tasks.py:303-310``).
Written from scratch for the B200 engine: adds the device-side physical widths of each type.
"""

from __future__ import annotations

import enum
from datetime import datetime
from pathlib import Path
from typing import Union

ROWS_PER_BLOCK = 1 << 21  # 2 097 152 rows: unit of ingest staging and of multi-GPU sharding
SHUFFLE_PARTITIONS = 10  # kept for interface parity; the GPU exchange partitions by world size
GLOBAL_TEMP_FOLDER = Path("tmp/")
SHUFFLE_FOLDER = Path("shuffle/")

MAX_INT = (1 << 31) - 1
MIN_INT = -(1 << 31)


class ColumnType(enum.Enum):
    """Logical column types; ``ordinal`` is the byte stored in a BlockFile schema header."""

    INTEGER = (0, int)
    STRING = (1, str)
    FLOAT = (2, float)
    TIMESTAMP = (3, int)
    UNKNOWN = (255, type(None))

    def __init__(self, ordinal: int, py_type: type) -> None:
        self.ordinal = ordinal
        self.type = py_type

    @classmethod
    def from_ordinal(cls, ordinal: int) -> "ColumnType":
        found = _BY_ORDINAL.get(ordinal)
        if found is None:
            raise NotImplementedError(ordinal)
        return found

    @classmethod
    def of(cls, value: object) -> "ColumnType":
        return _BY_PYTYPE.get(type(value), cls.UNKNOWN)

    @property
    def disk_width(self) -> int:
        """Bytes per value on disk (STRING: the u8 length prefix only)."""
        return {0: 4, 1: 1, 2: 4, 3: 8}[self.ordinal]

    def __str__(self) -> str:
        return self.name

    __repr__ = __str__


_BY_ORDINAL = {t.ordinal: t for t in ColumnType}
_BY_PYTYPE = {int: ColumnType.INTEGER, str: ColumnType.STRING, float: ColumnType.FLOAT,
              datetime: ColumnType.TIMESTAMP}

ColumnTypePython = Union[int, float, str, datetime]
NumericColumnTypes = {int, float}
Row = dict
Columns = tuple
Schema = list
