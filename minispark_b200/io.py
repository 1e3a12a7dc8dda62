"""BlockFile: the columnar row-block file used for tables, results and (in the reference) shuffles.

Same public surface and byte format as the reference's ``src/mini_spark/io.py`` (format defined at
``io.py:47-60`` schema header, ``io.py:74-109`` blocks, ``io.py:217-229`` footer), re-implemented
from scratch on numpy so whole columns are encoded/decoded with one ``frombuffer``/``tobytes``
instead of one ``f.read`` per value (``io.py:129-149``).  The GPU engine never goes through this
module for ingest (that is ``csrc/ingest.cu``); it is the host mirror used by tests, the generator
and ``collect_results``.

File layout (all little endian)::

    u8 ncols { u8 type-ordinal, u8 namelen, name }*          schema header
    { u32 rows, { u64 nbytes, payload }* }*                  blocks, one payload per column
    { u64 block_start }*  u32 nblocks                        footer

Payloads: INTEGER i32, FLOAT f32, TIMESTAMP i64 microseconds (naive local time),
STRING ``rows`` u8 lengths followed by the concatenated bytes.
"""

from __future__ import annotations

import os
import struct
from datetime import datetime
from pathlib import Path
from typing import Any, BinaryIO, Iterable, Iterator, Sequence

import numpy as np

from . import constants as _c
from .constants import ColumnType, Columns, Row, Schema

ROWS_PER_BLOCK = _c.ROWS_PER_BLOCK  # module-level so tests can patch it like the reference's
MAX_COLUMNS = 0xFF
MAX_STR_LENGTH = 0xFF

_U32 = struct.Struct("<I")
_U64 = struct.Struct("<Q")
_NP = {ColumnType.INTEGER: np.dtype("<i4"), ColumnType.FLOAT: np.dtype("<f4"),
       ColumnType.TIMESTAMP: np.dtype("<i8")}


def datetime_to_timestamp(dt: datetime) -> int:
    """Microseconds since the epoch of a naive-local datetime (reference ``io.py:34-35``)."""
    return int(dt.timestamp() * 1_000_000)


def timestamp_to_datetime(microseconds_since_epoch: int) -> datetime:
    """Inverse of :func:`datetime_to_timestamp` (reference ``io.py:38-39``)."""
    return datetime.fromtimestamp(microseconds_since_epoch / 1_000_000)


def encode_string(text: str) -> bytes:
    if len(text) >= MAX_STR_LENGTH:
        raise AssertionError(f"string longer than {MAX_STR_LENGTH - 1} chars")
    return bytes((len(text) & 0xFF,)) + text.encode("utf-8")


# ----------------------------------------------------------------------------- schema header
def schema_to_bytes(schema: Schema) -> bytes:
    if len(schema) >= MAX_COLUMNS:
        raise AssertionError("too many columns")
    out = bytearray((len(schema),))
    for name, col_type in schema:
        out.append(col_type.ordinal & 0xFF)
        out += encode_string(name)
    return bytes(out)


def _serialize_schema(schema: Schema, f: BinaryIO) -> None:
    f.write(schema_to_bytes(schema))


def _deserialize_schema(f: BinaryIO) -> Schema:
    ncols = f.read(1)[0]
    schema: Schema = []
    for _ in range(ncols):
        ordinal, name_len = f.read(2)
        schema.append((f.read(name_len).decode("utf-8"), ColumnType.from_ordinal(ordinal)))
    return schema


def _deserialize_block_starts(f: BinaryIO) -> list[int]:
    f.seek(-4, os.SEEK_END)
    (nblocks,) = _U32.unpack(f.read(4))
    if nblocks == 0:
        return []
    f.seek(-4 - 8 * nblocks, os.SEEK_END)
    return [int(v) for v in np.frombuffer(f.read(8 * nblocks), dtype="<u8")]


# ----------------------------------------------------------------------------- column payloads
def encode_column(col_type: ColumnType, values: Any) -> bytes:
    """One column payload of one block.  ``values`` is a Python sequence or a numpy array."""
    if col_type == ColumnType.INTEGER:
        if isinstance(values, np.ndarray) and values.dtype.kind in "iu":
            arr = values
        else:
            for v in values:
                if type(v) is not int:
                    raise AssertionError(f"INTEGER column holds {type(v).__name__}")
            arr = np.array(values, dtype=object) if len(values) else np.zeros(0, np.int64)
        if len(arr) and (max(arr) > _c.MAX_INT or min(arr) < _c.MIN_INT):
            raise OverflowError("int too big to convert")  # what int.to_bytes raises (io.py:90)
        return np.asarray(arr, dtype="<i4").tobytes()
    if col_type == ColumnType.FLOAT:
        if not isinstance(values, np.ndarray):
            for v in values:
                if type(v) is not float:
                    raise AssertionError(f"FLOAT column holds {type(v).__name__}")
        with np.errstate(over="ignore"):
            return np.asarray(values, dtype=np.float64).astype("<f4").tobytes()
    if col_type == ColumnType.TIMESTAMP:
        if isinstance(values, np.ndarray) and values.dtype.kind in "iu":
            return np.asarray(values, dtype="<i8").tobytes()
        micros = []
        for v in values:
            if type(v) is str:
                v = datetime.fromisoformat(v)
            if type(v) is not datetime:
                raise AssertionError(f"TIMESTAMP column holds {type(v).__name__}")
            micros.append(datetime_to_timestamp(v))
        return np.asarray(micros, dtype="<i8").tobytes()
    if col_type == ColumnType.STRING:
        if isinstance(values, tuple) and len(values) == 2 and isinstance(values[0], np.ndarray):
            lens, raw = values  # pre-encoded (lens u8, bytes) pair from the generator
            return np.asarray(lens, dtype=np.uint8).tobytes() + bytes(raw)
        for v in values:
            if type(v) is not str:
                raise AssertionError(f"STRING column holds {type(v).__name__}")
        lens = bytes(len(v) & 0xFF for v in values)
        return lens + "".join(values).encode("utf-8")
    raise ValueError(f"Unsupported column type {col_type}")


def decode_column(col_type: ColumnType, payload: memoryview | bytes, rows: int, *, raw: bool = False) -> Any:
    """Inverse of :func:`encode_column`.  ``raw=True`` keeps numpy arrays (strings as lens+bytes)."""
    if col_type in _NP:
        arr = np.frombuffer(payload, dtype=_NP[col_type], count=rows)
        if raw:
            return arr
        if col_type == ColumnType.TIMESTAMP:
            return [timestamp_to_datetime(int(v)) for v in arr]
        if col_type == ColumnType.FLOAT:
            return [float(v) for v in arr]
        return [int(v) for v in arr]
    if col_type == ColumnType.STRING:
        lens = np.frombuffer(payload, dtype=np.uint8, count=rows)
        body = bytes(payload[rows:])
        if raw:
            return lens, body
        ends = np.cumsum(lens, dtype=np.int64)
        out, start = [], 0
        for end in ends.tolist():
            out.append(body[start:end].decode("utf-8"))
            start = end
        return out
    raise ValueError(f"Unsupported column type {col_type}")


def _column_len(col: Any) -> int:
    if isinstance(col, tuple) and len(col) == 2 and isinstance(col[0], np.ndarray):
        return len(col[0])
    return len(col)


def _slice_column(col: Any, lo: int, hi: int) -> Any:
    if isinstance(col, tuple) and len(col) == 2 and isinstance(col[0], np.ndarray):
        lens, raw = col
        offs = np.concatenate(([0], np.cumsum(lens, dtype=np.int64)))
        return lens[lo:hi], raw[int(offs[lo]):int(offs[min(hi, len(lens))])]
    return col[lo:hi]


def _generate_data_blocks_for_columns(schema: Schema, columns: Columns) -> Iterator[bytes]:
    if len(columns) == 0 or _column_len(columns[0]) == 0:
        return
    if len(columns) != len(schema):
        raise ValueError("zip() argument mismatch: columns vs schema")
    total = _column_len(columns[0])
    for lo in range(0, total, ROWS_PER_BLOCK):
        hi = min(lo + ROWS_PER_BLOCK, total)
        parts = [_U32.pack(hi - lo)]
        for col, (_, col_type) in zip(columns, schema):
            payload = encode_column(col_type, _slice_column(col, lo, hi))
            parts.append(_U64.pack(len(payload)))
            parts.append(payload)
        yield b"".join(parts)


def _read_block_layout(f: BinaryIO, ncols: int) -> tuple[int, list[tuple[int, int]]]:
    """(rows, [(payload_offset, payload_nbytes)]) of the block starting at ``f.tell()``."""
    (rows,) = _U32.unpack(f.read(4))
    layout = []
    for _ in range(ncols):
        (nbytes,) = _U64.unpack(f.read(8))
        layout.append((f.tell(), nbytes))
        f.seek(nbytes, os.SEEK_CUR)
    return rows, layout


def _deserialize_block(f: BinaryIO, schema: Schema, *, raw: bool = False, only: Sequence[int] | None = None) -> Columns:
    rows, layout = _read_block_layout(f, len(schema))
    out = []
    for i, ((_, col_type), (off, nbytes)) in enumerate(zip(schema, layout)):
        if only is not None and i not in only:
            out.append(None)
            continue
        f.seek(off)
        out.append(decode_column(col_type, f.read(nbytes), rows, raw=raw))
    return tuple(out)


def merge_data_blocks(data_block_1: Columns, data_block_2: Columns) -> Columns:
    if len(data_block_1) != len(data_block_2):
        raise ValueError("column count mismatch")
    return tuple(list(a) + list(b) for a, b in zip(data_block_1, data_block_2))


_SCHEMA_CACHE: dict[tuple, Schema] = {}


class BlockFile:
    """Reader/writer with the reference's method names (``io.py:180-313``)."""

    def __init__(self, file: Path | str, schema: Schema | None = None) -> None:
        self.file = Path(file)
        self.schema: Schema = list(schema) if schema else []
        self._block_starts: list[int] | None = None
        self._file_schema: Schema | None = None

    def __repr__(self) -> str:
        return f"BlockFile(file={self.file!r}, schema={self.schema!r})"

    def __eq__(self, other: object) -> bool:
        return isinstance(other, BlockFile) and (self.file, self.schema) == (other.file, other.schema)

    __hash__ = None  # type: ignore[assignment]

    # -- metadata ---------------------------------------------------------------------------
    @property
    def block_starts(self) -> list[int]:
        if self._block_starts is None:
            with self.file.open("rb") as f:
                self._block_starts = _deserialize_block_starts(f)
        return self._block_starts

    @property
    def file_schema(self) -> Schema:
        if self._file_schema is None:
            # planning asks for a table's schema several times per query (validate_schema, the lowering, the engine), each
            # time through a fresh BlockFile object: remember the header per (path, mtime, size)
            try:
                st = self.file.stat()
                key = (str(self.file), st.st_mtime_ns, st.st_size)
            except OSError:
                key = None
            cached = _SCHEMA_CACHE.get(key) if key else None
            if cached is None:
                with self.file.open("rb") as f:
                    cached = _deserialize_schema(f)
                if key:
                    if len(_SCHEMA_CACHE) > 256:
                        _SCHEMA_CACHE.clear()
                    _SCHEMA_CACHE[key] = cached
            self._file_schema = list(cached)
        return self._file_schema

    def rows(self) -> int:
        total = 0
        with self.file.open("rb") as f:
            for start in self.block_starts:
                f.seek(start)
                total += _U32.unpack(f.read(4))[0]
        return total

    def block_layout(self, block_id: int) -> tuple[int, list[tuple[int, int]]]:
        """Rows and per-column (file offset, nbytes) of one block: what the ingest path preads."""
        with self.file.open("rb") as f:
            f.seek(self.block_starts[block_id])
            return _read_block_layout(f, len(self.file_schema))

    # -- writing ----------------------------------------------------------------------------
    def write_data(self, data: Columns) -> "BlockFile":
        return self._write_data_with_known_schema(data, self.schema)

    def write_tuples(self, tuples: list[tuple[Any, ...]]) -> "BlockFile":
        return self.write_data(_transpose(tuples))

    def write_rows(self, data: list[Row]) -> "BlockFile":
        if not data:
            if self.schema:
                with self.file.open("wb") as f:
                    _serialize_schema(self.schema, f)
                    f.write(_U32.pack(0))
                self._invalidate()
            return self
        self.schema = [(key, ColumnType.of(value)) for key, value in data[0].items()]
        return self.write_data(tuple([row[name] for row in data] for name, _ in self.schema))

    def _write_data_with_known_schema(self, columns_data: Columns, schema: Schema) -> "BlockFile":
        if not schema:
            raise AssertionError("BlockFile needs a schema to write")
        self._invalidate()
        with self.file.open("wb") as f:
            _serialize_schema(schema, f)
            self._write_blocks_and_footer(f, [], schema, columns_data)
        return self

    @staticmethod
    def _write_blocks_and_footer(f: BinaryIO, starts: list[int], schema: Schema, data: Columns) -> None:
        for block in _generate_data_blocks_for_columns(schema, data):
            starts.append(f.tell())
            f.write(block)
        f.write(np.asarray(starts, dtype="<u8").tobytes())
        f.write(_U32.pack(len(starts)))
        f.truncate()

    def append_data(self, data: Columns) -> "BlockFile":
        """Append rows; a partially filled last block is re-packed (reference ``io.py:231-252``)."""
        self._invalidate()
        if not self.file.exists() or not self.block_starts:
            return self.write_data(data)
        starts = list(self.block_starts)
        schema = self.file_schema
        if self.schema != schema:
            raise AssertionError((self.file, self.schema, schema))
        with self.file.open("rb+") as f:
            f.seek(starts[-1])
            last = _deserialize_block(f, schema)
            if len(last[0]) < ROWS_PER_BLOCK:
                data = merge_data_blocks(last, data)
                f.seek(starts.pop())
            else:
                f.seek(-(8 * len(starts) + 4), os.SEEK_END)
            self._write_blocks_and_footer(f, starts, schema, data)
        self._invalidate()
        return self

    def append_tuples(self, data: list[tuple[Any, ...]]) -> "BlockFile":
        return self.append_data(_transpose(data))

    def append_rows(self, data: list[Row]) -> "BlockFile":
        if not self.schema:
            raise AssertionError("append_rows needs a schema")
        names = list(data[0].keys())
        return self.append_data(tuple([row[name] for row in data] for name in names))

    def merge_files(self, files: list[Path]) -> "BlockFile":
        self.schema = BlockFile(files[0]).file_schema
        for path in files:
            other = BlockFile(path)
            if other.file_schema != self.schema:
                raise AssertionError("schema mismatch in merge_files")
            for block in other.read_block_data_columns_sequentially():
                self.append_data(block)
        return self

    # -- reading ----------------------------------------------------------------------------
    def read_block_data_columns_by_id(self, block_id: int, f: BinaryIO | None = None) -> Columns:
        if f is not None:
            f.seek(self.block_starts[block_id])
            return _deserialize_block(f, self.file_schema)
        with self.file.open("rb") as fh:
            fh.seek(self.block_starts[block_id])
            return _deserialize_block(fh, self.file_schema)

    def read_block_arrays(self, block_id: int, columns: Sequence[str] | None = None) -> dict[str, Any]:
        """numpy view of one block: ``{name: array}``; STRING columns as ``(lens u8, bytes)``."""
        schema = self.file_schema
        names = [n for n, _ in schema]
        only = None if columns is None else [names.index(c) for c in columns]
        with self.file.open("rb") as f:
            f.seek(self.block_starts[block_id])
            cols = _deserialize_block(f, schema, raw=True, only=only)
        return {n: c for n, c in zip(names, cols) if c is not None}

    def read_block_data(self, block_id: int) -> list[tuple[Any, ...]]:
        return list(zip(*self.read_block_data_columns_by_id(block_id)))

    def read_block_data_columns_sequentially(self) -> Iterable[Columns]:
        schema = self.file_schema
        with self.file.open("rb") as f:
            for start in self.block_starts:
                f.seek(start)
                yield _deserialize_block(f, schema)

    def read_blocks_sequentially(self) -> Iterable[list[Row]]:
        names = [n for n, _ in self.file_schema]
        for block in self.read_block_data_columns_sequentially():
            yield [dict(zip(names, values)) for values in zip(*block)]

    def read_data_rows(self) -> Iterable[Row]:
        for block in self.read_blocks_sequentially():
            yield from block

    def _invalidate(self) -> None:
        self._block_starts = None
        self._file_schema = None


def _transpose(tuples: Sequence[Sequence[Any]]) -> Columns:
    if not tuples:
        return ()
    width = len(tuples[0])
    for t in tuples:
        if len(t) != width:
            raise ValueError("ragged tuples")
    return tuple([t[i] for t in tuples] for i in range(width))
