"""Deterministic TPC-H-shaped synthetic data (lineitem / orders) written directly as BlockFiles.

There is no network and no dbgen here, so the benchmark tables are generated.  Shapes follow
SURVEY.md section 8(d): lineitem has the 16 columns of the reference's benchmark schema
(``examples/benchmark.py:24-41``), rows are clustered by order key, about 4 lines per order
(about 6.0 M rows per scale factor), dates at midnight UTC in microseconds, FLOAT stored as f32.

Generation is block-parallel and seed-stable: per-order attributes come from a counter-based hash
of the order index, per-line attributes from ``numpy.random.Generator(PCG64(seed + block_id))``.
``--columns`` restricts the file to the listed columns (Q1 needs 6 of the 16).

    python bench/gen_tpch.py --table lineitem --sf 1 --out /tmp/lineitem_sf1.bin
"""

from __future__ import annotations

import argparse
import io
import struct
import sys
from pathlib import Path
from typing import BinaryIO, Iterator, Optional, Sequence

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minispark_b200.constants import ColumnType  # noqa: E402

INT, STR, FLOAT, DATE = ColumnType.INTEGER, ColumnType.STRING, ColumnType.FLOAT, ColumnType.TIMESTAMP
ROWS_PER_BLOCK = 1 << 21
DAY_US = 86_400_000_000

LINEITEM_SCHEMA = [
    ("l_orderkey", INT), ("l_partkey", INT), ("l_suppkey", INT), ("l_linenumber", INT),
    ("l_quantity", FLOAT), ("l_extendedprice", FLOAT), ("l_discount", FLOAT), ("l_tax", FLOAT),
    ("l_returnflag", STR), ("l_linestatus", STR), ("l_shipdate", DATE), ("l_commitdate", DATE),
    ("l_receiptdate", DATE), ("l_shipinstruct", STR), ("l_shipmode", STR), ("l_comment", STR),
]
ORDERS_SCHEMA = [
    ("o_orderkey", INT), ("o_custkey", INT), ("o_orderstatus", STR), ("o_totalprice", FLOAT),
    ("o_orderdate", DATE), ("o_orderpriority", STR), ("o_clerk", STR), ("o_shippriority", INT),
    ("o_comment", STR),
]
Q1_COLUMNS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_shipdate"]

SHIPINSTRUCT = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]
SHIPMODE = ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"]
PRIORITY = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]


def _days(y: int, m: int, d: int) -> int:
    return int(np.datetime64(f"{y:04d}-{m:02d}-{d:02d}").astype("datetime64[D]").astype(np.int64))


D_START = _days(1992, 1, 1)
D_END = _days(1998, 8, 2)
D_CUTOFF = _days(1995, 6, 17)


def splitmix64(x: np.ndarray, salt: int) -> np.ndarray:
    """Counter-based hash: the same order index gives the same attributes in every block."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64((0x9E3779B97F4A7C15 * (salt + 1)) & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def n_orders(sf: float) -> int:
    return max(1, int(round(1_500_000 * sf)))


def order_key(o: np.ndarray) -> np.ndarray:
    return ((o // 8) * 32 + o % 8 + 1).astype(np.int64)  # TPC-H sparse order keys


def order_date_days(o: np.ndarray, seed: int) -> np.ndarray:
    return (D_START + (splitmix64(o, seed + 11) % np.uint64(D_END - D_START + 1)).astype(np.int64)).astype(np.int64)


def lines_per_order(sf: float, seed: int) -> np.ndarray:
    o = np.arange(n_orders(sf), dtype=np.uint64)
    return (1 + (splitmix64(o, seed + 7) % np.uint64(7))).astype(np.int64)


def strings_from_codes(codes: np.ndarray, values: Sequence[str]) -> tuple[np.ndarray, bytes]:
    """(u8 lengths, concatenated bytes) of ``values[codes]`` without a Python loop over rows."""
    enc = [v.encode("ascii") for v in values]
    lens_table = np.array([len(b) for b in enc], dtype=np.int64)
    lens = lens_table[codes]
    ends = np.cumsum(lens)
    starts = ends - lens
    out = np.empty(int(ends[-1]) if len(ends) else 0, dtype=np.uint8)
    for code, b in enumerate(enc):
        idx = np.nonzero(codes == code)[0]
        if len(idx) == 0 or not b:
            continue
        pos = starts[idx][:, None] + np.arange(len(b))[None, :]
        out[pos] = np.frombuffer(b, dtype=np.uint8)[None, :]
    return lens.astype(np.uint8), out.tobytes()


def random_text(rng: np.random.Generator, n: int, lo: int, hi: int) -> tuple[np.ndarray, bytes]:
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz      ", dtype=np.uint8)
    lens = rng.integers(lo, hi + 1, size=n, dtype=np.int64)
    body = alphabet[rng.integers(0, len(alphabet), size=int(lens.sum()), dtype=np.int64)]
    return lens.astype(np.uint8), body.tobytes()


def lineitem_block(sf: float, block_id: int, row_lo: int, row_hi: int, cum_lines: np.ndarray, seed: int,
                   columns: Sequence[str]) -> dict[str, object]:
    """Columns of lineitem rows [row_lo, row_hi)."""
    n = row_hi - row_lo
    rng = np.random.default_rng(np.random.PCG64(seed + block_id))
    rows = np.arange(row_lo, row_hi, dtype=np.int64)
    o = np.searchsorted(cum_lines, rows, side="right").astype(np.int64)  # order index of each row
    first_row = np.where(o > 0, cum_lines[np.maximum(o - 1, 0)], 0)
    want = set(columns)
    out: dict[str, object] = {}
    odate = order_date_days(o, seed)
    ship = odate + rng.integers(1, 122, size=n)
    commit = odate + rng.integers(30, 91, size=n)
    receipt = ship + rng.integers(1, 31, size=n)
    partkey = rng.integers(1, int(200_000 * max(sf, 0.005)) + 1, size=n)
    qty = rng.integers(1, 51, size=n)
    if "l_orderkey" in want:
        out["l_orderkey"] = order_key(o).astype(np.int32)
    if "l_partkey" in want:
        out["l_partkey"] = partkey.astype(np.int32)
    if "l_suppkey" in want:
        out["l_suppkey"] = rng.integers(1, int(10_000 * max(sf, 0.005)) + 1, size=n).astype(np.int32)
    if "l_linenumber" in want:
        out["l_linenumber"] = (rows - first_row + 1).astype(np.int32)
    if "l_quantity" in want:
        out["l_quantity"] = qty.astype(np.float32)
    if "l_extendedprice" in want:
        retail = (90000 + (partkey // 10) % 20001 + 100 * (partkey % 1000)) / 100.0
        out["l_extendedprice"] = (qty * retail).astype(np.float32)
    if "l_discount" in want:
        out["l_discount"] = (rng.integers(0, 11, size=n) / 100.0).astype(np.float32)
    if "l_tax" in want:
        out["l_tax"] = (rng.integers(0, 9, size=n) / 100.0).astype(np.float32)
    if "l_returnflag" in want:
        ra = rng.integers(0, 2, size=n)
        codes = np.where(receipt <= D_CUTOFF, ra, 2)
        out["l_returnflag"] = strings_from_codes(codes, ["R", "A", "N"])
    if "l_linestatus" in want:
        out["l_linestatus"] = strings_from_codes((ship > D_CUTOFF).astype(np.int64), ["F", "O"])
    if "l_shipdate" in want:
        out["l_shipdate"] = ship * DAY_US
    if "l_commitdate" in want:
        out["l_commitdate"] = commit * DAY_US
    if "l_receiptdate" in want:
        out["l_receiptdate"] = receipt * DAY_US
    if "l_shipinstruct" in want:
        out["l_shipinstruct"] = strings_from_codes(rng.integers(0, 4, size=n), SHIPINSTRUCT)
    if "l_shipmode" in want:
        out["l_shipmode"] = strings_from_codes(rng.integers(0, 7, size=n), SHIPMODE)
    if "l_comment" in want:
        out["l_comment"] = random_text(rng, n, 10, 43)
    return out


def orders_block(sf: float, block_id: int, row_lo: int, row_hi: int, seed: int, columns: Sequence[str]) -> dict[str, object]:
    n = row_hi - row_lo
    rng = np.random.default_rng(np.random.PCG64(seed + 100_000 + block_id))
    o = np.arange(row_lo, row_hi, dtype=np.int64)
    want = set(columns)
    out: dict[str, object] = {}
    if "o_orderkey" in want:
        out["o_orderkey"] = order_key(o).astype(np.int32)
    if "o_custkey" in want:
        out["o_custkey"] = rng.integers(1, int(150_000 * max(sf, 0.005)) + 1, size=n).astype(np.int32)
    if "o_orderstatus" in want:
        out["o_orderstatus"] = strings_from_codes(rng.integers(0, 3, size=n), ["O", "F", "P"])
    if "o_totalprice" in want:
        out["o_totalprice"] = (rng.integers(90_000, 50_000_000, size=n) / 100.0).astype(np.float32)
    if "o_orderdate" in want:
        out["o_orderdate"] = order_date_days(o, seed) * DAY_US
    if "o_orderpriority" in want:
        out["o_orderpriority"] = strings_from_codes(rng.integers(0, 5, size=n), PRIORITY)
    if "o_clerk" in want:
        clerk = rng.integers(1, int(1000 * max(sf, 0.01)) + 1, size=n)
        text = np.char.add("Clerk#", np.char.zfill(clerk.astype(str), 9))
        blob = "".join(text.tolist()).encode("ascii")
        out["o_clerk"] = (np.full(n, 15, dtype=np.uint8), blob)
    if "o_shippriority" in want:
        out["o_shippriority"] = np.zeros(n, dtype=np.int32)
    if "o_comment" in want:
        out["o_comment"] = random_text(rng, n, 19, 78)
    return out


def _payload(ctype: ColumnType, value: object) -> bytes:
    if ctype == STR:
        lens, body = value  # type: ignore[misc]
        return np.asarray(lens, dtype=np.uint8).tobytes() + bytes(body)
    dtype = {INT: "<i4", FLOAT: "<f4", DATE: "<i8"}[ctype]
    return np.ascontiguousarray(value, dtype=dtype).tobytes()


def table_blocks(table: str, sf: float, columns: Optional[Sequence[str]] = None, rows_per_block: int = ROWS_PER_BLOCK,
                 seed: int = 1234, max_rows: Optional[int] = None) -> tuple[list[tuple[str, ColumnType]], int, Iterator[tuple[int, dict]]]:
    """(schema, total rows, iterator of (rows, columns) per block)."""
    full = LINEITEM_SCHEMA if table == "lineitem" else ORDERS_SCHEMA
    names = [n for n, _ in full] if columns is None else list(columns)
    schema = [(n, t) for n, t in full if n in names]
    names = [n for n, _ in schema]
    if table == "lineitem":
        cum = np.cumsum(lines_per_order(sf, seed))
        total = int(cum[-1])
    else:
        cum = None
        total = n_orders(sf)
    if max_rows is not None:
        total = min(total, max_rows)

    def blocks() -> Iterator[tuple[int, dict]]:
        for block_id, lo in enumerate(range(0, total, rows_per_block)):
            hi = min(lo + rows_per_block, total)
            if table == "lineitem":
                yield hi - lo, lineitem_block(sf, block_id, lo, hi, cum, seed, names)
            else:
                yield hi - lo, orders_block(sf, block_id, lo, hi, seed, names)

    return schema, total, blocks()


def write_stream(f: BinaryIO, schema: list[tuple[str, ColumnType]], blocks: Iterator[tuple[int, dict]]) -> int:
    """Serialise in BlockFile format (reference io.py:47-60,74-109,217-229); returns rows written."""
    header = bytearray((len(schema),))
    for name, ctype in schema:
        raw = name.encode("utf-8")
        header += bytes((ctype.ordinal, len(raw))) + raw
    f.write(header)
    pos = len(header)
    starts = []
    total = 0
    for rows, cols in blocks:
        starts.append(pos)
        f.write(struct.pack("<I", rows))
        pos += 4
        for name, ctype in schema:
            payload = _payload(ctype, cols[name])
            f.write(struct.pack("<Q", len(payload)))
            f.write(payload)
            pos += 8 + len(payload)
        total += rows
    f.write(np.asarray(starts, dtype="<u8").tobytes())
    f.write(struct.pack("<I", len(starts)))
    return total


def _block_payloads(args: tuple) -> tuple[int, list[bytes]]:
    """Worker of the parallel writer: one block's column payloads, serialised (generation is seed-stable per block)."""
    table, sf, names, types, block_id, lo, hi, seed = args
    if table == "lineitem":
        cum = _block_payloads.cum  # type: ignore[attr-defined]
        cols = lineitem_block(sf, block_id, lo, hi, cum, seed, names)
    else:
        cols = orders_block(sf, block_id, lo, hi, seed, names)
    return hi - lo, [_payload(t, cols[n]) for n, t in zip(names, types)]


def _init_worker(table: str, sf: float, seed: int) -> None:
    _block_payloads.cum = np.cumsum(lines_per_order(sf, seed)) if table == "lineitem" else None  # type: ignore[attr-defined]


def write_table(path: Path | str, table: str = "lineitem", sf: float = 1.0, columns: Optional[Sequence[str]] = None,
                rows_per_block: int = ROWS_PER_BLOCK, seed: int = 1234, max_rows: Optional[int] = None, workers: int = 1) -> int:
    """Write the table as a BlockFile.  ``workers`` > 1 generates the row-blocks in that many processes (the bytes are the
    same: every block has its own seed) and writes them in order."""
    schema, total, blocks = table_blocks(table, sf, columns, rows_per_block, seed, max_rows)
    nblocks = -(-total // rows_per_block)
    if workers <= 1 or nblocks < 4:
        with open(path, "wb") as f:
            return write_stream(f, schema, blocks)
    import multiprocessing as mp

    names, types = [n for n, _ in schema], [t for _, t in schema]
    jobs = [(table, sf, names, types, b, lo, min(lo + rows_per_block, total), seed) for b, lo in enumerate(range(0, total, rows_per_block))]
    with open(path, "wb") as f, mp.get_context("fork").Pool(min(workers, nblocks), initializer=_init_worker, initargs=(table, sf, seed)) as pool:
        header = bytearray((len(schema),))
        for name, ctype in schema:
            raw = name.encode("utf-8")
            header += bytes((ctype.ordinal, len(raw))) + raw
        f.write(header)
        pos, starts, rows_total = len(header), [], 0
        for rows, payloads in pool.imap(_block_payloads, jobs):
            starts.append(pos)
            f.write(struct.pack("<I", rows))
            pos += 4
            for payload in payloads:
                f.write(struct.pack("<Q", len(payload)))
                f.write(payload)
                pos += 8 + len(payload)
            rows_total += rows
        f.write(np.asarray(starts, dtype="<u8").tobytes())
        f.write(struct.pack("<I", len(starts)))
    return rows_total


def table_image(table: str = "lineitem", sf: float = 1.0, columns: Optional[Sequence[str]] = None,
                rows_per_block: int = ROWS_PER_BLOCK, seed: int = 1234, max_rows: Optional[int] = None) -> tuple[bytes, int]:
    schema, _, blocks = table_blocks(table, sf, columns, rows_per_block, seed, max_rows)
    buf = io.BytesIO()
    rows = write_stream(buf, schema, blocks)
    return buf.getvalue(), rows


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--table", choices=["lineitem", "orders"], default="lineitem")
    ap.add_argument("--sf", type=float, default=1.0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--columns", default=None, help="comma separated subset, or 'q1'")
    ap.add_argument("--rows-per-block", type=int, default=ROWS_PER_BLOCK)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--max-rows", type=int, default=None)
    ap.add_argument("--workers", type=int, default=1)
    args = ap.parse_args()
    columns = None
    if args.columns:
        columns = Q1_COLUMNS if args.columns == "q1" else args.columns.split(",")
    rows = write_table(args.out, args.table, args.sf, columns, args.rows_per_block, args.seed, args.max_rows, args.workers)
    print(f"{args.out}: {rows} rows")


if __name__ == "__main__":
    main()
