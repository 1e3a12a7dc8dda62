"""The shape signature under which a compiled scan is reused (execution._CompiledScan.signature_of): equal for sources that differ
only in row count and pointers, different as soon as a column's physical type, access path, dictionary (identity or size), the
probe form or the translation targets change.  CPU only: the signature is plain host logic."""

from __future__ import annotations

from minispark_b200 import native as N
from minispark_b200.execution import DeviceColumn, _CompiledScan, _Source


class FakeDict:
    _next = 1000

    def __init__(self, size: int) -> None:
        FakeDict._next += 1
        self.serial, self.size = FakeDict._next, size


def _source(nrows: int, base: int, d: FakeDict, phys=N.P_F32) -> _Source:
    cols = {0: DeviceColumn(base, N.P_I32, "I"), 1: DeviceColumn(base + 4096, phys, "F"), 2: DeviceColumn(base + 8192, N.P_U8, "S", d)}
    return _Source(nrows, cols)


def test_rows_and_pointers_do_not_matter():
    d = FakeDict(7)
    a, _ = _CompiledScan.signature_of(_source(100, 1 << 20, d), None)
    b, pins = _CompiledScan.signature_of(_source(5_000_000, 9 << 20, d), None)
    assert a == b and pins == [d]


def test_shape_changes_are_seen():
    d = FakeDict(7)
    base, _ = _CompiledScan.signature_of(_source(100, 1 << 20, d), None)
    assert _CompiledScan.signature_of(_source(100, 1 << 20, d, phys=N.P_F64), None)[0] != base      # physical type
    assert _CompiledScan.signature_of(_source(100, 1 << 20, FakeDict(7)), None)[0] != base           # another dictionary of the same size
    grown = FakeDict(7)
    before, _ = _CompiledScan.signature_of(_source(100, 1 << 20, grown), None)
    grown.size = 8                                                                                # the same dictionary, grown: its lookup tables are stale
    assert _CompiledScan.signature_of(_source(100, 1 << 20, grown), None)[0] != before
    via = _source(100, 1 << 20, d)
    via.columns[1] = DeviceColumn(via.columns[1].ptr, N.P_F32, "F", None, via=0)                  # read through an index vector
    via.index_vectors = [DeviceColumn(1 << 30, N.P_U32, "I")]
    assert _CompiledScan.signature_of(via, None)[0] != base
    probing = _source(100, 1 << 20, d)
    probing.probe_table, probing.probe_compact = 123456, True                                     # a join probe inside the scan
    assert _CompiledScan.signature_of(probing, None)[0] != base
    table = _source(100, 1 << 20, d)
    table.table_columns = True                                                                    # the library may keep derived data
    assert _CompiledScan.signature_of(table, None)[0] != base
    target = FakeDict(11)
    with_target, pins = _CompiledScan.signature_of(_source(100, 1 << 20, d), {"join": target})
    assert with_target != base and target in pins
