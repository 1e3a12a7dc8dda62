"""Tracing: a Perfetto trace of what the engine did, like the reference's ``TRACER`` (``src/mini_spark/utils.py:85-166``).

The reference records slices ("execute full task", "Execution", "Stage i", per-worker "job" tracks) through the
``perfetto`` package and ``TRACER.save("trace.pftrace")`` (``examples/benchmark.py:72``) writes a file the Perfetto UI
opens.  That package is not available here, so the few protobuf messages involved are encoded by hand (they are stable,
public wire format): ``Trace.packet = 1``; ``TracePacket`` {timestamp = 8, trusted_packet_sequence_id = 10,
track_event = 11, track_descriptor = 60}; ``TrackEvent`` {type = 9, track_uuid = 11, name = 23};
``TrackDescriptor`` {uuid = 1, name = 2, parent_uuid = 5}.

Same surface as the reference: ``TRACER.new_track / define_custom_track / start / end / save`` and the ``@trace(name)``
decorator.  Added for the GPU engine: one child track per GPU ("GPU 0 (device time)") whose slices carry the CUDA-event
durations the library reports (``device_slice``).  Recording is off until ``TRACER.enable()`` or ``MINISPARK_TRACE=1``:
the hot path of a sub-millisecond query should not pay for it.
"""

from __future__ import annotations

import functools
import os
import time
from pathlib import Path
from typing import Any, Callable, TypeVar

F = TypeVar("F", bound=Callable[..., Any])

MAIN_SYSTEM_TRACK_UUID = 1
TRUSTED_PACKET_SEQUENCE_ID = 1
TYPE_SLICE_BEGIN = 1
TYPE_SLICE_END = 2


def _varint(value: int) -> bytes:
    value &= 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while True:
        byte = value & 0x7F
        value >>= 7
        if value:
            out.append(byte | 0x80)
        else:
            out.append(byte)
            return bytes(out)


def _field_varint(number: int, value: int) -> bytes:
    return _varint(number << 3) + _varint(value)


def _field_bytes(number: int, payload: bytes) -> bytes:
    return _varint((number << 3) | 2) + _varint(len(payload)) + payload


class Tracer:
    def __init__(self) -> None:
        self.packets: list[bytes] = []
        self.tracks: set[int] = set()
        self.enabled = os.environ.get("MINISPARK_TRACE", "0") not in ("", "0")
        self.define_custom_track(MAIN_SYSTEM_TRACK_UUID, "Main System")

    def enable(self, on: bool = True) -> None:
        self.enabled = on

    # ---- tracks (reference utils.py:91-105) ----------------------------------------------------------------------
    def new_track(self, name: str, parent_track_uuid: int = MAIN_SYSTEM_TRACK_UUID) -> int:
        track_uuid = sum((i + 1) * b for i, b in enumerate(name.encode("utf-8"))) % 100000 + 1000  # (stable across processes)
        self.define_custom_track(track_uuid, name, parent_track_uuid)
        return track_uuid

    def define_custom_track(self, track_uuid: int, name: str, parent_track_uuid: int | None = None) -> None:
        if track_uuid in self.tracks:
            return
        desc = _field_varint(1, track_uuid) + _field_bytes(2, name.encode("utf-8"))
        if parent_track_uuid:
            desc += _field_varint(5, parent_track_uuid)
        self.packets.append(_field_bytes(60, desc))
        self.tracks.add(track_uuid)

    # ---- slices (reference utils.py:107-121) ---------------------------------------------------------------------
    def _event(self, kind: int, track_uuid: int, name: str | None, timestamp_ns: int) -> None:
        event = _field_varint(9, kind) + _field_varint(11, MAIN_SYSTEM_TRACK_UUID if track_uuid == -1 else track_uuid)
        if name is not None:
            event += _field_bytes(23, name.encode("utf-8"))
        self.packets.append(_field_varint(8, timestamp_ns) + _field_varint(10, TRUSTED_PACKET_SEQUENCE_ID) + _field_bytes(11, event))

    def start(self, name: str, track_uuid: int = -1) -> None:
        if self.enabled:
            self._event(TYPE_SLICE_BEGIN, track_uuid, name, time.time_ns())

    def end(self, track_uuid: int = -1) -> None:
        if self.enabled:
            self._event(TYPE_SLICE_END, track_uuid, None, time.time_ns())

    def device_slice(self, name: str, milliseconds: float, track_uuid: int) -> None:
        """A slice of `milliseconds` (a CUDA-event duration the library measured) ending now, on a GPU's track."""
        if not self.enabled or milliseconds is None or milliseconds <= 0:
            return
        now = time.time_ns()
        self._event(TYPE_SLICE_BEGIN, track_uuid, name, now - int(milliseconds * 1e6))
        self._event(TYPE_SLICE_END, track_uuid, None, now)

    def serialize(self) -> bytes:
        return b"".join(_field_bytes(1, p) for p in self.packets)

    def save(self, filename: str) -> None:
        with Path(filename).open("wb") as f:
            f.write(self.serialize())


def trace(block_name: str) -> Callable[[F], F]:
    def decorator(func: F) -> F:
        @functools.wraps(func)
        def wrapper(*args, **kwargs):  # noqa: ANN002, ANN003, ANN202
            if not TRACER.enabled:
                return func(*args, **kwargs)
            TRACER.start(block_name)
            try:
                return func(*args, **kwargs)
            finally:
                TRACER.end()

        return wrapper  # type: ignore[return-value]

    return decorator


TRACER = Tracer()


def parse_trace(blob: bytes) -> list[dict]:
    """Decode a trace written by :class:`Tracer` back into dicts (tests; the Perfetto UI is the real consumer)."""

    def fields(buf: bytes):  # noqa: ANN202
        pos = 0
        while pos < len(buf):
            key, shift = 0, 0
            while True:
                b = buf[pos]
                pos += 1
                key |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            number, wire = key >> 3, key & 7
            if wire == 0:
                value, shift = 0, 0
                while True:
                    b = buf[pos]
                    pos += 1
                    value |= (b & 0x7F) << shift
                    shift += 7
                    if not b & 0x80:
                        break
                yield number, value
            elif wire == 2:
                length, shift = 0, 0
                while True:
                    b = buf[pos]
                    pos += 1
                    length |= (b & 0x7F) << shift
                    shift += 7
                    if not b & 0x80:
                        break
                yield number, buf[pos:pos + length]
                pos += length
            else:
                raise ValueError(f"unexpected wire type {wire}")

    out = []
    for number, packet in fields(blob):
        assert number == 1
        rec: dict = {}
        for n, v in fields(packet):
            if n == 8:
                rec["timestamp"] = v
            elif n == 10:
                rec["sequence"] = v
            elif n == 11:
                ev = dict(fields(v))
                rec["event"] = {"type": ev.get(9), "track": ev.get(11), "name": ev[23].decode() if 23 in ev else None}
            elif n == 60:
                d = dict(fields(v))
                rec["track"] = {"uuid": d.get(1), "name": d[2].decode() if 2 in d else None, "parent": d.get(5)}
        out.append(rec)
    return out
