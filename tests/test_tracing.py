"""The tracer writes a Perfetto trace with the reference's surface (src/mini_spark/utils.py:85-166): track descriptors,
slice begin / end packets on the main track and on child tracks, `@trace`, `save`.  CPU only; the GPU engine's use of it
is checked in tests/test_gpu_engine.py."""

from __future__ import annotations

from minispark_b200.utils import MAIN_SYSTEM_TRACK_UUID, TYPE_SLICE_BEGIN, TYPE_SLICE_END, Tracer, parse_trace, trace
from minispark_b200 import utils


def test_tracer_round_trip(tmp_path):
    t = Tracer()
    t.enable()
    gpu = t.new_track("GPU 0 (device time)")
    assert gpu == t.new_track("GPU 0 (device time)")  # defined once
    t.start("execute full task")
    t.start("Stage 0")
    t.device_slice("scan: filter + aggregate", 0.35, gpu)
    t.end()
    t.end()
    path = tmp_path / "trace.pftrace"
    t.save(str(path))
    packets = parse_trace(path.read_bytes())
    tracks = [p["track"] for p in packets if "track" in p]
    assert tracks[0] == {"uuid": MAIN_SYSTEM_TRACK_UUID, "name": "Main System", "parent": None}
    assert tracks[1] == {"uuid": gpu, "name": "GPU 0 (device time)", "parent": MAIN_SYSTEM_TRACK_UUID}
    events = [p for p in packets if "event" in p]
    assert [e["event"]["type"] for e in events] == [TYPE_SLICE_BEGIN, TYPE_SLICE_BEGIN, TYPE_SLICE_BEGIN, TYPE_SLICE_END, TYPE_SLICE_END, TYPE_SLICE_END]
    assert [e["event"]["name"] for e in events[:3]] == ["execute full task", "Stage 0", "scan: filter + aggregate"]
    assert events[2]["event"]["track"] == gpu and events[0]["event"]["track"] == MAIN_SYSTEM_TRACK_UUID
    assert all(e["sequence"] == 1 for e in events)
    dev_begin, dev_end = events[2]["timestamp"], events[3]["timestamp"]
    assert 300_000 <= dev_end - dev_begin <= 400_000  # 0.35 ms in nanoseconds
    assert events[0]["timestamp"] <= events[1]["timestamp"] <= events[-1]["timestamp"]


def test_tracer_is_off_by_default_and_decorator(monkeypatch):
    t = Tracer()
    monkeypatch.setattr(utils, "TRACER", t)
    t.enabled = False
    calls = []

    @trace("block")
    def work(x):
        calls.append(x)
        return x + 1

    assert work(1) == 2 and len(t.packets) == 1  # only the main track's descriptor
    t.enable()
    assert work(2) == 3
    names = [p["event"]["name"] for p in parse_trace(t.serialize()) if "event" in p]
    assert names == ["block", None] and calls == [1, 2]
