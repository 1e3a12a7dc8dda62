"""JSON encoding of result rows that keeps floats bit-exact (hex) and datetimes typed."""

from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Any


def encode_value(v: Any) -> Any:
    if isinstance(v, bool):
        return {"b": v}
    if isinstance(v, float):
        return {"f": v.hex()}
    if isinstance(v, datetime):
        return {"t": v.isoformat()}
    return v


def decode_value(v: Any) -> Any:
    if isinstance(v, dict):
        if "f" in v:
            return float.fromhex(v["f"])
        if "t" in v:
            return datetime.fromisoformat(v["t"])
        if "b" in v:
            return bool(v["b"])
    return v


def dump(path: Path, cases: dict[str, list[dict]]) -> None:
    # case names sorted for stable diffs; column order inside a row is significant and preserved
    doc = {name: [{k: encode_value(v) for k, v in row.items()} for row in cases[name]] for name in sorted(cases)}
    path.write_text(json.dumps(doc, indent=1) + "\n")


def load(path: Path) -> dict[str, list[dict]]:
    doc = json.loads(Path(path).read_text())
    return {name: [{k: decode_value(v) for k, v in row.items()} for row in rows] for name, rows in doc.items()}
