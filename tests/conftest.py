"""pytest configuration: the ``gpu`` marker, UTC timestamps, repo on sys.path, shared table fixtures."""

from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import pytest

os.environ["TZ"] = "UTC"  # TIMESTAMP literals use naive local time (reference io.py:34-39)
time.tzset()
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a B200 and the built libminispark_cuda.so")


@pytest.fixture
def tables(tmp_path: Path) -> dict[str, str]:
    import cases

    return cases.write_tables(tmp_path / "tables")


@pytest.fixture(scope="session")
def small_lineitem(tmp_path_factory: pytest.TempPathFactory) -> str:
    """A 3-block, ~12k-row lineitem (all 16 columns) from the committed generator."""
    sys.path.insert(0, str(ROOT / "bench"))
    import gen_tpch

    path = tmp_path_factory.mktemp("tpch") / "lineitem_small.bin"
    gen_tpch.write_table(path, "lineitem", sf=0.002, rows_per_block=4096)
    return str(path)
