"""The engine's input boundary is the reference's OWN task tree: every tests/cases.py query built with the real
``mini_spark`` classes must lower to the same logical plan as the mirror-built one (CPU; needs the reference checkout)."""

from __future__ import annotations

import re
import subprocess
import sys
from pathlib import Path

import pytest

RUNNER = Path(__file__).with_name("ref_boundary_runner.py")


def test_reference_built_task_trees_lower_like_mirror_built_ones(tmp_path):
    if not Path("/root/reference/src/mini_spark").exists():
        pytest.skip("reference checkout not present")
    out = subprocess.run([sys.executable, str(RUNNER), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    m = re.search(r"boundary ok (\d+)", out.stdout)
    assert m and int(m.group(1)) >= 29, out.stdout[-1000:]
