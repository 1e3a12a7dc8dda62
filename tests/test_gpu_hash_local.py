"""Hash-mode GROUP BY behind CTA-local pre-aggregation tables (scan_kernel.cuh, ScanParams::lcap) and with optimistic table
sizing (scan.cu, msc_scan_aggregate): few groups, more groups than local slots, more groups than the first table holds.

The reference's aggregate is a Python dict per block merged across blocks (tasks.py:347-375); the GPU's version of "a dict per
worker, merged" is a shared-memory table per CTA folded into one global table.  Checked bit-exactly against numpy (the sums
are exact in f64 by construction), each configuration in its own process because the switches are read once per process.
"""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
WORKER = Path(__file__).resolve().parent / "hash_local_worker.py"


def _run(tmp_path, nrows, distinct, attempts=None, **env):
    full = dict(os.environ)
    full.update({k: str(v) for k, v in env.items()})
    args = [sys.executable, str(WORKER), str(tmp_path), str(nrows), str(distinct)] + ([str(attempts)] if attempts is not None else [])
    done = subprocess.run(args, env=full, capture_output=True, text=True, timeout=600)
    assert done.returncode == 0 and done.stdout.strip().endswith("OK"), done.stdout + done.stderr
    return done.stdout


@pytest.mark.parametrize("distinct", [1, 50, 700])
def test_few_groups_stay_in_the_local_tables(tmp_path, distinct):
    out = _run(tmp_path, 60000, distinct, attempts=1)
    assert "local slots 0" not in out


def test_more_groups_than_local_slots_overflow_to_the_global_table(tmp_path):
    _run(tmp_path, 80000, 5000, attempts=1, MSC_HASH_LOCAL_SLOTS=64)


def test_more_groups_than_the_first_table_repeats_the_scan(tmp_path):
    out = _run(tmp_path, 400000, 150000, attempts=2)
    assert "attempts 2" in out


def test_switches_off_match(tmp_path):
    out = _run(tmp_path, 60000, 50, attempts=1, MSC_HASH_LOCAL_SLOTS=0, MSC_HASH_OPTIMISTIC=0)
    assert "local slots 0" in out
