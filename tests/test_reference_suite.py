"""The reference's OWN test files, run unmodified against this package's host mirror (parser, Col / DataFrame / Task
surface, BlockFile) -- where the reference checkout is present (the build container; it is absent on the GPU box, and
these are CPU tests).  ``tests/test_parser.py:18-414`` pins the hand-written SQL front end (the reference's needs
``parsimonious``, which is not installed), ``test_tasks.py`` schema propagation and output naming, ``test_io.py`` the
BlockFile round trips and block splitting, ``test_sql.py`` the expression classes (its Zig code-generation cases do not
apply: this engine lowers to expression programs instead)."""

from __future__ import annotations

import re
import subprocess
import sys
from pathlib import Path

import pytest

REF_TESTS = Path("/root/reference/tests")
RUNNER = Path(__file__).with_name("ref_suite_runner.py")

CASES = [
    ("test_parser.py", [], 25),
    ("test_tasks.py", [], 10),
    ("test_io.py", [], 4),
    ("test_sql.py", ["-k", "not zig"], 5),
]


@pytest.mark.parametrize(("name", "extra", "at_least"), CASES, ids=[c[0] for c in CASES])
def test_reference_test_file_passes_against_the_mirror(name, extra, at_least, tmp_path):
    path = REF_TESTS / name
    if not path.exists():
        pytest.skip("reference checkout not present")
    # (run from a scratch folder: the reference's tasks write their shuffle files relative to the working directory)
    out = subprocess.run([sys.executable, str(RUNNER), str(path), *extra], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    passed = re.search(r"(\d+) passed", out.stdout)
    assert passed and int(passed.group(1)) >= at_least, out.stdout[-1500:]
