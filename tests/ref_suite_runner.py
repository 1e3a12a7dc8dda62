"""Runs one of the REFERENCE's own test files against this package's host mirror (tests/test_reference_suite.py).

The reference's tests import ``mini_spark.<module>``; here those names resolve to the same-named modules of
``minispark_b200``, so the files run unmodified from where they lie (nothing is copied).  Their conftest is not loaded
(it patches the reference's tracer and shuffle folder); the one fixture the files below use from it is supplied here.

    python tests/ref_suite_runner.py /root/reference/tests/test_parser.py [pytest args...]
"""
import importlib
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import minispark_b200 as pkg  # noqa: E402

sys.modules["mini_spark"] = pkg
for name in ("dataframe", "parser", "sql", "tasks", "io", "constants", "jobs"):
    sys.modules["mini_spark." + name] = importlib.import_module("minispark_b200." + name)

import pytest  # noqa: E402


class _Fixtures:
    @pytest.fixture
    def temporary_file(self):  # a path that does not exist yet, removed afterwards
        with tempfile.TemporaryDirectory() as folder:
            yield Path(folder) / "file.bin"


if __name__ == "__main__":
    scratch = tempfile.mkdtemp()
    sys.exit(pytest.main(["-q", "--noconftest", "-p", "no:cacheprovider", "--rootdir", scratch, "-c", "/dev/null", *sys.argv[1:]], plugins=[_Fixtures()]))
