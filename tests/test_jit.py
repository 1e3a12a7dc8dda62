"""Query-specialised scan kernels (csrc/jit.cu), host side (no GPU): the generator turns the three-address program of
TPC-H Q1 into CUDA C++ and NVRTC compiles it for sm_100a.  The GPU behaviour of the compiled kernels is covered by
tests/test_gpu_jit.py."""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import pytest
from minispark_b200 import native as N
from test_lowering import _compile_q1


def _desc(prog, res) -> N.ScanDesc:
    d = N.ScanDesc()
    d.nrows = 1 << 20
    d.nstaged = len(res.staged)
    for i, index in enumerate(res.staged):
        d.staged[i].data = 0
        d.staged[i].phys = res.PHYS[res.ltypes[index]]
    words = prog.program.words()
    d.ncode = len(words)
    for i, w in enumerate(words):
        d.code[i] = w
    d.nconsts = len(prog.program.consts)
    for i, c in enumerate(prog.program.consts):
        d.consts[i] = c
    d.ntemps = prog.program.ntemps
    d.ncode2 = len(prog.program.regvm)
    d.count_slot2 = prog.program.regvm_count_slot
    for i, w in enumerate(prog.program.regvm):
        d.code2[i] = w
    return d


def q1_source(tmp_path, ngroups: int = 3, masked: bool = True) -> str:
    prog, res = _compile_q1(tmp_path)
    lib = N.load()
    d = _desc(prog, res)
    kinds = N.int32_array(prog.agg_kinds)
    n = C.c_size_t()
    buf = C.create_string_buffer(1 << 20)
    rc = lib.msc_jit_dense_source(C.byref(d), ngroups, kinds, len(prog.agg_kinds), int(masked), buf, len(buf), C.byref(n))
    assert rc == 0, buf.value.decode()
    return buf.value.decode()


def compile_source(source: str) -> bytes:
    lib = N.load()
    n = C.c_size_t()
    log = C.create_string_buffer(1 << 16)
    cubin = C.create_string_buffer(1 << 22)
    rc = lib.msc_jit_compile(source.encode(), cubin, len(cubin), C.byref(n), log, len(log))
    if rc != 0 and b"not found" in log.value:
        pytest.skip(log.value.decode())
    assert rc == 0, log.value.decode()
    return cubin.raw[: n.value]


def test_q1_source_is_straight_line_code_with_register_accumulators(tmp_path):
    src = q1_source(tmp_path)
    assert 'extern "C" __global__' in src and "msc_jit_dense" in src
    assert "constexpr int NG = 3, NGP = 4, STRIDE = 6" in src   # 8 requested aggregates share 6 accumulators
    assert src.count("// instruction") == 10                    # the 10 three-address instructions of Q1, once each
    assert "a2_5" in src and "n0_5" in src                      # accumulators per (group, slot); COUNT rows in a u32
    assert "brx" not in src and "switch (op" not in src         # no dispatch


def test_q1_source_compiles_for_sm100a(tmp_path):
    cubin = compile_source(q1_source(tmp_path))
    assert cubin[:4] == b"\x7fELF"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    path = tmp_path / "q1.cubin"
    path.write_bytes(cubin)
    sass = subprocess.run([cuobjdump, "-sass", str(path)], capture_output=True, text=True, check=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass                 # TMA bulk copies + mbarriers
    assert "DADD" in sass or "DFMA" in sass
    loop = sass[sass.index("TRYWAIT"):sass.index("UBLKCP", sass.index("TRYWAIT"))]  # the per-tile loop: wait .. refill
    assert "BRX" not in loop                                    # no dispatch left in it
    assert loop.count("DFMA") >= 8 * 3 * 5                      # one fma per row, group and SUM


def test_exact_variant_has_no_mask_arithmetic(tmp_path):
    src = q1_source(tmp_path, masked=False)
    assert "addf_if<2>" in src and "mlut + (" not in src
    compile_source(src)


def test_generator_refuses_too_many_cells(tmp_path):
    prog, res = _compile_q1(tmp_path)
    lib = N.load()
    d = _desc(prog, res)
    n = C.c_size_t()
    buf = C.create_string_buffer(4096)
    rc = lib.msc_jit_dense_source(C.byref(d), 40, N.int32_array(prog.agg_kinds), len(prog.agg_kinds), 1, buf, len(buf), C.byref(n))
    assert rc != 0 and b"register budget" in buf.value
