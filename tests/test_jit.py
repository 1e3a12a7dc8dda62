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


def test_more_than_32_cells_keep_their_accumulators_in_shared_memory(tmp_path):
    """8 groups x 6 accumulators: every thread owns a copy of each cell in shared memory and a row updates its own group's
    cells only -- no per-group predicated FP64 instruction stream (HBM- not FP64-bound)."""
    src = q1_source(tmp_path, ngroups=8)
    assert "constexpr int NG = 8" in src and "cellbase" in src and "cellrow + 0 * NT" in src
    assert "a0_0" not in src and "mlut" not in src and "addf_if<" not in src.split("msc_jit_dense")[1]
    cubin = compile_source(src)
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    path = tmp_path / "q1_smem.cubin"
    path.write_bytes(cubin)
    sass = subprocess.run([cuobjdump, "-sass", str(path)], capture_output=True, text=True, check=True).stdout
    loop = sass[sass.index("TRYWAIT"):sass.index("UBLKCP", sass.index("TRYWAIT"))]
    assert loop.count("DADD") + loop.count("DFMA") <= 8 * 5 * 2   # per row one add per SUM (and the expression arithmetic), not one per group
    assert "STS" in loop and "LDS" in loop


def test_generator_refuses_too_many_cells(tmp_path):
    prog, res = _compile_q1(tmp_path)
    lib = N.load()
    d = _desc(prog, res)
    n = C.c_size_t()
    buf = C.create_string_buffer(4096)
    rc = lib.msc_jit_dense_source(C.byref(d), 40, N.int32_array(prog.agg_kinds), len(prog.agg_kinds), 1, buf, len(buf), C.byref(n))
    assert rc != 0 and b"register budget" in buf.value


def _project_program(tmp_path, sql_tail: str):
    """Lower `SELECT ... FROM lineitem WHERE ...` with the stub resolver of test_lowering."""
    import sys
    from copy import deepcopy
    from pathlib import Path

    from minispark_b200 import lowering as L
    from minispark_b200.parser import parse_sql
    from test_lowering import StubDict, StubResolver

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))
    import gen_tpch

    table = tmp_path / "l.bin"
    gen_tpch.write_table(table, "lineitem", sf=0.0005, columns=gen_tpch.Q1_COLUMNS, rows_per_block=4096)
    task = deepcopy(parse_sql(sql_tail.format(table=str(table))).task)
    task.validate_schema()
    plan = L.lower_task(task)
    assert isinstance(plan, L.LSelect)
    base = plan.child
    res = StubResolver(base.schema)
    ltype_of = {"INTEGER": "I", "FLOAT": "F", "TIMESTAMP": "T", "STRING": "S"}
    res.ltypes = [ltype_of[t.name] for _, t in base.schema]
    res.dict_of = {i: StubDict(["A", "N", "R"]) for i, t in enumerate(res.ltypes) if t == "S"}
    return L.compile_project(res, list(plan.filters), list(plan.outputs)), res


def _project_source(prog, res, count_only: bool) -> str:
    lib = N.load()
    d = N.ScanDesc()
    d.nrows = 1 << 20
    d.nstaged = len(res.staged)
    for i, index in enumerate(res.staged):
        d.staged[i].phys = res.PHYS[res.ltypes[index]]
    words = prog.program.words()
    d.ncode = len(words)
    for i, w in enumerate(words):
        d.code[i] = w
    d.nconsts = len(prog.program.consts)
    for i, c in enumerate(prog.program.consts):
        d.consts[i] = c
    d.ntemps = prog.program.ntemps
    n = C.c_size_t()
    buf = C.create_string_buffer(1 << 20)
    rc = lib.msc_jit_project_source(C.byref(d), int(count_only), N.int32_array(prog.out_phys), len(prog.out_phys), buf, len(buf), C.byref(n))
    assert rc == 0, buf.value.decode()
    return buf.value.decode()


def test_filtered_projection_compiles_as_two_passes(tmp_path):
    sql = ("SELECT l_returnflag, l_extendedprice * (1 - l_discount) AS disc_price, l_quantity FROM '{table}' "
           "WHERE (l_quantity > 10) AND (l_tax < l_discount);")  # (the reference types AND by its operands: both FLOAT)
    prog, res = _project_program(tmp_path, sql)
    count_src = _project_source(prog, res, True)
    assert "p.tile_counts[tile]" in count_src and "p.out[" not in count_src
    proj_src = _project_source(prog, res, False)
    assert "p.tile_offsets[tile]" in proj_src and proj_src.count("p.out[") == 3
    assert "reinterpret_cast<u32*>(p.out[0])" in proj_src        # the STRING column leaves as dictionary codes
    for src in (count_src, proj_src):
        assert compile_source(src)[:4] == b"\x7fELF"


def test_unfiltered_projection_uses_vector_stores(tmp_path):
    prog, res = _project_program(tmp_path, "SELECT l_quantity * l_extendedprice AS v FROM '{table}';")
    src = _project_source(prog, res, False)
    assert "p.tile_offsets" not in src.split("msc_jit_scan")[1] and "make_longlong2" in src
    compile_source(src)


def test_q1_with_fused_finish_compiles(tmp_path):
    """scan + compaction + final projection (AVG = SUM / COUNT, output order) as ONE kernel: the last CTA finishes the query."""
    from minispark_b200 import lowering as L
    from test_lowering import StubDict, StubResolver, _q1_plan

    prog, res = _compile_q1(tmp_path)
    plan = _q1_plan(tmp_path)
    agg = plan.child

    class FinalResolver(StubResolver):  # columns of the compacted aggregate: key code (u32) + one f64 / i64 column per aggregate
        PHYS = {"I": N.P_I64, "F": N.P_F64, "T": N.P_I64, "S": N.P_U32}

    res2 = FinalResolver(agg.schema)
    ltype_of = {"INTEGER": "I", "FLOAT": "F", "TIMESTAMP": "T", "STRING": "S"}
    res2.ltypes = [ltype_of[t.name] for _, t in agg.schema]
    res2.dict_of = {0: StubDict(["A", "N", "R"])}
    prog2 = L.compile_project(res2, list(plan.filters), list(plan.outputs))
    d2 = N.ScanDesc()
    d2.nstaged = len(res2.staged)
    for i, index in enumerate(res2.staged):
        d2.staged[i].phys = res2.PHYS[res2.ltypes[index]]
    words = prog2.program.words()
    d2.ncode = len(words)
    for i, w in enumerate(words):
        d2.code[i] = w
    d2.nconsts = len(prog2.program.consts)
    for i, c in enumerate(prog2.program.consts):
        d2.consts[i] = c
    d2.ntemps = prog2.program.ntemps
    raw_cols = [0 if ci == 0 else 1 + prog.slot_of[ci - 1] for ci in res2.staged]
    lib = N.load()
    d = _desc(prog, res)
    n = C.c_size_t()
    buf = C.create_string_buffer(1 << 20)
    rc = lib.msc_jit_dense_fused_source(C.byref(d), 3, N.int32_array(prog.agg_kinds), len(prog.agg_kinds), 1, C.byref(d2), N.int32_array(raw_cols),
                                        N.int32_array(prog2.out_phys), len(prog2.out_phys), 0, buf, len(buf), C.byref(n))
    src = buf.value.decode()
    assert rc == 0, src
    assert "p.mailbox" not in src.split("msc_jit_dense(")[1]
    assert "atomicAdd(p.ticket, 1u) == gridDim.x - 1" in src and "p.fmeta[0] = __popc(keep)" in src
    assert src.count("p.out[") == len(prog2.out_phys)
    assert " / " in src.split("s_last")[-1]                     # AVG = SUM / COUNT happens in the finish
    compile_source(src)
    # the cross-rank variant: the same last CTA first exchanges the partial tables over NVLink peer memory
    rc = lib.msc_jit_dense_fused_source(C.byref(d), 3, N.int32_array(prog.agg_kinds), len(prog.agg_kinds), 1, C.byref(d2), N.int32_array(raw_cols),
                                        N.int32_array(prog2.out_phys), len(prog2.out_phys), 1, buf, len(buf), C.byref(n))
    peer_src = buf.value.decode()
    assert rc == 0, peer_src
    body = peer_src.split("msc_jit_dense(")[1]
    assert "st_release_sys(p.mailbox[lane]" in body and "ld_acquire_sys(flag) != p.epoch" in body and "p.inv[r * 32 + g]" in body
    compile_source(peer_src)


def test_streaming_aggregate_over_sorted_runs_compiles(tmp_path):
    """GROUP BY l_orderkey (a sorted key): run heads by comparison with the previous row, run numbers by a warp scan, one
    atomic per (run, accumulator) segment."""
    import sys
    from copy import deepcopy
    from pathlib import Path

    from minispark_b200 import lowering as L
    from test_lowering import StubResolver

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))
    import gen_tpch

    table = tmp_path / "l.bin"
    gen_tpch.write_table(table, "lineitem", sf=0.0005, columns=["l_orderkey", "l_quantity", "l_extendedprice"], rows_per_block=4096)
    from minispark_b200.parser import parse_sql

    task = deepcopy(parse_sql(f"SELECT l_orderkey, SUM(l_quantity) AS q, AVG(l_extendedprice) AS p, MAX(l_quantity) AS hi FROM '{table}' "
                              "GROUP BY l_orderkey;").task)
    task.validate_schema()
    plan = L.lower_task(task)
    agg = plan.child
    res = StubResolver(agg.child.schema)
    ltype_of = {"INTEGER": "I", "FLOAT": "F", "TIMESTAMP": "T", "STRING": "S"}
    res.ltypes = [ltype_of[t.name] for _, t in agg.child.schema]
    prog = L.compile_aggregate(res, [], agg.group, list(agg.aggs))
    d = _desc(prog, res)
    lib = N.load()
    n = C.c_size_t()
    buf = C.create_string_buffer(1 << 20)
    key_slot = res.staged.index(0)
    rc = lib.msc_jit_runs_source(C.byref(d), N.int32_array(prog.agg_kinds), len(prog.agg_kinds), key_slot, buf, len(buf), C.byref(n))
    src = buf.value.decode()
    assert rc == 0, src
    assert "msc_jit_runs" in src and "fold_store_heads<0>(" in src and "fold_store_heads<3>(" in src   # SUM_F and MAX_F accumulators: stored, then added to
    assert "p.tile_offsets[tile + nw]" in src and "out_key[run0] = key[r]" in src   # (run base fetched one tile ahead)
    compile_source(src)


class _ProbeResolver:
    """Probe side: l_orderkey (i32), l_shipmode (u8 code), l_extendedprice (f32) staged; build side: o_orderpriority (u32 code)
    read through the probe's match (Binding.probe) -- the consumer scan of BASELINE config 5."""

    def __init__(self):
        from test_lowering import StubDict

        self.staged: list[int] = []
        self.gather: list[int] = []
        self.dicts = {1: StubDict(["AIR", "MAIL", "REG AIR"]), 100: StubDict(["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"])}
        self.phys = {0: N.P_I32, 1: N.P_U8, 2: N.P_F32, 100: N.P_U32}

    def binding(self, index):
        from minispark_b200 import lowering as L

        if index >= 100:
            if index not in self.gather:
                self.gather.append(index)
            return L.Binding(self.phys[index], gather=self.gather.index(index), dict_id=self.dicts.get(index), probe=True)
        if index not in self.staged:
            self.staged.append(index)
        return L.Binding(self.phys[index], staged=self.staged.index(index), dict_id=self.dicts.get(index))

    def like_lut(self, dict_id, pattern):
        return 1

    def literal_code(self, dict_id, text):
        return dict_id.entries.index(text)

    def same_dict(self, a, b):
        return a is b


def _probe_desc(prog, res) -> N.ScanDesc:
    d = N.ScanDesc()
    d.nrows = 1 << 20
    d.nstaged, d.ngather, d.nluts = len(res.staged), len(res.gather), 2
    for i, index in enumerate(res.staged):
        d.staged[i].phys = res.phys[index]
    for i, index in enumerate(res.gather):
        d.gather[i].phys = res.phys[index]
    words = prog.program.words()
    d.ncode = len(words)
    for i, w in enumerate(words):
        d.code[i] = w
    d.nconsts = len(prog.program.consts)
    for i, c in enumerate(prog.program.consts):
        d.consts[i] = c
    d.ntemps = prog.program.ntemps
    return d


def test_fused_join_probe_compiles(tmp_path):
    """A join carried by the consuming scan (MSC_OP_PROBE + build-side columns through the matched row): the program the
    lowering emits and the specialised kernels generated from it, as a dense aggregate and as a filter / project scan."""
    from minispark_b200 import lowering as L

    res = _ProbeResolver()
    like = L.ELike(L.BOOL, L.EInput(L.STR, 1), "%AIR%")
    probe = L.ProbeSpec(L.EInput(L.INT, 0), 0)
    prog = L.compile_aggregate(res, [], L.EInput(L.STR, 100), [("count", L.EConst(L.INT, 1)), ("sum", L.EInput(L.FLOAT, 2))],
                               probe=probe, pre_filters=[like])
    text = prog.program.text
    assert text[0].startswith("filter <- LUT8(col0, lut1)") and text[1].startswith("t0 <- PROBE(col1, lut0)")
    assert text[2].startswith("filter <- GE_I(t0, const0)") and text[3].startswith("group <- MOV(gather0[t0], -)")
    assert prog.program.regvm == []  # the register interpreter has no probe: the C++ interpreter or a specialised kernel runs it
    lib = N.load()
    d = _probe_desc(prog, res)
    n = C.c_size_t()
    buf = C.create_string_buffer(1 << 20)
    for masked in (1, 0):
        rc = lib.msc_jit_dense_source(C.byref(d), 5, N.int32_array(prog.agg_kinds), len(prog.agg_kinds), masked, buf, len(buf), C.byref(n))
        src = buf.value.decode()
        assert rc == 0, src
        # the dense scan splits its row loop around the probe: all of a lane's table reads are issued, then resolved
        assert "join_probe_issue(p.luts[0], c1[r], valid, pq[r], keep_policy)" in src and "JoinProbe<false> pq[R]" in src
        assert "t0[r] = join_probe_resolve(p.luts[0], pq[r], valid, keep_policy)" in src
        # ... and the rows that found a partner are queued: the gathers and aggregates run once per survivor, a lane each
        assert "queue[2 * at + 1] = static_cast<u32>(t0[r])" in src and "for (u32 e = lane; e < survivors; e += 32)" in src
        assert "gather_at<2>(p.gather[0], w0, valid && w0 >= 0)" in src and "ldrow<5>(sb + " in src
        compile_source(src)
    res2 = _ProbeResolver()
    proj = L.compile_project(res2, [L.EBin(L.BOOL, "gt", L.EInput(L.FLOAT, 2), L.EConst(L.FLOAT, 10.0))],
                             [L.EInput(L.INT, 0), L.EInput(L.STR, 100), L.EInput(L.FLOAT, 2)], probe=L.ProbeSpec(L.EInput(L.INT, 0), 0, compact=True), pre_filters=[])
    assert any("RANK" in t for t in proj.program.text), proj.program.text
    d2 = _probe_desc(proj, res2)
    for count_only in (1, 0):
        rc = lib.msc_jit_project_source(C.byref(d2), count_only, N.int32_array(proj.out_phys), len(proj.out_phys), buf, len(buf), C.byref(n))
        src = buf.value.decode()
        assert rc == 0, src
        assert "join_probe<true>(p.luts[0]" in src
        compile_source(src)
