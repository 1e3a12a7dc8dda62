"""Host-side lowering (no GPU): SQL -> task tree -> fused logical plan -> expression programs.

A stub resolver stands in for the engine (device columns, dictionaries), so these tests pin the two
encodings the library receives: the three-address program of the C++ kernels and the postfix program
of the register-resident interpreter (csrc/gen_regvm.py), including its fused forms.
"""

from __future__ import annotations

from copy import deepcopy

import cases
from minispark_b200 import lowering as L
from minispark_b200 import native as N
from minispark_b200.parser import parse_sql


class StubDict:
    def __init__(self, entries):
        self.entries = list(entries)
        self.size = len(self.entries)


class StubResolver:
    """Columns of lineitem in native physical types; returnflag / linestatus dictionary coded."""

    PHYS = {"I": N.P_I32, "F": N.P_F32, "T": N.P_I64, "S": N.P_U8}

    def __init__(self, schema):
        self.schema = schema
        self.staged: list[int] = []
        self.ltypes: list[str] = []
        self.dict_of: dict[int, StubDict] = {}

    def binding(self, index: int) -> L.Binding:
        if index not in self.staged:
            self.staged.append(index)
        ltype = self.ltypes[index]
        return L.Binding(self.PHYS[ltype], staged=self.staged.index(index), dict_id=self.dict_of.get(index))

    def literal_code(self, dict_id, text):
        return dict_id.entries.index(text) if text in dict_id.entries else -1

    def like_lut(self, dict_id, pattern):
        return 0

    def translate_lut(self, dict_id, token):
        return -1, dict_id

    def same_dict(self, a, b):
        return a is b

    def recode_lut(self, src, dst):
        return 0


def _q1_plan(tmp_path):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))
    import gen_tpch

    table = tmp_path / "l.bin"
    gen_tpch.write_table(table, "lineitem", sf=0.0005, columns=gen_tpch.Q1_COLUMNS, rows_per_block=4096)
    task = parse_sql(cases.Q1_SQL.format(table=str(table))).task
    task = deepcopy(task)
    task.validate_schema()
    return L.lower_task(task)


def _compile_q1(tmp_path):
    plan = _q1_plan(tmp_path)
    assert isinstance(plan, L.LSelect) and isinstance(plan.child, L.LAggregate)
    agg = plan.child
    child = agg.child
    assert isinstance(child, L.LSelect)
    filters = list(child.filters)
    group = L.substitute(agg.group, child.outputs)
    aggs = [(k, L.substitute(e, child.outputs)) for k, e in agg.aggs]
    table = child.child
    res = StubResolver(table.schema)
    ltype_of = {"INTEGER": "I", "FLOAT": "F", "TIMESTAMP": "T", "STRING": "S"}
    res.ltypes = [ltype_of[t.name] for _, t in table.schema]
    res.dict_of = {i: StubDict(["A", "N", "R"]) for i, t in enumerate(res.ltypes) if t == "S"}
    return L.compile_aggregate(res, filters, group, aggs), res


def test_q1_lowers_to_one_fused_scan_with_ten_instructions(tmp_path):
    prog, res = _compile_q1(tmp_path)
    text = prog.program.text
    assert text[-1] == "END" and len(text) == 11
    assert text[0].startswith("filter <- LE_I(") and text[1].startswith("group <- MOV(")
    assert len(res.staged) == 6                     # the 6 Q1 columns, each staged once
    assert len(prog.agg_kinds) == 6                 # 8 requested aggregates share 6 accumulators (AVG = SUM / COUNT)
    assert prog.slot_of.count(prog.slot_of[-1]) >= 1
    assert all("[fast" in line for line in text[:-1])


def test_q1_regvm_program_uses_the_fused_forms(tmp_path):
    prog, _ = _compile_q1(tmp_path)
    rv = prog.program.regvm_text
    names = [line.split()[0] for line in rv]
    assert names[0] == "CMPCOL_LE_I64" and names[1] == "GROUP_U8" and names[-1] == "END"
    # SUM(l_quantity), SUM(l_extendedprice), SUM(l_discount): adjacent columns into adjacent slots, one instruction
    assert "AGGCOL3_F32" in names
    assert names.count("COUNT") == 1
    assert prog.program.regvm_count_slot >= 0
    # price * (1 - disc) is summed and kept for the charge expression in one instruction
    assert any(n.startswith("AGGT0_SUMF") for n in names)
    assert len(names) <= 11
    for word, line in zip(prog.program.regvm, rv):
        name, a1, a2 = line.split()
        assert word == N.RV[name] | (int(a1) << 8) | (int(a2) << 16)


def _lower_sql(sql: str):
    task = deepcopy(parse_sql(sql).task)
    task.validate_schema()
    return L.lower_task(task)


def test_filters_move_below_an_inner_join(tmp_path):
    """One-sided conjuncts of a WHERE above a join filter that side BEFORE the join (the reference filters after it,
    tasks.py:167-177 above :201-240; same rows for an inner join, a much smaller build / probe)."""
    paths = cases.write_tables(tmp_path)
    plan = _lower_sql("SELECT u.first_name, o.product, o.price FROM '{users}' AS u JOIN '{orders}' AS o ON u.user_id=o.user_id "
                      "WHERE (o.price > 100) AND (o.quantity < u.age);".format(**paths))
    assert isinstance(plan, L.LSelect) and isinstance(plan.child, L.LJoin)
    join = plan.child
    assert isinstance(join.left, L.LTable)                                   # nothing to push to the users side
    assert isinstance(join.right, L.LSelect) and len(join.right.filters) == 1  # price > 100 runs on orders alone
    assert len(plan.filters) == 1                                            # quantity < age needs both sides: stays above
    nl = len(join.left.schema)
    ins = L.expr_inputs(plan.filters[0])
    assert any(i < nl for i in ins) and any(i >= nl for i in ins)


def test_a_filter_that_could_raise_stays_above_the_join(tmp_path):
    paths = cases.write_tables(tmp_path)
    plan = _lower_sql("SELECT u.first_name, o.product FROM '{users}' AS u JOIN '{orders}' AS o ON u.user_id=o.user_id "
                      "WHERE 100 / o.quantity > 20;".format(**paths))
    join = plan.child
    assert isinstance(join, L.LJoin) and isinstance(join.right, L.LTable) and len(plan.filters) == 1


def test_build_scan_program_is_filters_rank_and_the_key(tmp_path):
    """The build half of a join as one scan (msc_scan_join_build; off by default): the side's own filters, RANK
    (so that a count pass can stop there) and GROUP <- key -- nothing else."""
    plan = _q1_plan(tmp_path)
    table = plan.child.child.child
    res = StubResolver(table.schema)
    ltype_of = {"INTEGER": "I", "FLOAT": "F", "TIMESTAMP": "T", "STRING": "S"}
    res.ltypes = [ltype_of[t.name] for _, t in table.schema]
    res.dict_of = {i: StubDict(["A", "N", "R"]) for i, t in enumerate(res.ltypes) if t == "S"}
    names = [n for n, _ in table.schema]
    qty, ts = names.index("l_quantity"), names.index("l_shipdate")
    filters = [L.EBin(L.BOOL, "gt", L.EInput(L.FLOAT, qty), L.EConst(L.FLOAT, 25.0)), L.EBin(L.BOOL, "le", L.EInput(L.TS, ts), L.EConst(L.TS, 10**15))]
    program, key_dict = L.compile_build(res, filters, L.EInput(L.TS, ts))
    text = program.text
    assert [line.split(" <- ")[0].split(",")[0] for line in text[:2]] == ["filter", "filter"]
    assert text[2].startswith("RANK") or "RANK" in text[2]
    assert text[3].startswith("group <- MOV(") and text[-1] == "END" and len(text) == 5
    assert key_dict is None
    unfiltered, _ = L.compile_build(res, [], L.EInput(L.TS, ts))
    assert len(unfiltered.text) == 2 and unfiltered.text[0].startswith("group <- MOV(")   # no filter: no RANK either
