#!/usr/bin/env python
"""Generator of the register-resident interpreter core ("regvm") of the dense aggregate scan.

nvcc cannot keep an interpreter's value stack in registers across a dispatch loop without drowning
it in moves (profiles/r01_scan_kernel_ncu_summary.json, prof_r1a: 70 % MOV).  This script therefore
emits the interpreter as hand-scheduled PTX: a postfix stack machine whose stack depth is static per
instruction, so every (operation, depth) pair is its own handler working on fixed registers, and
handlers are reached through one ``brx.idx`` jump table (threaded dispatch).  Outputs, both committed:

  regvm_ptx.inc      the PTX text, included as the body of one ``asm volatile`` block in scan_regvm.cu
  regvm_handlers.h   handler ids + operand metadata (validated / patched on the host, parsed by Python)

A lane owns R = 4 consecutive rows.  Instruction word: handler id | a1 << 8 | a2 << 20 (12 bits each).
Register conventions inside the block: s<d>_<r> stack slot d of row r (64-bit), t<k>_<r> temporaries,
go<r> byte offset of row r's group in the per-lane accumulator block, pv<r> row validity predicate.

    python minispark_b200/csrc/gen_regvm.py      # rewrites the two files next to this script
"""

from __future__ import annotations

from pathlib import Path

R = 4
DEPTH = 4
NTEMPS = 2
SLOT_STRIDE = 128 * 8  # bytes between accumulator slots: NT lanes x 8 B

# operand kinds for the host-side validator
A_NONE, A_COL, A_CONST, A_SLOT = 0, 1, 2, 3
PHYS = {"U8": 0, "U16": 1, "U32": 2, "I32": 3, "I64": 4, "F32": 5, "F64": 6}

handlers: list[dict] = []   # name, body lines, a1 kind, a2 kind, col phys, stack delta, min depth


AGG_KIND = {"SUMF": 0, "SUMI": 1, "MINF": 2, "MAXF": 3, "MINI": 4, "MAXI": 5}  # MSC_AGG_* of include/minispark_cuda.h


def add(name: str, body: list[str], a1: int = A_NONE, a2: int = A_NONE, phys: int = -1, delta: int = 0, depth: int = -1,
        agg: int = -1) -> None:
    """depth: stack depth required before the instruction (-1: any); agg: accumulator kind the slot operand must have."""
    handlers.append(dict(name=name, body=body, a1=a1, a2=a2, phys=phys, delta=delta, depth=depth, agg=agg))


def col_addr(width_bytes_per_lane: int) -> list[str]:
    return ["mad.lo.u32 ad, a1, 16, sb;", f"mad.lo.u32 ad, lane, {width_bytes_per_lane}, ad;"]


def load_rows(kind: str, dst: str) -> list[str]:
    """Load this lane's 4 rows of a staged column (a1 = offset/16) into registers <dst>_0..3 as 64-bit values."""
    if kind == "F32":
        return col_addr(16) + ["ld.shared.v4.f32 {f0, f1, f2, f3}, [ad];"] + [f"cvt.f64.f32 {dst}_{r}, f{r};" for r in range(R)]
    if kind in ("F64", "I64"):
        return col_addr(32) + [f"ld.shared.v2.b64 {{{dst}_0, {dst}_1}}, [ad];", f"ld.shared.v2.b64 {{{dst}_2, {dst}_3}}, [ad+16];"]
    if kind == "I32F":
        return col_addr(16) + ["ld.shared.v4.s32 {i0, i1, i2, i3}, [ad];"] + [f"cvt.rn.f64.s32 {dst}_{r}, i{r};" for r in range(R)]
    if kind == "I32":
        return col_addr(16) + ["ld.shared.v4.s32 {i0, i1, i2, i3}, [ad];"] + [f"cvt.s64.s32 {dst}_{r}, i{r};" for r in range(R)]
    raise ValueError(kind)


def const_load(reg: str, arg: str) -> list[str]:
    return [f"mad.lo.u32 ad, {arg}, 8, cstb;", f"ld.shared.b64 {reg}, [ad];"]


# ---- END -----------------------------------------------------------------------------------------
add("END", ["bra.uni RV_DONE;"])

# ---- column <cmp> const -> filter ------------------------------------------------------------------
CMPS = ["lt", "le", "gt", "ge", "eq", "ne"]
for ty, load_kind, setp_ty, phys in (("I64", "I64", "s64", "I64"), ("I32", "I32", "s64", "I32"), ("F32", "F32", "f64", "F32"), ("F64", "F64", "f64", "F64")):
    for cmp_ in CMPS:
        body = load_rows(load_kind, "x") + const_load("c64", "a2")
        for r in range(R):
            body += [f"setp.{cmp_}.{setp_ty} q, x_{r}, c64;", f"and.pred pv{r}, pv{r}, q;"]
        add(f"CMPCOL_{cmp_.upper()}_{ty}", body, A_COL, A_CONST, PHYS[phys])

# ---- dictionary code column -> dense group ---------------------------------------------------------
def group_body(codes: list[str]) -> list[str]:
    body = []
    for r in range(R):
        body += codes[r] if isinstance(codes[r], list) else [codes[r]]
        # a code outside [0, ngroups) cannot be aggregated: the row is dropped like a filtered one
        body += [f"setp.lt.u32 q, c{r}, ng;", f"and.pred pv{r}, pv{r}, q;", f"selp.b32 c{r}, c{r}, 0, q;", f"mul.lo.u32 go{r}, c{r}, gstride;"]
    return body


add("GROUP_U8", col_addr(4) + ["ld.shared.u32 w8, [ad];"] + group_body([f"bfe.u32 c{r}, w8, {8 * r}, 8;" for r in range(R)]),
    A_COL, A_NONE, PHYS["U8"])
add("GROUP_U16", col_addr(8) + ["ld.shared.v2.u32 {w8, w9}, [ad];"] + group_body(
    ["and.b32 c0, w8, 0xffff;", "shr.u32 c1, w8, 16;", "and.b32 c2, w9, 0xffff;", "shr.u32 c3, w9, 16;"]), A_COL, A_NONE, PHYS["U16"])
add("GROUP_U32", col_addr(16) + ["ld.shared.v4.u32 {c0, c1, c2, c3}, [ad];"] + group_body(["", "", "", ""]), A_COL, A_NONE, PHYS["U32"])

# ---- pushes ----------------------------------------------------------------------------------------
for d in range(DEPTH):
    for kind, phys in (("F32", "F32"), ("F64", "F64"), ("I32F", "I32"), ("I64", "I64"), ("I32", "I32")):
        add(f"LD_{kind}_D{d}", load_rows(kind, f"s{d}"), A_COL, A_NONE, PHYS[phys], +1, d)
    add(f"CONST_D{d}", const_load("c64", "a1") + [f"mov.b64 s{d}_{r}, c64;" for r in range(R)], A_CONST, A_NONE, -1, +1, d)
    for k in range(NTEMPS):
        add(f"GET{k}_D{d}", [f"mov.b64 s{d}_{r}, t{k}_{r};" for r in range(R)], A_NONE, A_NONE, -1, +1, d)

# ---- top-of-stack ops ------------------------------------------------------------------------------
for d in range(1, DEPTH + 1):
    top = d - 1
    for k in range(NTEMPS):
        add(f"TEE{k}_D{d}", [f"mov.b64 t{k}_{r}, s{top}_{r};" for r in range(R)], A_NONE, A_NONE, -1, 0, d)
    add(f"RSUBC_D{d}", const_load("c64", "a1") + [f"sub.f64 s{top}_{r}, c64, s{top}_{r};" for r in range(R)], A_CONST, A_NONE, -1, 0, d)
    add(f"ADDC_D{d}", const_load("c64", "a1") + [f"add.f64 s{top}_{r}, s{top}_{r}, c64;" for r in range(R)], A_CONST, A_NONE, -1, 0, d)
    add(f"MULC_D{d}", const_load("c64", "a1") + [f"mul.f64 s{top}_{r}, s{top}_{r}, c64;" for r in range(R)], A_CONST, A_NONE, -1, 0, d)

# ---- binary f64 arithmetic -------------------------------------------------------------------------
for d in range(2, DEPTH + 1):
    a, b = d - 2, d - 1
    for name, op in (("ADDF", "add"), ("SUBF", "sub"), ("MULF", "mul")):
        add(f"{name}_D{d}", [f"{op}.f64 s{a}_{r}, s{a}_{r}, s{b}_{r};" for r in range(R)], A_NONE, A_NONE, -1, -1, d)

# ---- aggregation -----------------------------------------------------------------------------------
def agg_rmw(kind: str, val: str) -> list[str]:
    body = [f"mad.lo.u32 base, a1, {SLOT_STRIDE}, accb;"]
    for r in range(R):
        v = val.format(r=r)
        body.append(f"add.u32 ad2, base, go{r};")
        pr = f"@pv{r} "  # rows that failed a filter (or lie past the end of the relation) touch nothing
        if kind == "SUMF":   # only the store is predicated: a predicated load + add would be if-converted into selects
            body += ["ld.shared.f64 xf, [ad2];", f"add.f64 xf, xf, {v};", pr + "st.shared.f64 [ad2], xf;"]
        elif kind == "SUMI":
            body += ["ld.shared.u64 xi, [ad2];", f"add.s64 xi, xi, {v};", pr + "st.shared.u64 [ad2], xi;"]
        elif kind in ("MINF", "MAXF"):
            cmp_ = "lt" if kind == "MINF" else "gt"
            body += ["ld.shared.f64 xf, [ad2];", f"setp.{cmp_}.f64 q, {v}, xf;", f"selp.f64 xf, {v}, xf, q;", pr + "st.shared.f64 [ad2], xf;"]
        else:
            cmp_ = "lt" if kind == "MINI" else "gt"
            body += ["ld.shared.u64 xi, [ad2];", f"setp.{cmp_}.s64 q, {v}, xi;", f"selp.b64 xi, {v}, xi, q;", pr + "st.shared.u64 [ad2], xi;"]
    return body


for d in range(1, DEPTH + 1):
    top = d - 1
    for kind in ("SUMF", "SUMI", "MINF", "MAXF", "MINI", "MAXI"):
        add(f"AGG_{kind}_D{d}", agg_rmw(kind, f"s{top}_{{r}}"), A_SLOT, A_NONE, -1, -1, d, AGG_KIND[kind])
    add(f"AGGK_SUMF_D{d}", agg_rmw("SUMF", f"s{top}_{{r}}"), A_SLOT, A_NONE, -1, 0, d, AGG_KIND["SUMF"])  # keep the value on the stack

# ---- fused forms (fewer dispatches for the common f64 expression shapes) --------------------------------
FLOAT_COLS = (("F32", "F32"), ("F64", "F64"), ("I32F", "I32"))
for d in range(DEPTH):  # push op(column, const): a1 = column, a2 = constant
    for kind, phys in FLOAT_COLS:
        for name, expr in (("LDADDC", "add.f64 s{d}_{r}, x_{r}, c64;"), ("LDRSUBC", "sub.f64 s{d}_{r}, c64, x_{r};"),
                           ("LDMULC", "mul.f64 s{d}_{r}, x_{r}, c64;")):
            add(f"{name}_{kind}_D{d}", load_rows(kind, "x") + const_load("c64", "a2") + [expr.format(d=d, r=r) for r in range(R)],
                A_COL, A_CONST, PHYS[phys], +1, d)
for d in range(1, DEPTH + 1):  # top = top op column / top = top op temporary
    top = d - 1
    forms = (("ADD", "add.f64 s{t}_{r}, s{t}_{r}, {v};"), ("SUB", "sub.f64 s{t}_{r}, s{t}_{r}, {v};"),
             ("RSUB", "sub.f64 s{t}_{r}, {v}, s{t}_{r};"), ("MUL", "mul.f64 s{t}_{r}, s{t}_{r}, {v};"))
    for kind, phys in FLOAT_COLS:
        for name, expr in forms:
            add(f"{name}COL_{kind}_D{d}", load_rows(kind, "x") + [expr.format(t=top, r=r, v=f"x_{r}") for r in range(R)],
                A_COL, A_NONE, PHYS[phys], 0, d)
    for k in range(NTEMPS):
        for name, expr in forms:
            add(f"{name}T{k}_D{d}", [expr.format(t=top, r=r, v=f"t{k}_{r}") for r in range(R)], A_NONE, A_NONE, -1, 0, d)
        # SUM the top of stack into slot a1, keep a copy in temporary k, pop
        add(f"AGGT{k}_SUMF_D{d}", [f"mov.b64 t{k}_{r}, s{top}_{r};" for r in range(R)] + agg_rmw("SUMF", f"s{top}_{{r}}"),
            A_SLOT, A_NONE, -1, -1, d, AGG_KIND["SUMF"])

count_body = [f"mad.lo.u32 base, a1, {SLOT_STRIDE}, accb;"]
for r in range(R):
    count_body += [f"add.u32 ad2, base, go{r};", "ld.shared.u64 xi, [ad2];", "add.u64 xi, xi, 1;", f"@pv{r} st.shared.u64 [ad2], xi;"]
add("COUNT", count_body, A_SLOT, agg=AGG_KIND["SUMI"])

for kind, phys in (("F32", "F32"), ("F64", "F64"), ("I32F", "I32")):  # fused load + SUM: a1 = column, a2 = slot
    body = load_rows(kind, "x") + [f"mad.lo.u32 base, a2, {SLOT_STRIDE}, accb;"]
    for r in range(R):
        body += [f"add.u32 ad2, base, go{r};", "ld.shared.f64 xf, [ad2];", f"add.f64 xf, xf, x_{r};", f"@pv{r} st.shared.f64 [ad2], xf;"]
    add(f"AGGCOL_{kind}", body, A_COL, A_SLOT, PHYS[phys], agg=AGG_KIND["SUMF"])


def ptx() -> str:
    regs = []
    for d in range(DEPTH):
        regs.append(".reg .b64 " + ", ".join(f"s{d}_{r}" for r in range(R)) + ";")
    for k in range(NTEMPS):
        regs.append(".reg .b64 " + ", ".join(f"t{k}_{r}" for r in range(R)) + ";")
    regs += [
        ".reg .b64 x_0, x_1, x_2, x_3, c64, xi;",
        ".reg .f64 xf;",
        ".reg .f32 f0, f1, f2, f3;",
        ".reg .b32 i0, i1, i2, i3, c0, c1, c2, c3, go0, go1, go2, go3, w8, w9;",
        ".reg .b32 sb, accb, pc, cstb, lane, ng, gstride, vm, w, wn, h, a1, a2, ad, ad2, base, tb;",
        ".reg .pred pv0, pv1, pv2, pv3, q;",
    ]
    lines = ["{"] + regs
    lines += ["mov.u32 sb, %0;", "mov.u32 accb, %1;", "mov.u32 pc, %2;", "mov.u32 cstb, %3;", "mov.u32 lane, %4;",
              "mov.u32 ng, %5;", "mov.u32 gstride, %6;", "mov.u32 vm, %7;", "ld.shared.u32 wn, [pc];"]
    for r in range(R):
        lines += [f"and.b32 tb, vm, {1 << r};", f"setp.ne.u32 pv{r}, tb, 0;", f"mov.u32 go{r}, 0;"]
    # stack slots / temporaries are written before they are read (the host validates the depth of every instruction)
    lines.append("RV_TABLE: .branchtargets " + ", ".join(f"RV_H{i}" for i in range(len(handlers))) + ";")
    # dispatch: the next instruction word is fetched one instruction ahead so its shared-memory latency overlaps the handler
    lines += ["RV_NEXT:", "mov.b32 w, wn;", "ld.shared.u32 wn, [pc+4];", "add.u32 pc, pc, 4;", "and.b32 h, w, 255;",
              "bfe.u32 a1, w, 8, 12;", "shr.u32 a2, w, 20;", "brx.idx h, RV_TABLE;"]
    for i, hnd in enumerate(handlers):
        lines.append(f"RV_H{i}:  // {hnd['name']}")
        lines += [ln for ln in hnd["body"] if ln]
        if hnd["name"] != "END":
            lines.append("bra.uni RV_NEXT;")
    lines += ["RV_DONE:", "}"]
    out = []
    for ln in lines:
        code = ln.split("//")[0].rstrip()
        out.append('"' + code.replace('"', '\\"') + '\\n\\t"' + ("  // " + ln.split("//", 1)[1].strip() if "//" in ln else ""))
    return "// GENERATED by gen_regvm.py -- do not edit.\n" + "\n".join(out) + "\n"


def header() -> str:
    assert len(handlers) <= 256
    lines = ["// GENERATED by gen_regvm.py -- do not edit.  Handler ids and operand metadata of the regvm interpreter.",
             "#pragma once", f"#define MSC_RV_ROWS {R}", f"#define MSC_RV_MAX_DEPTH {DEPTH}", f"#define MSC_RV_MAX_TEMPS {NTEMPS}",
             f"#define MSC_RV__COUNT {len(handlers)}"]
    for i, hnd in enumerate(handlers):
        lines.append(f"#define MSC_RV_{hnd['name']} {i}")
    lines += ["#define MSC_RV_ARG_NONE 0", "#define MSC_RV_ARG_COL 1", "#define MSC_RV_ARG_CONST 2", "#define MSC_RV_ARG_SLOT 3",
              "struct msc_rv_info { const char* name; signed char a1, a2, phys, delta, depth, agg; };",
              "static const msc_rv_info MSC_RV_INFO[MSC_RV__COUNT] = {"]
    for hnd in handlers:
        lines.append(f'  {{"{hnd["name"]}", {hnd["a1"]}, {hnd["a2"]}, {hnd["phys"]}, {hnd["delta"]}, {hnd["depth"]}, {hnd["agg"]}}},')
    lines.append("};")
    return "\n".join(lines) + "\n"


def main() -> None:
    here = Path(__file__).resolve().parent
    (here / "regvm_ptx.inc").write_text(ptx())
    (here / "regvm_handlers.h").write_text(header())
    print(f"{len(handlers)} handlers")


if __name__ == "__main__":
    main()
