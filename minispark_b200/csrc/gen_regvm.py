#!/usr/bin/env python
"""Generator of the register-resident interpreter core ("regvm") of the dense aggregate scan.

nvcc cannot keep an interpreter's value stack in registers across a dispatch loop without drowning
it in moves (profiles/r01_scan_kernel_ncu_summary.json, prof_r1a: 70 % MOV).  This script therefore
emits the interpreter as hand-scheduled PTX: a postfix stack machine whose stack depth is static per
instruction, so every (operation, depth) pair is its own handler working on fixed registers, and
handlers are reached through one ``brx.idx`` jump table (threaded dispatch).  Outputs, all committed:

  regvm_ptx_ng<N>.inc   the PTX text of variant N, the body of the ``asm volatile`` block in scan_regvm.cuh
  regvm_handlers.h      handler ids + operand metadata (validated / patched on the host, parsed by Python)

A lane owns R = 8 rows of a 256-row warp tile: rows 4*lane .. 4*lane+3 of each 128-row half, so a
32-bit column is read with one conflict-free 128-bit shared-memory load per half.  One dispatch
(about 10 SASS instructions) is amortised over those 8 rows (prof_r1f: with 4 rows per lane, dispatch
and per-tile bookkeeping were 54 % of the executed instructions).

Instruction word: handler id | a1 << 8 | a2 << 16 (8 bits each; column operands are the column's
offset in the warp stage / 128).  Register conventions inside the block: s<d>_<r> stack slot d of row
r (64-bit), t<k>_<r> temporaries, vm the 8-bit row-validity mask.

Variants (one PTX file each, same handler ids):
  NG = 0   generic.  GROUP turns each row's dictionary code into go<r>, the byte offset of its group in the
           lane's private accumulator block; rows that fail a filter (or lie past the end of the relation)
           go to a trash group, so the updates need no predicates.  One shared-memory read-modify-write
           per row and aggregate.
  NG 1..4  exactly NG groups.  prof_r1g showed the generic form bound by those updates (65 % of the
           shared-memory pipe, short-scoreboard stalls on the serial ld -> add -> st chains), so here GROUP
           builds one-hot f64 masks m<r>_<g> (1.0 when row r is valid and in group g) and the SUM / COUNT
           handlers first reduce the lane's 8 rows per group in registers (two fma chains of 4 rows each per
           group) -- fma(v, 1.0, s) == s + v and fma(v, 0.0, s) == s exactly for finite v -- then update each
           group's accumulator once: NG
           independent updates instead of 8 dependent ones, and no trash group.  A non-finite input would
           leak into the other groups through v * 0.0 = NaN; it always leaves a non-finite SUM behind, so
           the host detects it and reruns the scan on the generic kernel (scan.cu).  MIN / MAX / integer
           SUM exist only in the generic variant.

    python minispark_b200/csrc/gen_regvm.py      # rewrites the files next to this script
"""

from __future__ import annotations

from pathlib import Path

R = 8
HALF = R // 2          # rows per 128-row half of the tile
DEPTH = 3
NTEMPS = 2
MAX_NG = 4             # masked variants exist for 1..MAX_NG groups
SLOT_STRIDE = 128 * 8  # bytes between accumulator slots: NT lanes x 8 B
COL_UNIT = 128         # column operands address the warp stage in units of 128 bytes
TAIL_DISPATCH = False  # True: replicate the dispatch at the end of every handler (threaded code); measured no gain (0.502 vs 0.495 ms), 4x the code

# operand kinds for the host-side validator
A_NONE, A_COL, A_CONST, A_SLOT = 0, 1, 2, 3
PHYS = {"U8": 0, "U16": 1, "U32": 2, "I32": 3, "I64": 4, "F32": 5, "F64": 6}
ROWS = range(R)
AGG_KIND = {"SUMF": 0, "SUMI": 1, "MINF": 2, "MAXF": 3, "MINI": 4, "MAXI": 5}  # MSC_AGG_* of include/minispark_cuda.h
F_FILTER, F_GROUP, F_GENERIC_ONLY = 1, 2, 4  # narrows the row mask / fixes the rows' groups / not in the masked variants
ONE, ZERO = "0d3FF0000000000000", "0d0000000000000000"


def col_addr(lane_bytes: int, arg: str = "a1") -> list[str]:
    """ad = address of this lane's first row of the column operand (half 0); sbl<k> = stage base + lane * k."""
    return [f"mad.lo.u32 ad, {arg}, {COL_UNIT}, sbl{lane_bytes};"]


def load_rows(kind: str, dst: str, arg: str = "a1", off: int = 0) -> list[str]:
    """Load this lane's 8 rows of a staged column (`off` bytes past the operand's column) into registers <dst>_0..7 as
    64-bit values."""
    if kind in ("F32", "I32F", "I32"):
        half = 128 * 4
        if kind == "F32":
            out = col_addr(16, arg) + [f"ld.shared.v4.f32 {{f0, f1, f2, f3}}, [ad+{off}];", f"ld.shared.v4.f32 {{f4, f5, f6, f7}}, [ad+{off + half}];"]
            return out + [f"cvt.f64.f32 {dst}_{r}, f{r};" for r in ROWS]
        out = col_addr(16, arg) + [f"ld.shared.v4.s32 {{i0, i1, i2, i3}}, [ad+{off}];", f"ld.shared.v4.s32 {{i4, i5, i6, i7}}, [ad+{off + half}];"]
        cvt = "cvt.rn.f64.s32" if kind == "I32F" else "cvt.s64.s32"
        return out + [f"{cvt} {dst}_{r}, i{r};" for r in ROWS]
    if kind in ("F64", "I64"):
        half = 128 * 8
        return col_addr(32, arg) + [f"ld.shared.v2.b64 {{{dst}_0, {dst}_1}}, [ad+{off}];", f"ld.shared.v2.b64 {{{dst}_2, {dst}_3}}, [ad+{off + 16}];",
                                    f"ld.shared.v2.b64 {{{dst}_4, {dst}_5}}, [ad+{off + half}];",
                                    f"ld.shared.v2.b64 {{{dst}_6, {dst}_7}}, [ad+{off + half + 16}];"]
    raise ValueError(kind)


TILE_BYTES = {"F32": 256 * 4, "I32F": 256 * 4, "F64": 256 * 8}  # one column's bytes in a warp tile


def const_load(reg: str, arg: str) -> list[str]:
    return [f"mad.lo.u32 ad, {arg}, 8, cstb;", f"ld.shared.b64 {reg}, [ad];"]


def build(NG: int) -> list[dict]:
    """Handlers of variant NG, in id order (the order, names and metadata do not depend on NG)."""
    handlers: list[dict] = []
    masked = NG > 0

    def add(name: str, body: list[str], a1: int = A_NONE, a2: int = A_NONE, phys: int = -1, delta: int = 0, depth: int = -1,
            agg: int = -1, flags: int = 0, span: int = 1) -> None:
        """depth: stack depth required before the instruction (-1: any); agg: accumulator kind the slot operand must have;
        span: the column / slot operands each name `span` consecutive columns / slots."""
        if masked and (flags & F_GENERIC_ONLY):
            body = ["trap;"]  # the host never sends such a program to a masked variant
        handlers.append(dict(name=name, body=body, a1=a1, a2=a2, phys=phys, delta=delta, depth=depth, agg=agg, flags=flags, span=span))

    # ---- END -------------------------------------------------------------------------------------
    add("END", ["bra.uni RV_DONE;"])

    # ---- column <cmp> const -> filter --------------------------------------------------------------
    for ty, load_kind, setp_ty, phys in (("I64", "I64", "s64", "I64"), ("I32", "I32", "s64", "I32"), ("F32", "F32", "f64", "F32"),
                                         ("F64", "F64", "f64", "F64")):
        for cmp_ in ("lt", "le", "gt", "ge", "eq", "ne"):
            body = load_rows(load_kind, "x") + const_load("c64", "a2")
            for r in ROWS:
                body += [f"setp.{cmp_}.{setp_ty} q, x_{r}, c64;", f"@!q and.b32 vm, vm, {0xff ^ (1 << r)};"]
            add(f"CMPCOL_{cmp_.upper()}_{ty}", body, A_COL, A_CONST, PHYS[phys], flags=F_FILTER)

    # ---- dictionary code column -> dense group -----------------------------------------------------
    def group_body(extract: list[list[str]]) -> list[str]:
        body = []
        for r in ROWS:
            body += extract[r]
            body += [f"and.b32 tb, vm, {1 << r};"]
            if masked:  # one-hot masks; a masked-out row, or a code outside [0, NG), is in no group
                body += ["setp.ne.u32 qv, tb, 0;"]
                for g in range(NG):
                    body += [f"setp.eq.and.u32 q, c{r}, {g}, qv;", f"selp.f64 m{r}_{g}, {ONE}, {ZERO}, q;"]
            else:       # a masked-out row, or a code outside [0, ngroups), goes to the trash group (never exported)
                body += [f"setp.lt.u32 q, c{r}, ng;", "setp.ne.and.u32 q, tb, 0, q;",
                         f"mul.lo.u32 go{r}, c{r}, gstride;", f"selp.b32 go{r}, go{r}, trash, q;"]
        return body

    add("GROUP_U8", col_addr(4) + ["ld.shared.u32 w8, [ad];", "ld.shared.u32 w9, [ad+128];"] + group_body(
        [[f"bfe.u32 c{r}, {'w8' if r < HALF else 'w9'}, {8 * (r % HALF)}, 8;"] for r in ROWS]), A_COL, A_NONE, PHYS["U8"], flags=F_GROUP)
    add("GROUP_U16", col_addr(8) + ["ld.shared.v2.u32 {w8, w9}, [ad];", "ld.shared.v2.u32 {w10, w11}, [ad+256];"] + group_body(
        [["and.b32 c0, w8, 0xffff;"], ["shr.u32 c1, w8, 16;"], ["and.b32 c2, w9, 0xffff;"], ["shr.u32 c3, w9, 16;"],
         ["and.b32 c4, w10, 0xffff;"], ["shr.u32 c5, w10, 16;"], ["and.b32 c6, w11, 0xffff;"], ["shr.u32 c7, w11, 16;"]]),
        A_COL, A_NONE, PHYS["U16"], flags=F_GROUP)
    add("GROUP_U32", col_addr(16) + ["ld.shared.v4.u32 {c0, c1, c2, c3}, [ad];", "ld.shared.v4.u32 {c4, c5, c6, c7}, [ad+512];"] + group_body(
        [[] for _ in ROWS]), A_COL, A_NONE, PHYS["U32"], flags=F_GROUP)

    # ---- pushes ------------------------------------------------------------------------------------
    for d in range(DEPTH):
        for kind, phys in (("F32", "F32"), ("F64", "F64"), ("I32F", "I32"), ("I64", "I64"), ("I32", "I32")):
            add(f"LD_{kind}_D{d}", load_rows(kind, f"s{d}"), A_COL, A_NONE, PHYS[phys], +1, d)
        add(f"CONST_D{d}", const_load("c64", "a1") + [f"mov.b64 s{d}_{r}, c64;" for r in ROWS], A_CONST, A_NONE, -1, +1, d)
        for k in range(NTEMPS):
            add(f"GET{k}_D{d}", [f"mov.b64 s{d}_{r}, t{k}_{r};" for r in ROWS], A_NONE, A_NONE, -1, +1, d)

    # ---- top-of-stack ops --------------------------------------------------------------------------
    for d in range(1, DEPTH + 1):
        top = d - 1
        for k in range(NTEMPS):
            add(f"TEE{k}_D{d}", [f"mov.b64 t{k}_{r}, s{top}_{r};" for r in ROWS], A_NONE, A_NONE, -1, 0, d)
        add(f"RSUBC_D{d}", const_load("c64", "a1") + [f"sub.f64 s{top}_{r}, c64, s{top}_{r};" for r in ROWS], A_CONST, A_NONE, -1, 0, d)
        add(f"ADDC_D{d}", const_load("c64", "a1") + [f"add.f64 s{top}_{r}, s{top}_{r}, c64;" for r in ROWS], A_CONST, A_NONE, -1, 0, d)
        add(f"MULC_D{d}", const_load("c64", "a1") + [f"mul.f64 s{top}_{r}, s{top}_{r}, c64;" for r in ROWS], A_CONST, A_NONE, -1, 0, d)

    # ---- binary f64 arithmetic ---------------------------------------------------------------------
    for d in range(2, DEPTH + 1):
        a, b = d - 2, d - 1
        for name, op in (("ADDF", "add"), ("SUBF", "sub"), ("MULF", "mul")):
            add(f"{name}_D{d}", [f"{op}.f64 s{a}_{r}, s{a}_{r}, s{b}_{r};" for r in ROWS], A_NONE, A_NONE, -1, -1, d)

    # ---- aggregation -------------------------------------------------------------------------------
    def agg_rmw(kind: str, val: str, slot_arg: str = "a1", slot_off: int = 0) -> list[str]:
        """Generic variant: read-modify-write of the lane's private accumulator (slot, group of row r) for the 8 rows,
        in row order -- two rows of one lane may share a group, so the updates must not be reordered."""
        body = [f"mad.lo.u32 base, {slot_arg}, {SLOT_STRIDE}, accb;"] + ([f"add.u32 base, base, {slot_off};"] if slot_off else [])
        for r in ROWS:
            v = val.format(r=r)
            body.append(f"add.u32 ad2, base, go{r};")
            if kind == "SUMF":
                body += ["ld.shared.f64 xf, [ad2];", f"add.f64 xf, xf, {v};", "st.shared.f64 [ad2], xf;"]
            elif kind == "SUMI":
                body += ["ld.shared.u64 xi, [ad2];", f"add.s64 xi, xi, {v};", "st.shared.u64 [ad2], xi;"]
            elif kind == "COUNT":
                body += ["ld.shared.u64 xi, [ad2];", "add.u64 xi, xi, 1;", "st.shared.u64 [ad2], xi;"]
            elif kind in ("MINF", "MAXF"):
                cmp_ = "lt" if kind == "MINF" else "gt"
                body += ["ld.shared.f64 xf, [ad2];", f"setp.{cmp_}.f64 q, {v}, xf;", f"@q st.shared.f64 [ad2], {v};"]
            else:
                cmp_ = "lt" if kind == "MINI" else "gt"
                body += ["ld.shared.u64 xi, [ad2];", f"setp.{cmp_}.s64 q, {v}, xi;", f"@q st.shared.u64 [ad2], {v};"]
        return body

    def agg_masked(kind: str, val: str, slot_arg: str = "a1", slot_off: int = 0) -> list[str]:
        """Masked variant: per-group sums of the lane's 8 rows in registers (row order, one chain per group), then one
        update per group; the NG accumulator addresses are distinct, so loads, adds and stores are batched."""
        body = []
        if kind == "SUMF":  # two chains per group (rows 0-3, rows 4-7): 2 * NG independent fma chains of depth 4
            for r0, acc in ((0, "p"), (HALF, "ph")):
                for g in range(NG):
                    body.append(f"mul.f64 {acc}{g}, {val.format(r=r0)}, m{r0}_{g};")
            for j in range(1, HALF):
                for r0, acc in ((0, "p"), (HALF, "ph")):
                    for g in range(NG):
                        body.append(f"fma.rn.f64 {acc}{g}, {val.format(r=r0 + j)}, m{r0 + j}_{g}, {acc}{g};")
            for g in range(NG):
                body.append(f"add.f64 p{g}, p{g}, ph{g};")
        else:  # COUNT: the masks themselves, summed (exact: at most 8)
            for g in range(NG):
                body.append(f"add.f64 p{g}, m0_{g}, m1_{g};")
            for r in range(2, R):
                for g in range(NG):
                    body.append(f"add.f64 p{g}, p{g}, m{r}_{g};")
            for g in range(NG):
                body.append(f"cvt.rzi.s64.f64 n{g}, p{g};")
        body.append(f"mad.lo.u32 base, {slot_arg}, {SLOT_STRIDE}, accb;")
        if slot_off:
            body.append(f"add.u32 base, base, {slot_off};")
        for g in range(NG):
            body.append(f"mad.lo.u32 ag{g}, gstride, {g}, base;")
        if kind == "SUMF":
            body += [f"ld.shared.f64 q{g}f, [ag{g}];" for g in range(NG)]
            body += [f"add.f64 q{g}f, q{g}f, p{g};" for g in range(NG)]
            body += [f"st.shared.f64 [ag{g}], q{g}f;" for g in range(NG)]
        else:
            body += [f"ld.shared.u64 q{g}i, [ag{g}];" for g in range(NG)]
            body += [f"add.s64 q{g}i, q{g}i, n{g};" for g in range(NG)]
            body += [f"st.shared.u64 [ag{g}], q{g}i;" for g in range(NG)]
        return body

    def agg(kind: str, val: str, slot_arg: str = "a1", slot_off: int = 0) -> list[str]:
        return agg_masked(kind, val, slot_arg, slot_off) if masked and kind in ("SUMF", "COUNT") else agg_rmw(kind, val, slot_arg, slot_off)

    for d in range(1, DEPTH + 1):
        top = d - 1
        for kind in ("SUMF", "SUMI", "MINF", "MAXF", "MINI", "MAXI"):
            add(f"AGG_{kind}_D{d}", agg(kind, f"s{top}_{{r}}"), A_SLOT, A_NONE, -1, -1, d, AGG_KIND[kind],
                flags=0 if kind == "SUMF" else F_GENERIC_ONLY)
        add(f"AGGK_SUMF_D{d}", agg("SUMF", f"s{top}_{{r}}"), A_SLOT, A_NONE, -1, 0, d, AGG_KIND["SUMF"])  # keep the value on the stack

    # ---- fused forms (fewer dispatches for the common f64 expression shapes) ----------------------------
    float_cols = (("F32", "F32"), ("F64", "F64"), ("I32F", "I32"))
    for d in range(DEPTH):  # push op(column, const): a1 = column, a2 = constant
        for kind, phys in float_cols:
            for name, expr in (("LDADDC", "add.f64 s{d}_{r}, x_{r}, c64;"), ("LDRSUBC", "sub.f64 s{d}_{r}, c64, x_{r};"),
                               ("LDMULC", "mul.f64 s{d}_{r}, x_{r}, c64;")):
                add(f"{name}_{kind}_D{d}", load_rows(kind, "x") + const_load("c64", "a2") + [expr.format(d=d, r=r) for r in ROWS],
                    A_COL, A_CONST, PHYS[phys], +1, d)
    for d in range(1, DEPTH + 1):  # top = top op column / top = top op temporary
        top = d - 1
        forms = (("ADD", "add.f64 s{t}_{r}, s{t}_{r}, {v};"), ("SUB", "sub.f64 s{t}_{r}, s{t}_{r}, {v};"),
                 ("RSUB", "sub.f64 s{t}_{r}, {v}, s{t}_{r};"), ("MUL", "mul.f64 s{t}_{r}, s{t}_{r}, {v};"))
        for kind, phys in float_cols:
            for name, expr in forms:
                add(f"{name}COL_{kind}_D{d}", load_rows(kind, "x") + [expr.format(t=top, r=r, v=f"x_{r}") for r in ROWS],
                    A_COL, A_NONE, PHYS[phys], 0, d)
        for k in range(NTEMPS):
            for name, expr in forms:
                add(f"{name}T{k}_D{d}", [expr.format(t=top, r=r, v=f"t{k}_{r}") for r in ROWS], A_NONE, A_NONE, -1, 0, d)
            # SUM the top of stack into slot a1, keep a copy in temporary k, pop
            add(f"AGGT{k}_SUMF_D{d}", [f"mov.b64 t{k}_{r}, s{top}_{r};" for r in ROWS] + agg("SUMF", f"s{top}_{{r}}"),
                A_SLOT, A_NONE, -1, -1, d, AGG_KIND["SUMF"])

    add("COUNT", agg("COUNT", ""), A_SLOT, agg=AGG_KIND["SUMI"])

    for kind, phys in float_cols:  # fused load + SUM: a1 = column, a2 = slot
        add(f"AGGCOL_{kind}", load_rows(kind, "x") + agg("SUMF", "x_{r}", "a2"), A_COL, A_SLOT, PHYS[phys], agg=AGG_KIND["SUMF"])
    # SUM of 2..4 adjacent columns into adjacent slots in one dispatch (SUM(a), SUM(b), SUM(c) of one query): fewer
    # dispatches, and independent load / convert / reduce sequences for the scheduler to overlap
    for n in (2, 3, 4):
        for kind, phys in float_cols:
            body = []
            for j in range(n):
                body += load_rows(kind, "x", off=j * TILE_BYTES[kind]) + agg("SUMF", "x_{r}", "a2", j * SLOT_STRIDE)
            add(f"AGGCOL{n}_{kind}", body, A_COL, A_SLOT, PHYS[phys], agg=AGG_KIND["SUMF"], span=n)
    return handlers


def ptx(NG: int) -> str:
    handlers = build(NG)
    regs = []
    for d in range(DEPTH):
        regs.append(".reg .b64 " + ", ".join(f"s{d}_{r}" for r in ROWS) + ";")
    for k in range(NTEMPS):
        regs.append(".reg .b64 " + ", ".join(f"t{k}_{r}" for r in ROWS) + ";")
    regs += [
        ".reg .b64 " + ", ".join(f"x_{r}" for r in ROWS) + ", c64, xi;",
        ".reg .f64 xf;",
        ".reg .f32 " + ", ".join(f"f{r}" for r in ROWS) + ";",
        ".reg .b32 " + ", ".join(f"i{r}" for r in ROWS) + ";",
        ".reg .b32 " + ", ".join(f"c{r}" for r in ROWS) + ";",
        ".reg .b32 w8, w9, w10, w11;",
        ".reg .b32 sb, sbl4, sbl8, sbl16, sbl32, accb, pc, cstb, lane, ng, gstride, trash, vm, wn, h, a1, a2, ad, ad2, base, tb;",
        ".reg .pred q, qv;",
    ]
    if NG > 0:
        for r in ROWS:
            regs.append(".reg .f64 " + ", ".join(f"m{r}_{g}" for g in range(NG)) + ";")
        regs += [".reg .f64 " + ", ".join(f"p{g}, ph{g}, q{g}f" for g in range(NG)) + ";",
                 ".reg .b64 " + ", ".join(f"n{g}, q{g}i" for g in range(NG)) + ";",
                 ".reg .b32 " + ", ".join(f"ag{g}" for g in range(NG)) + ";"]
    else:
        regs.append(".reg .b32 " + ", ".join(f"go{r}" for r in ROWS) + ";")
    lines = ["{"] + regs
    lines += ["mov.u32 sb, %0;", "mov.u32 accb, %1;", "mov.u32 pc, %2;", "mov.u32 cstb, %3;", "mov.u32 lane, %4;",
              "mov.u32 ng, %5;", "mov.u32 gstride, %6;", "mov.u32 vm, %7;", "ld.shared.u32 wn, [pc];"]
    lines += [f"mad.lo.u32 sbl{k}, lane, {k}, sb;" for k in (4, 8, 16, 32)]
    if NG == 0:
        lines += ["mul.lo.u32 trash, ng, gstride;"] + [f"mov.u32 go{r}, trash;" for r in ROWS]
    # stack slots / temporaries / masks are written before they are read (the host validates depths and GROUP-before-aggregate)
    lines.append("RV_TABLE: .branchtargets " + ", ".join(f"RV_H{i}" for i in range(len(handlers))) + ";")
    # dispatch: the word of the NEXT instruction is already in wn (its shared-memory latency overlapped the previous
    # handler); its fields are extracted, then wn is reloaded in place -- no register moves (prof_r1j: the moves of a
    # two-deep prefetch cost 26 of 794 instructions per tile and did not remove the wait, which only moved onto them)
    dispatch = ["and.b32 h, wn, 255;", "bfe.u32 a1, wn, 8, 8;", "bfe.u32 a2, wn, 16, 8;", "ld.shared.u32 wn, [pc+4];",
                "add.u32 pc, pc, 4;", "brx.idx h, RV_TABLE;"]
    lines += ["RV_NEXT:"] + dispatch
    for i, hnd in enumerate(handlers):
        lines.append(f"RV_H{i}:  // {hnd['name']}")
        lines += [ln for ln in hnd["body"] if ln]
        if hnd["name"] != "END":
            # threaded code: every handler ends with its own copy of the dispatch, so the scheduler can start the
            # next instruction's decode and jump-table load while this handler's arithmetic is still in flight
            lines += dispatch if TAIL_DISPATCH else ["bra.uni RV_NEXT;"]
    lines += ["RV_DONE:", "}"]
    out = []
    for ln in lines:
        code = ln.split("//")[0].rstrip()
        out.append('"' + code.replace('"', '\\"') + '\\n\\t"' + ("  // " + ln.split("//", 1)[1].strip() if "//" in ln else ""))
    return f"// GENERATED by gen_regvm.py (variant NG = {NG}) -- do not edit.\n" + "\n".join(out) + "\n"


def header() -> str:
    handlers = build(0)
    assert len(handlers) <= 256
    for ng in range(1, MAX_NG + 1):
        meta = lambda hs: [tuple(h[k] for k in ("name", "a1", "a2", "phys", "delta", "depth", "agg", "flags", "span")) for h in hs]  # noqa: E731
        assert meta(build(ng)) == meta(handlers)
    lines = ["// GENERATED by gen_regvm.py -- do not edit.  Handler ids and operand metadata of the regvm interpreter.",
             "#pragma once", f"#define MSC_RV_ROWS {R}", f"#define MSC_RV_MAX_DEPTH {DEPTH}", f"#define MSC_RV_MAX_TEMPS {NTEMPS}",
             f"#define MSC_RV_COL_UNIT {COL_UNIT}", f"#define MSC_RV_MAX_NG {MAX_NG}", f"#define MSC_RV__COUNT {len(handlers)}"]
    for i, hnd in enumerate(handlers):
        lines.append(f"#define MSC_RV_{hnd['name']} {i}")
    lines += ["#define MSC_RV_ARG_NONE 0", "#define MSC_RV_ARG_COL 1", "#define MSC_RV_ARG_CONST 2", "#define MSC_RV_ARG_SLOT 3",
              f"#define MSC_RV_F_FILTER {F_FILTER}", f"#define MSC_RV_F_GROUP {F_GROUP}", f"#define MSC_RV_F_GENERIC_ONLY {F_GENERIC_ONLY}",
              "struct msc_rv_info { const char* name; signed char a1, a2, phys, delta, depth, agg, flags, span; };",
              "static const msc_rv_info MSC_RV_INFO[MSC_RV__COUNT] = {"]
    for hnd in handlers:
        lines.append(f'  {{"{hnd["name"]}", {hnd["a1"]}, {hnd["a2"]}, {hnd["phys"]}, {hnd["delta"]}, {hnd["depth"]}, {hnd["agg"]}, {hnd["flags"]}, {hnd["span"]}}},')
    lines.append("};")
    return "\n".join(lines) + "\n"


def main() -> None:
    here = Path(__file__).resolve().parent
    for ng in range(MAX_NG + 1):
        (here / f"regvm_ptx_ng{ng}.inc").write_text(ptx(ng))
    (here / "regvm_handlers.h").write_text(header())
    stale = here / "regvm_ptx.inc"
    if stale.exists():
        stale.unlink()
    print(f"{len(build(0))} handlers, variants NG = 0..{MAX_NG}")


if __name__ == "__main__":
    main()
