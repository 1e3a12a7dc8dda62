"""Execution engines: the ``ExecutionEngine`` contract and the B200 ``CudaExecutionEngine``.

``ExecutionEngine`` mirrors the reference's abstract base (``src/mini_spark/execution.py:40-62``):
``execute_full_task(task) -> list[JobResult]``, ``collect_results``, ``sql`` and the context-manager
protocol.  ``CudaExecutionEngine`` is the drop-in third engine next to the reference's
``PythonExecutionEngine`` (``execution.py:65-93``) and ``ThreadEngine`` (``execution.py:96-123``):

    with CudaExecutionEngine() as engine:
        rows = DataFrame(engine).table(path).group_by(Col("k")).agg(F.count()).collect()

It lowers the task tree (:mod:`minispark_b200.lowering`) and drives the CUDA operators of
``libminispark_cuda.so`` through ctypes.  Every row-level operation runs on the GPU; there is no CPU
execution path (construction fails without the library and a B200).
"""

from __future__ import annotations

import ctypes as C
import math
import os
import shutil
import tempfile
import time
import uuid
from abc import ABC, abstractmethod
from contextlib import AbstractContextManager
from copy import deepcopy
from dataclasses import dataclass, field
from pathlib import Path
from typing import TYPE_CHECKING, Any, Callable, Iterable, Optional

from . import lowering as L
from . import native as N
from .distributed import Comm, PeerShuffle, boundary_plan, invert_code_maps, range_bounds, shard_blocks, unify_keys
from .constants import ColumnType, Row, Schema
from .io import BlockFile
from .jobs import JobResult, OutputFile
from .utils import TRACER

if TYPE_CHECKING:
    from .dataframe import DataFrame


class ExecutionError(Exception):
    def __init__(self, message: str = "Execution failed") -> None:
        super().__init__(message)


class ExecutionEngine(AbstractContextManager, ABC):
    """The plugin interface a ``DataFrame`` talks to (reference ``execution.py:40-62``)."""

    @abstractmethod
    def execute_full_task(self, full_task: Any) -> list[JobResult]: ...

    def collect_results(self, results: list[JobResult], limit: float = math.inf) -> Iterable[Row]:
        seen: set[OutputFile] = set()
        for result in results:
            for out in result.output_files:
                if out in seen:
                    continue
                seen.add(out)
                for row in BlockFile(out.file_path).read_data_rows():
                    yield row
                    limit -= 1
                    if limit <= 0:
                        return

    def sql(self, query: str) -> "DataFrame":
        from .parser import parse_sql  # noqa: PLC0415

        # the same text parses to the same (immutable-style) DataFrame: keep it, and with it the fingerprint the engine's plan
        # cache leaves on its task tree -- a repeated engine.sql(text).collect() skips parsing and fingerprinting
        cache = self.__dict__.setdefault("_sql_cache", {})
        df = cache.get(query)
        if df is None:
            df = parse_sql(query)
            df.engine = self
            if len(cache) < 256:
                cache[query] = df
        return df


# ------------------------------------------------------------------------------------------------
# device-side objects
# ------------------------------------------------------------------------------------------------
class DictHandle:
    """A device string dictionary plus host-side caches of the tables derived from it."""

    _serials = 0

    def __init__(self, ctx: N.Context, handle: Optional[int] = None) -> None:
        DictHandle._serials += 1
        self.serial = DictHandle._serials  # never reused, unlike id(): cache keys of derived tables must not alias
        self.ctx = ctx
        if handle is None:
            h = C.c_void_p()
            ctx.call("msc_dict_create", C.byref(h))
            handle = h.value
        self.handle = handle
        self._luts: dict[tuple, int] = {}
        self._codes: dict[tuple, int] = {}
        self._translated_from: list[tuple["DictHandle", tuple]] = []  # translation tables other dictionaries cache INTO this one
        self.persistent = False  # a table column's dictionary (lives as long as the engine): results derived from it may be kept
        self._unified: Optional[tuple[int, "DictHandle"]] = None  # (size it was built at, the same entries on every rank)
        self.same_on_all_ranks = False  # built from the union of all ranks' entries (_unified_dictionary): same size and codes everywhere

    @property
    def size(self) -> int:
        n = C.c_uint32()
        nb = C.c_uint64()
        self.ctx.check(self.ctx.lib.msc_dict_size(C.c_void_p(self.handle), C.byref(n), C.byref(nb)))
        return n.value

    def literal_code(self, text: str, insert: bool = False) -> int:
        key = (text, self.size)
        if key not in self._codes or insert:
            raw = text.encode("utf-8")
            code = C.c_int64()
            self.ctx.call("msc_dict_lookup", C.c_void_p(self.handle), raw, len(raw), int(insert), C.byref(code))
            self._codes[(text, self.size)] = code.value
            return code.value
        return self._codes[key]

    def like_lut(self, pattern: str) -> int:
        key = ("like", pattern, self.size)
        if key not in self._luts:
            raw = pattern.encode("utf-8")
            out = C.c_void_p()
            self.ctx.call("msc_dict_like", C.c_void_p(self.handle), raw, len(raw), C.byref(out))
            self._luts[key] = out.value
        return self._luts[key]

    def _upload(self, array: Any) -> int:
        ptr = C.c_void_p()
        self.ctx.call("msc_dev_alloc", max(array.nbytes, 16), C.byref(ptr))
        if array.nbytes:
            self.ctx.call("msc_memcpy_h2d", ptr, array.ctypes.data_as(C.c_void_p), array.nbytes)
        return ptr.value

    def order_lut(self, op: str, text: str) -> int:
        """u8 LUT over the entries: 1 where ``entry <op> text`` (op in lt / le / gt / ge), Python string order."""
        import operator as _op

        import numpy as np  # noqa: PLC0415

        key = ("order", op, text, self.size)
        if key not in self._luts:
            fn = getattr(_op, op)
            self._luts[key] = self._upload(np.fromiter((fn(entry, text) for entry in self.export()), dtype=np.uint8, count=self.size))
        return self._luts[key]

    def rank_lut(self, other: "DictHandle") -> int:
        """u32 LUT: rank of every entry in the sorted union of this dictionary's and ``other``'s entries."""
        import numpy as np  # noqa: PLC0415

        key = ("rank", other.serial, self.size, other.size)
        if key not in self._luts:
            mine = self.export()
            order = {s: i for i, s in enumerate(sorted(set(mine) | set(other.export() if other is not self else mine)))}
            self._luts[key] = self._upload(np.fromiter((order[s] for s in mine), dtype=np.uint32, count=len(mine)))
            other._translated_from.append((self, key))  # dies with `other`
        return self._luts[key]

    def translate_lut(self, target: "DictHandle", insert: bool = False) -> int:
        """u32 LUT mapping this dictionary's codes to ``target``'s codes."""
        key = ("tr", target.serial, self.size, target.size, insert)
        if key not in self._luts or insert:
            out = C.c_void_p()
            self.ctx.call("msc_dict_translate", C.c_void_p(self.handle), C.c_void_p(target.handle), int(insert), C.byref(out))
            key = ("tr", target.serial, self.size, target.size, insert)
            stale = self._luts.pop(key, None)
            if stale:
                self.ctx.dev_free(stale)
            self._luts[key] = out.value
            target._translated_from.append((self, key))
            return out.value
        return self._luts[key]

    def export(self) -> list[str]:
        import numpy as np  # noqa: PLC0415

        n = self.size
        if n == 0:
            return []
        nb = C.c_uint64()
        self.ctx.check(self.ctx.lib.msc_dict_size(C.c_void_p(self.handle), None, C.byref(nb)))
        lens = np.zeros(n, dtype=np.uint32)
        raw = np.zeros(max(nb.value, 1), dtype=np.uint8)
        self.ctx.call("msc_dict_export", C.c_void_p(self.handle), lens.ctypes.data_as(C.c_void_p), raw.ctypes.data_as(C.c_void_p))
        out, pos, blob = [], 0, raw.tobytes()
        for ln in lens.tolist():
            out.append(blob[pos:pos + ln].decode("utf-8"))
            pos += ln
        return out

    def load(self, entries: list[str]) -> "DictHandle":
        """Fill this (empty) dictionary with ``entries``; entry i gets code i on every rank that loads the same list."""
        import numpy as np  # noqa: PLC0415

        raws = [t.encode("utf-8") for t in entries]
        lens = np.asarray([len(r) for r in raws], dtype=np.uint32)
        blob = np.frombuffer(b"".join(raws) or b"\0", dtype=np.uint8)
        self.ctx.call("msc_dict_load", C.c_void_p(self.handle), lens.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), len(raws))
        return self

    def free(self) -> None:
        if self.handle:
            for ptr in self._luts.values():
                self.ctx.dev_free(ptr)
            self._luts.clear()
            # translation tables that map INTO this dictionary die with it (a long-lived table dictionary would otherwise
            # keep one per query-scoped target)
            for source, key in self._translated_from:
                ptr = source._luts.pop(key, None) if source.handle else None
                if ptr:
                    self.ctx.dev_free(ptr)
            self._translated_from.clear()
            self.ctx.lib.msc_dict_free(C.c_void_p(self.handle))
            self.handle = None


@dataclass
class DeviceColumn:
    ptr: int
    phys: int
    ltype: str                      # lowering type tag: I / F / T / S / B
    dict: Optional[DictHandle] = None
    via: Optional[int] = None       # index-vector id (virtual join relation) or None for direct


class DeviceRel:
    """A device-resident relation (owned ``msc_rel`` handle) with host-side column metadata."""

    def __init__(self, ctx: N.Context, handle: Optional[int], nrows: int, cols: list[DeviceColumn],
                 keep: Optional[list] = None, owned: bool = True) -> None:
        self.ctx = ctx
        self.handle = handle
        self.owned = owned  # False: the msc_rel belongs to a library object (a prepared pass) that frees it
        self.nrows = nrows
        self.cols = cols
        self.keep = keep or []  # relations / dictionaries whose memory these columns reference
        # several ranks: True = this rank holds only its part of the relation (the rank-ordered union is the relation),
        # False = every rank holds all of it
        self.partitioned = False

    @classmethod
    def from_handle(cls, ctx: N.Context, handle: int, ltypes: list[str], dicts: list[Optional[DictHandle]]) -> "DeviceRel":
        nrows = C.c_uint64()
        ncols = C.c_int32()
        ctx.check(ctx.lib.msc_rel_info(C.c_void_p(handle), C.byref(nrows), C.byref(ncols)))
        binds = (N.ColBind * max(ncols.value, 1))()
        ctx.check(ctx.lib.msc_rel_cols(C.c_void_p(handle), binds, ncols.value))
        cols = [DeviceColumn(binds[i].data or 0, binds[i].phys, ltypes[i], dicts[i]) for i in range(ncols.value)]
        return cls(ctx, handle, nrows.value, cols)

    def free(self) -> None:
        if self.handle and self.owned:
            self.ctx.lib.msc_rel_free(C.c_void_p(self.handle))
        self.handle = None

    def column_numpy(self, i: int):  # noqa: ANN201
        """Full-precision host copy of one column (f64 / i64 / codes): the 1e-9 parity hook."""
        import numpy as np  # noqa: PLC0415

        col = self.cols[i]
        arr = np.zeros(self.nrows, dtype=N.PHYS_NUMPY[col.phys])
        if self.nrows:
            self.ctx.call("msc_memcpy_d2h", arr.ctypes.data_as(C.c_void_p), C.c_void_p(col.ptr), arr.nbytes)
        return arr


@dataclass
class TableEntry:
    path: Path
    stamp: tuple
    handle: int
    schema: Schema
    blocks: list[int]
    nrows: int
    total_rows: int = 0  # rows of the whole file (all ranks' blocks): known from the footer without asking anybody
    columns: dict[int, DeviceColumn] = field(default_factory=dict)
    rels: list[DeviceRel] = field(default_factory=list)
    keepalive: Any = None  # pinned image for in-memory tables


_PROBE_BASE = 1 << 20  # column indices of a fused join's build side inside the consuming scan's source


def _has_concat(e: L.Expr) -> bool:
    return isinstance(e, L.EConcat) or any(_has_concat(c) for c in L.expr_children(e))


def _path_stamp(path: str) -> tuple:
    try:
        st = os.stat(path)
        return (st.st_mtime_ns, st.st_size)
    except OSError:
        return ()


def _plan_tables(node: L.LNode) -> list[str]:
    if isinstance(node, L.LTable):
        return [str(node.path)]
    out: list[str] = []
    for attr in ("child", "left", "right"):
        sub = getattr(node, attr, None)
        if sub is not None:
            out += _plan_tables(sub)
    return out


def _file_stamp(entry: "TableEntry") -> tuple:
    if entry.stamp and entry.stamp[0] == "mem":
        return entry.stamp
    try:
        st = os.stat(entry.path)
        return (st.st_mtime_ns, st.st_size)
    except OSError:
        return ()


class _Source:
    """Input of one fused scan: ``nrows`` rows whose column ``i`` is ``columns[i]``."""

    def __init__(self, nrows: int, columns: dict[int, DeviceColumn], index_vectors: Optional[list[DeviceColumn]] = None,
                 keep: Optional[list] = None, partitioned: bool = False) -> None:
        self.nrows = nrows
        self.columns = columns
        self.index_vectors = index_vectors or []
        self.keep = keep or []
        self.partitioned = partitioned  # see DeviceRel.partitioned
        self.table_columns = False  # the columns are a loaded table's own (immutable while it is loaded): msc_scan_desc.table_columns
        # a join fused into this scan (CudaExecutionEngine._probe_source): the scan walks the probe side's rows, runs the
        # probe side's own filters, looks every row's key up in the build side's table and reads build-side columns
        # (via == "probe") through the matched row
        self.pre_filters: list[L.Expr] = []
        self.probe_key: Optional[L.Expr] = None
        self.probe_table: Optional[int] = None
        self.probe_compact = False
        self.translate_targets: Optional[dict[str, "DictHandle"]] = None


class _ScanResolver:
    """Allocates staged / gather / LUT slots of one ``msc_scan_desc`` while the program compiles."""

    def __init__(self, engine: "CudaExecutionEngine", source: _Source, translate_targets: Optional[dict[str, DictHandle]] = None) -> None:
        self.engine = engine
        self.source = source
        self.staged: list[DeviceColumn] = []
        self.gather: list[DeviceColumn] = []
        self.luts: list[int] = []
        self._staged_slot: dict[int, int] = {}
        self._gather_slot: dict[int, int] = {}
        self._bindings: dict[int, L.Binding] = {}
        self.translate_targets = translate_targets or source.translate_targets or {}
        # where every staged / gathered column came from -- ("col", input index) or ("ivec", k) -- so that a later run of the
        # same scan over a structurally identical source can rebind the pointers without compiling again (_CompiledScan)
        self.staged_origin: list[tuple[str, int]] = []
        self.gather_origin: list[tuple[str, int]] = []

    def _stage(self, col: DeviceColumn, origin: tuple[str, int]) -> int:
        if col.ptr not in self._staged_slot:
            if len(self.staged) >= N.K["MSC_VM_MAX_STAGED"]:
                raise L.LoweringError("query reads more columns than one scan can stage")
            self._staged_slot[col.ptr] = len(self.staged)
            self.staged.append(col)
            self.staged_origin.append(origin)
        return self._staged_slot[col.ptr]

    def binding(self, index: int) -> L.Binding:
        if index not in self._bindings:
            col = self.source.columns[index]
            if col.via is None:
                b = L.Binding(col.phys, staged=self._stage(col, ("col", index)), dict_id=col.dict)
            elif col.via == "probe":
                if col.ptr not in self._gather_slot:
                    if len(self.gather) >= N.K["MSC_VM_MAX_GATHER"]:
                        raise L.LoweringError("query gathers more columns than one scan supports")
                    self._gather_slot[col.ptr] = len(self.gather)
                    self.gather.append(col)
                    self.gather_origin.append(("col", index))
                b = L.Binding(col.phys, gather=self._gather_slot[col.ptr], dict_id=col.dict, probe=True)
            else:
                if col.ptr not in self._gather_slot:
                    if len(self.gather) >= N.K["MSC_VM_MAX_GATHER"]:
                        raise L.LoweringError("query gathers more columns than one scan supports")
                    self._gather_slot[col.ptr] = len(self.gather)
                    self.gather.append(col)
                    self.gather_origin.append(("col", index))
                b = L.Binding(col.phys, gather=self._gather_slot[col.ptr], index=self._stage(self.source.index_vectors[col.via], ("ivec", col.via)),
                              dict_id=col.dict)
            self._bindings[index] = b
        return self._bindings[index]

    def _lut(self, ptr: int) -> int:
        if ptr not in self.luts:
            if len(self.luts) >= N.K["MSC_VM_MAX_LUTS"]:
                raise L.LoweringError("too many string predicates in one scan")
            self.luts.append(ptr)
        return self.luts.index(ptr)

    def probe_spec(self) -> Optional[L.ProbeSpec]:
        if self.source.probe_key is None:
            return None
        return L.ProbeSpec(self.source.probe_key, self._lut(self.source.probe_table), self.source.probe_compact)

    def literal_code(self, dict_id: DictHandle, text: str) -> int:
        return dict_id.literal_code(text)

    def like_lut(self, dict_id: DictHandle, pattern: str) -> int:
        return self._lut(dict_id.like_lut(pattern))

    def same_dict(self, a: DictHandle, b: DictHandle) -> bool:
        return a is b

    def recode_lut(self, src: DictHandle, dst: DictHandle) -> int:
        return self._lut(src.translate_lut(dst, insert=False))

    def order_lut(self, dict_id: DictHandle, op: str, text: str) -> int:
        return self._lut(dict_id.order_lut(op, text))

    def rank_luts(self, a: DictHandle, b: DictHandle) -> tuple[int, int]:
        return self._lut(a.rank_lut(b)), self._lut(b.rank_lut(a))

    def translate_lut(self, dict_id: DictHandle, token: str) -> tuple[int, DictHandle]:
        target = self.translate_targets[token]
        if target is dict_id:
            return -1, target
        return self._lut(dict_id.translate_lut(target, insert=False)), target

    def desc(self, program: L.Program) -> N.ScanDesc:
        d = N.ScanDesc()
        d.nrows = self.source.nrows
        if not self.staged and self.source.nrows:
            # a program without column reads (e.g. COUNT(*) over a constant) still needs the row count only
            pass
        d.nstaged = len(self.staged)
        d.ngather = len(self.gather)
        for i, col in enumerate(self.staged):
            d.staged[i].data = col.ptr
            d.staged[i].phys = col.phys
        for i, col in enumerate(self.gather):
            d.gather[i].data = col.ptr
            d.gather[i].phys = col.phys
        words = program.words()
        d.ncode = len(words)
        for i, w in enumerate(words):
            d.code[i] = w
        d.nconsts = len(program.consts)
        for i, c in enumerate(program.consts):
            d.consts[i] = c
        d.nluts = len(self.luts)
        d.ntemps = program.ntemps
        d.want_jit = 1 if self.engine.jit == "always" or self.engine._specialise_now else 0
        d.table_columns = 1 if self.source.table_columns else 0
        d.ncode2 = len(program.regvm)
        d.count_slot2 = program.regvm_count_slot
        for i, w in enumerate(program.regvm):
            d.code2[i] = w
        for i, ptr in enumerate(self.luts):
            d.luts[i] = ptr
        return d


class _CompiledScan:
    """A scan that has been compiled once: its program, its ``msc_scan_desc`` and where every pointer in it came from.

    A repeated one-shot query lowers to the same plan (the plan cache), prepares structurally identical sources and would
    compile the same programs again -- about 50 us of Python per scan, a fifth of a 1 ms query.  The engine keeps the compiled
    scan per plan node instead and, when the source of the next run has the same shape (column types, access paths,
    dictionaries and their sizes, probe form) and the expressions are equal, only rebinds row count and pointers."""

    def __init__(self, node: Any, signature: tuple, pins: list, exprs: tuple, resolver: _ScanResolver, prog: Any, desc: N.ScanDesc) -> None:
        self.node, self.signature, self.pins, self.exprs = node, signature, pins, exprs  # (node / pins keep the ids in the signature unique)
        self.prog, self.desc = prog, desc
        self.staged_origin, self.gather_origin = list(resolver.staged_origin), list(resolver.gather_origin)
        probe_table = resolver.source.probe_table
        self.probe_lut = resolver.luts.index(probe_table) if probe_table is not None and probe_table in resolver.luts else -1

    @staticmethod
    def signature_of(source: _Source, translate_targets: Optional[dict[str, DictHandle]]) -> tuple[tuple, list]:
        pins: list = []
        cols = []
        for i, c in source.columns.items():
            if c.dict is not None:
                pins.append(c.dict)
                cols.append((i, c.phys, c.via, c.dict.serial, c.dict.size))  # (serial: never reused, unlike id())
            else:
                cols.append((i, c.phys, c.via, 0, 0))
        targets = translate_targets or source.translate_targets or {}
        tsig = tuple((k, v.serial, v.size) for k, v in targets.items())
        pins.extend(targets.values())
        return (tuple(cols), len(source.index_vectors), source.probe_compact, source.probe_table is not None, source.table_columns, tsig), pins

    def rebind(self, engine: "CudaExecutionEngine", source: _Source) -> N.ScanDesc:
        d = self.desc
        d.nrows = source.nrows
        for slot, (kind, at) in enumerate(self.staged_origin):
            d.staged[slot].data = (source.columns[at] if kind == "col" else source.index_vectors[at]).ptr
        for slot, (_, at) in enumerate(self.gather_origin):
            d.gather[slot].data = source.columns[at].ptr
        if self.probe_lut >= 0:
            d.luts[self.probe_lut] = source.probe_table
        d.want_jit = 1 if engine.jit == "always" or engine._specialise_now else 0
        return d


def _comm_from_torch(device: int) -> Comm:
    """Join the process group the launcher (torchrun) set up, if any; otherwise a single-rank Comm."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return Comm()
    try:
        import torch.distributed as dist
    except ImportError:
        return Comm()
    if not dist.is_initialized():
        return Comm()
    return Comm.from_env(device)


class _DevView:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _torch_dtype(phys: int):  # noqa: ANN202
    import torch  # noqa: PLC0415

    return {N.P_U8: torch.uint8, N.P_U16: torch.int16, N.P_U32: torch.int32, N.P_I32: torch.int32, N.P_I64: torch.int64,
            N.P_F32: torch.float32, N.P_F64: torch.float64}[phys]


class _Fingerprint:
    """A task tree's structural description with its hash computed once (tuples re-hash on every dictionary lookup)."""

    __slots__ = ("tree", "hash")

    def __init__(self, tree: Any) -> None:
        self.tree = tree
        self.hash = hash(tree)

    def __hash__(self) -> int:
        return self.hash

    def __eq__(self, other: Any) -> bool:
        return self is other or (isinstance(other, _Fingerprint) and self.hash == other.hash and self.tree == other.tree)


def task_fingerprint(obj: Any, _depth: int = 0) -> Any:
    """A hashable structural description of a task / column tree (class names + attribute values, recursively): two trees
    with the same fingerprint lower to the same plan.  Works on the reference's own classes as well as the mirror's."""
    if _depth > 200:
        raise RecursionError("task tree too deep")
    if obj is None or isinstance(obj, (bool, int, float, str, bytes)):
        return obj
    if isinstance(obj, Path):
        return ("path", str(obj))
    if isinstance(obj, (list, tuple)):
        return tuple(task_fingerprint(x, _depth + 1) for x in obj)
    if isinstance(obj, dict):
        return tuple(sorted((str(k), task_fingerprint(v, _depth + 1)) for k, v in obj.items()))
    if isinstance(obj, ExecutionEngine):
        return "engine"
    if hasattr(obj, "isoformat"):
        return ("time", obj.isoformat())
    if hasattr(obj, "name") and hasattr(obj, "value") and type(obj).__module__.endswith("constants"):
        return ("enum", type(obj).__name__, obj.name)
    if callable(obj) and not hasattr(obj, "__dict__"):
        return ("fn", getattr(obj, "__name__", repr(obj)))
    fields = getattr(obj, "__dict__", None)
    if fields is None:
        fields = {k: getattr(obj, k) for k in getattr(obj, "__slots__", ())}
    return (type(obj).__name__, tuple(sorted((k, task_fingerprint(v, _depth + 1)) for k, v in fields.items() if not k.startswith("_msc"))))


_LTYPE_OF = {ColumnType.INTEGER: L.INT, ColumnType.FLOAT: L.FLOAT, ColumnType.TIMESTAMP: L.TS, ColumnType.STRING: L.STR}
_MSC_TYPE = {ColumnType.INTEGER: N.K["MSC_T_INTEGER"], ColumnType.STRING: N.K["MSC_T_STRING"],
             ColumnType.FLOAT: N.K["MSC_T_FLOAT"], ColumnType.TIMESTAMP: N.K["MSC_T_TIMESTAMP"]}
# (groups + 1 trash) x (aggregates + 1 hidden counter) cells of per-thread shared-memory accumulators
DENSE_MAX_CELLS = 96


class CudaExecutionEngine(ExecutionEngine):
    """B200 engine: BlockFile ingest -> device columns -> fused scan kernels -> result BlockFile.

    Parameters (all optional, so ``CudaExecutionEngine()`` works like the reference engines):
      device       CUDA device ordinal (default: ``LOCAL_RANK`` or 0)
      work_folder  where result BlockFiles are written (default: a fresh temp dir, removed on exit)
      layout       ``"native"`` (i32/f32/i64/narrow codes, the disk widths) or ``"wide"`` (i64/f64)
      shard        ``(rank, world)``: this engine only ingests row-blocks ``b % world == rank``
    """

    def __init__(self, device: Optional[int] = None, work_folder: Optional[Path] = None, layout: str = "native",
                 shard: Optional[tuple[int, int]] = None, comm: Optional[Comm] = None, jit: Optional[str] = None) -> None:
        """``jit``: when scans run on kernels compiled for exactly their program (csrc/jit.cu), the counterpart of the
        reference ThreadEngine compiling one Zig program per query (execution.py:139-160).  "auto" (default): prepared
        queries, and any scan whose kernel this process has compiled already; "always": every dense aggregate and
        filter / project scan (0.1-0.3 s per new query shape, cached; MSC_JIT_CACHE=<dir> keeps the cubins on disk);
        "never": interpreters only.  The environment variable MINISPARK_JIT supplies the default."""
        self.jit = jit or os.environ.get("MINISPARK_JIT", "auto")
        if self.jit not in ("auto", "always", "never"):
            raise ValueError("jit must be 'auto', 'always' or 'never'")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.ctx = N.Context(device)  # raises when the library or the GPU is missing: no fallback
        self.device = device
        self.layout = N.K["MSC_LAYOUT_WIDE"] if layout == "wide" else N.K["MSC_LAYOUT_NATIVE"]
        if comm is None:
            comm = _comm_from_torch(device)
        self.comm = comm
        self.shard = shard if shard is not None else (comm.rank, comm.world)
        # several ranks: collect() / execute_full_task return the COMPLETE result on every rank (rows a rank does not hold
        # -- a sharded scan's, a partitioned GROUP BY's or join's -- are gathered in rank order); execute_to_device leaves
        # the rank-local part on the device unless asked (last_stats["result_partitioned"] says which it is)
        self.replicate_results = True
        self._shuffle: Any = None  # PeerShuffle once several ranks exchange rows; False when CUDA IPC is unavailable
        # Repeated queries: the reference's ThreadEngine keeps the executable it compiled for a query (execution.py:139-160);
        # here the second execution of an identical aggregate task tree is prepared once (lowering, validated programs, the
        # kernel specialised for it, result buffers) and every later one is a single launch.  fingerprint -> [runs, prepared]
        self._plan_cache: dict[Any, list] = {}
        self._scan_cache: dict[tuple, _CompiledScan] = {}  # compiled scans per plan node (_CompiledScan)
        self.scan_cache_enabled = os.environ.get("MINISPARK_SCAN_CACHE", "1") != "0"
        self.plan_cache_enabled = os.environ.get("MINISPARK_PLAN_CACHE", "1") != "0"
        # joins whose build side has no duplicate keys run as a lookup inside the consuming scan (MSC_OP_PROBE) instead of
        # materialising both sides and the pair list
        self.fused_probe = os.environ.get("MINISPARK_FUSED_PROBE", "1") != "0"
        self._probe_declined: Optional[DeviceRel] = None
        self.fused_build = os.environ.get("MINISPARK_FUSED_BUILD", "0") != "0"
        self._trace_track: Optional[int] = None
        self._dense_merges: dict[tuple, "_DenseMerge"] = {}
        # jit="auto": a task tree that comes back is worth kernels compiled for exactly its scans (0.1-0.3 s each, once per
        # process and shape) -- the reference's ThreadEngine compiles every query before it runs it (execution.py:139-160)
        self._specialise_now = False
        self._own_work = work_folder is None
        self.work_folder = Path(work_folder) if work_folder is not None else Path(tempfile.mkdtemp(prefix="minispark_cuda_"))
        self.work_folder.mkdir(parents=True, exist_ok=True)
        self._tables: dict[str, TableEntry] = {}
        self._query_rels: list[DeviceRel] = []
        self._query_dicts: list[DictHandle] = []
        self._table_dicts: list[DictHandle] = []
        self._result_files: list[Path] = []
        self._closers: list[Any] = []  # run at close(), before the context goes (peer mailboxes)
        self.last_stats: dict[str, Any] = {}

    # ---- ExecutionEngine contract ---------------------------------------------------------------
    def execute_full_task(self, full_task: Any) -> list[JobResult]:
        TRACER.start("execute full task")  # (the reference's slice names, execution.py:69-78)
        try:
            rel, schema = self.execute_to_device(full_task, replicate=self.replicate_results)
            job = JobResult(str(uuid.uuid4()), f"cuda:{self.device}", [], result_partitioned=rel.partitioned)
            if rel.nrows > 0:  # empty result -> no output file (reference tasks.py:405-406)
                path = self.work_folder / f"result_{job.job_id}.bin"
                TRACER.start("write result BlockFile")
                try:
                    self.write_blockfile(rel, schema, path)
                finally:
                    TRACER.end()
                self._result_files.append(path)
                job.output_files.append(OutputFile(path))
            return [job]
        finally:
            self.release_query()
            TRACER.end()

    def __exit__(self, exc_type, exc_value, traceback) -> None:  # noqa: ANN001
        self.close()

    def close(self) -> None:
        if getattr(self, "ctx", None) is None:
            return
        self.release_query()
        for closer in getattr(self, "_closers", []):
            closer()
        self._closers = []
        if getattr(self, "_shuffle", None):
            self._shuffle.close()
        self._shuffle = None
        for entry in self._tables.values():
            for rel in entry.rels:
                rel.free()
            self.ctx.lib.msc_table_close(C.c_void_p(entry.handle))
        self._tables.clear()
        for d in self._table_dicts:
            d.free()
        self._table_dicts.clear()
        for path in self._result_files:
            path.unlink(missing_ok=True)
        self._result_files.clear()
        if self._own_work:
            shutil.rmtree(self.work_folder, ignore_errors=True)
        self.ctx.close()
        self.ctx = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- public extras ----------------------------------------------------------------------------
    def execute_to_device(self, full_task: Any, replicate: bool = False) -> tuple[DeviceRel, Schema]:
        """Run the query and leave the (full-precision) result on the device.  Several ranks: a result that is spread
        over the ranks (``rel.partitioned``) stays so unless ``replicate`` asks for the rank-ordered gather."""
        try:
            t0 = time.perf_counter()
            entry = self._plan_entry(full_task)
            cached = self._cached_run(full_task, entry)
            if cached is not None:
                self.last_stats["result_partitioned"] = False
                self.last_stats["query_s"] = time.perf_counter() - t0
                return cached
            TRACER.start("lower task tree")
            try:
                plan = self._lowered(full_task, entry)
            finally:
                TRACER.end()
            self.last_plan = plan
            self.last_stats["exchange"] = None
            self.last_stats["exchanges"] = []  # every cross-rank step of this query, in order
            self.last_stats["scans_rebound"] = 0  # scans of this query that reused the program compiled by an earlier run (_CompiledScan)
            self.last_stats["plan"] = "one-shot"
            self._specialise_now = self.jit == "auto" and entry is not None and entry[0] >= 2
            if self._specialise_now:
                self.last_stats["plan"] = "one-shot, scans on kernels specialised for this task tree (it came back)"
            TRACER.start("Execution")
            try:
                rel = self._run(plan)
                if replicate and rel.partitioned:
                    rel = self._gather_rows(rel)
            finally:
                TRACER.end()
            self.last_stats["result_partitioned"] = rel.partitioned
            self.last_stats["query_s"] = time.perf_counter() - t0
            return rel, plan.schema
        except N.NativeError as e:
            self.release_query()
            raise ExecutionError(str(e)) from e
        except Exception:
            self.release_query()
            raise

    def _plan_entry(self, full_task: Any) -> Optional[list]:
        """The plan cache's entry for this task tree: [runs, prepared pass | None | False, lowered plan | None, table stamps].
        The fingerprint is remembered on the task object, so re-executing the same DataFrame costs a dictionary lookup."""
        if not self.plan_cache_enabled:
            return None
        try:
            key = getattr(full_task, "_msc_fingerprint", None)
            if key is None:
                key = _Fingerprint(task_fingerprint(full_task))
                try:
                    full_task._msc_fingerprint = key
                except Exception:  # noqa: BLE001  (slots / frozen classes: recompute next time)
                    pass
        except Exception:  # noqa: BLE001  (an exotic tree: just run it the one-shot way)
            return None
        entry = self._plan_cache.setdefault(key, [0, None, None, []])
        entry[0] += 1
        return entry

    def _lowered(self, full_task: Any, entry: Optional[list]) -> L.LNode:
        """deepcopy + the reference's own validation + lowering, once per task tree (and again when a table file changed)."""
        if entry is not None and entry[2] is not None and all(_path_stamp(p) == st for p, st in entry[3]):
            return entry[2]
        task = deepcopy(full_task)  # planning mutates the tree (reference plan.py:181-204)
        task.validate_schema()      # the reference's own validation and its errors
        plan = L.lower_task(task)
        if entry is not None:
            entry[2] = plan
            entry[3] = [(p, _path_stamp(p)) for p in _plan_tables(plan)]
        return plan

    def _cached_run(self, full_task: Any, entry: Optional[list]) -> Optional[tuple[DeviceRel, Schema]]:
        """Second and later executions of an identical aggregate query run as a prepared pass (see _plan_cache)."""
        if entry is None or self.jit == "never" or entry[0] < 2 or entry[1] is False:
            return None
        if entry[1] is None:
            entry[1] = False
            try:
                prepared = self.prepare(full_task)
                if prepared.ngroups and prepared.reusable:  # the dense (low-cardinality) form is what a prepared pass accelerates
                    entry[1] = prepared
            except (L.LoweringError, N.NativeError):
                self.release_query()
            if entry[1] is False:
                return None
        prepared = entry[1]
        if any(self._tables.get(k) is not e or e.stamp != _file_stamp(e) for k, e in prepared.tables):  # a table changed
            entry[0], entry[1] = 1, None
            return None
        rel, _ = prepared.run()
        self.last_plan = prepared.plan
        self.last_stats["plan"] = "prepared (repeat of an identical task tree)"
        self.last_stats["exchange"] = prepared.scan_stats.get("exchange")
        self._note_kernel()
        return rel, prepared.plan.schema

    def release_query(self) -> None:
        for rel in self._query_rels:
            rel.free()
        self._query_rels.clear()
        for d in self._query_dicts:
            d.free()
        self._query_dicts.clear()
        if self._shuffle:
            self._shuffle.release_all()

    def write_blockfile(self, rel: DeviceRel, schema: Schema, path: Path) -> None:
        from . import io as _io  # noqa: PLC0415

        cols = (N.OutCol * len(schema))()
        names = []
        for i, (name, ctype) in enumerate(schema):
            raw = name.encode("utf-8")
            names.append(raw)
            cols[i].name = raw
            cols[i].type = _MSC_TYPE[ctype]
            cols[i].rel_col = i
            cols[i].dict = rel.cols[i].dict.handle if rel.cols[i].dict is not None else None
        try:
            self.ctx.call("msc_write_blockfile", C.c_void_p(rel.handle), cols, len(schema), str(path).encode(), _io.ROWS_PER_BLOCK)
        except N.NativeError as e:
            if e.code == N.K["MSC_ERR_OVERFLOW"]:
                raise OverflowError("int too big to convert") from e  # what the reference raises (io.py:90)
            raise ExecutionError(str(e)) from e

    def register_table_image(self, name: str, image_ptr: int, nbytes: int, keepalive: Any = None) -> None:
        """Expose a BlockFile image in (pinned) host memory under a table name (used by the bench)."""
        h = C.c_void_p()
        self.ctx.call("msc_table_open_mem", C.c_void_p(image_ptr), nbytes, C.byref(h))
        self._tables[name] = self._make_entry(Path(name), ("mem", image_ptr, nbytes), h.value, keepalive)

    def drop_table_cache(self, name: Optional[str] = None) -> None:
        """Forget device-resident columns (all tables, or one) so the next query ingests again."""
        self._plan_cache.clear()  # prepared passes are bound to the columns that go away
        self._scan_cache.clear()  # (and compiled scans pin their dictionaries)
        for key, entry in list(self._tables.items()):
            if name is not None and key != name:
                continue
            for rel in entry.rels:
                rel.free()
            entry.rels.clear()
            entry.columns.clear()

    # ---- tables -----------------------------------------------------------------------------------
    def _make_entry(self, path: Path, stamp: tuple, handle: int, keepalive: Any = None) -> TableEntry:
        ncols, nblocks, nrows = C.c_int32(), C.c_int32(), C.c_uint64()
        self.ctx.check(self.ctx.lib.msc_table_info(C.c_void_p(handle), C.byref(ncols), C.byref(nblocks), C.byref(nrows)))
        schema: Schema = []
        for c in range(ncols.value):
            ty = C.c_int32()
            buf = C.create_string_buffer(256)
            self.ctx.check(self.ctx.lib.msc_table_col_info(C.c_void_p(handle), c, C.byref(ty), buf, 256))
            schema.append((buf.value.decode("utf-8"), ColumnType.from_ordinal(ty.value)))
        rank, world = self.shard
        blocks = shard_blocks(nblocks.value, rank, world)
        local_rows = 0
        for b in blocks:
            rows = C.c_uint32()
            self.ctx.check(self.ctx.lib.msc_table_block_rows(C.c_void_p(handle), b, C.byref(rows)))
            local_rows += rows.value
        return TableEntry(path, stamp, handle, schema, blocks, local_rows, total_rows=nrows.value, keepalive=keepalive)

    def _table(self, path: Path) -> TableEntry:
        key = str(path)
        entry = self._tables.get(key)
        if entry is not None and entry.stamp[0] == "mem":
            return entry
        st = os.stat(path)
        stamp = (st.st_mtime_ns, st.st_size)
        if entry is not None and entry.stamp == stamp:
            return entry
        if entry is not None:
            for rel in entry.rels:
                rel.free()
            self.ctx.lib.msc_table_close(C.c_void_p(entry.handle))
        h = C.c_void_p()
        self.ctx.call("msc_table_open", str(path).encode(), C.byref(h))
        entry = self._make_entry(Path(path), stamp, h.value)
        self._tables[key] = entry
        return entry

    def _ensure_columns(self, entry: TableEntry, needed: Iterable[int]) -> None:
        missing = sorted(set(needed) - set(entry.columns))
        if not missing:
            return
        cols = N.int32_array(missing)
        blocks = N.int32_array(entry.blocks)
        dict_slots = (C.c_void_p * len(missing))()
        out = C.c_void_p()
        self.ctx.call("msc_table_load", C.c_void_p(entry.handle), cols, len(missing), blocks, len(entry.blocks), self.layout,
                      dict_slots, C.byref(out))
        ltypes = [_LTYPE_OF[entry.schema[c][1]] for c in missing]
        dicts: list[Optional[DictHandle]] = []
        for i, c in enumerate(missing):
            if entry.schema[c][1] == ColumnType.STRING:
                d = DictHandle(self.ctx, dict_slots[i])
                d.persistent = True
                self._table_dicts.append(d)
                dicts.append(d)
            else:
                dicts.append(None)
        rel = DeviceRel.from_handle(self.ctx, out.value, ltypes, dicts)
        entry.rels.append(rel)
        for i, c in enumerate(missing):
            entry.columns[c] = rel.cols[i]
        st = self.ctx.stats()
        self.last_stats["ingest_ms"] = st.last_ingest_ms
        self.last_stats["ingest_bytes"] = st.last_ingest_bytes
        if TRACER.enabled:
            TRACER.device_slice(f"load table block columns {missing} of {entry.path.name} ({st.last_ingest_bytes} bytes)", st.last_ingest_ms, self._gpu_track())

    # ---- plan execution -----------------------------------------------------------------------------
    def _track(self, rel: DeviceRel) -> DeviceRel:
        self._query_rels.append(rel)
        return rel

    def _run(self, node: L.LNode) -> DeviceRel:
        if isinstance(node, L.LSelect):
            return self._run_select(node)
        if isinstance(node, L.LAggregate):
            return self._run_aggregate(node)
        return self._run_select(L.identity_select(node))

    def _source(self, node: L.LNode, needed: set[int]) -> _Source:
        if isinstance(node, L.LTable):
            entry = self._table(node.path)
            self._ensure_columns(entry, needed)
            source = _Source(entry.nrows, {i: entry.columns[i] for i in needed}, partitioned=self.comm.world > 1)
            source.table_columns = True
            return source
        if isinstance(node, L.LJoin):
            return self._join_source(node, needed)
        rel = self._run(node)
        return _Source(rel.nrows, dict(enumerate(rel.cols)), keep=[rel], partitioned=rel.partitioned)

    def _prepare(self, base: L.LNode, exprs: list[L.Expr], fuse_probe: bool = True) -> tuple[_Source, list[L.Expr]]:
        """Resolve the scan input and materialise string concatenations the program cannot compute.  A join directly below
        the scan is fused into it where possible (_probe_source): the returned expressions are then rewritten to the probe
        side's columns and source.pre_filters / probe_key describe the rest."""
        needed: set[int] = set()
        for e in exprs:
            needed |= L.expr_inputs(e)
        if fuse_probe and isinstance(base, L.LJoin) and self.fused_probe:
            fused = self._probe_source(base, exprs, needed)
            if fused is not None:
                return fused
        source = self._source(base, needed)
        concats: dict[L.EConcat, int] = {}

        def rewrite(e: L.Expr) -> L.Expr:
            if isinstance(e, L.EConcat):
                if e not in concats:
                    concats[e] = self._materialise_concat(source, e)
                return L.EInput(L.STR, concats[e])
            if isinstance(e, L.EBin):
                return L.EBin(e.type, e.op, rewrite(e.left), rewrite(e.right))
            if isinstance(e, L.ECast):
                return L.ECast(e.type, rewrite(e.child))
            if isinstance(e, L.ELike):
                return L.ELike(e.type, rewrite(e.child), e.pattern)
            if isinstance(e, L.ETranslate):
                return L.ETranslate(e.type, rewrite(e.child), e.token)
            if isinstance(e, L.ECode):
                return L.ECode(e.type, rewrite(e.child))
            return e

        return source, [rewrite(e) for e in exprs]

    def _materialise_concat(self, source: _Source, e: L.EConcat) -> int:
        """Evaluate ``a + b + ...`` over every source row with msc_str_concat; returns the new column index."""
        parts = (N.ConcatPart * len(e.parts))()
        keep = []
        for i, part in enumerate(e.parts):
            if isinstance(part, L.EConst):
                raw = str(part.value).encode("utf-8")
                keep.append(raw)
                parts[i].literal = raw
                parts[i].literal_len = len(raw)
                continue
            if not isinstance(part, L.EInput):
                raise L.LoweringError("string concatenation operands must be columns or literals")
            col = source.columns[part.index]
            if col.via is not None:  # virtual (joined) rows: gather the codes into a row-aligned column first
                col = self._gather_column(source, part.index)
            parts[i].codes.data = col.ptr
            parts[i].codes.phys = col.phys
            parts[i].dict = col.dict.handle
        out_dict = DictHandle(self.ctx)
        self._query_dicts.append(out_dict)
        out = C.c_void_p()
        self.ctx.call("msc_str_concat", parts, len(e.parts), source.nrows, C.c_void_p(out_dict.handle), C.byref(out))
        rel = self._track(DeviceRel.from_handle(self.ctx, out.value, [L.STR], [out_dict]))
        index = max(list(source.columns) + [-1]) + 1
        source.columns[index] = rel.cols[0]
        return index

    def _gather_column(self, source: _Source, index: int) -> DeviceColumn:
        col = source.columns[index]
        resolver = _ScanResolver(self, source)
        prog = L.compile_project(resolver, [], [L.EInput(col.ltype, index)])
        rel = self._scan_project(prog, resolver.desc(prog.program), [col.ltype], source.partitioned)
        return rel.cols[0]

    def _compiled_scan(self, key: Optional[tuple], node: Any, source: _Source, translate_targets: Optional[dict[str, DictHandle]], exprs: tuple,
                       compile_fn: Callable[[_ScanResolver], Any]) -> tuple[Any, N.ScanDesc]:
        """(program, descriptor) of one scan: compiled now, or the compiled scan of an earlier run of the same plan node rebound
        to this run's rows (_CompiledScan)."""
        if key is None or not self.scan_cache_enabled:
            resolver = _ScanResolver(self, source, translate_targets)
            prog = compile_fn(resolver)
            return prog, resolver.desc(prog.program)
        signature, pins = _CompiledScan.signature_of(source, translate_targets)
        hit = self._scan_cache.get(key)
        if hit is not None and hit.signature == signature and hit.exprs == exprs:
            self.last_stats["scans_rebound"] = self.last_stats.get("scans_rebound", 0) + 1
            return hit.prog, hit.rebind(self, source)
        resolver = _ScanResolver(self, source, translate_targets)
        prog = compile_fn(resolver)
        desc = resolver.desc(prog.program)
        if len(self._scan_cache) >= 256:
            self._scan_cache.clear()
        self._scan_cache[key] = _CompiledScan(node, signature, pins, exprs, resolver, prog, desc)
        return prog, desc

    def _scan_project(self, prog: L.ProjectProgram, desc: N.ScanDesc, ltypes: list[str], partitioned: bool) -> DeviceRel:
        out = C.c_void_p()
        phys = N.int32_array(prog.out_phys)
        self.ctx.call("msc_scan_project", C.byref(desc), phys, len(prog.out_phys), C.byref(out))
        self._note_kernel("scan: filter + project")
        rel = self._track(DeviceRel.from_handle(self.ctx, out.value, ltypes, prog.out_dicts))
        rel.partitioned = partitioned
        return rel

    def _gpu_track(self) -> int:
        if self._trace_track is None:
            self._trace_track = TRACER.new_track(f"GPU {self.device} (device time)")
        return self._trace_track

    def _note_kernel(self, what: str = "kernels") -> None:
        st = self.ctx.stats()
        if TRACER.enabled:
            TRACER.device_slice(what, st.last_kernel_ms, self._gpu_track())
        self.last_stats["kernel_ms"] = st.last_kernel_ms
        self.last_stats["scan_ms"] = st.last_scan_ms
        self.last_stats["scan_grid"] = st.last_scan_grid
        self.last_stats["scan_stages"] = st.last_scan_stages
        self.last_stats["scan_smem"] = st.last_scan_smem
        self.last_stats["scan_kind"] = st.last_scan_kind  # N.K["MSC_SCAN_KIND_*"]: which kernel ran this (aggregate) scan
        self.last_stats["launches"] = st.launches

    def _run_select(self, sel: L.LSelect, translate_targets: Optional[dict[str, DictHandle]] = None) -> DeviceRel:
        nf = len(sel.filters)
        source, exprs = self._prepare(sel.child, [*sel.filters, *sel.outputs])
        filters, outputs = exprs[:nf], exprs[nf:]
        rel = self._project(source, filters, outputs, translate_targets, node=sel)
        rel.keep.extend(source.keep)
        return rel

    def _project(self, source: _Source, filters: list[L.Expr], outputs: list[L.Expr], translate_targets: Optional[dict[str, DictHandle]],
                 node: Any = None) -> DeviceRel:
        """One filter + project scan -- or several over the same rows when the outputs exceed what one scan can bind
        (MSC_VM_MAX_OUT output columns, MSC_VM_MAX_STAGED staged / MSC_VM_MAX_GATHER gathered inputs: `SELECT *` over a wide
        join).  Every pass evaluates the same filters and compaction is stable, so the passes' columns line up row by row."""
        # An unfiltered projection keeps every row, so an output that is just an input column -- in the physical type results
        # have anyway -- IS that column: it is shared, not copied (the AVG projection over a GROUP BY's 15 M groups read and
        # wrote the key and the SUM only to pass them on: 56 -> 24 bytes per group).
        if not filters and not source.pre_filters and source.probe_key is None and not translate_targets and not source.translate_targets \
                and os.environ.get("MINISPARK_SHARE_COLUMNS", "1") != "0":
            result_phys = {L.INT: N.P_I64, L.TS: N.P_I64, L.FLOAT: N.P_F64, L.STR: N.P_U32}
            shared = {j: source.columns[e.index] for j, e in enumerate(outputs)
                      if isinstance(e, L.EInput) and source.columns[e.index].via is None and source.columns[e.index].phys == result_phys.get(e.type)}
            if shared:
                computed = [j for j in range(len(outputs)) if j not in shared]
                cols: list[Optional[DeviceColumn]] = [shared.get(j) for j in range(len(outputs))]
                keep: list = []
                if computed:
                    part = self._project(source, [], [outputs[j] for j in computed], None, node=node)
                    if part.nrows != source.nrows:
                        raise ExecutionError("an unfiltered projection changed the row count")
                    for j, c in zip(computed, part.cols):
                        cols[j] = c
                    keep.append(part)
                rel = DeviceRel(self.ctx, None, source.nrows, [c for c in cols if c is not None], keep=keep)
                rel.partitioned = source.partitioned
                return self._materialised(rel)
        try:
            if len(outputs) > N.K["MSC_VM_MAX_OUT"]:
                raise L.LoweringError("too many output columns in one projection")
            prog, desc = self._compiled_scan(
                ("project", id(node)) if node is not None else None, node, source, translate_targets, (tuple(filters), tuple(outputs), tuple(source.pre_filters), source.probe_key),
                lambda r: L.compile_project(r, filters, outputs, probe=r.probe_spec(), pre_filters=source.pre_filters))
            return self._scan_project(prog, desc, [e.type for e in outputs], source.partitioned)
        except L.LoweringError as err:
            splittable = any(t in str(err) for t in ("too many output columns", "more columns than one scan", "gathers more columns"))
            if len(outputs) < 2 or not splittable:
                raise
        half = len(outputs) // 2
        first = self._project(source, filters, outputs[:half], translate_targets)
        second = self._project(source, filters, outputs[half:], translate_targets)
        if first.nrows != second.nrows:
            raise ExecutionError("projection passes disagree on the surviving rows")
        rel = DeviceRel(self.ctx, None, first.nrows, [*first.cols, *second.cols], keep=[first, second])
        rel.partitioned = first.partitioned
        return self._materialised(rel)

    def _run_aggregate(self, agg: L.LAggregate) -> DeviceRel:
        child = agg.child
        if isinstance(child, L.LSelect):  # fuse the filter / projection below the aggregate into the same scan
            filters = list(child.filters)
            group = L.substitute(agg.group, child.outputs)
            aggs = [(k, L.substitute(e, child.outputs)) for k, e in agg.aggs]
            base = child.child
        else:
            filters, group, aggs, base = [], agg.group, list(agg.aggs), child
        exprs = [*filters, group, *[e for _, e in aggs]]
        source, exprs = self._prepare(base, exprs)
        nf = len(filters)
        filters, group = exprs[:nf], exprs[nf]
        aggs = [(k, e) for (k, _), e in zip(aggs, exprs[nf + 1:])]
        prog, desc = self._compiled_scan(
            ("aggregate", id(agg)), agg, source, None, (tuple(filters), group, tuple(aggs), tuple(source.pre_filters), source.probe_key),
            lambda r: L.compile_aggregate(r, filters, group, aggs, probe=r.probe_spec(), pre_filters=source.pre_filters))
        ngroups, hint = self._dense_groups(prog)
        kinds = N.int32_array(prog.agg_kinds)
        # raw result: key + one column per unique accumulator slot; expose it in the aggregate's schema order
        slot_types = [L.FLOAT if k in (N.K["MSC_AGG_SUM_F"], N.K["MSC_AGG_MIN_F"], N.K["MSC_AGG_MAX_F"]) else L.INT for k in prog.agg_kinds]
        self.last_stats["agg_mode"] = "dense" if ngroups else "hash"
        spread = source.partitioned and self.comm.world > 1  # the rows are spread over the ranks: partial results must merge
        partitioned = False
        if ngroups and spread:  # dense tables merge across ranks without moving rows (_DenseMerge)
            merge = self._dense_merge(prog.group_dict, desc, kinds, prog.agg_kinds)
            handle, _ = merge.run(desc)
            self._note_kernel()
            self.last_stats["exchange"] = merge.exchange_kind
            self.last_stats.setdefault("exchanges", []).append(merge.exchange_kind)
            raw = self._track(DeviceRel.from_handle(self.ctx, handle, [group.type, *slot_types], [merge.global_dict] + [None] * len(slot_types)))
        else:
            out = C.c_void_p()
            if not ngroups and prog.group_dict is None and agg.groups_seen:
                hint = agg.groups_seen | N.K["MSC_HASH_HINT_SOFT"]  # no dictionary bounds the groups, but this plan has run before
            self.ctx.call("msc_scan_aggregate", C.byref(desc), ngroups, kinds, len(prog.agg_kinds), hint, C.byref(out))
            self._note_kernel("scan: filter + aggregate")
            self.last_stats["agg_scan_kind"] = self.last_stats["scan_kind"]  # later scans (final projection) overwrite scan_kind
            if not ngroups:
                st = self.ctx.stats()
                self.last_stats["hash_local_slots"], self.last_stats["hash_attempts"] = st.last_hash_local_slots, st.last_hash_attempts
                self.last_stats["run_index_hit"] = bool(st.last_run_index_hit)
            self.last_stats["agg_scan_ms"] = self.last_stats["scan_ms"]
            raw = self._track(DeviceRel.from_handle(self.ctx, out.value, [group.type, *slot_types], [prog.group_dict] + [None] * len(slot_types)))
            if not ngroups:
                agg.groups_seen = raw.nrows
            if spread:  # merge the per-rank partial tables (reference: shuffle + final aggregate, plan.py:190-199)
                raw = self._merge_partials(raw, prog.agg_kinds, slot_types, group.type)
                partitioned = raw.partitioned
        key = raw.cols[0]
        if group.type == L.FLOAT and key.phys == N.P_I64:
            key = DeviceColumn(key.ptr, N.P_F64, L.FLOAT)  # hash mode stores the f64 bit pattern
        cols = [key] + [raw.cols[1 + s] for s in prog.slot_of]
        rel = DeviceRel(self.ctx, None, raw.nrows, cols, keep=[raw, *source.keep])
        rel.partitioned = partitioned
        return rel

    def _dense_merge(self, group_dict: DictHandle, desc: N.ScanDesc, kinds: Any, agg_kinds: list[int]) -> "_DenseMerge":
        """The cross-rank merge set-up of a dense GROUP BY (unified key dictionary, code permutations, exchange buffers) --
        kept per key dictionary and accumulator layout when the key is a table column (several host collectives otherwise
        repeat for every query)."""
        stride, count_slot = C.c_int32(), C.c_int32()
        self.ctx.call("msc_dense_layout", C.byref(desc), kinds, len(agg_kinds), C.byref(stride), C.byref(count_slot))
        key = (group_dict.serial, group_dict.size, stride.value, count_slot.value, tuple(agg_kinds))
        merge = self._dense_merges.get(key) if group_dict.persistent else None
        if merge is None:
            merge = _DenseMerge(self, group_dict, desc, kinds, len(agg_kinds), persistent=group_dict.persistent, jit=False)
            if group_dict.persistent:
                self._dense_merges[key] = merge
        return merge

    def _dense_groups(self, prog: L.AggregateProgram) -> tuple[int, int]:
        """(groups of the dense table or 0 for hash mode, capacity hint).  Dense mode needs a dictionary-coded key and a
        table that fits shared memory; with several ranks all of them must take the same decision."""
        if prog.group_dict is None:
            return 0, 0
        size = prog.group_dict.size
        limit = size
        if self.comm.world > 1 and not prog.group_dict.same_on_all_ranks:
            limit = max(r[0] for r in self._gather_counts([size]))
        dense = limit > 0 and (limit + 1) * (len(prog.agg_kinds) + 1) <= DENSE_MAX_CELLS
        if self.comm.world == 1:
            dense = dense and size > 0
        return (max(size, 1) if dense else 0), max(size, 1)

    def _join_side(self, node: L.LNode, idxs: list[int], key: L.Expr, targets: Optional[dict[str, DictHandle]]) -> DeviceRel:
        """One side of a join as a relation: the needed columns, then the key (a STR key as its dictionary code)."""
        memo = node.__dict__.setdefault("_side_selects", {})  # (the same node object next time: its compiled scan is found again)
        sel = memo.get((tuple(idxs), key))
        if sel is None:
            outs = [L.EInput(_LTYPE_OF[node.schema[i][1]], i) for i in idxs]
            key_out = L.ECode(L.INT, key) if key.type == L.STR else key
            schema = [node.schema[i] for i in idxs] + [("__key", ColumnType.INTEGER)]
            sel = memo[(tuple(idxs), key)] = L.fuse_selects(L.LSelect(schema, node, [], [*outs, key_out]))
        return self._run_select(sel, targets)

    def _rows_upper_bound(self, node: L.LNode) -> Optional[int]:
        """Rows the relation of `node` can have over ALL ranks, from file footers alone (every rank computes the same number
        without a message), or None when only the data can tell."""
        if isinstance(node, L.LTable):
            return self._table(node.path).total_rows
        if isinstance(node, (L.LSelect, L.LAggregate)):
            return self._rows_upper_bound(node.child)
        return None

    def _probe_source(self, join: L.LJoin, exprs: list[L.Expr], needed: set[int]) -> Optional[tuple[_Source, list[L.Expr]]]:
        """A join as a LOOKUP inside the scan that consumes it.  The reference builds key -> [left rows] and streams the right
        rows through it (tasks.py:201-240); when no key repeats on the left, every right row has at most one partner, so the
        consumer can walk the right side itself: its own filters, then the probe (MSC_OP_PROBE), then everything above the
        join, with left columns read through the matched row.  Neither side's rows nor the pair list are materialised beyond
        the (filtered) build side; output order is the right side's row order, which is the reference's right-row-major
        order.  None: not applicable here (duplicate build keys, expressions that must be evaluated on unmatched rows too, a
        build side too large to give every rank a copy) -- the caller runs the materialising join.

        Several ranks: the (filtered) build side is BROADCAST -- every rank receives all of its rows over NVLink
        (_gather_rows) and probes it with its own part of the probe side, which never moves.  The reference's task is
        called BroadcastHashJoinTask but shuffles both sides (plan.py:186-189); with the probe side several times the
        build side, sending the small side everywhere moves less than co-partitioning both."""
        right = join.right
        rsel = right if isinstance(right, L.LSelect) else L.identity_select(right)
        if not all(L._cannot_raise(e) for e in rsel.outputs) or any(_has_concat(e) for e in [*exprs, *rsel.outputs, *rsel.filters]):
            return None
        nl = len(join.left.schema)
        left_needed = sorted(i for i in needed if i < nl)
        fused = self._fused_build(join, left_needed) if self.fused_build else None
        if fused is not None:
            trel, left_columns, key_dict, keep_left = fused
            return self._probe_over(join, exprs, rsel, left_columns, trel, True, key_dict, keep_left, False, scanned_build=True)
        lrel = self._join_side(join.left, left_needed, join.left_key, None)
        key_dict = lrel.cols[-1].dict  # (a STR key is joined on its code in this dictionary)
        broadcast = False
        if self.comm.world > 1 and lrel.partitioned:
            limit = int(os.environ.get("MSC_BROADCAST_JOIN_MAX", str(32 << 20)))
            bound = self._rows_upper_bound(join.left)  # (a build side that cannot exceed the limit needs no count exchange)
            total = bound if bound is not None and bound <= limit else sum(r[0] for r in self._gather_counts([lrel.nrows]))
            if total > limit:
                self._probe_declined = lrel
                return None
            if join.left_key.type == L.STR:  # the key travels as a code: of a dictionary every rank shares
                key_dict = self._unified_dictionary(key_dict)
                lrel = self._join_side(join.left, left_needed, L.ETranslate(L.STR, join.left_key, "join"), {"join": key_dict})
            lrel = self._gather_rows(lrel)
            broadcast = True
        table, unique, slot_bytes = C.c_void_p(), C.c_int32(), C.c_int32()
        self.ctx.call("msc_join_build", C.c_void_p(lrel.cols[-1].ptr), lrel.nrows, C.byref(table), C.byref(unique), C.byref(slot_bytes))
        self._note_kernel("hash join: build")
        trel = self._track(DeviceRel.from_handle(self.ctx, table.value, [L.INT], [None]))
        if not unique.value:
            if not broadcast:
                self._probe_declined = lrel  # (the materialising join reuses the build side it already has)
            return None
        left_columns = {}
        for pos, i in enumerate(left_needed):
            c = lrel.cols[pos]
            left_columns[_PROBE_BASE + i] = DeviceColumn(c.ptr, c.phys, c.ltype, c.dict, via="probe")
        return self._probe_over(join, exprs, rsel, left_columns, trel, slot_bytes.value == 8, key_dict, [lrel], broadcast)

    def _probe_over(self, join: L.LJoin, exprs: list[L.Expr], rsel: L.LSelect, left_columns: dict[int, DeviceColumn], trel: DeviceRel,
                    compact: bool, key_dict: Optional[DictHandle], keep_left: list, broadcast: bool, scanned_build: bool = False) -> tuple[_Source, list[L.Expr]]:
        """The consuming scan's source for a join whose table is built: the probe side's base rows, its own filters, the probe
        and the build side's columns behind it."""
        inputs: list[L.Expr] = [L.EInput(_LTYPE_OF[t], _PROBE_BASE + i) for i, (_, t) in enumerate(join.left.schema)] + list(rsel.outputs)
        new_exprs = [L.substitute(e, inputs) for e in exprs]
        rkey = L.substitute(join.right_key, list(rsel.outputs))
        targets = None
        if rkey.type == L.STR:  # joined on the code in the LEFT key column's dictionary
            targets = {"join": key_dict}
            rkey = L.ECode(L.INT, L.ETranslate(L.STR, rkey, "join"))
        rneeded: set[int] = set()
        for e in [*new_exprs, *rsel.filters, rkey]:
            rneeded |= {i for i in L.expr_inputs(e) if i < _PROBE_BASE}
        rsource = self._source(rsel.child, rneeded)
        columns = dict(rsource.columns)
        columns.update(left_columns)
        source = _Source(rsource.nrows, columns, rsource.index_vectors, keep=[*rsource.keep, *keep_left, trel], partitioned=rsource.partitioned)
        source.pre_filters = list(rsel.filters)
        source.probe_key = rkey
        source.probe_table = trel.cols[0].ptr
        source.probe_compact = compact
        source.translate_targets = targets
        self.last_stats["join"] = ("lookup fused into the consuming scan (MSC_OP_PROBE)" + (", build side broadcast to every rank" if broadcast else "")
                                   + (", table built by one scan over the build side's base rows" if scanned_build else ""))
        return source, new_exprs

    def _fused_build(self, join: L.LJoin, left_needed: list[int]) -> Optional[tuple[DeviceRel, dict[int, DeviceColumn], Optional[DictHandle], list]]:
        """The join table straight from the build side's BASE rows (msc_scan_join_build): its filters and key in one scan, the
        rows that pass insert (key, base row number); build-side columns are then read from the base table through the match.
        For `filters over a table` build sides on one rank; None: use the materialising build (which also handles duplicate
        and wide keys, computed columns, relations that must travel between ranks)."""
        left = join.left
        lsel = left if isinstance(left, L.LSelect) else L.identity_select(left)
        base = lsel.child
        if not isinstance(base, L.LTable) or self.comm.world > 1:  # (several ranks: the build side's rows must travel)
            return None
        if not all(isinstance(lsel.outputs[i], L.EInput) for i in left_needed):
            return None
        key = L.substitute(join.left_key, list(lsel.outputs))
        if not L._cannot_raise(key) or any(_has_concat(e) for e in [key, *lsel.filters]):
            return None
        key_out = L.ECode(L.INT, key) if key.type == L.STR else key
        need: set[int] = {lsel.outputs[i].index for i in left_needed}
        for e in [key, *lsel.filters]:
            need |= L.expr_inputs(e)
        lsource = self._source(base, need)
        resolver = _ScanResolver(self, lsource)
        program, key_dict = L.compile_build(resolver, list(lsel.filters), key_out)
        desc = resolver.desc(program)
        table, usable, nkeys = C.c_void_p(), C.c_int32(), C.c_uint64()
        self.ctx.call("msc_scan_join_build", C.byref(desc), C.byref(table), C.byref(usable), C.byref(nkeys))
        self._note_kernel("hash join: filter + build in one scan")
        if not usable.value:
            return None
        trel = self._track(DeviceRel.from_handle(self.ctx, table.value, [L.INT], [None]))
        columns = {}
        for i in left_needed:
            c = lsource.columns[lsel.outputs[i].index]
            columns[_PROBE_BASE + i] = DeviceColumn(c.ptr, c.phys, c.ltype, c.dict, via="probe")
        return trel, columns, key_dict, list(lsource.keep)

    def _join_source(self, join: L.LJoin, needed: set[int]) -> _Source:
        nl = len(join.left.schema)
        left_needed = sorted(i for i in needed if i < nl)
        right_needed = sorted(i - nl for i in needed if i >= nl)
        side = self._join_side
        self.last_stats["join"] = "build / probe / emit pairs (msc_hash_join)"

        # A STR key is joined on its dictionary code (key_out above), so both sides must be coded in ONE dictionary:
        # the left key column's -- unified over all ranks first when the tables are sharded.
        declined, self._probe_declined = getattr(self, "_probe_declined", None), None
        if declined is not None and len(declined.cols) == len(left_needed) + 1:
            lrel = declined
        else:
            lrel = side(join.left, left_needed, join.left_key, None)
        key_dict = lrel.cols[-1].dict if join.left_key.type == L.STR else None
        if key_dict is not None and self.comm.world > 1:
            key_dict = self._unified_dictionary(key_dict)
            lrel = side(join.left, left_needed, L.ETranslate(L.STR, join.left_key, "join"), {"join": key_dict})
        targets = None
        rkey = join.right_key
        if rkey.type == L.STR:
            targets = {"join": key_dict}
            rkey = L.ETranslate(L.STR, rkey, "join")
        rrel = side(join.right, right_needed, rkey, targets)
        partitioned = False
        if self.comm.world > 1:  # the reference shuffles both sides on the key (plan.py:186-189): same here, over NVLink
            lrel = self._exchange_on_key(lrel)
            rrel = self._exchange_on_key(rrel)
            partitioned = True
        pairs = C.c_void_p()
        self.ctx.call("msc_hash_join", C.c_void_p(lrel.cols[-1].ptr), lrel.nrows, C.c_void_p(rrel.cols[-1].ptr), rrel.nrows, C.byref(pairs))
        self._note_kernel("hash join: build + probe + emit")
        prel = self._track(DeviceRel.from_handle(self.ctx, pairs.value, [L.INT, L.INT], [None, None]))
        columns: dict[int, DeviceColumn] = {}
        for pos, i in enumerate(left_needed):
            c = lrel.cols[pos]
            columns[i] = DeviceColumn(c.ptr, c.phys, c.ltype, c.dict, via=0)
        for pos, i in enumerate(right_needed):
            c = rrel.cols[pos]
            columns[nl + i] = DeviceColumn(c.ptr, c.phys, c.ltype, c.dict, via=1)
        return _Source(prel.nrows, columns, index_vectors=[prel.cols[0], prel.cols[1]], keep=[lrel, rrel, prel], partitioned=partitioned)

    def _unified_dictionary(self, local: DictHandle) -> DictHandle:
        """The same dictionary on every rank: the sorted union of all ranks' entries (code = position).  For a table column's
        dictionary the result is kept (every rank runs the same queries on the same tables, so all keep or rebuild alike)."""
        if local.persistent and local._unified is not None and local._unified[0] == local.size and local._unified[1].handle:
            return local._unified[1]
        universe, _ = unify_keys(self.comm.all_gather_object(local.export()))
        unified = DictHandle(self.ctx).load(universe)
        unified.same_on_all_ranks = True
        if local.persistent:
            unified.persistent = True
            local._unified = (local.size, unified)
            self._table_dicts.append(unified)
        else:
            self._query_dicts.append(unified)
        return unified

    def _rank_independent(self, rel: DeviceRel) -> DeviceRel:
        """STR columns re-coded into dictionaries that are the same on every rank, so that codes can travel."""
        recode = [c.dict is not None and c.ltype == L.STR for c in rel.cols]
        if not any(recode):
            return rel
        ltypes = [c.ltype for c in rel.cols]
        targets: dict[str, DictHandle] = {}
        outs: list[L.Expr] = []
        for i, c in enumerate(rel.cols):
            if not recode[i]:
                outs.append(L.EInput(c.ltype, i))
                continue
            targets[f"x{i}"] = self._unified_dictionary(c.dict)
            outs.append(L.ETranslate(L.STR, L.EInput(L.STR, i), f"x{i}"))
        source = _Source(rel.nrows, dict(enumerate(rel.cols)), keep=[rel], partitioned=rel.partitioned)
        signature, _ = _CompiledScan.signature_of(source, targets)
        prog, desc = self._compiled_scan(("rank-independent", signature), None, source, targets, tuple(outs), lambda r: L.compile_project(r, [], outs))
        return self._scan_project(prog, desc, ltypes, source.partitioned)

    def _local_part(self, rel: DeviceRel) -> DeviceRel:
        """A relation every rank holds in full enters an exchange from rank 0 only (the other ranks contribute no rows)."""
        if rel.partitioned or self.comm.rank == 0:
            return rel
        out = C.c_void_p()
        self.ctx.call("msc_rel_alloc", 0, N.int32_array([c.phys for c in rel.cols]), len(rel.cols), C.byref(out))
        return self._track(DeviceRel.from_handle(self.ctx, out.value, [c.ltype for c in rel.cols], [c.dict for c in rel.cols]))

    def _materialised(self, rel: DeviceRel) -> DeviceRel:
        """The relation as ONE library relation of directly readable columns (exchanges take an msc_rel)."""
        if rel.handle and all(c.via is None for c in rel.cols):
            return rel
        binds = (N.ColBind * max(len(rel.cols), 1))()
        for i, c in enumerate(rel.cols):
            binds[i].data, binds[i].phys = c.ptr, c.phys
        out = C.c_void_p()
        self.ctx.call("msc_rel_wrap", rel.nrows, binds, len(rel.cols), C.byref(out))
        wrapped = self._track(DeviceRel.from_handle(self.ctx, out.value, [c.ltype for c in rel.cols], [c.dict for c in rel.cols]))
        wrapped.keep.append(rel)
        wrapped.partitioned = rel.partitioned
        return wrapped

    def _gather_counts(self, values: list[int]) -> list[list[int]]:
        """[rank][i] = values[i] of that rank: through the peer control blocks where available, else a host collective."""
        sh = self._peer_shuffle()
        return sh.allgather_ints(values) if sh else self.comm.all_gather_counts(values)

    def _peer_shuffle(self) -> Any:
        """The library's exchange over NVLink peer memory, set up collectively on first use (False: not available)."""
        if self._shuffle is None:
            self._shuffle = False
            if os.environ.get("MINISPARK_PEER_EXCHANGE", "1") != "0":
                sh = PeerShuffle(self.ctx, self.comm)
                if sh.ok:
                    self._shuffle = sh
        return self._shuffle

    def _exchange_rows(self, rel: DeviceRel, key_col: Optional[int], lower_bounds: Optional[list[int]] = None) -> DeviceRel:
        """One shuffle: every row to rank hash(column key_col) % world, or with key_col None every row to every rank
        (rows arrive ordered by sending rank, then input order).  The reference writes and re-reads shuffle files here
        (tasks.py:347-375, 144-150); this is a push over NVLink into the receiver's memory (csrc/shuffle.cu), or one group
        of NCCL sends / receives where CUDA IPC is unavailable."""
        rel = self._materialised(rel)
        ltypes, dicts = [c.ltype for c in rel.cols], [c.dict for c in rel.cols]
        sh = self._peer_shuffle()
        t0 = time.perf_counter()
        how = " (all rows to all ranks)" if key_col is None else (" (range partitioned: sorted partial results)" if lower_bounds else " (hash partitioned)")
        if sh:
            handle, _ = sh.exchange(rel.handle, key_col, lower_bounds)
            out = self._track(DeviceRel.from_handle(self.ctx, handle, ltypes, dicts))
            out.keep.append(rel)
            self.last_stats["exchange"] = "nvlink peer push" + how
            matrix = sh.last_matrix
        else:
            out, matrix = self._exchange_rows_nccl(rel, key_col, lower_bounds)
            self.last_stats["exchange"] = "nccl send/recv group" + how
        self.last_stats.setdefault("exchanges", []).append(self.last_stats["exchange"])
        widths = sum(N.PHYS_WIDTH[c.phys] for c in rel.cols)
        me = self.comm.rank
        rows_sent = sum(n for d, n in enumerate(matrix[me]) if d != me)
        first = len(self.last_stats["exchanges"]) == 1 or "exchange_rows_sent" not in self.last_stats
        self.last_stats["exchange_rows_sent"] = rows_sent + (0 if first else self.last_stats["exchange_rows_sent"])
        self.last_stats["exchange_bytes_sent"] = rows_sent * widths + (0 if first else self.last_stats["exchange_bytes_sent"])
        self.last_stats["exchange_host_s"] = time.perf_counter() - t0 + (0.0 if first else self.last_stats["exchange_host_s"])
        out.partitioned = key_col is not None
        return out

    def _exchange_rows_nccl(self, rel: DeviceRel, key_col: Optional[int], lower_bounds: Optional[list[int]] = None) -> tuple[DeviceRel, list[list[int]]]:
        import torch  # noqa: PLC0415

        comm = self.comm
        if key_col is None:
            prel, counts = rel, [rel.nrows] * comm.world
            cols = [self._torch_column(c, rel.nrows).repeat(comm.world) for c in rel.cols]
        else:
            host_counts = (C.c_uint64 * comm.world)()
            part = C.c_void_p()
            if lower_bounds is not None:
                self.ctx.call("msc_partition_range", C.c_void_p(rel.handle), key_col, comm.world, (C.c_int64 * comm.world)(*lower_bounds), host_counts,
                              C.byref(part))
            else:
                self.ctx.call("msc_partition", C.c_void_p(rel.handle), key_col, comm.world, host_counts, C.byref(part))
            prel = self._track(DeviceRel.from_handle(self.ctx, part.value, [c.ltype for c in rel.cols], [c.dict for c in rel.cols]))
            counts = [int(c) for c in host_counts]
            cols = [self._torch_column(c, prel.nrows) for c in prel.cols]
        stream = C.c_void_p()
        self.ctx.call("msc_stream_handle", C.byref(stream))
        dev = torch.device("cuda", self.device)
        matrix = comm.all_gather_counts(counts)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream.value, device=dev)):  # ordered with the library's own work
            pad = 2 * 8192  # tile padding the scan kernels may read past the last row (common.cuh MSC_ROW_PAD)
            received, _ = comm.all_to_all_rows(cols, counts, pad_rows=pad, matrix=matrix)
        nrows = sum(row[comm.rank] for row in matrix)
        binds = (N.ColBind * max(len(received), 1))()
        for i, (t, c) in enumerate(zip(received, prel.cols)):
            binds[i].data, binds[i].phys = t.data_ptr(), c.phys
        out = C.c_void_p()
        self.ctx.call("msc_rel_wrap", nrows, binds, len(received), C.byref(out))
        got = self._track(DeviceRel.from_handle(self.ctx, out.value, [c.ltype for c in prel.cols], [c.dict for c in prel.cols]))
        got.keep.extend([received, prel])
        return got, matrix

    def _exchange_on_key(self, rel: DeviceRel) -> DeviceRel:
        """Multi-rank join: send every row to rank hash(key) % world (key = last column, an integer), so that equal
        keys of both sides meet on one GPU.  STR columns are re-coded into rank-independent dictionaries first.
        (the key column is an integer already: a STR key was coded in the unified key dictionary by the caller)"""
        rel = self._local_part(self._rank_independent(rel))
        return self._exchange_rows(rel, len(rel.cols) - 1)

    def _gather_rows(self, rel: DeviceRel) -> DeviceRel:
        """The complete relation on every rank: the rank-local parts concatenated in rank order (for sharded scans that is
        the table's row order, see distributed.shard_blocks)."""
        if self.comm.world == 1 or not rel.partitioned:
            return rel
        out = self._exchange_rows(self._rank_independent(rel), None)
        out.partitioned = False
        return out

    # ---- prepared queries ------------------------------------------------------------------------------
    def prepare(self, full_task: Any) -> "PreparedAggregate":
        """Compile an aggregate query once; ``run()`` then re-executes it with no Python lowering.

        Supported shape: ``table -> [filter/project] -> GROUP BY -> [having / final select]`` (e.g. TPC-H Q1)."""
        task = deepcopy(full_task)
        task.validate_schema()
        plan = L.lower_task(task)
        if not (isinstance(plan, L.LSelect) and isinstance(plan.child, L.LAggregate)):
            raise L.LoweringError("prepare() supports aggregate queries only")
        return PreparedAggregate(self, plan)

    # ---- multi-GPU exchange ---------------------------------------------------------------------------
    def _torch_column(self, col: DeviceColumn, nrows: int):  # noqa: ANN202
        import torch  # noqa: PLC0415

        width = N.PHYS_WIDTH[col.phys]
        raw = torch.as_tensor(_DevView(col.ptr, nrows * width), device=f"cuda:{self.device}")
        return raw[: nrows * width].view(_torch_dtype(col.phys))

    def _merge_partials(self, raw: DeviceRel, agg_kinds: list[int], slot_types: list[str], group_type: str) -> DeviceRel:
        """Combine per-rank partial aggregates (the reference's shuffle + final aggregate, plan.py:190-199): small tables
        go to every rank and are re-aggregated there; large ones are hash-partitioned on the key and exchanged
        (_exchange_rows), which leaves every rank with the final groups of its share of the keys."""
        comm = self.comm
        key_dict = raw.cols[0].dict
        global_dict = None
        if key_dict is not None:  # unify string keys through their dictionary entries
            global_dict = self._unified_dictionary(key_dict)
            source = _Source(raw.nrows, dict(enumerate(raw.cols)), keep=[raw])
            resolver = _ScanResolver(self, source, {"global": global_dict})
            outs = [L.ECode(L.INT, L.ETranslate(L.STR, L.EInput(L.STR, 0), "global"))]
            outs += [L.EInput(t, i + 1) for i, t in enumerate(slot_types)]
            prog = L.compile_project(resolver, [], outs)
            raw = self._scan_project(prog, resolver.desc(prog.program), [L.INT, *slot_types], source.partitioned)
        # what every rank holds: rows, and -- when its pre-aggregation streamed over the runs of a sorted key (msc_stats.last_agg_runs),
        # so that its partial rows ascend and no key repeats on the rank -- the first key and the whole last row
        ascending = bool(self.ctx.stats().last_agg_runs) and group_type in (L.INT, L.TS) and raw.nrows > 0 and raw.handle is not None
        ncols = len(raw.cols)
        first, last_row = 0, [0] * ncols
        if ascending:
            out = (C.c_int64 * (2 * ncols))()
            self.ctx.call("msc_rel_read_rows", C.c_void_p(raw.handle), (C.c_uint64 * 2)(0, raw.nrows - 1), 2, out)
            first, last_row = int(out[0]), [int(out[ncols + i]) for i in range(ncols)]
        everyone = self._gather_counts([int(ascending or raw.nrows == 0), raw.nrows, first, *last_row])
        per_rank = [(r[0], r[1], r[2], r[3]) for r in everyone]
        small = max(r[1] for r in per_rank) <= int(os.environ.get("MSC_EXCHANGE_GATHER_MAX", "65536"))
        bounds = None if small or os.environ.get("MSC_EXCHANGE_RANGE", "1") == "0" else range_bounds([(bool(r[0]), r[1], r[2], r[3]) for r in per_rank])
        if bounds is not None and os.environ.get("MSC_EXCHANGE_BOUNDARY", "1") != "0":
            # Sorted partial results whose key ranges follow each other in rank order: only a group that straddles two ranks has a
            # partner anywhere.  Its row -- the last row of the earlier rank -- is folded into the owner's first row and dropped
            # where it came from; no other row moves and nothing is re-aggregated (the reference would shuffle every partial row
            # and aggregate them again, plan.py:190-199).  Every rank decides from the same gathered numbers.
            drop_last, incoming = boundary_plan([(r[1], r[2], r[3]) for r in everyone], bounds, comm.rank)
            kinds = N.int32_array([-1, *agg_kinds])
            for s in incoming:
                values = (C.c_int64 * ncols)(*everyone[s][3:3 + ncols])
                self.ctx.call("msc_rel_fold_row", C.c_void_p(raw.handle), 0, values, kinds)
            if drop_last:
                raw.nrows -= 1
            self.last_stats["exchange"] = "boundary rows over nvlink peer control blocks (sorted partial results: no other row moves)"
            self.last_stats.setdefault("exchanges", []).append(self.last_stats["exchange"])
            self.last_stats["exchange_rows_sent"] = int(drop_last)
            self.last_stats["exchange_bytes_sent"] = int(drop_last) * 8 * ncols
            self.last_stats["exchange_host_s"] = 0.0
            if global_dict is not None:
                raw.cols[0] = DeviceColumn(raw.cols[0].ptr, raw.cols[0].phys, group_type, global_dict)
            raw.partitioned = True
            return raw
        gathered = self._exchange_rows(raw, None if small else 0, bounds)
        # final aggregate over the partial rows: SUM of sums / counts, MIN of mins, MAX of maxes
        merge_kind = {N.K["MSC_AGG_SUM_F"]: "sum", N.K["MSC_AGG_SUM_I"]: "sum", N.K["MSC_AGG_MIN_F"]: "min", N.K["MSC_AGG_MIN_I"]: "min",
                      N.K["MSC_AGG_MAX_F"]: "max", N.K["MSC_AGG_MAX_I"]: "max"}
        source = _Source(gathered.nrows, dict(enumerate(gathered.cols)), keep=[gathered])
        resolver = _ScanResolver(self, source)
        prog2 = L.compile_aggregate(resolver, [], L.EInput(L.INT, 0), [(merge_kind[k], L.EInput(t, i + 1)) for i, (k, t) in enumerate(zip(agg_kinds, slot_types))])
        desc = resolver.desc(prog2.program)
        out = C.c_void_p()
        # (no capacity hint after a range-partitioned exchange: the rows a rank received ascend, so the library's run count
        # finds every run to be one group and streams over them instead of probing a table)
        hint = 0 if bounds is not None else max(gathered.nrows, 1)
        self.ctx.call("msc_scan_aggregate", C.byref(desc), 0, N.int32_array(prog2.agg_kinds), len(prog2.agg_kinds), hint, C.byref(out))
        self._note_kernel()
        # prog2 may carry one accumulator more than asked for (the hidden COUNT of the regvm encoding): type every
        # column the library returns, then keep the requested ones in order
        types2 = [L.FLOAT if k in (N.K["MSC_AGG_SUM_F"], N.K["MSC_AGG_MIN_F"], N.K["MSC_AGG_MAX_F"]) else L.INT for k in prog2.agg_kinds]
        merged = self._track(DeviceRel.from_handle(self.ctx, out.value, [group_type, *types2], [global_dict] + [None] * len(types2)))
        merged.cols = [merged.cols[0]] + [merged.cols[1 + s] for s in prog2.slot_of]
        merged.partitioned = not small  # the hash-partitioned exchange leaves every rank with a disjoint set of keys
        return merged


class _DenseMerge:
    """Cross-rank merge of a low-cardinality (dense) GROUP BY, set up once per query.

    Every rank scans its row-blocks into a [groups][stride] table of its own dictionary codes; the tables are
    all-gathered (a few hundred bytes per rank: one small kernel that stores into the peers' control blocks over NVLink,
    msc_shuffle_allgather; NCCL where CUDA IPC is unavailable), folded in rank order by ``msc_dense_merge`` through
    per-rank code permutations into the unified dictionary, and compacted.  Every rank ends up with the complete,
    bit-identical result.  This is the reference's pre-aggregate -> shuffle -> final aggregate (plan.py:190-199)
    without a row ever leaving its GPU."""

    def __init__(self, engine: "CudaExecutionEngine", group_dict: DictHandle, desc: N.ScanDesc, kinds, naggs: int,  # noqa: ANN001
                 persistent: bool = False, jit: bool = True) -> None:
        import torch  # noqa: PLC0415

        self.engine = engine
        self.kinds, self.naggs = N.int32_array(list(kinds)[:naggs]), naggs  # (own copy: the object may outlive the caller's array)
        self.persistent = persistent  # the unified dictionary lives as long as the engine (prepared queries, cached set-ups)
        self.jit = jit and persistent  # prepared queries always run on a kernel specialised for them
        comm = engine.comm
        stride, count_slot = C.c_int32(), C.c_int32()
        engine.ctx.call("msc_dense_layout", C.byref(desc), kinds, naggs, C.byref(stride), C.byref(count_slot))
        self.stride, self.count_slot = stride.value, count_slot.value
        layouts = comm.all_gather_object((self.stride, self.count_slot))
        if len(set(layouts)) != 1:
            raise ExecutionError("ranks disagree on the aggregate table layout")
        per_rank = comm.all_gather_object(group_dict.export())
        universe, maps = unify_keys(per_rank)
        self.nlocal = len(per_rank[comm.rank])
        self.gmax = max(max(len(e) for e in per_rank), 1)
        self.nglobal = len(universe)
        flat = []
        for m in maps:
            flat.extend(list(m) + [-1] * (self.gmax - len(m)))
        self.perm = N.int32_array(flat)
        self.global_dict = DictHandle(engine.ctx).load(universe)
        # a prepared query outlives release_query(): its dictionary is freed with the engine instead
        (engine._table_dicts if persistent else engine._query_dicts).append(self.global_dict)
        dev = torch.device("cuda", engine.device)
        cells = self.gmax * self.stride
        self.local = torch.zeros(cells, dtype=torch.int64, device=dev)
        self.gathered = torch.zeros(comm.world * cells, dtype=torch.int64, device=dev)
        self.merged = torch.zeros(max(self.nglobal, 1) * self.stride, dtype=torch.int64, device=dev)
        self.perm_dev = torch.tensor(flat, dtype=torch.int32, device=dev)
        # the collective runs stream-ordered between the library's launches: torch sees the library's compute stream
        stream = C.c_void_p()
        engine.ctx.call("msc_stream_handle", C.byref(stream))
        self.stream = torch.cuda.ExternalStream(stream.value, device=dev)
        torch.cuda.synchronize(dev)
        shuffle = engine._peer_shuffle()  # (collective on first use: every rank builds its _DenseMerge at the same point)
        self.shuffle = shuffle if shuffle and cells * 8 <= N.K["MSC_SHUFFLE_TABLE_BYTES"] else None
        self.exchange_kind = "nvlink peer stores (partial tables)" if self.shuffle else "nccl all-gather (partial tables)"

    def enqueue(self, desc: N.ScanDesc, exact: bool = False) -> int:
        """scan (enqueued) -> all-gather of the tables -> merge + compaction, all ordered on one stream and none waited
        for: returns the handle of a PENDING relation (msc_rel_settle)."""
        import torch  # noqa: PLC0415
        import torch.distributed as dist  # noqa: PLC0415

        e = self.engine
        want_jit = e.jit != "never" and (self.jit or e._specialise_now or e.jit == "always")
        flags = N.K["MSC_DENSE_ASYNC"] | (N.K["MSC_DENSE_EXACT"] if exact else 0) | (N.K["MSC_DENSE_JIT"] if want_jit else 0)
        if self.nlocal > 0:
            e.ctx.call("msc_scan_dense_table", C.byref(desc), self.nlocal, self.kinds, self.naggs, C.c_void_p(self.local.data_ptr()), flags)
        if self.shuffle:
            self.shuffle.allgather_table(self.local.data_ptr(), self.local.numel() * 8, self.gathered.data_ptr())
        else:
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.gathered, self.local)
        out = C.c_void_p()
        e.ctx.call("msc_dense_merge_compact_async", C.c_void_p(self.gathered.data_ptr()), e.comm.world, self.gmax, self.stride, self.kinds,
                   self.naggs, C.c_void_p(self.perm_dev.data_ptr()), self.nglobal, self.count_slot, C.c_void_p(self.merged.data_ptr()),
                   C.byref(out))
        return out.value

    def run(self, desc: N.ScanDesc) -> tuple[Optional[int], float]:
        """One pass with a single host wait; returns (handle of the merged result relation, device ms from the scan's
        first launch to the result).  A non-finite SUM in the merged table -- identical on every rank, so all ranks
        agree -- means the register-reduction kernel must not be trusted (gen_regvm.py): the pass is repeated with the
        exact kernel."""
        e = self.engine
        for exact in (False, True):
            handle = self.enqueue(desc, exact)
            rels = (C.c_void_p * 1)(handle)
            nonfinite = C.c_int32()
            e.ctx.call("msc_rel_settle", rels, 1, C.byref(nonfinite))
            if not nonfinite.value or exact:
                break
            e.ctx.lib.msc_rel_free(C.c_void_p(handle))
        return handle, e.ctx.stats().last_kernel_ms


class _PreparedPass:
    """A library-side prepared dense aggregate (msc_prepared_*): one kernel launch and one host wait per pass.  The
    result relation belongs to the library object and is overwritten by the next pass."""

    def __init__(self, engine: "CudaExecutionEngine", handle: int, out_types: list[str], out_dicts: list) -> None:
        self.engine, self.handle = engine, handle
        self.out_types, self.out_dicts = out_types, out_dicts
        self._cols_of: dict[int, list[DeviceColumn]] = {}
        self.in_flight = 0
        engine._closers.append(self.close)

    def close(self) -> None:
        e = self.engine
        if self.handle and getattr(e, "ctx", None) is not None:
            e.ctx.lib.msc_prepared_free(C.c_void_p(self.handle))
        self.handle = 0

    def _rel(self, res: C.c_void_p, nrows: int) -> DeviceRel:
        cols = self._cols_of.get(res.value)
        if cols is None:  # (the library rotates a few result relations: remember each one's column pointers)
            n = len(self.out_types)
            binds = (N.ColBind * n)()
            self.engine.ctx.check(self.engine.ctx.lib.msc_rel_cols(res, binds, n))
            cols = self._cols_of[res.value] = [DeviceColumn(binds[i].data or 0, binds[i].phys, self.out_types[i], self.out_dicts[i]) for i in range(n)]
        return DeviceRel(self.engine.ctx, res.value, nrows, cols, owned=False)

    def run(self, epoch: int, bump: Any = None) -> DeviceRel:
        """One pass; repeats it with the exact kernel when the masked one met a non-finite value.  `bump` supplies the
        next epoch for such a repeat (every rank repeats: the merged sums are identical everywhere)."""
        e = self.engine
        res, nrows, nonfinite = C.c_void_p(), C.c_uint64(), C.c_int32()
        for exact in (False, True):
            flags = N.K["MSC_DENSE_EXACT"] if exact else 0
            e.ctx.check(e.ctx.lib.msc_prepared_run(C.c_void_p(self.handle), flags, epoch, C.byref(res), C.byref(nrows), C.byref(nonfinite)))
            if not nonfinite.value or exact:
                break
            if bump is not None:
                epoch = bump()
        return self._rel(res, nrows.value)

    def enqueue(self, epoch: int) -> None:
        """Launch a pass without waiting for it (at most MSC_PREPARED_RING in flight); collect it with wait()."""
        self.engine.ctx.check(self.engine.ctx.lib.msc_prepared_enqueue(C.c_void_p(self.handle), 0, epoch))
        self.in_flight += 1

    def wait(self) -> tuple[DeviceRel, bool]:
        """The oldest pass in flight: (its result, whether a SUM came out non-finite -- then the result must not be used)."""
        e = self.engine
        res, nrows, nonfinite = C.c_void_p(), C.c_uint64(), C.c_int32()
        self.in_flight -= 1
        e.ctx.check(e.ctx.lib.msc_prepared_wait(C.c_void_p(self.handle), C.byref(res), C.byref(nrows), C.byref(nonfinite)))
        return self._rel(res, nrows.value), bool(nonfinite.value)


class _PeerExchange:
    """Cross-rank merge of a prepared dense aggregate INSIDE the scan kernel, over NVLink peer memory.

    Every rank owns a mailbox (CUDA IPC device memory its peers map); the scan's last CTA stores the rank's partial table
    into all mailboxes, publishes the pass number, waits for its peers' and folds the tables in rank order before it
    evaluates the final projection (csrc/jit.cu emit_finish, include/minispark_cuda.h msc_dense_fused_peer).  One kernel
    per rank and pass where the NCCL path (_DenseMerge.enqueue) needs scan + all-gather + merge + compaction +
    projection.  Set up collectively; any rank that cannot (no IPC, kernel not generated, more than 32 merged groups,
    a rank without rows) sends everyone back to the NCCL path."""

    def __init__(self, prepared: "PreparedAggregate") -> None:
        import torch  # noqa: PLC0415

        self.ok = False
        e, m = prepared.engine, prepared.merge
        comm = e.comm
        self.engine = e
        world, rank = comm.world, comm.rank
        desc2, prog2, staged_cols = prepared._final[0], prepared._final[1], prepared._final[2]
        self.desc2, self.prog2 = desc2, prog2
        self.raw_cols = N.int32_array([0 if ci == 0 else 1 + prepared.prog.slot_of[ci - 1] for ci in staged_cols])
        self.out_phys, self.nout = N.int32_array(prog2.out_phys), len(prog2.out_phys)
        self.out_types = [x.type for x in prepared.plan.outputs]
        cells = m.gmax * m.stride
        self.own = C.c_void_p()
        handle = C.create_string_buffer(64)
        able = e.jit != "never" and m.nlocal > 0 and 0 < m.nglobal <= 32 and world <= N.K["MSC_PEER_MAX_WORLD"] and m.gmax <= 32
        if able:
            try:
                e.ctx.call("msc_peer_alloc", 8 * (2 * world + 2 * world * cells), C.byref(self.own), handle)
            except N.NativeError:
                able = False
        everyone = comm.all_gather_object((able, bytes(handle.raw)))
        self.opened: list[int] = []
        spec = N.PeerSpec()
        if all(a for a, _ in everyone):
            try:
                for r, (_, h) in enumerate(everyone):
                    if r == rank:
                        spec.mailbox[r] = self.own.value
                    else:
                        ptr = C.c_void_p()
                        e.ctx.call("msc_peer_open", C.create_string_buffer(h, 64), C.byref(ptr))
                        self.opened.append(ptr.value)
                        spec.mailbox[r] = ptr.value
            except N.NativeError:
                able = False
        else:
            able = False
        inv = invert_code_maps([[m.perm[r * m.gmax + g] for g in range(m.gmax)] for r in range(world)]) if able else [-1] * (world * 32)
        self.inv_dev = torch.tensor(inv, dtype=torch.int32, device=torch.device("cuda", e.device))
        spec.inv = self.inv_dev.data_ptr()
        spec.rank, spec.world, spec.nlocal, spec.gmax, spec.nglobal = rank, world, m.nlocal, m.gmax, m.nglobal
        self.spec = spec
        self.epoch = 0
        self.prepared = prepared
        self.pass_ = None
        if able:  # compile now, so that no rank waits in the kernel for a peer that is still in NVRTC (or gave up)
            handle = C.c_void_p()
            try:
                e.ctx.call("msc_prepared_create", C.byref(prepared.desc), m.nlocal, prepared.kinds, len(prepared.prog.agg_kinds), C.byref(desc2),
                           self.raw_cols, self.out_phys, self.nout, C.byref(spec), C.byref(handle))
                able = bool(handle.value)
                if handle.value:
                    self.pass_ = _PreparedPass(e, handle.value, self.out_types, prog2.out_dicts)
            except N.NativeError:
                able = False
        self.ok = all(comm.all_gather_object(able))
        e._closers.append(self.close)
        if not self.ok:
            if self.pass_ is not None:
                self.pass_.close()
            self.close()

    def close(self) -> None:
        e = self.engine
        if getattr(e, "ctx", None) is None:
            return
        for ptr in self.opened:
            e.ctx.lib.msc_peer_close(e.ctx.handle, C.c_void_p(ptr))
        self.opened = []
        if self.own.value:
            e.ctx.lib.msc_peer_free(e.ctx.handle, self.own)
            self.own = C.c_void_p()

    def _next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def run(self) -> DeviceRel:
        """One pass on this rank (every rank must call it); returns the final relation (owned by the library object)."""
        return self.pass_.run(self._next_epoch(), self._next_epoch)


class PreparedAggregate:
    """A compiled ``scan -> filter -> GROUP BY -> final projection`` query bound to device-resident columns."""

    def __init__(self, engine: CudaExecutionEngine, plan: L.LSelect) -> None:
        self.engine = engine
        self.plan = plan
        agg = plan.child
        child = agg.child
        if isinstance(child, L.LSelect):
            filters = list(child.filters)
            group = L.substitute(agg.group, child.outputs)
            aggs = [(k, L.substitute(e, child.outputs)) for k, e in agg.aggs]
            base = child.child
        else:
            filters, group, aggs, base = [], agg.group, list(agg.aggs), child
        tracked = len(engine._query_rels)
        self.source, exprs = engine._prepare(base, [*filters, group, *[e for _, e in aggs]], fuse_probe=False)
        # what the pass is bound to: device-resident table columns outlive a query; anything the preparation had to compute
        # (a join below the aggregate, a concatenation) is released with the query, so such a plan must not be kept
        self.tables = [(str(base.path), engine._tables[str(base.path)])] if isinstance(base, L.LTable) else []
        self.reusable = isinstance(base, L.LTable) and len(engine._query_rels) == tracked
        nf = len(filters)
        self.group_type = exprs[nf].type
        self.resolver = _ScanResolver(engine, self.source)
        self.prog = L.compile_aggregate(self.resolver, exprs[:nf], exprs[nf], [(k, e) for (k, _), e in zip(aggs, exprs[nf + 1:])])
        self.desc = self.resolver.desc(self.prog.program)
        self.kinds = N.int32_array(self.prog.agg_kinds)
        self.slot_types = [L.FLOAT if k in (N.K["MSC_AGG_SUM_F"], N.K["MSC_AGG_MIN_F"], N.K["MSC_AGG_MAX_F"]) else L.INT
                           for k in self.prog.agg_kinds]
        self.ngroups, self.hint = engine._dense_groups(self.prog)
        self.merge = None
        self._table = None
        if self.ngroups and engine.comm.world == 1:  # the dense accumulator table of every pass (identities + scan + compaction)
            stride, count_slot = C.c_int32(), C.c_int32()
            engine.ctx.call("msc_dense_layout", C.byref(self.desc), self.kinds, len(self.prog.agg_kinds), C.byref(stride), C.byref(count_slot))
            self._stride, self._count_slot = stride.value, count_slot.value
            table = C.c_void_p()
            engine.ctx.call("msc_dev_alloc", self.ngroups * self._stride * 8, C.byref(table))
            self._table = table.value
        if self.ngroups and engine.comm.world > 1:
            self.merge = _DenseMerge(engine, self.prog.group_dict, self.desc, self.kinds, len(self.prog.agg_kinds), persistent=True)
        self.bytes_per_row = sum(N.PHYS_WIDTH[c.phys] for c in self.resolver.staged)
        self.nrows = self.source.nrows
        self.scan_stats: dict[str, Any] = {}
        self._final: Optional[tuple] = None  # compiled final projection, re-bound to every pass's aggregate result
        self._chain_args: Optional[tuple] = None
        self._fusable = True  # until msc_dense_fused says otherwise
        self._prep: Any = None  # _PreparedPass (msc_prepared_*) once created, False when this query cannot use it
        self._peer: Any = None  # _PeerExchange once set up, False when unavailable

    def run(self) -> tuple[DeviceRel, float]:
        """One pass of the hot path; returns (result relation, device milliseconds from the first launch to the result).

        Dense mode chains scan-aggregate (-> cross-rank merge) -> compaction -> final projection on the stream as PENDING
        relations and waits for the device once (msc_rel_settle); hash mode runs call by call."""
        e = self.engine
        naggs = len(self.prog.agg_kinds)
        if not self.ngroups:
            return self._run_stepwise()
        key_dict = self.merge.global_dict if self.merge is not None else self.prog.group_dict
        upper = self.merge.nglobal if self.merge is not None else self.ngroups
        if self.merge is None and self._final is not None:
            return self._run_chain()
        if self.merge is not None and self._final is not None and self.merge.persistent:
            if self._peer is None:  # collective: every rank gets here on its second pass
                self._peer = _PeerExchange(self) if os.environ.get("MINISPARK_PEER_MERGE", "1") != "0" else False
            if self._peer and self._peer.ok:
                final = self._peer.run()
                st = e.ctx.stats()
                self.scan_stats = {"scan_ms": st.last_scan_ms, "grid": st.last_scan_grid, "stages": st.last_scan_stages,
                                   "smem": st.last_scan_smem, "rows_per_thread": st.last_scan_rows_per_thread, "kind": st.last_scan_kind,
                                   "regs": st.last_scan_regs, "exchange": "nvlink peer stores inside the scan kernel"}
                return final, st.last_kernel_ms
        for exact in (False, True):
            if self.merge is not None:
                raw_h = self.merge.enqueue(self.desc, exact)
            else:
                # a prepared query runs many times: worth a kernel compiled for exactly this program (csrc/jit.cu)
                flags = N.K["MSC_DENSE_ASYNC"] | (N.K["MSC_DENSE_JIT"] if e.jit != "never" else 0) | (N.K["MSC_DENSE_EXACT"] if exact else 0)
                e.ctx.call("msc_scan_dense_table", C.byref(self.desc), self.ngroups, self.kinds, naggs, C.c_void_p(self._table), flags)
                out = C.c_void_p()
                e.ctx.call("msc_dense_compact_async", C.c_void_p(self._table), self.ngroups, self._stride, self.kinds, naggs, self._count_slot,
                           C.byref(out))
                raw_h = out.value
            binds = (N.ColBind * (1 + naggs))()
            e.ctx.check(e.ctx.lib.msc_rel_cols(C.c_void_p(raw_h), binds, 1 + naggs))
            cols = [DeviceColumn(binds[0].data, binds[0].phys, self.group_type, key_dict)]
            cols += [DeviceColumn(binds[1 + s].data, binds[1 + s].phys, self.slot_types[s]) for s in self.prog.slot_of]
            desc2, prog2, staged_cols = self._final_projection(cols, upper)
            nrows_dev = C.c_void_p()
            e.ctx.check(e.ctx.lib.msc_rel_nrows_dev(C.c_void_p(raw_h), C.byref(nrows_dev)))  # (no ctx argument)
            desc2.nrows = upper
            desc2.nrows_dev = nrows_dev.value
            for slot, ci in enumerate(staged_cols):
                desc2.staged[slot].data = cols[ci].ptr
            out2 = C.c_void_p()
            e.ctx.call("msc_scan_project", C.byref(desc2), N.int32_array(prog2.out_phys), len(prog2.out_phys), C.byref(out2))
            rels = (C.c_void_p * 2)(raw_h, out2.value)
            nonfinite = C.c_int32()
            e.ctx.call("msc_rel_settle", rels, 2, C.byref(nonfinite))
            if not nonfinite.value or exact:
                break
            e.ctx.lib.msc_rel_free(C.c_void_p(raw_h))  # non-finite SUM out of a masked variant: redo the pass the exact way
            e.ctx.lib.msc_rel_free(C.c_void_p(out2.value))
        st = e.ctx.stats()
        self.scan_stats = {"scan_ms": st.last_scan_ms, "grid": st.last_scan_grid, "stages": st.last_scan_stages,
                           "smem": st.last_scan_smem, "rows_per_thread": st.last_scan_rows_per_thread, "kind": st.last_scan_kind,
                           "regs": st.last_scan_regs}
        e._track(DeviceRel(e.ctx, raw_h, 0, []))
        final = e._track(DeviceRel.from_handle(e.ctx, out2.value, [x.type for x in self.plan.outputs], prog2.out_dicts))
        return final, st.last_kernel_ms

    def _async_pass(self) -> Optional[_PreparedPass]:
        """The library-side prepared pass that can run ahead of the host: after the passes that set it up (two on one GPU,
        three with several ranks: NCCL merge, then the collective set-up of the in-kernel exchange)."""
        if self.merge is None:
            return self._prep if self._prep else None
        return self._peer.pass_ if self._peer and self._peer.ok else None

    def enqueue(self) -> bool:
        """Launch one pass without waiting for its result; False when this query has no asynchronous form (yet): use run()."""
        p = self._async_pass()
        if p is None or p.in_flight >= N.K["MSC_PREPARED_RING"]:
            return False
        p.enqueue(self._peer._next_epoch() if self.merge is not None else 0)
        return True

    def wait(self) -> DeviceRel:
        """Result of the oldest pass in flight.  A non-finite SUM out of the masked kernel (identical on every rank) drains
        the passes still in flight and repeats the pass the exact way."""
        p = self._async_pass()
        final, nonfinite = p.wait()
        if nonfinite:
            while p.in_flight:
                p.wait()
            final, _ = self.run()
        return final

    def _run_chain(self) -> tuple[DeviceRel, float]:
        """Second and later passes on one GPU: the whole chain in one library call (msc_dense_chain)."""
        e = self.engine
        naggs = len(self.prog.agg_kinds)
        desc2, prog2, staged_cols = self._final[0], self._final[1], self._final[2]
        if self._chain_args is None:
            raw_cols = [0 if ci == 0 else 1 + self.prog.slot_of[ci - 1] for ci in staged_cols]
            self._chain_args = (N.int32_array(raw_cols), N.int32_array(prog2.out_phys), len(prog2.out_phys),
                                [x.type for x in self.plan.outputs])
        raw_cols, out_phys, nout, out_types = self._chain_args
        if self._prep is None:  # everything that does not change between passes, once (msc_prepared_create)
            self._prep = False
            if e.jit != "never":
                handle = C.c_void_p()
                e.ctx.call("msc_prepared_create", C.byref(self.desc), self.ngroups, self.kinds, naggs, C.byref(desc2), raw_cols, out_phys, nout,
                           None, C.byref(handle))
                if handle.value:
                    self._prep = _PreparedPass(e, handle.value, out_types, prog2.out_dicts)
        if self._prep:
            final = self._prep.run(0)
            st = e.ctx.stats()
            self.scan_stats = {"scan_ms": st.last_scan_ms, "grid": st.last_scan_grid, "stages": st.last_scan_stages,
                               "smem": st.last_scan_smem, "rows_per_thread": st.last_scan_rows_per_thread, "kind": st.last_scan_kind,
                               "regs": st.last_scan_regs}
            return final, st.last_kernel_ms
        raw, fin, nonfinite = C.c_void_p(), C.c_void_p(), C.c_int32()
        for exact in (False, True):
            flags = (N.K["MSC_DENSE_JIT"] if e.jit != "never" else 0) | (N.K["MSC_DENSE_EXACT"] if exact else 0)
            if self._fusable:  # scan + compaction + final projection in ONE kernel (the scan's last CTA finishes the query)
                e.ctx.call("msc_dense_fused", C.byref(self.desc), self.ngroups, self.kinds, naggs, C.c_void_p(self._table), flags,
                           C.byref(desc2), raw_cols, out_phys, nout, C.byref(fin), C.byref(nonfinite))
                if fin.value:
                    if not nonfinite.value or exact:
                        break
                    e.ctx.lib.msc_rel_free(fin)
                    continue
                self._fusable = False
            e.ctx.call("msc_dense_chain", C.byref(self.desc), self.ngroups, self.kinds, naggs, C.c_void_p(self._table), self._stride,
                       self._count_slot, flags, C.byref(desc2), raw_cols, out_phys, nout, C.byref(raw), C.byref(fin), C.byref(nonfinite))
            if not nonfinite.value or exact:
                break
            e.ctx.lib.msc_rel_free(raw)  # non-finite SUM out of a masked variant: redo the pass the exact way
            e.ctx.lib.msc_rel_free(fin)
        st = e.ctx.stats()
        self.scan_stats = {"scan_ms": st.last_scan_ms, "grid": st.last_scan_grid, "stages": st.last_scan_stages,
                           "smem": st.last_scan_smem, "rows_per_thread": st.last_scan_rows_per_thread, "kind": st.last_scan_kind,
                           "regs": st.last_scan_regs}
        if raw.value:
            e._track(DeviceRel(e.ctx, raw.value, 0, []))
        final = e._track(DeviceRel.from_handle(e.ctx, fin.value, out_types, prog2.out_dicts))
        return final, st.last_kernel_ms

    def _final_projection(self, cols: list[DeviceColumn], nrows: int) -> tuple:
        """The final projection (AVG = SUM / COUNT, HAVING, output order), compiled once; later passes only re-bind it."""
        e = self.engine
        sig = [(c.phys, id(c.dict)) for c in cols]
        if self._final is None or self._final[3] != sig:
            src2 = _Source(nrows, dict(enumerate(cols)))
            res2 = _ScanResolver(e, src2)
            prog2 = L.compile_project(res2, self.plan.filters, self.plan.outputs)
            if res2.gather:
                raise L.LoweringError("prepared final projection must not gather")
            index_of = {c.ptr: i for i, c in enumerate(cols)}
            self._final = (res2.desc(prog2.program), prog2, [index_of[c.ptr] for c in res2.staged], sig)
        return self._final[0], self._final[1], self._final[2]

    def _run_stepwise(self) -> tuple[DeviceRel, float]:
        e = self.engine
        naggs = len(self.prog.agg_kinds)
        out = C.c_void_p()
        e.ctx.call("msc_scan_aggregate", C.byref(self.desc), self.ngroups, self.kinds, naggs, self.hint, C.byref(out))
        st = e.ctx.stats()
        dev_ms = st.last_kernel_ms
        self.scan_stats = {"scan_ms": st.last_scan_ms, "grid": st.last_scan_grid, "stages": st.last_scan_stages,
                           "smem": st.last_scan_smem, "rows_per_thread": st.last_scan_rows_per_thread, "kind": st.last_scan_kind,
                           "regs": st.last_scan_regs}
        raw = e._track(DeviceRel.from_handle(e.ctx, out.value, [self.group_type, *self.slot_types], [self.prog.group_dict] + [None] * len(self.slot_types)))
        if e.comm.world > 1:  # hash mode: partition + all-to-all (or all-gather when small)
            raw = e._merge_partials(raw, self.prog.agg_kinds, self.slot_types, self.group_type)
            dev_ms += e.ctx.stats().last_kernel_ms
        key = raw.cols[0]
        if self.group_type == L.FLOAT and key.phys == N.P_I64:
            key = DeviceColumn(key.ptr, N.P_F64, L.FLOAT)
        cols = [key] + [raw.cols[1 + s] for s in self.prog.slot_of]
        desc2, prog2, staged_cols = self._final_projection(cols, raw.nrows)
        desc2.nrows = raw.nrows
        desc2.nrows_dev = None
        for slot, ci in enumerate(staged_cols):
            desc2.staged[slot].data = cols[ci].ptr
        out2 = C.c_void_p()
        e.ctx.call("msc_scan_project", C.byref(desc2), N.int32_array(prog2.out_phys), len(prog2.out_phys), C.byref(out2))
        dev_ms += e.ctx.stats().last_kernel_ms
        final = e._track(DeviceRel.from_handle(e.ctx, out2.value, [x.type for x in self.plan.outputs], prog2.out_dicts))
        return final, dev_ms
