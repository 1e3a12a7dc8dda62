"""ctypes binding of ``libminispark_cuda.so`` (the C-ABI declared in ``include/minispark_cuda.h``).

This is the only module that touches the shared library.  There is deliberately no fallback: if
the library is missing or no B200 is visible, :func:`load` / :class:`Context` raise.
"""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path
from typing import Sequence

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "minispark_cuda.h"
LIB_PATH = PKG / "lib" / "libminispark_cuda.so"


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"[msc {code}] {message}")
        self.code = code


def _parse_header() -> dict[str, int]:
    text = HEADER.read_text()
    consts: dict[str, int] = {}
    for name, value in re.findall(r"#define\s+(MSC_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text):
        consts[name] = int(value)
    for name, value in re.findall(r"(MSC_OP_[A-Z0-9_]+)\s*=\s*(\d+)", text):
        consts[name] = int(value)
    return consts


def _parse_regvm() -> dict[str, int]:
    """Handler ids of the register-resident interpreter (generated header next to the CUDA sources)."""
    text = (PKG / "csrc" / "regvm_handlers.h").read_text()
    return {name: int(value) for name, value in re.findall(r"#define\s+MSC_RV_([A-Z0-9_]+)\s+(\d+)", text)}


K = _parse_header()
RV = _parse_regvm()
OP = {name[len("MSC_OP_"):]: value for name, value in K.items() if name.startswith("MSC_OP_")}
P_U8, P_U16, P_U32, P_I32, P_I64, P_F32, P_F64 = (K[f"MSC_P_{n}"] for n in ("U8", "U16", "U32", "I32", "I64", "F32", "F64"))
PHYS_WIDTH = {P_U8: 1, P_U16: 2, P_U32: 4, P_I32: 4, P_I64: 8, P_F32: 4, P_F64: 8}
PHYS_NAME = {P_U8: "U8", P_U16: "U16", P_U32: "U32", P_I32: "I32", P_I64: "I64", P_F32: "F32", P_F64: "F64"}
PHYS_NUMPY = {P_U8: "<u1", P_U16: "<u2", P_U32: "<u4", P_I32: "<i4", P_I64: "<i8", P_F32: "<f4", P_F64: "<f8"}


class ColBind(C.Structure):
    _fields_ = [("data", C.c_void_p), ("phys", C.c_int32), ("_pad", C.c_int32)]


class ScanDesc(C.Structure):
    _fields_ = [
        ("nrows", C.c_uint64),
        ("nstaged", C.c_int32),
        ("ngather", C.c_int32),
        ("staged", ColBind * K["MSC_VM_MAX_STAGED"]),
        ("gather", ColBind * K["MSC_VM_MAX_GATHER"]),
        ("ncode", C.c_int32),
        ("nconsts", C.c_int32),
        ("code", C.c_uint32 * K["MSC_VM_MAX_CODE"]),
        ("consts", C.c_int64 * K["MSC_VM_MAX_CONSTS"]),
        ("nluts", C.c_int32),
        ("ntemps", C.c_int32),
        ("luts", C.c_void_p * K["MSC_VM_MAX_LUTS"]),
        ("ncode2", C.c_int32),
        ("count_slot2", C.c_int32),
        ("code2", C.c_uint32 * K["MSC_VM_MAX_CODE2"]),
        ("nrows_dev", C.c_void_p),
        ("want_jit", C.c_int32),
        ("table_columns", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("last_kernel_ms", C.c_double),
        ("last_ingest_ms", C.c_double),
        ("last_ingest_bytes", C.c_uint64),
        ("launches", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("last_scan_ms", C.c_double),
        ("last_scan_grid", C.c_int32),
        ("last_scan_stages", C.c_int32),
        ("last_scan_smem", C.c_int32),
        ("last_scan_rows_per_thread", C.c_int32),
        ("last_scan_kind", C.c_int32),
        ("last_scan_regs", C.c_int32),
        ("jit_compiles", C.c_uint64),
        ("last_jit_compile_ms", C.c_double),
        ("last_agg_runs", C.c_int32),
        ("last_hash_local_slots", C.c_int32),
        ("last_hash_attempts", C.c_int32),
        ("last_run_index_hit", C.c_int32),
    ]


class PeerSpec(C.Structure):
    _fields_ = [
        ("mailbox", C.c_void_p * 8),
        ("inv", C.c_void_p),
        ("epoch", C.c_uint64),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("nlocal", C.c_int32),
        ("gmax", C.c_int32),
        ("nglobal", C.c_int32),
        ("compile_only", C.c_int32),
    ]


class ConcatPart(C.Structure):
    _fields_ = [("codes", ColBind), ("dict", C.c_void_p), ("literal", C.c_char_p), ("literal_len", C.c_uint64)]


class OutCol(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_int32), ("rel_col", C.c_int32), ("dict", C.c_void_p)]


_SIGNATURES = {
    "msc_abi_version": (C.c_int, []),
    "msc_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "msc_destroy": (None, [C.c_void_p]),
    "msc_last_error": (C.c_char_p, [C.c_void_p]),
    "msc_sync": (C.c_int, [C.c_void_p]),
    "msc_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "msc_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "msc_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msc_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "msc_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "msc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "msc_table_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "msc_table_open_mem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "msc_table_close": (None, [C.c_void_p]),
    "msc_table_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "msc_table_col_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_char_p, C.c_int32]),
    "msc_table_block_rows": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_uint32)]),
    "msc_table_load": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                 C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "msc_rel_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "msc_rel_col": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "msc_rel_free": (None, [C.c_void_p]),
    "msc_rel_cols": (C.c_int, [C.c_void_p, C.POINTER(ColBind), C.c_int32]),
    "msc_timer_start": (C.c_int, [C.c_void_p]),
    "msc_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "msc_dense_layout": (C.c_int, [C.c_void_p, C.POINTER(ScanDesc), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32)]),
    "msc_scan_dense_table": (C.c_int, [C.c_void_p, C.POINTER(ScanDesc), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_int32]),
    "msc_dense_merge_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "msc_stream_handle": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_jit_dense_source": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_char_p, C.c_size_t,
                             C.POINTER(C.c_size_t)]),
    "msc_jit_dense_fused_source": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "msc_jit_runs_source": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "msc_jit_project_source": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "msc_jit_compile": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]),
    "msc_dense_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                        C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                        C.POINTER(C.c_int32)]),
    "msc_dense_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "msc_peer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "msc_peer_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msc_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msc_dense_fused_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(PeerSpec), C.POINTER(C.c_void_p),
                             C.POINTER(C.c_int32)]),
    "msc_prepared_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.POINTER(C.c_int32),
                            C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_prepared_run": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "msc_prepared_enqueue": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64]),
    "msc_prepared_wait": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "msc_prepared_free": (None, [C.c_void_p]),
    "msc_rel_nrows_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_rel_settle": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_int32)]),
    "msc_dense_merge_compact_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                                C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_dense_compact_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                          C.POINTER(C.c_void_p)]),
    "msc_dense_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                  C.POINTER(C.c_int32), C.c_int32, C.c_void_p]),
    "msc_dense_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                    C.POINTER(C.c_void_p)]),
    "msc_rel_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    "msc_rel_wrap": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(ColBind), C.c_int32, C.POINTER(C.c_void_p)]),
    "msc_scan_aggregate": (C.c_int, [C.c_void_p, C.POINTER(ScanDesc), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_uint64,
                                     C.POINTER(C.c_void_p)]),
    "msc_scan_project": (C.c_int, [C.c_void_p, C.POINTER(ScanDesc), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    "msc_dict_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_dict_free": (None, [C.c_void_p]),
    "msc_dict_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "msc_dict_lookup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t, C.c_int32, C.POINTER(C.c_int64)]),
    "msc_dict_like": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "msc_dict_translate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "msc_dict_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "msc_dict_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "msc_str_concat": (C.c_int, [C.c_void_p, C.POINTER(ConcatPart), C.c_int32, C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "msc_hash_join": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "msc_join_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "msc_scan_join_build": (C.c_int, [C.c_void_p, C.POINTER(ScanDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "msc_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]),
    "msc_shuffle_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]),
    "msc_shuffle_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msc_shuffle_slot_alloc": (C.c_int, [C.c_void_p, C.c_int32, C.c_size_t, C.c_void_p]),
    "msc_shuffle_slot_attach": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "msc_shuffle_slot_detach": (C.c_int, [C.c_void_p, C.c_int32]),
    "msc_shuffle_slot_bytes": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]),
    "msc_shuffle_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "msc_shuffle_begin_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.c_uint64, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64)]),
    "msc_partition_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_void_p)]),
    "msc_shuffle_finish": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "msc_shuffle_wait": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "msc_shuffle_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "msc_shuffle_free": (None, [C.c_void_p]),
    "msc_rel_copy_column": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "msc_rel_read_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int32, C.POINTER(C.c_int64)]),
    "msc_rel_fold_row": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "msc_write_blockfile": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(OutCol), C.c_int32, C.c_char_p, C.c_uint32]),
}

_lib: C.CDLL | None = None


def declared_symbols() -> list[str]:
    """Every function the header declares (used by the CPU test that the library exports them all)."""
    return re.findall(r"^MSC_API\s+[\w\s\*]+?\b(msc_[a-z0-9_]+)\(", HEADER.read_text(), flags=re.M)


def _preload_cudart() -> None:
    """The library links the CUDA runtime dynamically (build.py): make sure libcudart.so.12 is in the process before it is
    dlopen'ed, wherever this installation keeps it (the dynamic linker's own search path comes first)."""
    candidates = ["libcudart.so.12"]
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.cuda_runtime")
        for root in (spec.submodule_search_locations or []) if spec else []:
            candidates.append(str(Path(root) / "lib" / "libcudart.so.12"))
    except Exception:  # noqa: BLE001
        pass
    candidates += ["/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libcudart.so.12"]
    for cand in candidates:
        try:
            C.CDLL(cand, mode=C.RTLD_GLOBAL)
            return
        except OSError:
            continue


def load() -> C.CDLL:
    """dlopen the library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeError(-1, f"{LIB_PATH} is missing: run `python -m minispark_b200.build` (there is no CPU fallback)")
    _preload_cudart()
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.msc_abi_version() != K["MSC_ABI_VERSION"]:
        raise NativeError(-1, "libminispark_cuda.so ABI version does not match include/minispark_cuda.h")
    _lib = lib
    return lib


class Context:
    """One ``msc_ctx``: a device, its streams and its memory.  All calls raise :class:`NativeError`."""

    def __init__(self, device: int = 0) -> None:
        self.lib = load()
        handle = C.c_void_p()
        rc = self.lib.msc_create(device, C.byref(handle))
        if rc != 0 or not handle:
            raise NativeError(rc, f"msc_create(device={device}) failed: no sm_100 GPU visible? (there is no CPU fallback)")
        self.handle = handle
        self.device = device

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.msc_destroy(self.handle)
            self.handle = None

    def check(self, rc: int) -> None:
        if rc != 0:
            raise NativeError(rc, (self.lib.msc_last_error(self.handle) or b"").decode("utf-8", "replace"))

    def call(self, name: str, *args):  # noqa: ANN002, ANN201
        """``name(ctx, *args)`` for the entry points that take the context first; raises on a non-zero return."""
        fn = getattr(self.lib, name)
        if len(fn.argtypes) != len(args) + 1:  # ctypes itself lets extra arguments through
            raise TypeError(f"{name} takes {len(fn.argtypes)} arguments, {len(args) + 1} given (does it take the context?)")
        self.check(fn(self.handle, *args))

    def stats(self) -> Stats:
        st = Stats()
        self.call("msc_get_stats", C.byref(st))
        return st

    # ---- small conveniences used by the engine -------------------------------------------------
    def d2h(self, dev_ptr: int, nbytes: int) -> bytes:
        buf = C.create_string_buffer(max(nbytes, 1))
        self.call("msc_memcpy_d2h", buf, C.c_void_p(dev_ptr), nbytes)
        return buf.raw[:nbytes]

    def dev_free(self, dev_ptr: int | None) -> None:
        if dev_ptr:
            self.call("msc_dev_free", C.c_void_p(dev_ptr))


def int32_array(values: Sequence[int]):  # noqa: ANN201
    return (C.c_int32 * max(len(values), 1))(*values)
