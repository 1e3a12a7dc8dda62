"""Build libminispark_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libminispark_cuda.so"
OBJ_DIR = PKG / "build"
SOURCES = ["scan_inst_r4_dense.cu", "scan_inst_r8_dense.cu", "scan_inst_r4_hash.cu", "scan_inst_r8_hash.cu", "scan_inst_r8_runs.cu", "scan_inst_r4_count.cu", "scan_inst_r4_build.cu",
           "scan_inst_r4_project.cu", "scan_regvm_ng0.cu", "scan_regvm_ng1.cu", "scan_regvm_ng2.cu", "scan_regvm_ng3.cu", "scan_regvm_ng4.cu", "core.cu", "scan.cu", "jit.cu", "strings.cu", "ingest.cu", "join.cu", "shuffle.cu", "result.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--diag-suppress", "177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    for dep in [src, *sorted(CSRC.glob("*.cuh")), *sorted(CSRC.glob("*.h")), *sorted(CSRC.glob("*.inc")), PKG.parent / "include" / "minispark_cuda.h"]:
        h.update(dep.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src_name: str, verbose: bool) -> Path:
    src = CSRC / src_name
    obj = OBJ_DIR / (src.stem + ".o")
    stamp = OBJ_DIR / (src.stem + ".sha")
    digest = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    stamp.write_text(digest)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every CUDA source and link the shared library; incremental by content hash."""
    OBJ_DIR.mkdir(exist_ok=True)
    LIB_DIR.mkdir(exist_ok=True)
    if force:
        for f in OBJ_DIR.glob("*.sha"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose), SOURCES))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        # (--cudart shared: the runtime stays in libcudart.so instead of being embedded, entry-point name table and all)
        cmd = [_nvcc(), "-shared", "--cudart", "shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        subprocess.run(["strip", "--strip-unneeded", str(LIB)], check=False)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
