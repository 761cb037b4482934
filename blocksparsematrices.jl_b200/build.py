"""Builds libbsm_b200.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

    python blocksparsematrices.jl_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libbsm_b200.so"
SOURCES = ["abi.cu", "dist.cu", "sparse.cu", "krylov.cu", "construct.cu", "pack.cpp"]
DEPS = ["kernels.cuh", "persist.cuh", "spmm.cuh", "spmm_tma.cuh", "spmm_layout.h", "plan.h", "../../include/bsm_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-pthread",
    "-shared",
    "-ldl",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libbsm_b200.so cannot be built")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / f).stat().st_mtime > t for f in SOURCES + DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compiles every source to an object file (in parallel, only the stale ones) and links the shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = HERE / "_obj"
    objdir.mkdir(exist_ok=True)
    newest_dep = max((CSRC / f).stat().st_mtime for f in DEPS)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "-ldl")]

    def compile_one(src):
        obj = objdir / (Path(src).stem + ".o")
        stamp = max((CSRC / src).stat().st_mtime, newest_dep)
        if not force and obj.exists() and obj.stat().st_mtime > stamp:
            return obj, ""
        cmd = [nvcc_path(), *flags, "-c"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += ["-o", str(obj), str(CSRC / src)]
        res = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB),
           *[str(o) for o, _ in results], "-ldl", "-lpthread"]
    res = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log for _, log in results))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
