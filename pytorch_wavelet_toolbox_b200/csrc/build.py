"""Build libwtb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pytorch_wavelet_toolbox_b200.csrc.build [--force] [--verbose]

The shared object is plain C ABI (include/wtb200.h): no torch / pybind dependency, the
CUDA runtime is linked statically, so the file built here runs unchanged on the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE / "libwtb200.so"
SOURCES = [HERE / "wtb200.cu"]
HEADERS = sorted(HERE.glob("*.cuh")) + [ROOT / "include" / "wtb200.h"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libwtb200.so")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [
        _nvcc(), "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo",
        "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
        "--expt-relaxed-constexpr",
        "-shared", "-cudart", "static",
        "-o", str(LIB),
    ] + [str(s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if extra:
        cmd += extra
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(out)
