"""In-tree build of libvoroffset_b200.so (hand-written CUDA for sm_100a + the extern "C" layer).

`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`: cross-compiles without a GPU; the .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libvoroffset_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "--fmad=false", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, "vo_lib.cu")]


def _deps():
    out = [os.path.join(ROOT, "include", "voroffset_b200.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """out / defines: a variant build next to the library (A/B runs of compile-time knobs, scripts/ only)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + [
        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", out or LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libvoroffset_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return out or LIB


if __name__ == "__main__":
    # build.py [--force] [-v] [--out path -DNAME=value ...]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
