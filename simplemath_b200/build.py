"""Build libsmb200.so (the sm_100a engine + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU-only
build container; the resulting .so travels to the GPU box with the repo.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # IEEE semantics the parity contract needs: no FMA contraction of a*b+c,
    # no flush-to-zero, correctly rounded div/sqrt (all nvcc defaults except fmad).
    "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "--shared", "-Xcompiler", "-fPIC",
    # The CUDA runtime is linked statically and stays PRIVATE to this library: its symbols are not exported
    # (--exclude-libs) and the library's own calls bind to its own definitions (-Bsymbolic).  Otherwise a program
    # that links this library AND carries its own static runtime (any nvcc-built user plugin, tests/plugin) has
    # half of its <<<>>> launch sequence resolved into our copy and half into its own.
    "-Xlinker", "--exclude-libs,ALL", "-Xlinker", "-Bsymbolic",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsmb200.so cannot be built")


def sources() -> list[str]:
    return [os.path.join(CSRC, "smb_api.cu")]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "smb200.h"), os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc without OpenMP specs; nvcc only
    # needs a plain host compiler
    subprocess.run(cmd, check=True, env=env)
    return LIB


def build_sweep(force: bool = False) -> str:
    """tools/sweep: the kernel-variant microbenchmark (same kernels header)."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tools", "sweep.cu")
    out = os.path.join(root, "tools", "sweep")
    if not os.path.exists(src):
        return ""
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if force or _stale(out, deps):
        flags = [f for f in NVCC_FLAGS if f not in ("--shared",)]
        # drop the "-Xcompiler -fPIC" and "-Xlinker ..." pairs too (shared-library options)
        flags = [f for i, f in enumerate(flags) if not (f in ("-Xcompiler", "-Xlinker") or (i > 0 and flags[i - 1] in ("-Xcompiler", "-Xlinker")))]
        subprocess.run([_nvcc(), *flags, "-o", out, src], check=True)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
