"""Build recipes: the C-ABI shared library (CUDA kernels for sm_100a + host VLC/driver) and the `p64b` CLI.

Everything is built IN-TREE (p64_b200/libp64b200.so, p64_b200/p64b) with explicit nvcc command lines so the
artefacts travel with a snapshot of the repository; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libp64b200.so")
CLI = os.path.join(HERE, "p64b")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
LIB_SOURCES = ["device.cu", "bits.cpp", "encoder.cpp", "y4m.cpp", "decoder.cpp"]
HEADERS = ["kernels.cuh", "vlc_kernels.cuh", "ingest.cuh", "vlc_dev.h", "vlc_tables.h", os.path.join("..", "..", "include", "p64_b200.h")]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the p64_b200 CUDA library cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    """`out` / `defines` (-D...) build experiment variants next to the product library (selected with P64B_LIB)."""
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(out, deps):
        cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall", "-shared",
               "-Xptxas", "-v" if verbose else "-warn-spills", *[f"-D{d}" for d in defines], "-o", out, *srcs, "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building libp64b200.so")
    return out


BOUNDS_LIB = os.path.join(HERE, "libp64b200_bounds.so")


def build_bounds_lib(force: bool = False) -> str:
    """the -DP64B_BOUNDS_CHECK debug build (every shared-memory access of the ME / macroblock kernels checked), next to the
    product library so that it travels with a snapshot; used by tests/test_bounds_check.py through P64B_LIB"""
    return build_lib(force, out=BOUNDS_LIB, defines=["P64B_BOUNDS_CHECK"])


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "cli.cpp")
    build_lib(force)
    if not os.path.exists(src):
        return CLI
    if force or _stale(CLI, [src, LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI, src, "-L" + HERE, "-lp64b200",
               "-Wl,-rpath,$ORIGIN", "-lpthread"]
        subprocess.run(cmd, check=True)
    return CLI


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_cli(force="--force" in sys.argv)
    print(LIB)
