"""Build lib/libbhsr.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libbhsr.so")
SOURCES = ["common.cu", "conv_tc.cu", "plumbing.cu", "rrdbnet.cu", "head.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "bhsr.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, timing: bool = False, epi_swz: bool = False) -> str:
    """timing=True builds lib/libbhsr_timing.so with the in-kernel cycle counters (-DBHSR_TIMING);
    select it at run time with BHSR_LIB=<path> BHSR_DEBUG_TIMING=1 (profiling only).  epi_swz adds
    -DBHSR_EPI_SWZ (experimental conflict-free epilogue staging, DESIGN.md §8 Finding 3) and writes
    lib/libbhsr_timing_episwz.so / lib/libbhsr_episwz.so — never the product library."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if timing or epi_swz:
        name = "libbhsr" + ("_timing" if timing else "") + ("_episwz" if epi_swz else "") + ".so"
        out = LIB.replace("libbhsr.so", name)
        defs = (["-DBHSR_TIMING"] if timing else []) + (["-DBHSR_EPI_SWZ"] if epi_swz else [])
        cmd = [os.environ.get("NVCC", "nvcc")] + FLAGS + defs + ["-o", out] + srcs
        print("[bhsr build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return out
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    print("[bhsr build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, timing="--timing" in sys.argv,
          epi_swz="--epi-swz" in sys.argv)
