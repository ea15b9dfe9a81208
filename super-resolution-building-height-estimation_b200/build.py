"""Build lib/libbhsr.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libbhsr.so")
SOURCES = ["common.cu", "conv_tc.cu", "plumbing.cu", "rrdbnet.cu", "head.cu", "head_tc.cu", "post.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "bhsr.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(cmd):
    subprocess.check_call(cmd)


def build(force: bool = False, verbose: bool = False, timing: bool = False, epi_swz: bool = False) -> str:
    """timing=True builds lib/libbhsr_timing.so with the in-kernel cycle counters (-DBHSR_TIMING);
    select it at run time with BHSR_LIB=<path> BHSR_DEBUG_TIMING=1 (profiling only).  epi_swz adds
    -DBHSR_EPI_SWZ (experimental epilogue staging of round 1, measured no gain in round 2) and writes
    lib/libbhsr_timing_episwz.so / lib/libbhsr_episwz.so — never the product library.
    Translation units are compiled in parallel into build/<variant>/*.o (re-used while newer than every
    source / header), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    variant = ("_timing" if timing else "") + ("_episwz" if epi_swz else "")
    out = LIB.replace("libbhsr.so", "libbhsr" + variant + ".so")
    if not (timing or epi_swz) and not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(HERE, "build", "obj" + variant)
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    defs = (["-DBHSR_TIMING"] if timing else []) + (["-DBHSR_EPI_SWZ"] if epi_swz else [])
    flags = [f for f in FLAGS if f != "-shared"] + defs + (["-Xptxas", "-v"] if verbose else [])
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "bhsr.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(path), newest_header):
            cmd = [nvcc] + flags + ["-c", "-o", obj, path]
            print("[bhsr build]", " ".join(cmd), flush=True)
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(_compile, jobs))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs
    print("[bhsr build]", " ".join(link), flush=True)
    subprocess.check_call(link)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, timing="--timing" in sys.argv,
          epi_swz="--epi-swz" in sys.argv)
