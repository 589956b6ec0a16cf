"""Build libragraph_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m ragraph_b200.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/ragraph_b200.h); it links the CUDA runtime
statically and resolves the one driver entry point it needs (cuTensorMapEncodeTiled) at run
time, so it has no link-time dependency on libcuda / torch.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libragraph_b200.so")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-DRAG_BUILDING=1"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "ragraph_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True compiles the pipeline-trace instrumentation into the tensor-core kernels (diagnostics only)."""
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr_m = _deps_mtime()
    jobs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src[:-3] + ".o")
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *ARCH, *CFLAGS, *(["-DRAG_TC_TRACE_BUILD=1"] if trace else []), "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJDIR, s[:-3] + ".o") for s in sources()]
    if jobs or force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs, "-Xlinker", "--no-undefined", "-Xlinker", "--exclude-libs,ALL", "-Xlinker", "-Bsymbolic",
               "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--trace" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv))
