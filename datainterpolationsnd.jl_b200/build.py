"""Builds libdind.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python datainterpolationsnd.jl_b200/build.py [--force] [--verbose]

One object per .cu (compiled in parallel), linked into
datainterpolationsnd.jl_b200/libdind.so.  The CUDA runtime is linked statically, so the
library has no dependency besides the driver.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libdind.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
         "-ccbin", "/usr/bin/g++"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "dind.h"))
    return max(os.path.getmtime(h) for h in hs)


def compile_one(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(spath)
            and os.path.getmtime(obj) > headers_mtime()):
        return obj, ""
    cmd = [NVCC] + ARCH + FLAGS + ["-c", spath, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = p.stdout + p.stderr
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, log


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = [o for o, _ in ex.map(lambda s: compile_one(s, force, verbose), srcs)]
    if (force or not os.path.exists(SO) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs)):
        cmd = [NVCC] + ARCH + ["-shared", "-o", SO] + objs + ["-ccbin", "/usr/bin/g++"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    return SO


if __name__ == "__main__":
    so = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print("built", so)
