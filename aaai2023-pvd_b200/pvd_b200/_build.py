"""Build recipe for libpvd_b200.so (nvcc, sm_100a only, in-tree).

The library is plain CUDA behind a C ABI (include/pvd_b200.h): no torch headers, so every translation
unit compiles in seconds and the result has no dependency on the PyTorch ABI.  `build_native()` is what
`__graft_entry__.build()` calls; the shared object lands next to this file (git-ignored, shipped to the
GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(PKG_DIR, "..", "csrc"))
INCLUDE = os.path.normpath(os.path.join(PKG_DIR, "..", "..", "include"))
TRACE = os.environ.get("PVD_TRACE", "0") == "1"   # diagnostic timeline build (tools/trace_kernels.py), never the default
BUILD_DIR = os.path.join(CSRC, "build_trace" if TRACE else "build")
LIB_PATH = os.path.join(PKG_DIR, "libpvd_b200_trace.so" if TRACE else "libpvd_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC",
    "-DPVD_BUILDING",
] + (["-DPVD_TRACE"] if TRACE else [])


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return hs


def build_native(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in sources()]
    stamp = os.path.join(BUILD_DIR, "stamp.txt")
    digest = _digest(srcs + _headers())
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = ["nvcc", "-c", src, "-o", obj, f"-I{INCLUDE}"] + NVCC_FLAGS
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = ["nvcc", "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
