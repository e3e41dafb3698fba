"""Recipe: build the UNMODIFIED reference CUDA extensions as the GPU oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
reference / cpu_baseline legs may use it.

What it does
------------
Compiles, *from where they lie* under ``/root/reference`` (never copied):

    raymarching/src/{raymarching.cu,bindings.cpp}  -> oracle/_ref/_raymarching<ext>.so
    gridencoder/src/{gridencoder.cu,bindings.cpp}  -> oracle/_ref/_gridencoder<ext>.so
    shencoder/src/{shencoder.cu,bindings.cpp}      -> oracle/_ref/_shencoder<ext>.so

with ``nvcc -O3 -std=c++17 --expt-relaxed-constexpr -U__CUDA_NO_HALF_*`` for
``sm_100a``.  The reference's own flag ``-std=c++14``
(``raymarching/setup.py:9``) does not compile against torch 2.11 (which demands
C++17), so that one flag is overridden; the sources are untouched.  The module
names are the ones the reference's ``setup.py`` files give them
(``raymarching/setup.py:58``, ``gridencoder/setup.py:46``, ``shencoder/setup.py:46``).

``oracle/_ref/`` is git-ignored (binary artefacts stay out of history) but not
gpurun-ignored, so the built ``.so`` files travel to the GPU box, where
``/root/reference`` does not exist.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

REF = os.environ.get("PVD_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

MODULES = {
    "_raymarching": "raymarching/src",
    "_gridencoder": "gridencoder/src",
    "_shencoder": "shencoder/src",
}
CU_NAME = {"_raymarching": "raymarching.cu", "_gridencoder": "gridencoder.cu", "_shencoder": "shencoder.cu"}


def _torch_flags():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    lib = ce.library_paths()
    abi = int(getattr(__import__("torch")._C, "_GLIBCXX_USE_CXX11_ABI", True))
    return inc, lib, abi


def target_path(mod: str) -> str:
    return os.path.join(OUT, mod + sysconfig.get_config_var("EXT_SUFFIX"))


def build_one(mod: str, force: bool = False) -> str:
    src_dir = os.path.join(REF, MODULES[mod])
    cu = os.path.join(src_dir, CU_NAME[mod])
    bind = os.path.join(src_dir, "bindings.cpp")
    out = target_path(mod)
    if not os.path.exists(cu):
        raise FileNotFoundError(cu)
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) >= max(os.path.getmtime(cu), os.path.getmtime(bind))):
        return out
    os.makedirs(OUT, exist_ok=True)
    inc, lib, abi = _torch_flags()
    common = [f"-I{p}" for p in inc] + [
        f"-DTORCH_EXTENSION_NAME={mod}", "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
    ]
    obj_cu = os.path.join(OUT, mod + "_cu.o")
    obj_cpp = os.path.join(OUT, mod + "_bind.o")
    nvcc = [
        "nvcc", "-c", cu, "-o", obj_cu, "-O3", "-std=c++17", "--expt-relaxed-constexpr",
        "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
        "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w",
    ] + common
    gxx = ["g++", "-c", bind, "-o", obj_cpp, "-O3", "-std=c++17", "-fPIC", "-w"] + common
    subprocess.check_call(nvcc)
    subprocess.check_call(gxx)
    link = [
        "g++", "-shared", obj_cu, obj_cpp, "-o", out,
    ] + [f"-L{p}" for p in lib] + [
        "-L/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
        "-lc10_cuda", "-ltorch_cuda", "-lcudart",
    ] + [f"-Wl,-rpath,{p}" for p in lib]
    subprocess.check_call(link)
    for o in (obj_cu, obj_cpp):
        os.remove(o)
    return out


def build_all(force: bool = False, parallel: bool = True):
    if not os.path.isdir(REF):
        return {}
    mods = list(MODULES)
    if parallel:
        with ThreadPoolExecutor(max_workers=3) as ex:
            outs = list(ex.map(lambda m: build_one(m, force), mods))
    else:
        outs = [build_one(m, force) for m in mods]
    return dict(zip(mods, outs))


if __name__ == "__main__":
    res = build_all(force="--force" in sys.argv)
    for k, v in res.items():
        print(k, "->", v)
    if not res:
        print("reference tree not present; nothing built")
