"""ctypes loader for libpvd_b200.so -- the ONLY way the Python host side reaches the kernels.

There is deliberately no fallback: if the shared library is missing or an entry point is absent the import / call
raises, so a GPU test can never silently pass on a PyTorch or CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "libpvd_b200_trace.so" if os.environ.get("PVD_TRACE", "0") == "1" else "libpvd_b200.so")
ABI_VERSION = 1

_lib = None

F16 = 1
F32 = 0


class NativeLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        l.pvd_error_string.restype = C.c_char_p
        l.pvd_error_string.argtypes = [C.c_int]
        l.pvd_abi_version.restype = C.c_int
        l.pvd_march_rays_train_workspace_words.restype = C.c_uint64
        l.pvd_march_rays_train_workspace_words.argtypes = [C.c_uint32, C.c_uint32]
        got = l.pvd_abi_version()
        if got != ABI_VERSION:
            raise NativeLibraryError(f"libpvd_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().pvd_error_string(int(rc))
        raise RuntimeError(f"pvd_b200 native call failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


@contextlib.contextmanager
def on_device(t: torch.Tensor):
    """Make t's device current for the launch (the reference launches on whatever device is current)."""
    if not t.is_cuda:
        raise RuntimeError("pvd_b200: expected a CUDA tensor; there is no CPU path")
    if t.device.index == torch.cuda.current_device():
        yield
    else:
        with torch.cuda.device(t.device):
            yield


u32 = C.c_uint32
f32 = C.c_float
i32 = C.c_int
