"""ctypes binding of libdpfnets_b200.so (the C ABI declared in include/dpfnets_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails the error is raised.
"""
import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPF_LIB_PATH") or os.path.join(_PKG, "_C", "libdpfnets_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "dpfnets_b200.h")

_lib = None


class DpfNativeError(RuntimeError):
    pass


def declared_symbols(header=HEADER_PATH):
    """Names of all functions declared in the public header."""
    with open(header) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dpf_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DpfNativeError(
                "libdpfnets_b200.so not found at %s - run `python -m dpf_nets_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dpf_last_error.restype = ctypes.c_char_p
        for name in declared_symbols():
            fn = getattr(_lib, name)  # AttributeError if the symbol is not exported
            if name != "dpf_last_error":
                fn.restype = ctypes.c_int
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dpf_last_error().decode("utf-8", "replace")
        raise DpfNativeError("%s failed (rc=%d): %s" % (what or "native call", rc, msg))


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DpfNativeError("expected a CUDA tensor (no CPU fallback in dpf_nets_b200)")
        if not t.is_contiguous():
            raise DpfNativeError("expected a contiguous tensor")


def call(name, *args, device=None):
    """Invoke `name` with the current torch stream appended as the trailing void* argument."""
    fn = getattr(lib(), name)
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor) or a is None:
            conv.append(ptr(a))
        elif isinstance(a, bool):
            conv.append(ctypes.c_int(int(a)))
        elif isinstance(a, int):
            conv.append(ctypes.c_int(a))
        elif isinstance(a, float):
            conv.append(ctypes.c_float(a))
        else:
            conv.append(a)
    conv.append(stream_ptr(device))
    check(fn(*conv), name)
