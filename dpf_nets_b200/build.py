"""In-tree nvcc build of libdpfnets_b200.so (sm_100a only, no torch headers).

`python -m dpf_nets_b200.build` or `build_library()`; objects are cached by source mtime.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only dev container.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# DPF_STAMPS=1: development build with in-kernel phase timers (tools/stamp_probe.py), kept apart from
# the product library (_C_stamps/, loaded only when DPF_LIB_PATH points at it)
STAMPS = os.environ.get("DPF_STAMPS", "0") == "1"
OUT_DIR = os.path.join(PKG_DIR, "_C_stamps" if STAMPS else "_C")
LIB_PATH = os.path.join(OUT_DIR, "libdpfnets_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"] + (["-DDPF_STAMPS"] if STAMPS else []) + (
    ["-DDPF_EXP_NOATOMICS"] if os.environ.get("DPF_EXP_NOATOMICS") == "1" else [])


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = 0.0
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "dpfnets_b200.h")
    if os.path.exists(inc):
        m = max(m, os.path.getmtime(inc))
    return m


def _compile_one(src, verbose):
    obj = os.path.join(OUT_DIR, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), _headers_mtime()):
        return obj, ""
    cmd = [NVCC] + ARCH + CFLAGS + ["-I", CSRC, "-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    log = r.stdout + r.stderr
    with open(obj[:-2] + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        sys.stderr.write(log)
    return obj, log


def build_library(verbose=False, force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OUT_DIR):
            if f.endswith(".o"):
                os.remove(os.path.join(OUT_DIR, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    objs = [o for o, _ in results]
    need_link = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)
    if need_link:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(verbose="-v" in sys.argv, force="-f" in sys.argv))
