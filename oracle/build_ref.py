"""TEST INFRASTRUCTURE ONLY. Compiles the reference's own CUDA kernels, from the sources where
they lie under /root/reference (nothing is copied into the repo), into
oracle/_ref/libref_structural.so.  The launchers keep their C++-mangled names
(nndistance.cuh:1-2, approxmatch.cuh:6-8); oracle/structural.py binds them through ctypes.

Only runs where /root/reference exists (the dev container); the GPU box uses the prebuilt .so
that travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/lib/metrics/pytorch_structural_losses"
OUT = os.path.join(HERE, "_ref", "libref_structural.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build_ref(force=False):
    srcs = [os.path.join(REF_SRC, "src", f) for f in ("nndistance.cu", "approxmatch.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-I", os.path.join(REF_SRC, "src"), "-I", REF_SRC, "-shared", "-o", OUT] + srcs
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build_ref(force=True))
