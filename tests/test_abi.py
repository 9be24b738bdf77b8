"""The C-ABI library builds, loads and exports every symbol include/dpfnets_b200.h declares
(no compute calls here - those are the -m gpu tests)."""
import ctypes
import os
import subprocess

from dpf_nets_b200 import _lib


def test_header_symbols_exported(native_lib):
    names = _lib.declared_symbols()
    assert "dpf_nndistance" in names and "dpf_pairwise_cd" in names
    for n in names:
        assert hasattr(native_lib, n), "missing export: " + n


def test_no_torch_or_python_dependency(native_lib):
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out


def test_argument_errors_are_reported(native_lib):
    native_lib.dpf_nndistance.restype = ctypes.c_int
    rc = native_lib.dpf_nndistance(-1, 0, None, 0, None, None, None, None, None, None)
    assert rc == -1
    assert b"negative" in native_lib.dpf_last_error()
    assert native_lib.dpf_version() >= 100


def test_built_for_sm100a_only(native_lib):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert all("sm_100a" in line for line in out.splitlines() if "ELF file" in line)
