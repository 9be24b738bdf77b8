"""GPU: the tcgen05 operand layouts / descriptor encodings used by the tensor-core coupling kernels
reproduce a host matmul (K-major SW128 operands, TMA bulk staging, MN-major operands for wgrad).
Runs the probe in a subprocess so that a bad descriptor cannot poison this process's CUDA context."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg", ["kmajor", "kmajor_bulk", "mnmajor"])
def test_umma_layouts(native_lib, cuda, cfg):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "umma_probe.py"), cfg], capture_output=True,
                       text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines and all(x["ok"] and x["rc"] == 0 for x in lines), lines
