"""bench.py contract on CPU: the reference arm runs here (oracle port on the host cores) and prints ONE JSON line
with the keys the driver reads; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, **kw):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT, **kw)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "2", "--points", "128")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("decoder points/s") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_non_zero_ranks_exit_quietly():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "2", "--points", "128",
             env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return                                               # covered by the GPU tests / the bench itself
    r = _run("--steps", "1", "--warmup", "0", "--no-extras")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_reference_arm_uses_no_product_code_and_all_host_threads():
    """The reference arm must not import the product package and, launched like a torchrun rank 0
    (OMP_NUM_THREADS=1 in the environment), must still use every host core."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '2', "
            "'--points', '64']; runpy.run_path(%r, run_name='__main__'); import torch; "
            "sys.stderr.write('PRODUCT_IMPORTED=%%s THREADS=%%d\\n' %% (any(m.startswith('dpf_nets_b200') for m in sys.modules), torch.get_num_threads()))"
            % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "PRODUCT_IMPORTED=False" in r.stderr
    assert "THREADS=%d" % (os.cpu_count() or 1) in r.stderr
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert "host CPU" in d["config"]["parallelism"] and d["config"]["global_batch"] == 2
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
