"""Fused eval-mode PointNet encoder + max-pool (dpf_pointnet_eval_forward) against the module's own
torch path = the reference's layer sequence (lib/networks/encoders.py:9-28, models.py:130-131)."""
import os
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_encoder(cuda, seed):
    from dpf_nets_b200.lib.networks.encoders import PointNetCloudEncoder
    torch.manual_seed(seed)
    enc = PointNetCloudEncoder(3, 64, [128, 256, 512]).to(cuda)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # non-trivial BatchNorm state: negative scales, shifted / scaled running statistics
        for m in enc.features:
            if isinstance(m, torch.nn.BatchNorm1d):
                n = m.num_features
                m.weight.copy_((torch.rand(n, generator=gen) * 1.5 + 0.25) * torch.where(torch.rand(n, generator=gen) < 0.2, -1.0, 1.0))
                m.bias.copy_(torch.randn(n, generator=gen) * 0.3)
                m.running_mean.copy_(torch.randn(n, generator=gen) * 0.2)
                m.running_var.copy_(torch.rand(n, generator=gen) * 0.5 + 0.05)
    enc.eval()
    return enc


@pytest.mark.parametrize("B,N", [(32, 2048), (3, 1000), (5, 77), (1, 128), (200, 300), (2, 2500)])
def test_fused_eval_encoder_matches_torch_path(native_lib, cuda, B, N):
    enc = make_encoder(cuda, 3)
    gen = torch.Generator().manual_seed(B * 1000 + N)
    x = (torch.rand((B, 3, N), generator=gen) - 0.5).to(cuda)
    with torch.no_grad():
        got = enc.global_features(x)
        enc.precision = "fp32"
        want = enc.global_features(x)          # torch path (library GEMMs, fp32)
        assert torch.equal(want, torch.max(enc(x), dim=2)[0])
    assert got.shape == (B, 512) and torch.isfinite(got).all()
    err = rel(got, want)
    print("pointnet eval", B, N, "rel err", err)
    assert err < 2e-2          # bf16 tolerance of BASELINE.json north_star; measured ~3e-3
    assert (got >= 0).all()


def test_fused_eval_encoder_is_used_and_fast(native_lib, cuda):
    import ctypes
    enc = make_encoder(cuda, 5)
    x = (torch.rand((32, 3, 2048)) - 0.5).to(cuda)
    n0, n1 = ctypes.c_longlong(0), ctypes.c_longlong(0)
    native_lib.dpf_launch_count(ctypes.byref(n0))
    with torch.no_grad():
        enc.global_features(x)
    native_lib.dpf_launch_count(ctypes.byref(n1))
    assert n1.value - n0.value == 2       # pack + fused kernel
    torch.cuda.synchronize()
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(20):
            enc.global_features(x)
        torch.cuda.synchronize()
        fused = (time.perf_counter() - t0) / 20
        enc.precision = "fp32"
        enc.global_features(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            enc.global_features(x)
        torch.cuda.synchronize()
        lib_path = (time.perf_counter() - t0) / 5
    print("pointnet eval 32x2048: fused %.3f ms, library path %.3f ms" % (fused * 1e3, lib_path * 1e3))
    assert fused < lib_path


def test_train_mode_and_autograd_keep_the_library_path(native_lib, cuda):
    enc = make_encoder(cuda, 7)
    x = (torch.rand((4, 3, 256)) - 0.5).to(cuda)
    out = enc.global_features(x)            # grad enabled -> torch path, differentiable
    assert out.requires_grad
    enc.train()
    with torch.no_grad():
        assert torch.equal(enc.global_features(x), torch.max(enc(x), dim=2)[0])
