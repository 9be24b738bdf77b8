"""GPU-side generation scores: COV / MMD / 1-NNA from the device-resident all-pairs
matrices in one native call, and the voxel-occupancy histogram behind JSD (replaces the torch / numpy
post-processing of lib/networks/utils.py:45-87,120-144; nothing but the final scalars is copied to the host)."""
import ctypes

import numpy as np
import torch

from .. import _lib


def cd_scores(gg, gt, tt):
    """gg (S1,S1), gt (S1,S2), tt (S2,S2) CUDA fp32 -> device tensor [COV, MMD, 1-NN accuracy]."""
    _lib.require_cuda(gg, gt, tt)
    S1, S2 = gt.shape
    if gg.shape != (S1, S1) or tt.shape != (S2, S2) or any(t.dtype != torch.float32 for t in (gg, gt, tt)):
        raise _lib.DpfNativeError("cd_scores expects fp32 gg (S1,S1), gt (S1,S2), tt (S2,S2); got %s %s %s"
                                  % (tuple(gg.shape), tuple(gt.shape), tuple(tt.shape)))
    nb = ctypes.c_longlong(0)
    _lib.check(_lib.lib().dpf_cd_scores_scratch_bytes(S1, S2, ctypes.byref(nb)), "dpf_cd_scores_scratch_bytes")
    dev = gt.device
    scratch = torch.empty(int(nb.value), dtype=torch.uint8, device=dev)
    out = torch.empty(3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_cd_scores", gg, gt, tt, S1, S2, scratch, out, device=dev)
    return out


def voxel_hist(clouds, res=28):
    """clouds (..., 3) CUDA fp32 -> (res,res,res) int64 occupancy counts over [-0.5, 0.5)^3."""
    _lib.require_cuda(clouds)
    if clouds.dtype != torch.float32 or clouds.shape[-1] != 3:
        raise _lib.DpfNativeError("voxel_hist expects fp32 clouds (...,3)")
    dev = clouds.device
    edges = torch.from_numpy(-0.5 + np.arange(res + 1) * (1. / res)).to(dev)      # the reference's float64 edges
    hist = torch.empty((res, res, res), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_voxel_hist", clouds, ctypes.c_longlong(clouds.numel() // 3), res, edges, hist, device=dev)
    return hist


def jsd_from_hists(h1, h2):
    """Jensen-Shannon divergence (base 2) of two occupancy histograms, on their device, as a 0-dim tensor
    (utils.py:82-87: entropy((d1+d2)/2) - (entropy(d1)+entropy(d2))/2 with d = hist / hist.sum())."""
    d1 = h1.reshape(-1).double() / h1.sum().double()
    d2 = h2.reshape(-1).double() / h2.sum().double()

    def H(p):
        return -(torch.where(p > 0, p * torch.log2(p.clamp_min(1e-300)), torch.zeros_like(p))).sum()
    return H((d1 + d2) / 2.0) - 0.5 * (H(d1) + H(d2))
