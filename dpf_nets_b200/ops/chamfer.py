"""Fused all-pairs Chamfer matrix (replaces the loop in lib/networks/utils.py:90-117)."""
import torch

from .. import _lib


def pairwise_cd(clouds1, clouds2, out=None, row_start=0, row_step=1, n_rows=None, symmetric=False):
    """clouds1 (S1,n,3), clouds2 (S2,m,3) CUDA fp32 -> (S1,S2) matrix of
    dl.mean(1)+dr.mean(1).  Rows row_start::row_step (n_rows of them) are computed; with
    `symmetric` (clouds1 is clouds2) only the upper triangle is evaluated and then mirrored
    when the whole matrix was requested."""
    _lib.require_cuda(clouds1, clouds2)
    if clouds1.dtype != torch.float32 or clouds2.dtype != torch.float32:
        raise _lib.DpfNativeError("pairwise_cd expects float32 clouds")
    S1, n = clouds1.shape[0], clouds1.shape[1]
    S2, m = clouds2.shape[0], clouds2.shape[1]
    whole = out is None and row_start == 0 and row_step == 1 and n_rows is None
    if out is None:
        out = torch.zeros((S1, S2), dtype=torch.float32, device=clouds1.device)
    if n_rows is None:
        n_rows = max(0, (S1 - row_start + row_step - 1) // row_step)
    dev = clouds1.device
    with torch.cuda.device(dev):
        _lib.call("dpf_pairwise_cd", S1, S2, n, m, clouds1, clouds2, out, row_start, row_step, n_rows,
                  bool(symmetric), device=dev)
        if symmetric and whole:
            _lib.call("dpf_symmetrize_upper", out, S1, device=dev)
    return out


def pairwise_emd(clouds1, clouds2, out=None, row_start=0, row_step=1, n_rows=None):
    """(S1,n,3),(S2,m,3) -> (S1,S2) un-normalised approximate EMD costs (MatchCost per pair) with the
    dense match never materialised; divide by n for the reference's emd_approx."""
    _lib.require_cuda(clouds1, clouds2)
    S1, n = clouds1.shape[0], clouds1.shape[1]
    S2, m = clouds2.shape[0], clouds2.shape[1]
    if out is None:
        out = torch.zeros((S1, S2), dtype=torch.float32, device=clouds1.device)
    if n_rows is None:
        n_rows = max(0, (S1 - row_start + row_step - 1) // row_step)
    with torch.cuda.device(clouds1.device):
        _lib.call("dpf_pairwise_emd", S1, S2, n, m, clouds1, clouds2, out, row_start, row_step, n_rows,
                  device=clouds1.device)
    return out
