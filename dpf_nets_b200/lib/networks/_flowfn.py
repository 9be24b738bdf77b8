"""autograd bridge between the arena-backed coupling stack and the C ABI
(dpf_decoder_forward / dpf_decoder_backward in include/dpfnets_b200.h)."""
import ctypes

import torch

from ... import _lib

MODES = {"direct": 0, "inverse": 1}
PRECISIONS = {"fp32": 0, "bf16": 1, "bf16x3": 2}


def _workspace(L, G, B, N, device):
    n = ctypes.c_longlong(0)
    _lib.check(_lib.lib().dpf_decoder_workspace_bytes(L, G, B, N, ctypes.byref(n)), "dpf_decoder_workspace_bytes")
    return torch.empty(int(n.value), dtype=torch.uint8, device=device)


def _meta_host_ptr(stack):
    return ctypes.c_void_p(stack.layout.meta.ctypes.data)


class CouplingStackFunction(torch.autograd.Function):
    """(p, g, arena) -> stacked (P, MU, LV) of shape (L, B, 3, N), indexed by layer like the reference's
    output lists, plus the two tensors the flow NLL consumes (losses.py:7-15) as SEPARATE outputs:
    Z = P[0] (samples[0]) and SLV = sum_l LV[l] (accumulated in the kernels' epilogues).  A loss that
    touches only Z and SLV - the reference's - back-propagates through two (B,3,N) cotangents: no dense
    (L,B,3,N) zero-filled gradients, no 63-way reduction.  Every stacked output stays differentiable, so
    any other use of the lists is still correct (and costs what it used to)."""

    @staticmethod
    def forward(ctx, p, g, arena, stack, mode, training, precision, want_mu):
        _lib.require_cuda(p, g, arena)
        L, G = stack.layout.L, stack.g_n_features
        B, C, N = p.shape
        if C != 3 or g.shape != (B, G) or p.dtype != torch.float32 or g.dtype != torch.float32:
            raise _lib.DpfNativeError("coupling stack expects p (B,3,N) and g (B,%d) float32; got %s, %s"
                                      % (G, tuple(p.shape), tuple(g.shape)))
        dev = p.device
        P = torch.empty((L, B, 3, N), dtype=torch.float32, device=dev)
        LV = torch.empty((L, B, 3, N), dtype=torch.float32, device=dev)
        MU = torch.empty((L, B, 3, N), dtype=torch.float32, device=dev) if want_mu else None
        SLV = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
        ws = _workspace(L, G, B, N, dev)
        update = bool(training)
        with torch.cuda.device(dev):
            _lib.call("dpf_decoder_forward_ex", _meta_host_ptr(stack), stack.layer_meta, arena, stack.stats, p, g,
                      P, MU, LV, SLV, ws, L, G, B, N, MODES[mode], bool(training), update,
                      PRECISIONS[precision], ctypes.c_float(stack.eps_value), device=dev)
        if training:
            from ...ops._counters import bump
            bump(stack.num_batches_tracked)
        ctx.stack, ctx.mode, ctx.training, ctx.ws, ctx.precision = stack, mode, bool(training), ws, precision
        stack._last_pass = (ws, L, G, B, N)
        ctx.save_for_backward(p, g, arena, P, LV)
        ctx.set_materialize_grads(False)
        Z = P[0].clone()
        if MU is None:
            MU = P.new_empty(0)
            ctx.mark_non_differentiable(MU)
        return P, MU, LV, Z, SLV

    @staticmethod
    def backward(ctx, dP, dMU, dLV, dZ, dSLV):
        from ._flowbwd import run_backward
        return run_backward(ctx, dP, dMU, dLV, dZ, dSLV)


def resolve_precision(stack):
    """'auto' = the tensor-core path: bf16x3 (split operands, fp32-class accuracy) when BatchNorm
    uses batch statistics, plain bf16 in eval mode (sampling)."""
    prec = stack.precision
    if prec == "auto":
        prec = "bf16x3" if stack.training else "bf16"
    if prec not in PRECISIONS:
        raise ValueError("precision must be one of %s or 'auto', got %r" % (sorted(PRECISIONS), prec))
    return prec


def run_stack(stack, p, g, mode, want_mu=True):
    """-> (P, MU, LV, Z, SLV); MU is an empty tensor when want_mu is False."""
    if mode not in MODES:
        raise ValueError("mode must be 'direct' or 'inverse', got %r" % (mode,))
    p = p.contiguous()
    g = g.contiguous()
    return CouplingStackFunction.apply(p, g, stack.arena, stack, mode, stack.training, resolve_precision(stack), want_mu)


def last_pass_status(stack):
    """Synchronous health check of the stack's most recent forward workspace (0 = fine): number of
    layers whose merged-forward grid barrier timed out."""
    ws, L, G, B, N = stack._last_pass
    flag = ctypes.c_int(-1)
    _lib.check(_lib.lib().dpf_decoder_status(_lib.ptr(ws), L, G, B, N, ctypes.byref(flag)), "dpf_decoder_status")
    return flag.value
