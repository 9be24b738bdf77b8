"""Backward of the coupling stack through dpf_decoder_backward."""
import ctypes

import torch

from ... import _lib
from ._flowfn import MODES, PRECISIONS, _meta_host_ptr


def _cotangent(t, L, plane):
    """(tensor or None) -> (pointer-able tensor or None, layer stride in elements).  A cotangent
    that is the same (B,3,N) block for every layer (stride 0 on dim 0, e.g. from `.sum(0)`) is
    passed without materialising L copies."""
    if t is None:
        return None, 0
    if t.dim() == 4 and t.stride(0) == 0 and t[0].is_contiguous():
        return t[0], 0
    t = t.contiguous()
    return t, plane


def run_backward(ctx, dP, dMU, dLV, dZ=None, dSLV=None):
    stack, mode = ctx.stack, ctx.mode
    p, g, arena, P, LV = ctx.saved_tensors
    L, G = stack.layout.L, stack.g_n_features
    B, _, N = p.shape
    dev = p.device
    plane = B * 3 * N
    # the flow-NLL outputs: dZ belongs to layer 0 only, dSLV is shared by every layer (stride 0)
    if dZ is not None:
        if dP is None:
            dP_t, dP_s = dZ.contiguous(), -1
        else:
            dP = dP.clone() if dP.stride(0) != 0 else dP.expand(L, B, 3, N).clone()
            dP[0] += dZ
            dP_t, dP_s = dP, plane
    else:
        dP_t, dP_s = _cotangent(dP, L, plane)
    if dSLV is not None:
        if dLV is None:
            dLV_t, dLV_s = dSLV.contiguous(), 0
        else:
            dLV_t, dLV_s = (dLV + dSLV.unsqueeze(0)).contiguous(), plane
    else:
        dLV_t, dLV_s = _cotangent(dLV, L, plane)
    dMU_t, dMU_s = _cotangent(dMU, L, plane)
    darena = torch.empty_like(arena)
    dg = torch.empty_like(g)
    dp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
    scratch = None
    if PRECISIONS[ctx.precision] >= 1:
        nb = ctypes.c_longlong(0)
        _lib.check(_lib.lib().dpf_decoder_backward_scratch_bytes(L, B, N, ctypes.byref(nb)), "dpf_decoder_backward_scratch_bytes")
        scratch = torch.empty(int(nb.value), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_decoder_backward", _meta_host_ptr(stack), stack.layer_meta, arena, stack.stats, p, g,
                  P, LV, dP_t, ctypes.c_longlong(dP_s), dMU_t, ctypes.c_longlong(dMU_s),
                  dLV_t, ctypes.c_longlong(dLV_s), darena, ctypes.c_longlong(arena.numel()), dg, dp, ctx.ws, scratch,
                  L, G, B, N, MODES[mode], ctx.training, PRECISIONS[ctx.precision],
                  ctypes.c_float(stack.eps_value), device=dev)
    return dp, dg, darena, None, None, None, None, None
