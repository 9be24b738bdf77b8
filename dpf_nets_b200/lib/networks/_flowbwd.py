"""Backward of the coupling stack (filled in with dpf_decoder_backward)."""


def run_backward(ctx, dP, dMU, dLV):
    raise NotImplementedError("coupling-stack backward is not built yet")
