"""Tensor-level wrappers over the C ABI (device pointers in, device tensors out)."""
from .chamfer import pairwise_cd, pairwise_emd  # noqa: F401
