"""dpf_nets_b200 - B200-native (sm_100a) hot path of DPF-Nets behind the reference's module surface.

`dpf_nets_b200.lib` mirrors the reference's `lib` package (lib.networks.*, lib.metrics.*); put this
directory on sys.path to use the reference's own import lines unchanged.  All compute goes through
libdpfnets_b200.so (include/dpfnets_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import DpfNativeError  # noqa: F401

__all__ = ["DpfNativeError"]
