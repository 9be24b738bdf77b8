import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpf_nets_b200 import _lib
lib = _lib.lib(); torch.cuda.init(); torch.zeros(1, device="cuda")
for which in (1, 0):
    res = []
    for smem in (16384, 32768, 49152, 65536, 81920, 91632, 98304, 110000):
        o = ctypes.c_int(-5); lib.dpf_debug_occupancy(which, smem, ctypes.byref(o)); res.append((smem, o.value))
    print("which", which, res)
p = torch.cuda.get_device_properties(0)
print(p.name, getattr(p, "shared_memory_per_multiprocessor", None), getattr(p, "shared_memory_per_block_optin", None), p.regs_per_multiprocessor if hasattr(p,"regs_per_multiprocessor") else None)
