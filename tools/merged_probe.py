import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpf_nets_b200 import _lib
from dpf_nets_b200.lib.networks.decoders import LocalCondRNVPDecoder
lib = _lib.lib()
dev = torch.device("cuda:0")
m = LocalCondRNVPDecoder(2, 64, 16).to(dev).train()
m.precision = "bf16x3"
p = (torch.rand((32, 3, 2048)) - 0.5).to(dev); g = torch.randn((32, 16)).to(dev)
lib.dpf_profile_enable(1)
with torch.no_grad():
    m(p, g, mode="inverse")
torch.cuda.synchronize()
ms = (ctypes.c_double * 8)(); cnt = (ctypes.c_longlong * 8)()
lib.dpf_profile_collect(ms, cnt, 8)
print("counts", list(cnt), "ms", [round(x, 3) for x in ms])
print("last_error:", lib.dpf_last_error())
