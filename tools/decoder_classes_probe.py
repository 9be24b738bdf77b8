"""Per-kernel-class CUDA-event times of one decoder train step for a given latent width G (FiLM nets scale with G)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200 import _lib  # noqa: E402
from dpf_nets_b200.lib.networks.decoders import LocalCondRNVPDecoder  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B, N = 32, 2048
dev = torch.device("cuda:0")
lib = _lib.lib()
torch.manual_seed(0)
m = LocalCondRNVPDecoder(21, 64, G).to(dev).train()
p = (torch.rand((B, 3, N)) - 0.5).to(dev)
g = torch.randn((B, G)).to(dev).requires_grad_(True)


def step():
    m.arena.grad = None
    g.grad = None
    z, slv = m.nll_terms(p, g)
    (0.5 * (slv.sum() + (z * z).sum()) / B).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    step()
b.record()
torch.cuda.synchronize()
print("G=%d decoder train step: %.3f ms" % (G, a.elapsed_time(b) / 10))
lib.dpf_profile_enable(1)
for _ in range(3):
    step()
torch.cuda.synchronize()
ms = (ctypes.c_double * 8)()
cnt = (ctypes.c_longlong * 8)()
lib.dpf_profile_collect(ms, cnt, 8)
lib.dpf_profile_enable(0)
for i, k in enumerate(["film_fwd", "moments", "fwd_stats", "fwd_apply", "bwd_p1", "bwd_p2", "bwd_final", "film_bwd"]):
    print("%-10s %8.3f ms/step  %s launches" % (k, ms[i] / 3, cnt[i] // 3))
