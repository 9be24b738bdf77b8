"""Kernel-time breakdown of the train-mode PointNet encoder forward + backward (32 x 2048), fused last layer vs library path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200.lib.networks.encoders import PointNetCloudEncoder  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = PointNetCloudEncoder(3, 64, [128, 256, 512]).to(dev).train()
x = (torch.rand((32, 3, 2048)) - 0.5).to(dev)
cot = torch.randn((32, 512), device=dev)


def step():
    enc.zero_grad()
    (enc.global_features(x) * cot).sum().backward()


for prec in ("fp32", "auto"):
    enc.precision = prec
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    print("==== encoder precision", prec)
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
    tot = sum(e.device_time_total for e in prof.key_averages())
    print("total device time per step: %.3f ms" % (tot / 5 / 1e3))
    for e in rows:
        print("%8.1f us/step  x%-3d %s" % (e.device_time_total / 5, e.count // 5, e.key[:110]))
