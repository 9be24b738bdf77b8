"""Kernel-time breakdown of the whole-model training step (generation/chair, 32 x 2048, eager launches) by kernel name."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200 import configs  # noqa: E402
from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss  # noqa: E402
from dpf_nets_b200.lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE  # noqa: E402
from dpf_nets_b200.lib.networks.optimizers import Adam  # noqa: E402

dev = torch.device("cuda:0")
config = configs.load(sys.argv[1] if len(sys.argv) > 1 else "generation/chair")
torch.manual_seed(0)
model = Local_Cond_RNVP_MC_Global_RNVP_VAE(**config).to(dev).train()
crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config).to(dev)
opt = Adam(model.parameters(), lr=config["max_lr"], weight_decay=config["wd"], betas=(config["beta1"], config["max_beta2"]), amsgrad=True)
g = torch.Generator().manual_seed(1)
a = (torch.rand((32, 3, 2048), generator=g) - 0.5).to(dev)
b = (torch.rand((32, 3, 2048), generator=g) - 0.5).to(dev)


def step():
    out = model(a, b)
    loss = crit(a, b, out)[0]
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
dec = sum(e.device_time_total for e in ev if "coupling" in e.key or "film_" in e.key or "dw1_reduce" in e.key or "bwd_tables" in e.key or "pack_w1" in e.key or "moments_kernel" in e.key)
print("device time per step %.3f ms over %d launches; decoder kernels %.3f ms" % (tot / 3e3, sum(e.count for e in ev) // 3, dec / 3e3))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:28]:
    if "coupling" in e.key:
        continue
    print("%8.1f us/step  x%-4d %s" % (e.device_time_total / 3, e.count // 3, e.key[:120]))
