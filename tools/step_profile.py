"""Kernel-time breakdown of the whole-model training step (torch.profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dpf_nets_b200 import configs
from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
from dpf_nets_b200.lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE
from dpf_nets_b200.lib.networks.optimizers import Adam

dev = torch.device("cuda", 0)
config = configs.load(sys.argv[1] if len(sys.argv) > 1 else "generation/chair")
B, N = 32, 2048
torch.manual_seed(0)
model = Local_Cond_RNVP_MC_Global_RNVP_VAE(**config).to(dev)
model.train()
crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config).to(dev)
opt = Adam(model.parameters(), lr=config["max_lr"], weight_decay=config["wd"], betas=(config["beta1"], config["max_beta2"]), amsgrad=True)
gen = torch.Generator().manual_seed(99)
a = (torch.rand((B, 3, N), generator=gen) - 0.5).to(dev)
b = (torch.rand((B, 3, N), generator=gen) - 0.5).to(dev)


def step():
    out = model(a, b)
    loss, *_ = crit(a, b, out)
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
