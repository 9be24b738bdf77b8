"""Where the whole-model training step (generation/chair, 32 x 2048) spends its time: CUDA-event time and host
wall time per phase, and the kernel table of one step from torch.profiler."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dpf_nets_b200 import configs
from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
from dpf_nets_b200.lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE
from dpf_nets_b200.lib.networks.optimizers import Adam


def main():
    dev = torch.device("cuda", 0)
    B, N = 32, 2048
    config = configs.load("generation/chair")
    torch.manual_seed(0)
    model = Local_Cond_RNVP_MC_Global_RNVP_VAE(**config).to(dev).train()
    crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config).to(dev)
    opt = Adam(model.parameters(), lr=config["max_lr"], weight_decay=config["wd"], betas=(config["beta1"], config["max_beta2"]), amsgrad=True)
    gen = torch.Generator().manual_seed(99)
    g_clouds = (torch.rand((B, 3, N), generator=gen) - 0.5).to(dev)
    p_clouds = (torch.rand((B, 3, N), generator=gen) - 0.5).to(dev)
    names = ["forward", "loss", "zero_grad", "backward", "opt_step"]
    acc_gpu = {k: 0.0 for k in names}
    acc_cpu = {k: 0.0 for k in names}

    def timed(name, fn, rec):
        if not rec:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        r = fn()
        b.record()
        acc_cpu[name] += time.perf_counter() - t0
        torch.cuda.synchronize()
        acc_gpu[name] += a.elapsed_time(b)
        return r

    def step(rec):
        out = timed("forward", lambda: model(g_clouds, p_clouds), rec)
        loss = timed("loss", lambda: crit(g_clouds, p_clouds, out)[0], rec)
        timed("zero_grad", opt.zero_grad, rec)
        timed("backward", loss.backward, rec)
        timed("opt_step", opt.step, rec)

    for _ in range(3):
        step(False)
    K = 5
    for _ in range(K):
        step(True)
    for k in names:
        print("%-10s gpu %7.3f ms   host-issue %7.3f ms" % (k, acc_gpu[k] / K, 1e3 * acc_cpu[k] / K))
    # unsynchronised wall time per step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        step(False)
    torch.cuda.synchronize()
    print("free-running step: %.3f ms" % (1e3 * (time.perf_counter() - t0) / K))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step(False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
    # sub-module timing of the forward
    with torch.no_grad():
        pass
    enc = model.pc_encoder
    x = g_clouds.clone().requires_grad_(False)
    for _ in range(2):
        f = enc.global_features(x); f.sum().backward()
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); f = enc.global_features(x); b.record(); f.sum().backward(); c.record(); torch.cuda.synchronize()
    print("encoder train fwd %.3f ms, bwd %.3f ms" % (a.elapsed_time(b), b.elapsed_time(c)))


if __name__ == "__main__":
    main()
