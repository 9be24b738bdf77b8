"""Eager vs eager vs graphed training trajectories (noise floor of the comparison in tests/test_model_gpu.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpf_nets_b200 import configs
from dpf_nets_b200.lib.networks import models as models_mod
from dpf_nets_b200.lib.networks._graphstep import GraphedTrainStep
from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
from dpf_nets_b200.lib.networks.optimizers import Adam

models_mod._reparameterize = lambda mu, logvar: mu + 0.3 * torch.exp(0.5 * logvar)
cuda = torch.device("cuda", 0)
cfg = configs.get('generation/chair')
cfg.update(p_decoder_n_flows=2, g_latent_space_size=16, g_prior_n_flows=2, g_prior_n_features=16)
gen = torch.Generator().manual_seed(3)
batches = [((torch.rand((6, 3, 300), generator=gen) - 0.5).to(cuda), (torch.rand((6, 3, 300), generator=gen) - 0.5).to(cuda)) for _ in range(7)]
runs = {}
for kind in ("eager", "eager2", "graphed"):
    torch.manual_seed(0)
    model = models_mod.Local_Cond_RNVP_MC_Global_RNVP_VAE(**cfg).to(cuda).train()
    crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**cfg)
    opt = Adam(model.parameters(), lr=1e-3, weight_decay=1e-6, betas=(0.9, 0.99), amsgrad=True)
    init = {n: p.detach().clone() for n, p in model.named_parameters()}
    step = GraphedTrainStep(model, crit, opt, eager_steps=2 if kind == "graphed" else 10 ** 9)
    losses = [float(step(c, e)[0].detach()) for c, e in batches]
    runs[kind] = (losses, {n: p.detach().clone() for n, p in model.named_parameters()}, init,
                  {k: v.clone() for k, v in model.state_dict().items() if 'running' in k})
for other in ("eager2", "graphed"):
    print("==", other, "vs eager; losses", runs[other][0], runs["eager"][0])
    groups = {}
    for n in runs["eager"][1]:
        top = n.split(".")[0]
        de = (runs["eager"][1][n] - runs["eager"][2][n]).flatten()
        do = (runs[other][1][n] - runs[other][2][n]).flatten()
        groups.setdefault(top, []).append((de, do))
    for top, lst in groups.items():
        de = torch.cat([a for a, _ in lst]); do = torch.cat([b for _, b in lst])
        print("   %-14s cos %.4f  |de| %.4e |do| %.4e" % (top, float(torch.dot(de, do) / (de.norm() * do.norm())), float(de.norm()), float(do.norm())))
    worst = sorted((((runs["eager"][3][k] - runs[other][3][k]).abs().max().item() / (runs["eager"][3][k].abs().max().item() + 1e-12), k) for k in runs["eager"][3]), reverse=True)[:4]
    print("   worst running stats", worst)
