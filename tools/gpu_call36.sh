#!/bin/bash
export PYTHONPATH=$PWD
S=$(mktemp -d)
echo "== no resume, bs 4"; python train_ae.py generation/chair smoke 1 0.000256 --synthetic 20 --batch_size 4 --path2save $S --cuda_graph 2>&1 | grep -v "^Epoch" | grep "Error\|saved" | head -3 | cut -c1-200
echo "== no resume, bs 32"; python train_ae.py generation/chair smoke 1 0.000256 --synthetic 160 --batch_size 32 --path2save $S --cuda_graph 2>&1 | grep -v "^Epoch" | grep "Error\|saved" | head -3 | cut -c1-200
echo "== no resume, bs 4, shuffle off via probe"; python - <<'PY' 2>&1 | grep -v "^Epoch" | grep "Error\|ok" | head -3 | cut -c1-200
import torch
from dpf_nets_b200 import configs
from dpf_nets_b200.lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE
from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
from dpf_nets_b200.lib.networks.optimizers import Adam
from dpf_nets_b200.lib.networks._graphstep import GraphedTrainStep
dev = torch.device("cuda", 0)
cfg = configs.load("generation/chair")
torch.manual_seed(0)
m = Local_Cond_RNVP_MC_Global_RNVP_VAE(**cfg).to(dev).train()
crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**cfg).to(dev)
opt = Adam(m.parameters(), lr=1e-4, weight_decay=1e-6, betas=(0.9, 0.99), amsgrad=True)
st = GraphedTrainStep(m, crit, opt, eager_steps=2)
for B in (4,):
    a = (torch.rand((B, 3, 2048)) - 0.5).to(dev); b = (torch.rand((B, 3, 2048)) - 0.5).to(dev)
    for i in range(5):
        l = st(a, b)[0]
    print("ok", float(l.detach()))
PY
