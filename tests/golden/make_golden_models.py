"""Generates tests/golden/model_keys.json and latent_modules.pt from the UNMODIFIED reference
(/root/reference): state_dict key -> shape listings of the three model families, and small
input/output vectors of the torch-level modules (latent flows, feature heads, PointNet encoder,
ResNet-18 tail, custom Adam)."""
import io
import json
import os
import sys

import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE, Local_Cond_RNVP_MC_Global_RNVP_VAE_IC  # noqa: E402
from lib.networks.decoders import GlobalRNVPDecoder  # noqa: E402
from lib.networks.encoders import FeatureEncoder, PointNetCloudEncoder  # noqa: E402
from lib.networks.optimizers import Adam, LRUpdater  # noqa: E402
from lib.networks.resnet import resnet18  # noqa: E402


def cfg(path):
    with io.open(os.path.join(REF, path), "r") as f:
        return yaml.safe_load(f)


def main():
    keys = {}
    for name, path, cls in [("generation_airplane", "configs/generation/airplane.yaml", Local_Cond_RNVP_MC_Global_RNVP_VAE),
                            ("autoencoding_all_scaled", "configs/autoencoding/all_scaled.yaml", Local_Cond_RNVP_MC_Global_RNVP_VAE),
                            ("svr_all", "configs/svr/all.yaml", Local_Cond_RNVP_MC_Global_RNVP_VAE_IC)]:
        torch.manual_seed(0)
        m = cls(**cfg(path))
        keys[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
        keys[name + "__n_params"] = sum(p.numel() for p in m.parameters())
    with open(os.path.join(HERE, "model_keys.json"), "w") as f:
        json.dump(keys, f)

    fx = {}
    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(1)
    gp = GlobalRNVPDecoder(3, 32, 16, weight_std=0.3)
    gp.train()
    with torch.no_grad():
        for _ in range(2):
            gp(torch.randn((6, 16), generator=gen), mode="inverse")
    gp.eval()
    g = torch.randn((5, 16), generator=gen)
    with torch.no_grad():
        fx["g_prior"] = {"state": {k: v.clone() for k, v in gp.state_dict().items()}, "g": g,
                         "direct": [torch.stack(x) for x in gp(g, mode="direct")],
                         "inverse": [torch.stack(x) for x in gp(g, mode="inverse")]}
    pe = PointNetCloudEncoder(3, 64, [128, 256, 512])
    x = torch.rand((2, 3, 50), generator=gen) - 0.5
    pe.train()
    with torch.no_grad():
        fx["pc_encoder"] = {"state": {k: v.clone() for k, v in pe.state_dict().items()}, "x": x, "train_out": pe(x).clone()}
    fe = FeatureEncoder(1, 32, 8, deterministic=False)
    fe.eval()
    f_in = torch.randn((4, 32), generator=gen)
    with torch.no_grad():
        fx["feature_encoder"] = {"state": {k: v.clone() for k, v in fe.state_dict().items()}, "x": f_in,
                                 "out": [t.clone() for t in fe(f_in)]}
    rn = resnet18(num_classes=16)
    rn.eval()
    img = torch.randn((2, 4, 64, 64), generator=gen)
    with torch.no_grad():
        fx["resnet18"] = {"keys": {k: list(v.shape) for k, v in rn.state_dict().items()}, "out_shape": list(rn(img).shape)}
    # custom Adam: 4 steps on a tiny quadratic with weight decay + amsgrad + LR updater
    w = torch.nn.Parameter(torch.linspace(-1, 1, 12).view(3, 4).clone())
    opt = Adam([w], lr=1e-2, weight_decay=1e-3, betas=(0.9, 0.99), amsgrad=True)
    upd = LRUpdater(10, cycle_length=2, min_lr=1e-3, max_lr=1e-2, beta1=0.9, min_beta2=0.99, max_beta2=0.999)
    traj = []
    for it in range(4):
        upd(opt, 0, it)
        opt.zero_grad()
        ((w ** 2).sum() + w.sum() * (it + 1)).backward()
        opt.step()
        traj.append(w.detach().clone())
    fx["adam"] = {"traj": torch.stack(traj), "lr": opt.param_groups[0]["lr"], "betas": list(opt.param_groups[0]["betas"])}
    torch.save(fx, os.path.join(HERE, "latent_modules.pt"))
    print("ok", {k: len(v) if isinstance(v, dict) else v for k, v in keys.items()})


if __name__ == "__main__":
    main()
