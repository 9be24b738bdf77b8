"""Generates tests/golden/wholemodel.pt by executing the UNMODIFIED reference models from /root/reference
(dev container only): whole-model training-mode forward + VAE loss + backward of the three BASELINE model
families - generation (airplane/chair, G=128), autoencoding all_original (G=512, BASELINE config 3) and
SVR all (image-conditioned, G=512, BASELINE config 4) - with weights regenerated from key names
(tests/_detstate.py) and torch.randn_like replaced by a seeded stream, so the GPU test rebuilds the identical
model and noise without shipping checkpoints.  Stored: (loss, pnll, gnll, gent), z = p_prior_samples[0],
g_posterior_mus, and the autograd gradients of a few parameters of every sub-module.

    python tests/golden/make_golden_wholemodel.py
"""
import io
import os
import sys

import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

from _detstate import DetRandn, GRAD_KEYS, det_state, whole_model_inputs  # noqa: E402
from lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss  # noqa: E402
from lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE, Local_Cond_RNVP_MC_Global_RNVP_VAE_IC  # noqa: E402

# batch sizes: BatchNorm over the batch dimension (FiLM nets, latent flows, ResNet) is ill-conditioned in fp32 for B = 3..4
# (the reference's own fp32 gradients are then several per cent away from its float64 run); 8 / 8 / 6 shapes keep the
# reference's fp32-vs-fp64 distance small enough for tight gates
CASES = [("generation_chair", "configs/generation/chair.yaml", False, 8, 512),
         ("autoencoding_all_original", "configs/autoencoding/all_original.yaml", False, 8, 384),
         ("svr_all", "configs/svr/all.yaml", True, 6, 2500 // 10)]


def cfg(path):
    with io.open(os.path.join(REF, path), "r") as f:
        return yaml.safe_load(f)


def main():
    torch.set_num_threads(8)
    fx = {}
    for name, path, ic, B, N in CASES:
        config = cfg(path)
        config["util_mode"] = "training"
        torch.manual_seed(0)
        model = (Local_Cond_RNVP_MC_Global_RNVP_VAE_IC if ic else Local_Cond_RNVP_MC_Global_RNVP_VAE)(**config)
        shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
        state = det_state(shapes)
        inp = whole_model_inputs(B, N, 77, ic)
        res = {}
        for prec in ("f32", "f64"):      # f64 = the same reference modules in double: the truth both fp32 results are measured against
            model.load_state_dict(state)
            model = model.double() if prec == "f64" else model.float()
            model.train()
            model.zero_grad()
            crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config)
            x = {k: (v.double() if prec == "f64" else v) for k, v in inp.items()}
            with DetRandn(5):
                out = model(x["cloud"], x["eval_cloud"], x["image"]) if ic else model(x["cloud"], x["eval_cloud"])
            loss, pnll, gnll, gent = crit(x["cloud"], x["eval_cloud"], out)
            loss.backward()
            named = dict(model.named_parameters())
            res[prec] = {"losses": torch.stack([loss.detach(), pnll.detach(), gnll.detach(), gent.detach()]).double(),
                         "z": out["p_prior_samples"][0].detach().clone(),
                         "sum_logvar": sum(out["p_prior_logvars"][1:]).detach().clone(),
                         "g_posterior_mus": out["g_posterior_mus"].detach().clone(),
                         "grads": {k: named[k].grad.clone() for k in GRAD_KEYS[ic]}}
        fx[name] = dict(res["f32"], config_path=path, B=B, N=N, shapes=shapes,
                        truth64={"losses": res["f64"]["losses"], "z": res["f64"]["z"].float(), "sum_logvar": res["f64"]["sum_logvar"].float(),
                                 "grads": {k: v.float() for k, v in res["f64"]["grads"].items()}})
        print(name, [float(v) for v in fx[name]["losses"]], [float(v) for v in fx[name]["truth64"]["losses"]])
        for k in GRAD_KEYS[ic]:
            a, b = res["f32"]["grads"][k].double(), res["f64"]["grads"][k]
            print("   reference fp32 vs its float64 run: %-70s %.2e" % (k, float((a - b).abs().max() / b.abs().max())))
    # the key -> shape listing is regenerated on the test side from this package's own model (it is checked against
    # the reference's in tests/test_models_host.py), so it is not stored
    for v in fx.values():
        del v["shapes"]
    torch.save(fx, os.path.join(HERE, "wholemodel.pt"))
    print("wholemodel.pt", os.path.getsize(os.path.join(HERE, "wholemodel.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
