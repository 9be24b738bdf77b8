"""CPU: the mirrored torch-level modules reproduce the reference (goldens generated from
/root/reference by tests/golden/make_golden_models.py): state_dict key/shape layout of the three
model families (checkpoint contract), latent flows, feature heads, PointNet encoder, custom Adam."""
import io
import json
import os

import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def cfg(path):
    from dpf_nets_b200 import configs
    return configs.get(path[len("configs/"):-len(".yaml")])


@pytest.mark.parametrize("name,path,cls", [
    ("generation_airplane", "configs/generation/airplane.yaml", "Local_Cond_RNVP_MC_Global_RNVP_VAE"),
    ("autoencoding_all_scaled", "configs/autoencoding/all_scaled.yaml", "Local_Cond_RNVP_MC_Global_RNVP_VAE"),
    ("svr_all", "configs/svr/all.yaml", "Local_Cond_RNVP_MC_Global_RNVP_VAE_IC")])
def test_state_dict_layout_matches_reference(name, path, cls):
    from dpf_nets_b200.lib.networks import models
    with open(os.path.join(GOLD, "model_keys.json")) as f:
        want = json.load(f)
    m = getattr(models, cls)(**cfg(path))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert set(got) == set(want[name]), (sorted(set(got) ^ set(want[name]))[:10])
    assert all(got[k] == want[name][k] for k in got)
    assert sum(p.numel() for p in m.parameters()) == want[name + "__n_params"]
    # a checkpoint written with the reference's key layout loads
    sd = {k: torch.zeros(v) if "num_batches" not in k else torch.tensor(0) for k, v in want[name].items()}
    for k in sd:
        if k.endswith(".eps"):
            sd[k] = torch.tensor([1e-6])
    assert not any(m.load_state_dict(sd, strict=True))


def test_latent_modules_vs_reference():
    from dpf_nets_b200.lib.networks.decoders import GlobalRNVPDecoder
    from dpf_nets_b200.lib.networks.encoders import FeatureEncoder, PointNetCloudEncoder
    from dpf_nets_b200.lib.networks.resnet import resnet18
    fx = torch.load(os.path.join(GOLD, "latent_modules.pt"), weights_only=False)
    gp = GlobalRNVPDecoder(3, 32, 16)
    gp.load_state_dict(fx["g_prior"]["state"])
    gp.eval()
    with torch.no_grad():
        for mode in ("direct", "inverse"):
            out = gp(fx["g_prior"]["g"], mode=mode)
            for a, b in zip(out, fx["g_prior"][mode]):
                assert rel(torch.stack(a), b) < 1e-5
    pe = PointNetCloudEncoder(3, 64, [128, 256, 512])
    pe.load_state_dict(fx["pc_encoder"]["state"])
    pe.train()
    with torch.no_grad():
        assert rel(pe(fx["pc_encoder"]["x"]), fx["pc_encoder"]["train_out"]) < 1e-5
    fe = FeatureEncoder(1, 32, 8)
    fe.load_state_dict(fx["feature_encoder"]["state"])
    fe.eval()
    with torch.no_grad():
        for a, b in zip(fe(fx["feature_encoder"]["x"]), fx["feature_encoder"]["out"]):
            assert rel(a, b) < 1e-6
    rn = resnet18(num_classes=16)
    assert {k: list(v.shape) for k, v in rn.state_dict().items()} == fx["resnet18"]["keys"]
    rn.eval()
    with torch.no_grad():
        assert list(rn(torch.zeros(2, 4, 64, 64)).shape) == fx["resnet18"]["out_shape"]


def test_custom_adam_and_lr_updater_vs_reference():
    from dpf_nets_b200.lib.networks.optimizers import Adam, LRUpdater
    fx = torch.load(os.path.join(GOLD, "latent_modules.pt"), weights_only=False)["adam"]
    w = torch.nn.Parameter(torch.linspace(-1, 1, 12).view(3, 4).clone())
    opt = Adam([w], lr=1e-2, weight_decay=1e-3, betas=(0.9, 0.99), amsgrad=True)
    upd = LRUpdater(10, cycle_length=2, min_lr=1e-3, max_lr=1e-2, beta1=0.9, min_beta2=0.99, max_beta2=0.999)
    for it in range(4):
        upd(opt, 0, it)
        opt.zero_grad()
        ((w ** 2).sum() + w.sum() * (it + 1)).backward()
        opt.step()
        assert rel(w.detach(), fx["traj"][it]) < 1e-6
    assert opt.param_groups[0]["lr"] == pytest.approx(fx["lr"])
    assert list(opt.param_groups[0]["betas"]) == pytest.approx(fx["betas"])
    assert set(opt.state[w].keys()) == {"step", "exp_avg", "exp_avg_sq", "max_exp_avg_sq"}


def test_capturable_adam_hyper_table_matches_eager_formulas():
    """Host side of the graph-replayable Adam (optimizers.Adam.prepare_replay / _hyper_row): the row written to the pinned
    table for step t must hold exactly the scalars the eager path passes by value - lr, betas, eps, wd, 1 - beta1^t and
    sqrt(1 - beta2^t) (reference optimizers.py:53-66) - also after a learning-rate / beta2 change by LRUpdater."""
    import math
    from dpf_nets_b200.lib.networks.optimizers import Adam, LRUpdater
    p = torch.nn.Parameter(torch.randn(5))
    opt = Adam([p], lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-6, amsgrad=True)
    for _ in range(3):                       # eager CPU steps create the state and advance the counter
        p.grad = torch.randn(5)
        opt.step()
    assert opt.state[p]['step'] == 3
    sched = LRUpdater(10, cycle_length=2, min_lr=1e-4, max_lr=1e-3, beta1=0.9, min_beta2=0.9, max_beta2=0.99)
    sched(opt, 0, 7)
    group = opt.param_groups[0]
    opt._hyper_host = torch.zeros((1, 8))    # what enable_capture() allocates (pinned on a CUDA box)
    opt.prepare_replay()                     # the host work of one replayed step
    assert opt.state[p]['step'] == 4
    b1, b2 = group['betas']
    want = [group['lr'], b1, b2, 1e-8, 1e-6, 1 - b1 ** 4, math.sqrt(1 - b2 ** 4)]
    assert torch.allclose(opt._hyper_host[0, :7], torch.tensor(want, dtype=torch.float32), rtol=1e-6, atol=0)
