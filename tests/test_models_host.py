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


def test_optimizer_state_round_trips_through_the_reference_layout():
    """Checkpoint boundary of the optimizer state (reference: training.py:75-80, optimizers.py:32-40): one entry per
    reference parameter on disk, one arena entry per coupling stack in memory; the conversion is lossless both ways
    and the on-disk layout has exactly the reference model's parameter count and tensor shapes."""
    from dpf_nets_b200.lib.networks import models
    from dpf_nets_b200.lib.networks.optimizers import Adam, optimizer_state_from_reference, optimizer_state_to_reference
    from dpf_nets_b200 import configs
    c = configs.get('generation/chair')
    c.update(p_decoder_n_flows=2, g_latent_space_size=16, g_prior_n_flows=2, g_prior_n_features=16)
    torch.manual_seed(0)
    m = models.Local_Cond_RNVP_MC_Global_RNVP_VAE(**c)
    opt = Adam(m.parameters(), lr=1e-3, weight_decay=1e-6, betas=(0.9, 0.99), amsgrad=True)
    g = torch.Generator().manual_seed(1)
    for p in m.parameters():                      # a fake step: state without running the (GPU-only) decoder
        opt.state[p] = {'step': 7, 'exp_avg': torch.randn(p.shape, generator=g), 'exp_avg_sq': torch.rand(p.shape, generator=g),
                        'max_exp_avg_sq': torch.rand(p.shape, generator=g)}
    ours = opt.state_dict()
    ref = optimizer_state_to_reference(m, ours)
    # the reference model has one nn.Parameter per state_dict entry that is not a buffer
    sd = m.state_dict()
    ref_params = [k for k in sd if not (k.endswith('running_mean') or k.endswith('running_var') or k.endswith('num_batches_tracked')
                                        or k.endswith('.eps') or k.endswith('p_prior_mus') or k.endswith('p_prior_logvar'))]
    assert len(ref['param_groups'][0]['params']) == len(ref_params) == len(ref['state'])
    assert len(ref['param_groups'][0]['params']) > len(ours['param_groups'][0]['params'])
    n_dec = len(m.pc_decoder.layout.param_index)
    views = list(m.pc_decoder.layout.param_index.values())
    names = [n for n, _ in m.named_parameters()]
    first = names.index('pc_decoder.arena')     # reference entries first .. first + n_dec - 1 belong to the decoder
    for j, (off, shape) in enumerate(views[:5] + views[-3:]):
        jj = j if j < 5 else n_dec - 3 + (j - 5)
        assert tuple(ref['state'][first + jj]['exp_avg'].shape) == tuple(shape)
    back = optimizer_state_from_reference(m, ref)
    assert back['param_groups'][0]['params'] == ours['param_groups'][0]['params']
    for i, st in ours['state'].items():
        assert back['state'][i]['step'] == st['step']
        for k in ('exp_avg', 'exp_avg_sq', 'max_exp_avg_sq'):
            assert torch.equal(back['state'][i][k].reshape(-1), st[k].reshape(-1)), (i, k)
    opt2 = Adam(m.parameters(), lr=1e-3, amsgrad=True)
    opt2.load_state_dict(back)                   # loads into a fresh optimizer
    assert optimizer_state_from_reference(m, ours) is ours       # already in this package's layout: untouched
    bad = {'state': {}, 'param_groups': [dict(ref['param_groups'][0], params=list(range(5)))]}
    with pytest.raises(ValueError):
        optimizer_state_from_reference(m, bad)


@pytest.mark.parametrize("name,cls,ic", [("generation_chair", "Local_Cond_RNVP_MC_Global_RNVP_VAE", False),
                                         ("autoencoding_all_original", "Local_Cond_RNVP_MC_Global_RNVP_VAE", False),
                                         ("svr_all", "Local_Cond_RNVP_MC_Global_RNVP_VAE_IC", True)])
def test_wholemodel_encoder_side_vs_reference(name, cls, ic):
    """CPU half of the whole-model parity (the decoder half needs the GPU: tests/test_model_gpu.py): with the
    key-derived deterministic weights, the train-mode PointNet encoder -> g_posterior head reproduces the
    reference's g_posterior_mus (golden: tests/golden/make_golden_wholemodel.py)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _detstate import det_state, whole_model_inputs
    from dpf_nets_b200 import configs
    from dpf_nets_b200.lib.networks import models
    fx = torch.load(os.path.join(GOLD, "wholemodel.pt"), weights_only=False)[name]
    c = configs.get(fx["config_path"][len("configs/"):-len(".yaml")])
    c["util_mode"] = "training"
    m = getattr(models, cls)(**c)
    m.load_state_dict(det_state({k: list(v.shape) for k, v in m.state_dict().items()}))
    m.train()
    inp = whole_model_inputs(fx["B"], fx["N"], 77, ic)
    with torch.no_grad():
        mus, logvars, _ = m._posterior(inp["cloud"], sample=False)
    assert rel(mus, fx["g_posterior_mus"]) < 1e-5


def test_pooled_last_layer_backward_algebra_vs_autograd(monkeypatch):
    """Train-mode PointNet last layer + max-pool (ops/pointnet_pool.py): the analytic backward - sparse max-pool gradient,
    BatchNorm batch terms through the Gram matrix of h2 - against torch.autograd of the plain module chain
    SharedDot -> BatchNorm1d(train) -> ReLU -> max (reference encoders.py:9-28, models.py:130-131) in float64, with the
    kernel's statistics pass replaced by its torch definition (the kernel itself is checked on the GPU)."""
    from dpf_nets_b200.ops import pointnet_pool as pp

    def stats(h2, W):
        h = torch.matmul(W, h2)
        vmax, imax = h.max(2)
        vmin, imin = h.min(2)
        return h.mean((0, 2)), h.var((0, 2), unbiased=False), vmax, vmin, imax, imin
    monkeypatch.setattr(pp, "_pool_stats", stats)
    torch.manual_seed(0)
    B, Cin, C, N = 3, 256, 512, 40
    h2 = torch.relu(torch.randn(B, Cin, N, dtype=torch.float64)).requires_grad_(True)
    W = (torch.randn(C, Cin, dtype=torch.float64) * 0.1).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(C).double()
    with torch.no_grad():
        bn.weight.copy_(torch.randn(C))             # both signs: max- and min-selected channels
        bn.bias.copy_(0.3 * torch.randn(C))
    bn2 = torch.nn.BatchNorm1d(C).double()
    bn2.load_state_dict(bn.state_dict())
    ref = torch.max(torch.relu(bn(torch.matmul(W, h2))), dim=2)[0]
    cot = torch.randn_like(ref)
    gr = torch.autograd.grad((ref * cot).sum(), [h2, W, bn.weight, bn.bias])
    out = pp.pooled_bn_relu_max(h2, W, bn2)
    go = torch.autograd.grad((out * cot).sum(), [h2, W, bn2.weight, bn2.bias])
    assert rel(out, ref) < 1e-12
    for a, b in zip(go, gr):
        assert rel(a, b) < 1e-10
    assert rel(bn2.running_mean, bn.running_mean) < 1e-12 and rel(bn2.running_var, bn.running_var) < 1e-12
    assert int(bn2.num_batches_tracked) == int(bn.num_batches_tracked) == 1


def test_deferred_batchnorm_counters():
    """ops/_counters.py: inside `deferred()` the BatchNorm step counters of the fused paths are collected and bumped once by the
    outermost context (one multi-tensor add); outside any context `bump` adds immediately; values equal either way."""
    from dpf_nets_b200.ops import _counters
    a, b, c = (torch.zeros((), dtype=torch.long) for _ in range(3))
    _counters.bump(a)
    assert int(a) == 1
    with _counters.deferred():
        _counters.bump(a)
        _counters.bump(b)
        with _counters.deferred():          # nested: still pending
            _counters.bump(c)
        assert (int(a), int(b), int(c)) == (1, 0, 0)
    assert (int(a), int(b), int(c)) == (2, 1, 1)
    try:
        with _counters.deferred():
            _counters.bump(b)
            raise RuntimeError("x")
    except RuntimeError:
        pass
    assert int(b) == 2                      # flushed on the way out, and the context is reset
    _counters.bump(b)
    assert int(b) == 3


def test_pointnet_stats_merge_host():
    """ops/pointnet_train.py::_merge_stats (Chan et al. merge of per-work-item {count, mean, M2}) against the direct statistics,
    including empty work items."""
    from dpf_nets_b200.ops.pointnet_train import _merge_stats
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((1000, 7), generator=gen, dtype=torch.float64) * 3 + 5
    cuts = [0, 0, 130, 131, 600, 1000]
    rows = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        seg = x[lo:hi]
        n = float(hi - lo)
        mean = seg.mean(0) if hi > lo else torch.zeros(7, dtype=torch.float64)
        m2 = ((seg - mean) ** 2).sum(0) if hi > lo else torch.zeros(7, dtype=torch.float64)
        rows.append(torch.stack([torch.full((7,), n, dtype=torch.float64), mean, m2], 1))
    mean, var = _merge_stats(torch.stack(rows).float())
    assert torch.allclose(mean, x.mean(0), rtol=1e-6, atol=1e-6)
    assert torch.allclose(var, x.var(0, unbiased=False), rtol=1e-5)
