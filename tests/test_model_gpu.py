"""GPU: whole-model wiring around the fused decoder - training step, fused AMSGrad, checkpoint
round trip, generating mode, evaluation sweep, entry-point smoke runs on synthetic data."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(**over):
    from dpf_nets_b200 import configs
    c = configs.get('generation/chair')
    c.update(p_decoder_n_flows=2, g_latent_space_size=16, g_prior_n_flows=2, g_prior_n_features=16, **over)
    return c


def test_train_step_checkpoint_and_generate(native_lib, cuda, tmp_path):
    from dpf_nets_b200.lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE
    from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
    from dpf_nets_b200.lib.networks.optimizers import Adam
    cfg = _cfg()
    torch.manual_seed(0)
    model = Local_Cond_RNVP_MC_Global_RNVP_VAE(**cfg).to(cuda)
    crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**cfg)
    opt = Adam(model.parameters(), lr=1e-3, weight_decay=1e-6, betas=(0.9, 0.99), amsgrad=True)
    gen = torch.Generator().manual_seed(1)
    cloud = (torch.rand((6, 3, 300), generator=gen) - 0.5).to(cuda)
    evalc = (torch.rand((6, 3, 300), generator=gen) - 0.5).to(cuda)
    model.train()
    losses = []
    before = model.pc_decoder.arena.detach().clone()
    for _ in range(3):
        out = model(cloud, evalc)
        assert len(out['p_prior_samples']) == 7 and len(out['p_prior_logvars']) == 7 and len(out['g_prior_samples']) == 5
        loss, pnll, gnll, gent = crit(cloud, evalc, out)
        assert torch.isfinite(loss)
        opt.zero_grad()
        loss.backward()
        for n, p in model.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
        opt.step()
        losses.append(loss.item())
    assert not torch.equal(before, model.pc_decoder.arena.detach())
    assert losses[-1] < losses[0]                      # a few Adam steps on one batch reduce the bound
    path = tmp_path / "ck.pkl"
    torch.save({'epoch': 1, 'iter': 0, 'model_state': model.state_dict(), 'optimizer_state': opt.state_dict()}, path,
               pickle_protocol=4)
    ck = torch.load(path, weights_only=False)
    m2 = Local_Cond_RNVP_MC_Global_RNVP_VAE(**_cfg(util_mode='generating')).to(cuda)
    m2.load_state_dict(ck['model_state'])
    assert torch.equal(m2.pc_decoder.arena, model.pc_decoder.arena) and torch.equal(m2.pc_decoder.stats, model.pc_decoder.stats)
    m2.eval()
    with torch.no_grad():
        out = m2(cloud, evalc, n_sampled_points=500)
    assert out['p_prior_samples'][-1].shape == (6, 3, 500) and torch.isfinite(out['p_prior_samples'][-1]).all()


def test_fused_adam_matches_host_formula(native_lib, cuda):
    """Multi-tensor AMSGrad kernel (several launches: > 48 tensors, slabs of 4096, ragged sizes, one
    parameter without gradient) vs the reference formula evaluated on the CPU by the same class."""
    from dpf_nets_b200.lib.networks.optimizers import Adam
    torch.manual_seed(0)
    sizes = [1, 7, 4096, 4097, 10000, 64, 3, 12289] + [5 + 11 * i for i in range(55)]
    w0 = [torch.randn(n) for n in sizes]
    sets = [[torch.nn.Parameter(w.clone().to(d)) for w in w0] for d in ("cpu", cuda)]
    opts = [Adam(ws, lr=1e-2, weight_decay=1e-3, betas=(0.9, 0.99), amsgrad=True) for ws in sets]
    for it in range(5):
        for ws, o in zip(sets, opts):
            o.zero_grad()
            loss = sum((w ** 2).sum() * 0.5 + (w * (it + 1 + 0.1 * j)).sum() for j, w in enumerate(ws) if j != 5)
            loss.backward()
            o.step()
    for a, b in zip(*sets):
        assert torch.allclose(a.detach(), b.detach().cpu(), rtol=1e-5, atol=1e-6)
    assert torch.equal(sets[1][5].detach().cpu(), w0[5])


def test_generation_sweep_metrics(native_lib, cuda):
    from dpf_nets_b200.lib.networks.evaluating import generation_metrics
    gen = torch.Generator().manual_seed(0)
    a = (torch.rand((24, 512, 3), generator=gen) - 0.5).to(cuda)
    b = (torch.rand((24, 512, 3), generator=gen) - 0.5).to(cuda)
    r = generation_metrics(a, b)
    assert 0 <= r['COV-CD'] <= 1 and 0 <= r['1NN-CD'] <= 1 and r['MMD-CD'] > 0 and 0 <= r['JSD'] <= 1
    same = generation_metrics(a, a.clone())
    assert same['MMD-CD'] == 0 and same['COV-CD'] == 1.0


def test_entry_points_synthetic(native_lib, cuda, tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    save = str(tmp_path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train_ae.py"), "generation/chair", "smoke", "1", "0.000256",
                        "--synthetic", "8", "--batch_size", "4", "--path2save", save], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert os.path.exists(os.path.join(save, "models", "DPFNets", "smoke.pkl"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train_ae.py"), "generation/chair", "smoke", "2", "0.000256",
                        "--synthetic", "20", "--batch_size", "4", "--path2save", save, "--resume", "--resume_optimizer",
                        "--cuda_graph"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]      # resumed run: 5 iterations, 3 of them replayed graphs
    r = subprocess.run([sys.executable, os.path.join(ROOT, "evaluate_ae.py"), "generation/chair", "smoke", "test", "2048",
                        "512", "generating", "--synthetic", "8", "--path2save", save], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "1NN-CD" in r.stdout and "COV-CD" in r.stdout


def test_graphed_train_step_matches_eager(native_lib, cuda, monkeypatch):
    """The two-graph training step (_graphstep.py) vs the eager loop on the same batches (sampling noise replaced
    by a deterministic offset): same losses; the parameter trajectory agrees with the eager one as well as a
    SECOND eager run does (training itself is not bit-reproducible: float atomics in the decoder backward, and
    Adam normalises noise-level gradient entries to +-lr - tools/graph_probe.py); learning-rate changes between
    replays are honoured; optimizer step counters advance."""
    from dpf_nets_b200.lib.networks import models as models_mod
    from dpf_nets_b200.lib.networks._graphstep import GraphedTrainStep
    from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
    from dpf_nets_b200.lib.networks.optimizers import Adam
    monkeypatch.setattr(models_mod, "_reparameterize", lambda mu, logvar: mu + 0.3 * torch.exp(0.5 * logvar))
    cfg = _cfg()
    gen = torch.Generator().manual_seed(3)
    batches = [((torch.rand((6, 3, 300), generator=gen) - 0.5).to(cuda), (torch.rand((6, 3, 300), generator=gen) - 0.5).to(cuda))
               for _ in range(7)]
    lrs = [1e-3, 1e-3, 1e-3, 1e-3, 5e-4, 5e-4, 2e-4]
    runs = {}
    for kind in ("eager", "eager2", "graphed"):
        torch.manual_seed(0)
        model = models_mod.Local_Cond_RNVP_MC_Global_RNVP_VAE(**cfg).to(cuda).train()
        crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**cfg)
        opt = Adam(model.parameters(), lr=1e-3, weight_decay=1e-6, betas=(0.9, 0.99), amsgrad=True)
        init = {n: p.detach().clone() for n, p in model.named_parameters()}
        step = GraphedTrainStep(model, crit, opt, eager_steps=2 if kind == "graphed" else 10 ** 9)
        losses = []
        for (c, e), lr in zip(batches, lrs):
            for group in opt.param_groups:
                group['lr'] = lr
            losses.append(float(step(c, e)[0].detach()))
        if kind == "graphed":
            assert step.graph_a is not None and step.graph_b is not None
        assert {opt.state[p]['step'] for p in model.parameters()} == {len(batches)}
        upd = torch.cat([(p.detach() - init[n]).flatten() for n, p in model.named_parameters()])
        nbt = {k: v.clone() for k, v in model.state_dict().items() if 'num_batches' in k}
        runs[kind] = (losses, upd, nbt)

    def cos(a, b):
        return float(torch.dot(a, b) / (a.norm() * b.norm()))
    le, l2, lg = runs["eager"][0], runs["eager2"][0], runs["graphed"][0]
    floor = cos(runs["eager"][1], runs["eager2"][1])
    got = cos(runs["eager"][1], runs["graphed"][1])
    print("losses eager", le, "graphed", lg, "update cosine: eager-eager2 %.4f, eager-graphed %.4f" % (floor, got))
    for a, b, c in zip(le, lg, l2):
        assert abs(a - b) <= 2e-3 * abs(a) + 3 * abs(a - c)
    assert got > floor - 0.06 and got > 0.9
    assert abs(float(runs["eager"][1].norm()) - float(runs["graphed"][1].norm())) < 0.03 * float(runs["eager"][1].norm())
    for k, v in runs["eager"][2].items():
        assert torch.equal(v, runs["graphed"][2][k]), k


# ---- whole-model parity against the reference (BASELINE configs 2, 3, 4) -------------------------------------
@pytest.mark.parametrize("name,cls,ic", [("generation_chair", "Local_Cond_RNVP_MC_Global_RNVP_VAE", False),
                                         ("autoencoding_all_original", "Local_Cond_RNVP_MC_Global_RNVP_VAE", False),
                                         ("svr_all", "Local_Cond_RNVP_MC_Global_RNVP_VAE_IC", True)])
@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_wholemodel_training_loss_and_gradients_vs_reference(native_lib, cuda, monkeypatch, name, cls, ic, precision):
    """Whole-model training-mode forward + VAE loss + backward of the generation (G=128), AE all_original (G=512) and
    SVR all (image encoder + G=512) models against the UNMODIFIED reference (tests/golden/make_golden_wholemodel.py:
    weights regenerated from key names, torch.randn_like replaced by the same seeded stream on both sides).
    Tolerances: (loss, pnll, gnll, gent) 1e-4 relative in fp32 mode / 0.5 % on the tensor path (north star: total
    NLL within 0.5 %), z and sum-logvar 1e-3 / 2e-2, gradients of parameters of every sub-module 2e-2 / 1e-1."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _detstate import DetRandn, GRAD_KEYS, det_state, whole_model_inputs
    from dpf_nets_b200 import configs
    from dpf_nets_b200.lib.networks import models
    from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "wholemodel.pt"), weights_only=False)[name]
    c = configs.get(fx["config_path"][len("configs/"):-len(".yaml")])
    c["util_mode"] = "training"
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # the ResNet-18 convolutions in true fp32, like the CPU reference
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    m = getattr(models, cls)(**c)
    m.load_state_dict(det_state({k: list(v.shape) for k, v in m.state_dict().items()}))
    m = m.to(cuda).train()
    m.pc_decoder.precision = precision
    m.pc_encoder.precision = "fp32" if precision == "fp32" else "auto"     # fp32 = the exact paths everywhere; auto = tensor paths
    crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**c)
    inp = {k: v.to(cuda) for k, v in whole_model_inputs(fx["B"], fx["N"], 77, ic).items()}
    with DetRandn(5):
        out = m(inp["cloud"], inp["eval_cloud"], inp["image"]) if ic else m(inp["cloud"], inp["eval_cloud"])
    losses = crit(inp["cloud"], inp["eval_cloud"], out)
    losses[0].backward()
    fp32 = precision == "fp32"
    t64 = fx["truth64"]

    def rel(a, b):
        return float((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max())

    def gate(a, ref32, truth, tol, what):
        """within `tol` of the reference's fp32 result, or no further from the fp64 truth (the same reference modules run
        in double) than three times the reference's own fp32 result is: batch-statistics BatchNorm over B = 6..8 shapes
        (FiLM nets, latent flows, ResNet) makes some gradients ill-conditioned in fp32 on BOTH sides."""
        e = rel(a, ref32)
        assert e < tol or rel(a, truth) < 3 * rel(ref32, truth) + 1e-6, (what, e, rel(a, truth), rel(ref32, truth))

    got = torch.stack([l.detach().double().cpu() for l in losses])
    tol_loss = 1e-4 if fp32 else 5e-3
    assert ((got - fx["losses"]).abs() <= tol_loss * fx["losses"].abs()).all(), (got, fx["losses"])
    # tensor path: the whole PointNet encoder runs on split-bf16 GEMMs (~1e-5 per layer), then the posterior head's
    # batch-statistics BatchNorm over B = 6..8 shapes amplifies it (measured 4e-5 .. 1.1e-4)
    assert rel(out["g_posterior_mus"], fx["g_posterior_mus"]) < (1e-4 if fp32 else 1e-3)
    gate(out["p_prior_samples"][0], fx["z"], t64["z"], 1e-3 if fp32 else 2e-2, "z")
    gate(out["p_prior_logvars"].tail_total, fx["sum_logvar"], t64["sum_logvar"], 1e-3 if fp32 else 2e-2, "sum_logvar")
    named = dict(m.named_parameters())
    dec = m.pc_decoder.named_views(grad=True)
    for k in GRAD_KEYS[ic]:
        gk = dec[k[len("pc_decoder."):]] if k.startswith("pc_decoder.") else named[k].grad
        # tensor path: the float atomics of the backward reorder sums from run to run, and batch-statistics BatchNorm over
        # B = 3..4 shapes amplifies that to a few 1e-2 on single gradient entries (measured run-to-run spread) - gated at 1e-1
        # there; the exact fp32 path carries the tight gate
        gate(gk, fx["grads"][k], t64["grads"][k], 2e-2 if fp32 else 1e-1, k)


def test_entry_points_svr_and_predicting(native_lib, cuda, tmp_path):
    """train_svr.py on synthetic image + cloud pairs (BASELINE config 4 model: ResNet-18 image encoder, G = 512,
    2500-point clouds), then evaluate_ae.py in 'predicting' mode (per-batch CD + F1, evaluating.py:144-205) and
    'evaluating' mode with --orig_scale_evaluation."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    save = str(tmp_path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train_svr.py"), "svr/all", "svr_smoke", "1", "0.000256",
                        "--synthetic", "8", "--batch_size", "4", "--path2save", save], capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert os.path.exists(os.path.join(save, "models", "DPFNets", "svr_smoke.pkl"))
    ck = torch.load(os.path.join(save, "models", "DPFNets", "svr_smoke.pkl"), weights_only=False)
    n_ref_params = 25203197          # svr/all.yaml model
    assert sum(v['exp_avg'].numel() for v in ck['optimizer_state']['state'].values()) == n_ref_params
    assert len(ck['optimizer_state']['state']) > 2000       # the reference's per-tensor layout, not one arena entry
    r = subprocess.run([sys.executable, os.path.join(ROOT, "evaluate_ae.py"), "svr/all", "svr_smoke", "test", "2500",
                        "2500", "predicting", "--synthetic", "6", "--path2save", save], capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "CD:" in r.stdout and "F1:" in r.stdout
    r = subprocess.run([sys.executable, os.path.join(ROOT, "evaluate_ae.py"), "svr/all", "svr_smoke", "test", "2500",
                        "1024", "evaluating", "--orig_scale_evaluation", "--synthetic", "6", "--path2save", save],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "CD:" in r.stdout


def test_f_score_vs_oracle(native_lib, cuda):
    """f_score (utils.py:38-42) through the native NN search against the numpy oracle."""
    import numpy as np
    from oracle import metrics_oracle as mo
    from oracle import structural as so
    from dpf_nets_b200.lib.networks.utils import f_score
    g = torch.Generator().manual_seed(9)
    a = (torch.rand((5, 700, 3), generator=g) - 0.5) * 0.2
    b = a + 0.02 * torch.randn((5, 700, 3), generator=g)
    got = f_score(a.to(cuda), b.to(cuda)).cpu().numpy()
    d1, _, d2, _ = so.nndistance(a.numpy(), b.numpy())
    want = mo.f_score(d1, d2)
    assert np.allclose(got, want, rtol=1e-5) and (got > 0).all() and (got < 100).all()
