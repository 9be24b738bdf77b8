"""Deterministic, architecture-independent weights for whole-model parity: every tensor of a state_dict is
regenerated from its KEY NAME and shape alone, so the reference model (in the dev container, when the goldens
are generated) and this package's model (on the GPU box) hold identical parameters without shipping a
100 MB checkpoint.  Also: a drop-in deterministic `torch.randn_like` for both sides."""
import zlib

import torch


def det_tensor(key, shape, is_bn_affine=False):
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()))
    shape = tuple(shape)
    if key.endswith("num_batches_tracked"):
        return torch.tensor(3, dtype=torch.long)
    if key.endswith(".eps"):
        return torch.tensor([1e-6], dtype=torch.float32)
    if key.endswith("running_var"):
        return 0.5 + torch.rand(shape, generator=g)
    if key.endswith("running_mean"):
        return 0.1 * torch.randn(shape, generator=g)
    if is_bn_affine and key.endswith("weight"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if key.endswith("g0_prior_mus") or key.endswith("g0_prior_logvars"):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith("bias") or len(shape) <= 1:
        return 0.05 * torch.randn(shape, generator=g)
    # weights: N(0, 1/fan_in); SharedDot weights are (1, out, in), Linear (out, in), Conv (out, in, kh, kw)
    dims = shape[2:] if (len(shape) == 3 and shape[0] == 1) else shape[1:]
    fan_in = 1
    for d in dims:
        fan_in *= d
    scale = fan_in ** -0.5
    # last layers of the point-flow conditioners: smaller, so 63 stacked layers stay well-conditioned
    if any(key.endswith(sfx) for sfx in ("sd2.weight", "film_w1.weight", "film_b1.weight")) and "pc_decoder" in key:
        scale *= 0.3
    return scale * torch.randn(shape, generator=g)


def det_state(key_shapes):
    """{key: shape} (e.g. tests/golden/model_keys.json[...]) -> state_dict."""
    out = {}
    for k, v in key_shapes.items():
        leaf = k.rsplit(".", 1)[0]
        out[k] = det_tensor(k, v, is_bn_affine=(leaf + ".running_mean") in key_shapes)
    return out


class DetRandn:
    """Context manager replacing torch.randn_like by a seeded CPU-generated stream (moved to the argument's
    device), identical wherever it runs as long as the calls come in the same order with the same shapes."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)

    def __call__(self, t, *a, **k):
        return torch.randn(t.shape, generator=self.gen, dtype=torch.float32).to(device=t.device, dtype=t.dtype)

    def __enter__(self):
        self._orig = torch.randn_like
        torch.randn_like = self
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig
        return False


def whole_model_inputs(B, N, seed, with_image):
    g = torch.Generator().manual_seed(seed)
    out = {"cloud": torch.rand((B, 3, N), generator=g) - 0.5, "eval_cloud": torch.rand((B, 3, N), generator=g) - 0.5}
    if with_image:
        out["image"] = torch.randn((B, 4, 224, 224), generator=g)
    return out


DECODER_GRAD_KEYS = ["pc_decoder.flows.0.nvp1.T_mu_0.mu_sd1.weight", "pc_decoder.flows.20.nvp3.T_logvar_1.logvar_sd2.weight",
                     "pc_decoder.flows.10.nvp2.T_mu_0_cond_w.mu_sd1_film_w0.weight"]
GRAD_KEYS = {
    False: ["g0_prior_mus", "pc_encoder.features.init_sd.weight", "g_posterior.mus.mu_mlp0.weight"] + DECODER_GRAD_KEYS,
    True: ["img_encoder.conv1.weight", "pc_encoder.features.init_sd.weight", "g_posterior.mus.mu_mlp0.weight"] + DECODER_GRAD_KEYS,
}
