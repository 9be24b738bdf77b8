"""Flat parameter arena of a stack of conditional coupling layers.

B200-first storage for the reference's CondRealNVPFlow3D parameters (lib/networks/flows.py:10-93):
ONE fp32 tensor holds every trainable parameter of every layer in the order the kernels read them
(csrc/coupling.cuh `branch_layout`), one more holds the BatchNorm running statistics.  The
reference's state_dict key names / shapes are kept at the checkpoint boundary
(`_save_to_state_dict` / `_load_from_state_dict`), so reference checkpoints load unchanged.
"""
import math

import numpy as np
import torch
import torch.nn as nn

F = 64  # conditioner width the kernels are specialised on (p_decoder_n_features in every config)
BRANCHES = ("mu", "logvar")


def branch_param_fields(br, k, w, G):
    """(reference key suffix, shape) in arena order - must match csrc/coupling.cuh::branch_layout."""
    t0 = "T_%s_0.%s_" % (br, br)
    out = [(t0 + "sd0.weight", (1, F, k)), (t0 + "sd0_bn.weight", (F,)), (t0 + "sd0_bn.bias", (F,)),
           (t0 + "sd1.weight", (1, F, F))]
    for kind in ("w", "b"):
        f = "T_%s_0_cond_%s.%s_sd1_film_%s" % (br, kind, br, kind)
        out += [(f + "0.weight", (F, G)), (f + "0_bn.weight", (F,)), (f + "0_bn.bias", (F,)),
                (f + "1.weight", (F, F)), (f + "1.bias", (F,))]
    t1 = "T_%s_1.%s_sd2." % (br, br)
    out += [(t1 + "weight", (1, w, F)), (t1 + "bias", (1, w))]
    return out


def branch_stat_fields(br):
    """BatchNorm modules of a branch in stats-arena order (running_mean, running_var each)."""
    t0 = "T_%s_0.%s_" % (br, br)
    return [t0 + "sd0_bn", t0 + "sd1_bn",
            "T_%s_0_cond_w.%s_sd1_film_w0_bn" % (br, br), "T_%s_0_cond_b.%s_sd1_film_b0_bn" % (br, br)]


class ArenaLayout:
    """Offsets of every named tensor of every layer inside the param / stats arenas."""

    def __init__(self, layer_specs, G):
        self.G = G
        self.layer_specs = [(pre, list(w)) for pre, w in layer_specs]
        self.L = len(self.layer_specs)
        self.param_index = {}   # full key -> (offset, shape)
        self.stat_index = {}    # full key (running_mean / running_var) -> (offset, shape)
        self.nbt_index = {}     # full key (num_batches_tracked) -> slot
        self.eps_keys = []
        meta = np.zeros((self.L, 8), dtype=np.int64)
        poff = soff = 0
        slot = 0
        for l, (pre, warp) in enumerate(self.layer_specs):
            keep = [c for c in (0, 1, 2) if c not in warp]
            k, w = len(keep), len(warp)
            assert sorted(warp) == list(warp) and k + w == 3 and k in (1, 2), "bad warp_inds %r" % (warp,)
            meta[l] = [poff, soff, k, w, keep[0], keep[1] if k == 2 else -1, warp[0], warp[1] if w == 2 else -1]
            for br in BRANCHES:
                for key, shape in branch_param_fields(br, k, w, G):
                    self.param_index[pre + key] = (poff, shape)
                    poff += int(np.prod(shape))
                for bn in branch_stat_fields(br):
                    for suffix in (".running_mean", ".running_var"):
                        self.stat_index[pre + bn + suffix] = (soff, (F,))
                        soff += F
                    self.nbt_index[pre + bn + ".num_batches_tracked"] = slot
                    slot += 1
            self.eps_keys.append(pre + "eps")
        self.n_params, self.n_stats, self.n_bn = poff, soff, slot
        self.meta = meta

    def layer_param_range(self, l):
        start = int(self.meta[l, 0])
        end = int(self.meta[l + 1, 0]) if l + 1 < self.L else self.n_params
        return start, end


def init_arena(layout, weight_std, generator=None):
    """Reference initialisation (flows.py:25-93, layers.py:29-39): kaiming-uniform SharedDots and
    first FiLM Linears, N(0, weight_std) final layers with zero bias, BN weight 1 / bias 0."""
    flat = torch.zeros(layout.n_params)
    for key, (off, shape) in layout.param_index.items():
        n = int(np.prod(shape))
        v = flat[off:off + n].view(shape)
        leaf = key.rsplit(".", 2)[-2] if key.count(".") >= 2 else key
        if key.endswith("_bn.weight"):
            v.fill_(1.0)
        elif key.endswith("_bn.bias") or key.endswith(".bias"):
            v.zero_()
        elif leaf.endswith("sd2") or leaf.endswith("film_w1") or leaf.endswith("film_b1"):
            v.normal_(std=weight_std, generator=generator)
        elif leaf.endswith("sd0") or leaf.endswith("sd1"):
            fan_in = shape[1] * shape[2]            # kaiming_uniform_(a=0) on (1,out,in): fan_in = out*in... see note
            bound = math.sqrt(6.0 / fan_in)
            v.uniform_(-bound, bound, generator=generator)
        else:                                        # nn.Linear default: kaiming_uniform_(a=sqrt(5)) -> U(+-1/sqrt(in))
            bound = 1.0 / math.sqrt(shape[1])
            v.uniform_(-bound, bound, generator=generator)
    return flat


def init_stats(layout):
    flat = torch.zeros(layout.n_stats)
    for key, (off, _) in layout.stat_index.items():
        if key.endswith("running_var"):
            flat[off:off + F] = 1.0
    return flat


class CouplingStack(nn.Module):
    """Arena-backed stack of conditional coupling layers; base of CondRealNVPFlow3D, ...Triple and
    LocalCondRNVPDecoder in this package."""

    def __init__(self, layer_specs, f_n_features, g_n_features, weight_std=0.01, eps=1e-6):
        super().__init__()
        if f_n_features != F:
            raise ValueError("dpf_nets_b200 coupling kernels are specialised on f_n_features == %d (got %d)"
                             % (F, f_n_features))
        self.f_n_features = f_n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.eps_value = float(np.float32(eps))
        self.layout = ArenaLayout(layer_specs, g_n_features)
        self.arena = nn.Parameter(init_arena(self.layout, weight_std))
        self.register_buffer("stats", init_stats(self.layout), persistent=False)
        self.register_buffer("num_batches_tracked", torch.zeros(self.layout.n_bn, dtype=torch.long), persistent=False)
        self.register_buffer("layer_meta", torch.from_numpy(self.layout.meta.copy()), persistent=False)
        # 'fp32' (CUDA cores, exact), 'bf16' / 'bf16x3' (tcgen05 tensor cores), or 'auto'
        # (= bf16x3 while training, bf16 in eval mode); see _flowfn.resolve_precision
        self.precision = "auto"

    # ---- named views ---------------------------------------------------------------------
    def named_views(self, grad=False):
        """{reference key: view} over the parameter arena (or its .grad) and the stats arena."""
        src = self.arena.grad if grad else self.arena.data
        out = {}
        for key, (off, shape) in self.layout.param_index.items():
            out[key] = src[off:off + int(np.prod(shape))].view(shape)
        if not grad:
            for key, (off, shape) in self.layout.stat_index.items():
                out[key] = self.stats[off:off + F]
        return out

    # ---- checkpoint boundary: the reference's key layout ----------------------------------
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for key, v in self.named_views().items():
            destination[prefix + key] = v if keep_vars else v.detach()
        for key, slot in self.layout.nbt_index.items():
            destination[prefix + key] = self.num_batches_tracked[slot]
        for key in self.layout.eps_keys:
            destination[prefix + key] = torch.tensor([self.eps_value], dtype=torch.float32, device=self.arena.device)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        views = self.named_views()
        expected = set()
        with torch.no_grad():
            for key, dst in views.items():
                full = prefix + key
                expected.add(full)
                if full not in state_dict:
                    missing_keys.append(full)
                    continue
                src = state_dict[full]
                if tuple(src.shape) != tuple(dst.shape):
                    error_msgs.append("size mismatch for %s: checkpoint %s vs model %s"
                                      % (full, tuple(src.shape), tuple(dst.shape)))
                    continue
                dst.copy_(src)
            for key, slot in self.layout.nbt_index.items():
                full = prefix + key
                expected.add(full)
                if full in state_dict:
                    self.num_batches_tracked[slot] = int(state_dict[full])
                elif strict:
                    missing_keys.append(full)
            for key in self.layout.eps_keys:
                full = prefix + key
                expected.add(full)
                if full in state_dict:
                    val = float(state_dict[full].reshape(-1)[0])
                    if abs(val - self.eps_value) > 1e-12:
                        error_msgs.append("%s = %g differs from this module's eps %g" % (full, val, self.eps_value))
                elif strict:
                    missing_keys.append(full)
        if strict:
            for k in state_dict.keys():
                if k.startswith(prefix) and k not in expected:
                    unexpected_keys.append(k)

    def extra_repr(self):
        return "layers=%d, F=%d, G=%d, params=%d, precision=%s" % (
            self.layout.L, F, self.g_n_features, self.layout.n_params, self.precision)
