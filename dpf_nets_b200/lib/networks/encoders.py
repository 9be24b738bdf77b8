"""PointNet cloud encoder and latent feature heads with the reference's names / state_dict keys
(lib/networks/encoders.py:9-83).  Eval mode: encoder + max-pool run as ONE fused tcgen05 kernel
(dpf_pointnet_eval_forward, csrc/pointnet.cu).  Train mode (batch statistics, autograd): the whole encoder + max-pool is
one autograd function over the library's own tcgen05 kernels (ops/pointnet_train.py: layer 0 analytic from the input
moments, layers 1-2 as GEMM kernels with the BatchNorm + ReLU of the previous layer applied on load and the output
statistics in the epilogue, the last layer + max-pool without materialising the (B,512,N) activation, fused dgrad / wgrad
kernels backward); other widths, or an input that needs a gradient, keep the narrow layers on the library path."""
import ctypes

import torch
import torch.nn as nn

from ... import _lib
from .layers import SharedDot, Swish


class PointNetCloudEncoder(nn.Module):
    def __init__(self, init_n_channels, init_n_features, n_features):
        super().__init__()
        self.init_n_channels, self.init_n_features, self.n_features = init_n_channels, init_n_features, n_features
        widths = [init_n_channels, init_n_features] + list(n_features)
        names = ['init_sd'] + ['sd%d' % i for i in range(len(n_features))]
        self.features = nn.Sequential()
        for name, cin, cout in zip(names, widths[:-1], widths[1:]):
            self.features.add_module(name, SharedDot(cin, cout, 1, bias=False))
            self.features.add_module(name + '_bn', nn.BatchNorm1d(cout))
            self.features.add_module(name + '_relu', nn.ReLU(inplace=True))

    def forward(self, input):
        return self.features(input)

    precision = 'auto'      # 'fp32' keeps the library path in eval mode too

    def _fused_ok(self, input):
        return (not self.training and not torch.is_grad_enabled() and input.is_cuda and input.dtype == torch.float32
                and self.precision != 'fp32' and self.init_n_channels == 3 and self.init_n_features == 64
                and list(self.n_features) == [128, 256, 512] and input.dim() == 3 and input.shape[1] == 3)

    def _fused_train_ok(self, input):
        return (self.training and torch.is_grad_enabled() and input.is_cuda and input.dtype == torch.float32
                and self.precision != 'fp32' and len(self.n_features) >= 2 and self.n_features[-2] == 256
                and self.n_features[-1] == 512 and input.dim() == 3 and hasattr(self.features, 'sd2')
                and len(self.features) == 12)

    fused_layers = True     # class switch: False keeps init_sd .. sd1 on the library path in train mode (tests compare both)

    def _fused_all_layers_ok(self, input):
        return (self.fused_layers and not input.requires_grad and self.init_n_channels == 3 and self.init_n_features == 64
                and list(self.n_features) == [128, 256, 512])

    def global_features(self, input):
        """max over the points of forward(input): (B, 3, N) -> (B, n_features[-1]); what the models take
        from the encoder (reference models.py:130-131).  Eval mode without autograd runs the fused kernel
        (bf16 tensor cores, fp32 accumulation)."""
        if self._fused_train_ok(input):
            f = self.features
            if self._fused_all_layers_ok(input):
                from ...ops.pointnet_train import pointnet_train_forward
                return pointnet_train_forward(input, [f.init_sd, f.sd0, f.sd1, f.sd2], [f.init_sd_bn, f.sd0_bn, f.sd1_bn, f.sd2_bn])
            from ...ops.pointnet_pool import pooled_bn_relu_max
            h2 = f[:-3](input)                                   # init_sd .. sd1_relu: (B, 256, N)
            return pooled_bn_relu_max(h2, f.sd2.weight[0], f.sd2_bn)
        if not self._fused_ok(input):
            return torch.max(self.forward(input), dim=2)[0]
        x = input.contiguous()
        B, _, N = x.shape
        f = self.features
        sds = [f.init_sd, f.sd0, f.sd1, f.sd2]
        bns = [f.init_sd_bn, f.sd0_bn, f.sd1_bn, f.sd2_bn]
        nbytes = ctypes.c_longlong(0)
        _lib.check(_lib.lib().dpf_pointnet_workspace_bytes(ctypes.byref(nbytes)), "dpf_pointnet_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        out = torch.empty((B, self.n_features[-1]), dtype=torch.float32, device=x.device)
        weights = [sd.weight.detach().reshape(sd.weight.shape[-2], sd.weight.shape[-1]).contiguous() for sd in sds]
        bn = []
        for m in bns:
            bn += [m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var]
        wp = (ctypes.c_void_p * 4)(*[w.data_ptr() for w in weights])
        bp = (ctypes.c_void_p * 16)(*[t.data_ptr() for t in bn])
        with torch.cuda.device(x.device):
            _lib.call("dpf_pointnet_eval_forward", x, int(B), int(N), wp, bp, float(bns[0].eps), ws, out, device=x.device)
        return out


class FeatureEncoder(nn.Module):
    def __init__(self, n_layers, in_features, latent_space_size, deterministic=False, batch_norm=True,
                 mu_weight_std=0.001, mu_bias=0.0, logvar_weight_std=0.01, logvar_bias=0.0, easy_init=False):
        super().__init__()
        self.n_layers, self.in_features, self.latent_space_size = n_layers, in_features, latent_space_size
        self.deterministic, self.batch_norm = deterministic, batch_norm
        if n_layers > 0:
            self.features = nn.Sequential()
            for i in range(n_layers):
                self.features.add_module('mlp%d' % i, nn.Linear(in_features, in_features, bias=False))
                if batch_norm:
                    self.features.add_module('mlp%d_bn' % i, nn.BatchNorm1d(in_features))
                self.features.add_module('mlp%d_swish' % i, Swish())
        self.mus = nn.Sequential()
        self.mus.add_module('mu_mlp0', nn.Linear(in_features, latent_space_size, bias=True))
        heads = [(self.mus[-1], mu_weight_std, mu_bias)]
        if not deterministic:
            self.logvars = nn.Sequential()
            self.logvars.add_module('logvar_mlp0', nn.Linear(in_features, latent_space_size, bias=True))
            heads.append((self.logvars[-1], logvar_weight_std, logvar_bias))
        if not easy_init:
            with torch.no_grad():
                for lin, std, bias in heads:
                    lin.weight.normal_(std=std)
                    lin.bias.fill_(bias)

    fused = True      # class switch: False keeps the plain module chain on the GPU too (tests compare both)

    def _features(self, x):
        if not (self.fused and self.batch_norm and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
            return self.features(x)
        from ...ops.latent import bn_swish      # Linear -> [BatchNorm1d + Swish as one fused kernel] per layer
        for i in range(self.n_layers):
            x = bn_swish(getattr(self.features, 'mlp%d' % i)(x), getattr(self.features, 'mlp%d_bn' % i))
        return x

    def forward(self, input):
        feats = self._features(input) if self.n_layers > 0 else input
        if self.deterministic:
            return self.mus(feats)
        return self.mus(feats), self.logvars(feats)
