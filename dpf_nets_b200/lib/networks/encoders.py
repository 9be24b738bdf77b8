"""PointNet cloud encoder and latent feature heads with the reference's names / state_dict keys
(lib/networks/encoders.py:9-83).  Round 1: the shared-MLP GEMMs of the encoder go through the
library path (cuBLAS bmm + ATen batch-norm); the fused tcgen05 tiles are DESIGN.md section 7 item 3."""
import torch
import torch.nn as nn

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

    def forward(self, input):
        feats = self.features(input) if self.n_layers > 0 else input
        if self.deterministic:
            return self.mus(feats)
        return self.mus(feats), self.logvars(feats)
