"""SharedDot / Swish with the reference's constructor and parameter names
(lib/networks/layers.py:5-45)."""
import math

import torch
import torch.nn as nn


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


class SharedDot(nn.Module):
    """Point-wise linear map over channels: out[b,:,n] = W[0] @ in[b,:,n] (+ bias)."""

    def __init__(self, in_features, out_features, n_channels, bias=False, init_weight=None, init_bias=None):
        super().__init__()
        self.in_features, self.out_features, self.n_channels = in_features, out_features, n_channels
        self.init_weight, self.init_bias = init_weight, init_bias
        self.weight = nn.Parameter(torch.empty(n_channels, out_features, in_features))
        self.bias = nn.Parameter(torch.empty(n_channels, out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            if self.init_weight:
                self.weight.uniform_(-self.init_weight, self.init_weight)
            else:
                bound = math.sqrt(6.0 / (self.out_features * self.in_features))  # kaiming_uniform_(a=0) on (1,out,in)
                self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.fill_(self.init_bias if self.init_bias else 0.0)

    def forward(self, input):
        out = torch.matmul(self.weight, input.unsqueeze(1)).squeeze(1)
        if self.bias is not None:
            out = out + self.bias.unsqueeze(0).unsqueeze(3).squeeze(1)
        return out
