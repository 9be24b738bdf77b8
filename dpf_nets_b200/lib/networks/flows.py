"""Conditional affine-coupling flow layers with the reference's class names, constructor
arguments, forward contract and state_dict keys (lib/networks/flows.py:10-160); the per-point
conditioner, scale/shift and log-det all run in the fused sm_100a kernels.

mode='direct'  : p_out = sqrt(eps + exp(logvar)) * p + mu      (noise -> data, sampling)
mode='inverse' : p_out = (p - mu) / sqrt(eps + exp(logvar))    (data -> noise, training NLL)
"""
from ._arena import CouplingStack
from ._flowfn import run_stack


def _triple_warps(pattern):
    if pattern == 0:
        return [[0], [1], [2]]
    if pattern == 1:
        return [[0, 1], [0, 2], [1, 2]]
    raise ValueError("pattern must be 0 or 1")


class CondRealNVPFlow3D(CouplingStack):
    """One coupling layer: forward(p (B,3,N), g (B,G), mode) -> (p_out, mu, logvar)."""

    def __init__(self, f_n_features, g_n_features, weight_std=0.01, warp_inds=[0],
                 centered_translation=False, eps=1e-6):
        super().__init__([("", list(warp_inds))], f_n_features, g_n_features, weight_std=weight_std, eps=eps)
        self.warp_inds = list(warp_inds)
        self.keep_inds = [c for c in (0, 1, 2) if c not in self.warp_inds]
        self.centered_translation = centered_translation  # accepted and ignored, like the reference

    def forward(self, p, g, mode="direct"):
        P, MU, LV = run_stack(self, p, g, mode)
        return P[0], MU[0], LV[0]


class CondRealNVPFlow3DTriple(CouplingStack):
    """Three coupling layers nvp1..nvp3; returns lists ordered [1, 2, 3] in both modes."""

    def __init__(self, f_n_features, g_n_features, weight_std=0.02, pattern=0, centered_translation=False):
        specs = [("nvp%d." % (j + 1), w) for j, w in enumerate(_triple_warps(pattern))]
        super().__init__(specs, f_n_features, g_n_features, weight_std=weight_std)
        self.pattern = pattern
        self.centered_translation = centered_translation

    def forward(self, p, g, mode="direct"):
        P, MU, LV = run_stack(self, p, g, mode)
        return list(P.unbind(0)), list(MU.unbind(0)), list(LV.unbind(0))
