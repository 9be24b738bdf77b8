"""GPU parity of the approximate-EMD path through the C ABI vs the reference's own CUDA kernels
(oracle/_ref) and the C oracle.  The reference uses __expf and sequential fp32 sums; ours keeps the
per-thread ascending summation order, so match agrees to rounding of the cross-tile structure; the
gate on the cost is 1e-4 relative (north_star gives no EMD tolerance)."""
import numpy as np
import pytest
import torch

from oracle import structural as so

pytestmark = pytest.mark.gpu


def _clouds(b, n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((b, n, 3), generator=g) - 0.5).to(dev).contiguous()


@pytest.fixture(scope="module")
def B(native_lib, cuda):
    from dpf_nets_b200.lib.metrics.StructuralLosses import StructuralLossesBackend
    return StructuralLossesBackend


@pytest.mark.parametrize("b,n,m", [(3, 256, 256), (2, 1024, 1024), (2, 300, 150), (1, 2048, 2048), (2, 100, 400)])
def test_approxmatch_matchcost_vs_reference_kernels(B, cuda, b, n, m):
    ref = so.RefCuda()
    x, y = _clouds(b, n, n + 1, cuda), _clouds(b, m, m + 2, cuda)
    match, _ = B.ApproxMatch(x, y)
    rmatch = ref.approxmatch(x, y)
    torch.cuda.synchronize()
    assert match.shape == (b, m, n)
    assert ((match - rmatch).abs().max() / rmatch.abs().max()).item() < 1e-4
    cost, rcost = B.MatchCost(x, y, match), ref.matchcost(x, y, rmatch)
    assert ((cost - rcost).abs() / rcost.abs()).max().item() < 1e-4
    # row / column marginals of a transport plan: every point ships (about) its mass
    assert (match.sum(1) <= max(1.0, m / n) * 1.001).all() and (match.sum(2) <= max(1.0, n / m) * 1.001).all()
    g1, g2 = B.MatchCostGrad(x, y, match)
    r1, r2 = ref.matchcost_grad(x, y, match)
    assert ((g1 - r1).abs().max() / r1.abs().max()).item() < 1e-4
    assert ((g2 - r2).abs().max() / r2.abs().max()).item() < 1e-4


def test_emd_vs_c_oracle_small(B, cuda):
    x, y = _clouds(2, 64, 5, cuda), _clouds(2, 64, 6, cuda)
    match, _ = B.ApproxMatch(x, y)
    om = so.approxmatch(x.cpu().numpy(), y.cpu().numpy())
    assert np.abs(match.cpu().numpy() - om).max() / np.abs(om).max() < 5e-4      # expf vs __expf
    oc = so.matchcost(x.cpu().numpy(), y.cpu().numpy(), om)
    cost = B.MatchCost(x, y, match).cpu().numpy()
    assert np.abs(cost - oc).max() / np.abs(oc).max() < 5e-4


def test_match_cost_autograd_and_pairwise_emd(B, cuda):
    from dpf_nets_b200.lib.metrics.StructuralLosses.match_cost import match_cost
    from dpf_nets_b200.lib.metrics.evaluation_metrics import emd_approx, _pairwise_EMD_CD_
    from dpf_nets_b200.ops import pairwise_emd
    x = _clouds(3, 128, 1, cuda).requires_grad_(True)
    y = _clouds(3, 128, 2, cuda)
    c = match_cost(x, y)
    c.sum().backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0
    A, Bc = _clouds(4, 200, 3, cuda), _clouds(5, 200, 4, cuda)
    pe = pairwise_emd(A, Bc)
    for i in range(4):
        ci = A[i].unsqueeze(0).expand(5, -1, -1).contiguous()
        ref_row = match_cost(ci, Bc)
        assert ((pe[i] - ref_row).abs() / ref_row.abs()).max().item() < 1e-4
    cd, emd = _pairwise_EMD_CD_(A, Bc, 32)
    assert cd.shape == emd.shape == (4, 5)
    assert torch.allclose(emd[0, :1], emd_approx(A[:1], Bc[:1]), rtol=1e-4)
