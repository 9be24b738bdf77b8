"""One launch each of the evaluation-side kernels at BASELINE sizes, for `ncu` captures (tools/gpu_evidence_r02.sh):
the fused sampling decoder (63 layers, 32 x 2048), the one-evaluation Chamfer all-pairs kernel (64 x 64 clouds of 2048),
the fused all-pairs EMD (8 x 8) and the score reduction (1000 x 1000 matrices)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200.lib.networks.decoders import LocalCondRNVPDecoder  # noqa: E402
from dpf_nets_b200.ops import pairwise_cd, pairwise_emd  # noqa: E402
from dpf_nets_b200.ops.metrics import cd_scores  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
torch.manual_seed(0)
dec = LocalCondRNVPDecoder(21, 64, 128).to(dev).eval()
z = torch.randn((32, 3, 2048), generator=g).to(dev)
lat = torch.randn((32, 128), generator=g).to(dev)
with torch.no_grad():
    dec(z, lat, mode="direct")
A = (torch.rand((64, 2048, 3), generator=g) - 0.5).to(dev)
B = (torch.rand((64, 2048, 3), generator=g) - 0.5).to(dev)
pairwise_cd(A, B)
pairwise_cd(A, A, symmetric=True)
pairwise_emd(A[:8].contiguous(), B[:8].contiguous())
M = [torch.rand((1000, 1000), generator=g).to(dev) for _ in range(3)]
print(cd_scores((M[0] + M[0].t()) / 2, M[1], (M[2] + M[2].t()) / 2).tolist())
# train-mode PointNet last layer + max-pool (statistics / max kernel) and the fused latent blocks
from dpf_nets_b200.lib.networks.decoders import GlobalRNVPDecoder  # noqa: E402
from dpf_nets_b200.ops.pointnet_pool import _pool_stats  # noqa: E402
h2 = torch.relu(torch.randn((32, 256, 2048), generator=g)).to(dev)
W3 = (torch.randn((512, 256), generator=g) * 0.08).to(dev)
_pool_stats(h2, W3)
gp = GlobalRNVPDecoder(7, 128, 128).to(dev).train()
lat = torch.randn((32, 128), generator=g).to(dev).requires_grad_(True)
out = gp(lat, mode="inverse")
(out[0][0].sum() + sum(t.sum() for t in out[2])).backward()
torch.cuda.synchronize()
