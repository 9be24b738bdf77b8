"""Test helper: the oracle decoder (oracle/flow_oracle.py) running on the SAME parameters as the product module -
one flat tensor in the product's arena layout, addressed through the reference's state_dict key names - so that
outputs, the NLL and every gradient of the two can be compared tensor by tensor (on the CPU or on the GPU)."""
import torch

from oracle import flow_oracle as fo

BASE_LOGVAR = -3.6990          # configs/generation/chair.yaml:51 p_decoder_base_var


def synth_inputs(B, N, G, rank):
    gen = torch.Generator().manual_seed(1234 + rank)
    return torch.rand((B, 3, N), generator=gen) - 0.5, torch.randn((B, G), generator=gen)


def oracle_train_step_factory(B, N, n_flows=21, G=128, rank=0, device="cpu"):
    """-> step() = fwd + PointFlowNLL + bwd of the oracle; returns (loss, state) with the arena, z = samples[0],
    sum_l logvar_l, d(arena) and dg."""
    from dpf_nets_b200.lib.networks._arena import ArenaLayout, init_arena, init_stats
    specs = fo.decoder_layer_names(n_flows)
    lay = ArenaLayout(specs, G)
    torch.manual_seed(0)
    arena, stats = init_arena(lay, 0.01).to(device), init_stats(lay).to(device)
    arena.requires_grad_(True)
    layers = []
    for pre, warp in specs:
        P = {"eps": torch.tensor([1e-6], device=device)}
        for key, (off, shape) in lay.param_index.items():
            if key.startswith(pre):
                n = 1
                for s in shape:
                    n *= s
                P[key[len(pre):]] = arena[off:off + n].view(shape)
        for key, (off, shape) in lay.stat_index.items():
            if key.startswith(pre):
                P[key[len(pre):]] = stats[off:off + 64]
        layers.append((P, warp))
    p, g = synth_inputs(B, N, G, rank)
    p, g = p.to(device), g.to(device)
    g.requires_grad_(True)
    base_mu, base_lv = torch.zeros_like(p), torch.full_like(p, BASE_LOGVAR)

    def step():
        arena.grad = None
        g.grad = None
        ps, mus, lvs = fo.decoder_forward(layers, p, g, "inverse", training=True)
        nll = fo.point_flow_nll(ps + [p], [base_mu] + mus, [base_lv] + lvs)
        nll.backward()
        return float(nll.detach()), {"arena": arena.detach(), "z": ps[0].detach(), "sum_logvar": sum(lvs).detach(),
                                     "darena": arena.grad.detach(), "dg": g.grad.detach()}
    return step
