"""Phase timing of the per-layer tensor-path kernels from in-kernel clock64() stamps.

Build the instrumented library first:  DPF_STAMPS=1 python -m dpf_nets_b200.build
Run:  DPF_LIB_PATH=dpf_nets_b200/_C_stamps/libdpfnets_b200.so python tools/stamp_probe.py
Prints, per kernel class, the mean / max over CTAs of the cycles between consecutive stamps of the
LAST launch (one train step of the bench workload: 63 layers, 32 x 2048 points, bf16x3)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dpf_nets_b200 import _lib
from dpf_nets_b200.lib.networks.decoders import LocalCondRNVPDecoder

NAMES = {
    0: ["start", "tmem_alloc", "tables", "weights_ready", "stats_phase(tiles)", "flush_stats", "grid_barrier", "bn_b_tables",
        "apply_phase(tiles)", "flush_moments", "end"],
    1: ["start", "tmem_alloc", "tables", "pending", "weights_ready", "t0:load_point", "t0:h1+gemm(br0)", "t0:epilogue(br0)",
        "t0:rest(br1)", "other_tiles", "final_flush", "end"],
    2: ["start", "setup+tmem_alloc", "pdl_wait", "tables+m1m2", "t0:load_point+finish", "t0:h1_write", "t0:gemm_wait",
        "t0:epilogue_A", "t0:dgrad_wgrad_wait", "t0:epilogue_B", "t0:dx + other tiles", "last BN wait + cta sync", "cta_epilogue"],
}
KERNEL = {0: "coupling_fwd_train_tc2_kernel", 1: "coupling_bwd_p1_tc2_kernel", 2: "coupling_bwd_p2_tc4_kernel"}


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.lib()
    buf = torch.zeros(3 * 512 * 16, dtype=torch.int64, device=dev)
    assert lib.dpf_debug_stamps(ctypes.c_void_p(buf.data_ptr())) == 0
    torch.manual_seed(0)
    m = LocalCondRNVPDecoder(21, 64, 128).to(dev)
    m.precision = "bf16x3"
    m.train()
    gen = torch.Generator().manual_seed(1)
    p = (torch.rand((32, 3, 2048), generator=gen) - 0.5).to(dev)
    g = torch.randn((32, 128), generator=gen).to(dev).requires_grad_(True)
    for _ in range(3):
        ps, mus, lvs = m(p, g, mode="inverse")
        (0.5 * (lvs.stacked.sum() + (ps.stacked[0] ** 2).sum())).backward()
    torch.cuda.synchronize()
    st = buf.view(3, 512, 16).cpu()
    # the last launches of a step are: ... fwd(l=0) [class 0], then backward P1/P2 of layer L-1 ... 0
    g = st[:, :, 14:16].double()
    print("order check (ns, relative): fwd last start %.0f; P1 last start %.0f end %.0f; P2 last start %.0f end %.0f" % (
        0.0, g[1][g[1][:, 0] > 0][:, 0].min() - g[0][g[0][:, 0] > 0][:, 0].min(), g[1][:, 1].max() - g[0][g[0][:, 0] > 0][:, 0].min(),
        g[2][g[2][:, 0] > 0][:, 0].min() - g[0][g[0][:, 0] > 0][:, 0].min(), g[2][:, 1].max() - g[0][g[0][:, 0] > 0][:, 0].min()))
    # launch-boundary latency: pass 1's last CTA end (globaltimer) -> pass 2's CTAs leaving griddepcontrol.wait
    p1_end = st[1][:, 15][st[1][:, 15] > 0].double().max()
    w = st[2][:, 13][st[2][:, 13] > 0].double()
    p2_end = st[2][:, 15][st[2][:, 15] > 0].double()
    print("boundary P1 -> P2: last P1 CTA end -> pdl_wait return in P2 CTAs: min %.2f us mean %.2f us max %.2f us; "
          "pdl_wait return -> CTA end: mean %.2f us max %.2f us; last P1 end -> last P2 end %.2f us" % (
              (w.min() - p1_end) / 1e3, (w.mean() - p1_end) / 1e3, (w.max() - p1_end) / 1e3,
              (p2_end - w).mean() / 1e3, (p2_end - w).max() / 1e3, (p2_end.max() - p1_end) / 1e3))
    for cls in range(3):
        names = NAMES[cls]
        s = st[cls]
        live = s[:, 0] > 0
        s = s[live].double()
        n = len(names)
        print("== %s: %d CTAs, total %.1f us mean / %.1f us max (per-CTA clock64 at 1.965 GHz)" % (
            KERNEL[cls], s.shape[0], (s[:, n - 1] - s[:, 0]).mean() / 1965, (s[:, n - 1] - s[:, 0]).max() / 1965))
        g0, g1 = st[cls][live][:, 14].double(), st[cls][live][:, 15].double()
        print("   globaltimer: CTA starts spread over %.2f us, first start -> last end %.2f us, CTA lifetimes mean %.2f us" % (
            (g0.max() - g0.min()) / 1e3, (g1.max() - g0.min()) / 1e3, (g1 - g0).mean() / 1e3))
        for i in range(1, n):
            ok = s[:, i] > 0
            d = (s[ok, i] - s[ok, i - 1]) / 1965.0
            if d.numel():
                print("   %-26s mean %7.2f us   max %7.2f us   (%d CTAs)" % (names[i], d.mean(), d.max(), d.numel()))


if __name__ == "__main__":
    main()
