"""GPU probe of the tcgen05 operand layouts / descriptors through dpf_umma_selftest.
Prints one JSON line per configuration: {"name":..., "max_err":..., "ok":...}."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200 import _lib  # noqa: E402

SW128 = 2


def desc_template(lbo, sbo, layout=SW128):
    return (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def idesc_bf16(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def sw128_tile(x):
    """x: (rows, 64) bf16 tensor -> uint8 image of the 128B-swizzled tile (rows*128 bytes)."""
    rows = x.shape[0]
    src = x.contiguous().view(torch.int16).numpy().view(np.uint8).reshape(rows, 8, 16)
    out = np.zeros((rows, 8, 16), np.uint8)
    for r in range(rows):
        for q in range(8):
            out[r, q ^ (r & 7)] = src[r, q]
    return out.reshape(-1)


def run(name, a_img, b_img, a_t, b_t, idesc, num_k, a_ks, b_ks, ncols, want, use_bulk=0):
    dev = torch.device("cuda:0")
    a = torch.from_numpy(a_img.copy()).to(dev)
    b = torch.from_numpy(b_img.copy()).to(dev)
    d = torch.zeros((128, ncols), device=dev)
    lib = _lib.lib()
    rc = lib.dpf_umma_selftest(ctypes.c_void_p(a.data_ptr()), a.numel(), ctypes.c_void_p(b.data_ptr()), b.numel(),
                               ctypes.c_ulonglong(a_t), ctypes.c_ulonglong(b_t), ctypes.c_uint(idesc), num_k, a_ks, b_ks,
                               ncols, use_bulk, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(0))
    torch.cuda.synchronize()
    err = (d.cpu() - want).abs().max().item()
    print(json.dumps({"name": name, "rc": rc, "max_err": err, "ref_max": want.abs().max().item(), "ok": bool(err < 2e-2)}), flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    g = torch.Generator().manual_seed(0)
    if which in ("all", "kmajor", "kmajor_bulk"):
        A = torch.randn((128, 64), generator=g).to(torch.bfloat16)
        B = torch.randn((64, 64), generator=g).to(torch.bfloat16)
        want = A.float() @ B.float().t()
        t = desc_template(16, 1024)
        if which != "kmajor_bulk":
            run("kmajor_m128n64k64", sw128_tile(A), sw128_tile(B), t, t, idesc_bf16(128, 64, 0, 0), 4, 32, 32, 64, want)
        if which != "kmajor":
            run("kmajor_m128n64k64_bulk", sw128_tile(A), sw128_tile(B), t, t, idesc_bf16(128, 64, 0, 0), 4, 32, 32, 64, want, 1)
    if which in ("all", "mnmajor", "mnmajor_swapped"):
        # A_all[p][m], B_all[p][n]: p = 128 points (K), m / n = 128 (two 64-wide tiles each)
        Aall = torch.randn((128, 128), generator=g).to(torch.bfloat16)
        Ball = torch.randn((128, 128), generator=g).to(torch.bfloat16)
        want = Aall.float().t() @ Ball.float()
        a_img = np.concatenate([sw128_tile(Aall[:, :64]), sw128_tile(Aall[:, 64:])])
        b_img = np.concatenate([sw128_tile(Ball[:, :64]), sw128_tile(Ball[:, 64:])])
        if which != "mnmajor_swapped":
            t = desc_template(16384, 1024)
            run("mnmajor_m128n128k128_lbo16384_sbo1024", a_img, b_img, t, t, idesc_bf16(128, 128, 1, 1), 8, 2048, 2048, 128, want)
        if which != "mnmajor":
            t = desc_template(1024, 16384)
            run("mnmajor_m128n128k128_lbo1024_sbo16384", a_img, b_img, t, t, idesc_bf16(128, 128, 1, 1), 8, 2048, 2048, 128, want)


if __name__ == "__main__":
    main()
