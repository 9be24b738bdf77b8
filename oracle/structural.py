"""TEST INFRASTRUCTURE ONLY - Python handles on the two Chamfer/EMD checkers:

* `liboracle.so`  (oracle/structural_oracle.c, CPU, numpy in/out)   -> functions `nndistance`, ...
* `libref_structural.so` (the reference's own .cu files built by oracle/build_ref.py, GPU,
  torch CUDA tensors in/out) -> class `RefCuda`.

Plus a numpy brute force (`brute_nn`) used to cross-check the C restatement itself.
"""
import ctypes
import os

import numpy as np

from .build_oracle import build_oracle, OUT as _ORACLE_SO
from .build_ref import OUT as REF_SO

_c = None
_F = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _lib():
    global _c
    if _c is None:
        _c = ctypes.CDLL(build_oracle())
        ci = ctypes.c_int
        _c.oracle_nndistance.argtypes = [ci, ci, _F, ci, _F, _F, _I, _F, _I]
        _c.oracle_nndistance_grad.argtypes = [ci, ci, _F, ci, _F, _F, _I, _F, _I, _F, _F]
        _c.oracle_pairwise_cd.argtypes = [ci, ci, ci, ci, _F, _F, _F]
        _c.oracle_approxmatch.argtypes = [ci, ci, ci, _F, _F, _F, _F]
        _c.oracle_matchcost.argtypes = [ci, ci, ci, _F, _F, _F, _F]
        _c.oracle_matchcost_grad.argtypes = [ci, ci, ci, _F, _F, _F, _F, _F]
        for f in ("oracle_nndistance", "oracle_nndistance_grad", "oracle_pairwise_cd", "oracle_approxmatch",
                  "oracle_matchcost", "oracle_matchcost_grad"):
            getattr(_c, f).restype = None
    return _c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def nndistance(xyz, xyz2):
    """(b,n,3),(b,m,3) -> dist1 (b,n), idx1, dist2 (b,m), idx2 - restates nndistance.cu:2-128."""
    xyz, xyz2 = _f32(xyz), _f32(xyz2)
    b, n, m = xyz.shape[0], xyz.shape[1], xyz2.shape[1]
    d1 = np.zeros((b, n), np.float32); i1 = np.zeros((b, n), np.int32)
    d2 = np.zeros((b, m), np.float32); i2 = np.zeros((b, m), np.int32)
    _lib().oracle_nndistance(b, n, xyz, m, xyz2, d1, i1, d2, i2)
    return d1, i1, d2, i2


def nndistance_grad(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, xyz2, g1, g2 = _f32(xyz1), _f32(xyz2), _f32(g1), _f32(g2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    o1 = np.zeros((b, n, 3), np.float32); o2 = np.zeros((b, m, 3), np.float32)
    _lib().oracle_nndistance_grad(b, n, xyz1, m, xyz2, g1, np.ascontiguousarray(idx1, np.int32), g2,
                                  np.ascontiguousarray(idx2, np.int32), o1, o2)
    return o1, o2


def pairwise_cd(A, B):
    """(S1,n,3),(S2,m,3) -> (S1,S2) - restates lib/networks/utils.py:90-117."""
    A, B = _f32(A), _f32(B)
    out = np.zeros((A.shape[0], B.shape[0]), np.float32)
    _lib().oracle_pairwise_cd(A.shape[0], B.shape[0], A.shape[1], B.shape[1], A, B, out)
    return out


def approxmatch(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = np.zeros((b, m, n), np.float32)
    temp = np.zeros((max(b, 32), (n + m) * 2), np.float32)
    _lib().oracle_approxmatch(b, n, m, xyz1, xyz2, match, temp)
    return match


def matchcost(xyz1, xyz2, match):
    xyz1, xyz2, match = _f32(xyz1), _f32(xyz2), _f32(match)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    out = np.zeros((b,), np.float32)
    _lib().oracle_matchcost(b, n, m, xyz1, xyz2, match, out)
    return out


def matchcost_grad(xyz1, xyz2, match):
    xyz1, xyz2, match = _f32(xyz1), _f32(xyz2), _f32(match)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    g1 = np.zeros((b, n, 3), np.float32); g2 = np.zeros((b, m, 3), np.float32)
    _lib().oracle_matchcost_grad(b, n, m, xyz1, xyz2, match, g1, g2)
    return g1, g2


def brute_nn(x, y):
    """Independent numpy check: direct-difference squared distances in float64, argmin lowest index."""
    x = np.asarray(x, np.float32).astype(np.float64)
    y = np.asarray(y, np.float32).astype(np.float64)
    d = ((y[:, None, :, :] - x[:, :, None, :]) ** 2).sum(-1)  # (b, n, m)
    return d.min(2), d.argmin(2), d.min(1), d.argmin(1)


class RefCuda:
    """The reference's own CUDA launchers (mangled C++ names) from oracle/_ref/libref_structural.so."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " missing - run `python oracle/build_ref.py` in the dev container")
        self.lib = ctypes.CDLL(REF_SO)
        self.nnd = getattr(self.lib, "_Z10nndistanceiiPKfiS0_PfPiS1_S2_P11CUstream_st")
        self.nndg = getattr(self.lib, "_Z14nndistancegradiiPKfiS0_S0_PKiS0_S2_PfS3_P11CUstream_st")
        self.am = getattr(self.lib, "_Z11approxmatchiiiPKfS0_PfS1_P11CUstream_st")
        self.mc = getattr(self.lib, "_Z9matchcostiiiPKfS0_PfS1_P11CUstream_st")
        self.mcg = getattr(self.lib, "_Z13matchcostgradiiiPKfS0_S0_PfS1_P11CUstream_st")
        for f in (self.nnd, self.nndg, self.am, self.mc, self.mcg):
            f.restype = None

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def _s():
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def nndistance(self, a, b):
        import torch
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        d1 = torch.empty((B, n), device=a.device); i1 = torch.empty((B, n), dtype=torch.int32, device=a.device)
        d2 = torch.empty((B, m), device=a.device); i2 = torch.empty((B, m), dtype=torch.int32, device=a.device)
        self.nnd(B, n, self._p(a), m, self._p(b), self._p(d1), self._p(i1), self._p(d2), self._p(i2), self._s())
        return d1, i1, d2, i2

    def approxmatch(self, a, b):
        import torch
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        match = torch.empty((B, m, n), device=a.device)
        temp = torch.empty((B, (n + m) * 2), device=a.device)
        self.am(B, n, m, self._p(a), self._p(b), self._p(match), self._p(temp), self._s())
        return match

    def matchcost(self, a, b, match):
        import torch
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        out = torch.empty((B,), device=a.device)
        self.mc(B, n, m, self._p(a), self._p(b), self._p(match), self._p(out), self._s())
        return out

    def matchcost_grad(self, a, b, match):
        import torch
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        g1 = torch.empty((B, n, 3), device=a.device); g2 = torch.empty((B, m, 3), device=a.device)
        self.mcg(B, n, m, self._p(a), self._p(b), self._p(match), self._p(g1), self._p(g2), self._s())
        return g1, g2

    def pairwise_cd(self, c1, c2):
        """The reference's Python loop (lib/networks/utils.py:90-117) over its own kernel."""
        import torch
        N1, N2 = c1.shape[0], c2.shape[0]
        cds = torch.zeros((N1, N2), device=c1.device)
        for i in range(N1):
            ci = c1[i].unsqueeze(0).expand(N2, -1, -1).contiguous()
            dl, _, dr, _ = self.nndistance(ci, c2)
            cds[i] = dl.mean(dim=1) + dr.mean(dim=1)
        return cds
