"""Metric helpers with the reference's names and return conventions (lib/networks/utils.py:8-144).
`pairwise_CD` is ONE fused kernel launch per matrix (row-sharded across ranks when
torch.distributed is initialised) instead of the reference's Python loop over clouds."""
import numpy as np
import torch

from ... import dist as _dist
from ...ops import pairwise_cd as _pairwise_cd_op
from ..metrics.StructuralLosses.nn_distance import nn_distance


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def save_model(state, model_name):
    torch.save(state, model_name, pickle_protocol=4)
    print('Model saved to ' + model_name)


def cnt_params(params):
    return sum(p.numel() for p in params if p.requires_grad)


def distChamferCUDA(x, y):
    return nn_distance(x, y)


def f_score(predicted_clouds, true_clouds, threshold=0.001):
    """F1 from the squared NN distances (utils.py:38-42)."""
    ld, rd = distChamferCUDA(predicted_clouds, true_clouds)
    precision = 100. * (rd < threshold).float().mean(1)
    recall = 100. * (ld < threshold).float().mean(1)
    return 2. * precision * recall / (precision + recall + 1e-7)


def get_voxel_occ_dist(all_clouds, clouds_flag='gen', res=28, bound=0.5, bs=128, warning=True):
    """Occupancy histogram of points over a res^3 grid on [-0.5, 0.5)^3 (utils.py:45-79); points
    outside the cube are dropped.  Vectorised restatement (one np.add.at over all clouds)."""
    all_clouds = np.asarray(all_clouds)
    if warning and np.any(np.fabs(all_clouds) > bound):
        print('{} clouds out of cube bounds: [-{}; {}]'.format(clouds_flag, bound, bound))
    n_nans = int(np.isnan(all_clouds).sum())
    if n_nans > 0:
        print('{} NaN values in point cloud tensors.'.format(n_nans))
    edges = -0.5 + np.arange(res + 1) * (1. / res)
    pts = all_clouds.reshape(-1, 3)
    idx = np.empty(pts.shape, dtype=np.int64)
    ok = np.ones(pts.shape[0], dtype=bool)
    for c in range(3):
        inside = np.logical_and(edges[:res, None] <= pts[None, :, c], pts[None, :, c] < edges[1:, None])
        idx[:, c] = inside.argmax(0)
        ok &= inside.any(0)
    hist = np.zeros((res, res, res), dtype=np.uint64)
    np.add.at(hist, (idx[ok, 0], idx[ok, 1], idx[ok, 2]), np.uint64(1))
    return np.float64(hist) / hist.sum()


def _entropy2(p):
    p = p[p > 0]
    return float(-(p * np.log2(p)).sum())


def JSD(clouds1, clouds2, clouds1_flag='gen', clouds2_flag='ref', warning=True):
    d1 = get_voxel_occ_dist(clouds1, clouds_flag=clouds1_flag, warning=warning).flatten()
    d2 = get_voxel_occ_dist(clouds2, clouds_flag=clouds2_flag, warning=warning).flatten()
    return _entropy2((d1 + d2) / 2.0) - 0.5 * (_entropy2(d1) + _entropy2(d2))


def pairwise_CD(clouds1, clouds2, bs=2048):
    """(N1,n,3),(N2,m,3) -> (N1,N2) matrix of dl.mean(1)+dr.mean(1) (utils.py:90-117).  `bs` is
    accepted for signature compatibility; there is no chunking to do."""
    symmetric = clouds1 is clouds2 or (clouds1.data_ptr() == clouds2.data_ptr() and clouds1.shape == clouds2.shape)

    def rows(out, row_start, row_step, n_rows, sym):
        _pairwise_cd_op(clouds1, clouds2, out=out, row_start=row_start, row_step=row_step, n_rows=n_rows, symmetric=sym)
    return _dist.sharded_pairwise(rows, clouds1.shape[0], clouds2.shape[0], clouds1.device, symmetric=symmetric)


def COV(dists, axis=1):
    return float(dists.min(axis)[1].unique().shape[0]) / float(dists.shape[axis])


def MMD(dists, axis=1):
    return float(dists.min((axis + 1) % 2)[0].mean().float())


def KNN(Mxx, Mxy, Myy, k, sqrt=False):
    """Leave-one-out k-NN accuracy on the (n0+n1)^2 block matrix (utils.py:128-144)."""
    n0, n1 = Mxx.size(0), Myy.size(0)
    label = torch.cat((-torch.ones(n0), torch.ones(n1))).to(Mxx)
    M = torch.cat((torch.cat((Mxx, Mxy), 1), torch.cat((Mxy.t(), Myy), 1)), 0)
    if sqrt:
        M = M.abs().sqrt()
    M = M + torch.diag(torch.full((n0 + n1,), float('inf'), device=M.device, dtype=M.dtype))
    _, idx = M.topk(k, 0, False)
    count = torch.zeros(n0 + n1).to(Mxx)
    for i in range(k):
        count = count + label.index_select(0, idx[i])
    pred = torch.where(count >= 0, torch.ones_like(count), -torch.ones_like(count))
    return float(torch.eq(label, pred).float().mean())
