"""PointFlow-style metric suite with the reference's function names
(lib/metrics/evaluation_metrics.py:22-200): CD and approximate EMD all-pairs matrices come from the
fused kernels (one launch per matrix, dense EMD match never materialised)."""
import torch

from ...ops import pairwise_cd, pairwise_emd
from .StructuralLosses.match_cost import match_cost
from .StructuralLosses.nn_distance import nn_distance


def distChamferCUDA(x, y):
    return nn_distance(x, y)


def emd_approx(sample, ref):
    B, N, N_ref = sample.size(0), sample.size(1), ref.size(1)
    assert N == N_ref, "Not sure what would EMD do in this case"
    return match_cost(sample, ref) / float(N)


def distChamfer(a, b):
    """Pure-torch bmm form kept for signature parity (evaluation_metrics.py:35-45); returns
    (min over a per b-point, min over b per a-point) like the reference."""
    xx = (a * a).sum(-1)
    yy = (b * b).sum(-1)
    P = xx.unsqueeze(2) + yy.unsqueeze(1) - 2 * torch.bmm(a, b.transpose(2, 1))
    return P.min(1)[0], P.min(2)[0]


def EMD_CD(sample_pcs, ref_pcs, batch_size, accelerated_cd=False, reduced=True):
    assert sample_pcs.shape[0] == ref_pcs.shape[0]
    cd_lst = []
    for s in range(0, sample_pcs.shape[0], batch_size):
        a, b = sample_pcs[s:s + batch_size].contiguous(), ref_pcs[s:s + batch_size].contiguous()
        dl, dr = distChamferCUDA(a, b) if accelerated_cd else distChamfer(a, b)
        cd_lst.append(dl.mean(dim=1) + dr.mean(dim=1))
    cd = torch.cat(cd_lst)
    return {'MMD-CD': cd.mean() if reduced else cd}


def _pairwise_EMD_CD_(sample_pcs, ref_pcs, batch_size, accelerated_cd=True):
    """-> (all_cd, all_emd), each (N_sample, N_ref); batch_size is accepted and irrelevant."""
    sample_pcs, ref_pcs = sample_pcs.contiguous(), ref_pcs.contiguous()
    all_cd = pairwise_cd(sample_pcs, ref_pcs)
    all_emd = pairwise_emd(sample_pcs, ref_pcs) / float(sample_pcs.shape[1])
    return all_cd, all_emd


def knn(Mxx, Mxy, Myy, k, sqrt=False):
    n0, n1 = Mxx.size(0), Myy.size(0)
    label = torch.cat((torch.ones(n0), torch.zeros(n1))).to(Mxx)
    M = torch.cat((torch.cat((Mxx, Mxy), 1), torch.cat((Mxy.t(), Myy), 1)), 0)
    if sqrt:
        M = M.abs().sqrt()
    M = M + torch.diag(torch.full((n0 + n1,), float('inf'), device=M.device, dtype=M.dtype))
    _, idx = M.topk(k, 0, False)
    count = torch.zeros(n0 + n1).to(Mxx)
    for i in range(k):
        count = count + label.index_select(0, idx[i])
    pred = torch.ge(count, (float(k) / 2) * torch.ones(n0 + n1).to(Mxx)).float()
    s = {'tp': (pred * label).sum(), 'fp': (pred * (1 - label)).sum(),
         'fn': ((1 - pred) * label).sum(), 'tn': ((1 - pred) * (1 - label)).sum()}
    s.update({'precision': s['tp'] / (s['tp'] + s['fp'] + 1e-10), 'recall': s['tp'] / (s['tp'] + s['fn'] + 1e-10),
              'acc_t': s['tp'] / (s['tp'] + s['fn'] + 1e-10), 'acc_f': s['tn'] / (s['tn'] + s['fp'] + 1e-10),
              'acc': torch.eq(label, pred).float().mean()})
    return s


def lgan_mmd_cov(all_dist):
    N_ref = all_dist.size(1)
    min_val_fromsmp, min_idx = torch.min(all_dist, dim=1)
    min_val, _ = torch.min(all_dist, dim=0)
    cov = torch.tensor(float(min_idx.unique().view(-1).size(0)) / float(N_ref)).to(all_dist)
    return {'lgan_mmd': min_val.mean(), 'lgan_cov': cov, 'lgan_mmd_smp': min_val_fromsmp.mean()}


def compute_all_metrics(sample_pcs, ref_pcs, batch_size, accelerated_cd=False):
    results = {}
    M_rs_cd, M_rs_emd = _pairwise_EMD_CD_(ref_pcs, sample_pcs, batch_size, accelerated_cd=accelerated_cd)
    results.update({"%s-CD" % k: v for k, v in lgan_mmd_cov(M_rs_cd.t()).items()})
    results.update({"%s-EMD" % k: v for k, v in lgan_mmd_cov(M_rs_emd.t()).items()})
    M_rr_cd, M_rr_emd = _pairwise_EMD_CD_(ref_pcs, ref_pcs, batch_size, accelerated_cd=accelerated_cd)
    M_ss_cd, M_ss_emd = _pairwise_EMD_CD_(sample_pcs, sample_pcs, batch_size, accelerated_cd=accelerated_cd)
    results.update({"1-NN-CD-%s" % k: v for k, v in knn(M_rr_cd, M_rs_cd, M_ss_cd, 1).items() if 'acc' in k})
    results.update({"1-NN-EMD-%s" % k: v for k, v in knn(M_rr_emd, M_rs_emd, M_ss_emd, 1).items() if 'acc' in k})
    return results
