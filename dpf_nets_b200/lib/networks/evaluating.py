"""Evaluation loop with the reference's signature, modes and printed metrics
(lib/networks/evaluating.py:13-321).  'generating' accumulates generated / reference clouds and runs
the all-pairs sweep (three fused Chamfer matrices, row-sharded across ranks) -> JSD, COV, MMD, 1-NNA;
'evaluating' / 'predicting' report per-batch Chamfer (and F1).  Returns the metrics as a dict too."""
from time import time

import numpy as np
import torch

from ... import dist as _dist
from .utils import AverageMeter, distChamferCUDA, f_score, pairwise_CD


def _to_eval_scale(r_clouds, p_clouds, batch, dev, kwargs, util_mode):
    """The reference's in-place rescaling of generated and ground-truth clouds before the metrics
    (evaluating.py:88-103 'evaluating', :118-133 'generating', :147-189 'predicting'):

    * orig_scale_evaluation: undo ScaleCloud (`*= cloud_scale_scale`), undo the constant translation
      (`+= cloud_translate_shift`), and - only when the dataset did NOT already hand out original-scale
      clouds (cloud_rescale2orig / cloud_recenter2orig False) - apply the per-shape `orig_s` / `orig_c`;
    * predicting only: unit_scale_evaluation (`*= cloud_scale_scale`) and bbox_scale_evaluation
      (back to the mesh's bounding-box frame with `bbox_c` / `bbox_s`).
    Both clouds get the identical affine map.  (B,3,N) tensors; returns new tensors."""
    k = kwargs.get

    def per_shape(name, shape):
        if name not in batch:
            raise KeyError("'%s' is missing from the batch: the dataset must be built with "
                           "return_original_scale / return_bbox_scale for this evaluation mode" % name)
        return batch[name].to(dev).float().reshape(shape)

    def undo_scale_and_shift(r, p):
        if k('cloud_scale'):
            sc = float(k('cloud_scale_scale'))
            r, p = r * sc, p * sc
        if k('cloud_translate'):
            shift = torch.as_tensor(np.array(k('cloud_translate_shift'), dtype=np.float32).reshape(1, -1, 1), device=dev)
            r, p = r + shift, p + shift
        return r, p

    r, p = r_clouds, p_clouds
    if util_mode == 'predicting' and k('unit_scale_evaluation'):
        if k('cloud_scale'):
            sc = float(k('cloud_scale_scale'))
            r, p = r * sc, p * sc
    if k('orig_scale_evaluation'):
        r, p = undo_scale_and_shift(r, p)
        if not k('cloud_rescale2orig'):
            s = per_shape('orig_s', (-1, 1, 1))
            r, p = r * s, p * s
        if not k('cloud_recenter2orig'):
            c = per_shape('orig_c', (-1, 3, 1))
            r, p = r + c, p + c
    if util_mode == 'predicting' and k('bbox_scale_evaluation'):
        r, p = undo_scale_and_shift(r, p)
        if k('cloud_recenter2orig'):
            c = per_shape('orig_c', (-1, 3, 1))
            r, p = r - c, p - c
        if k('cloud_rescale2orig'):
            s = per_shape('orig_s', (-1, 1, 1))
            r, p = r / s, p / s
        bc, bs = per_shape('bbox_c', (-1, 3, 1)), per_shape('bbox_s', (-1, 1, 1))
        r, p = (r - bc) / bs, (p - bc) / bs
    return r, p


def evaluate(iterator, model, loss_func, **kwargs):
    """Reference signature (evaluating.py:13).  Gradient recording is switched off for the evaluation like in the
    reference (:59) and - unlike it - switched back to its previous state on return."""
    grad_was = torch.is_grad_enabled()
    try:
        return _evaluate(iterator, model, loss_func, **kwargs)
    finally:
        torch.set_grad_enabled(grad_was)


def _evaluate(iterator, model, loss_func, **kwargs):
    train_mode, util_mode = kwargs.get('train_mode'), kwargs.get('util_mode')
    if kwargs.get('saving'):
        try:
            import h5py  # noqa: F401
        except ImportError:
            raise RuntimeError("saving=True needs h5py (not installed here); run with saving disabled")
    model.eval()
    torch.set_grad_enabled(False)
    dev = next(model.parameters()).device
    inf_time, CD, F1 = AverageMeter(), AverageMeter(), AverageMeter()
    meters = {k: AverageMeter() for k in ('LB', 'PNLL', 'GNLL', 'GENT')}
    gen_buf, ref_buf = [], []
    n_sampled = kwargs.get('sampled_cloud_size') or kwargs.get('cloud_size')
    for batch in iterator:
        g_clouds = batch['cloud'].to(dev, non_blocking=True)
        p_clouds = batch['eval_cloud'].to(dev, non_blocking=True)
        torch.cuda.synchronize(dev) if dev.type == 'cuda' else None
        t0 = time()
        if train_mode == 'p_rnvp_mc_g_rnvp_vae_ic':
            outputs = model(g_clouds, p_clouds, batch['image'].to(dev, non_blocking=True), n_sampled_points=n_sampled)
        else:
            outputs = model(g_clouds, p_clouds, n_sampled_points=n_sampled)
        torch.cuda.synchronize(dev) if dev.type == 'cuda' else None
        inf_time.update((time() - t0) / g_clouds.shape[0], g_clouds.shape[0])
        if util_mode == 'training':
            loss, pnll, gnll, gent = loss_func(g_clouds, p_clouds, outputs)
            for k, v in (('PNLL', pnll), ('GNLL', gnll), ('GENT', gent), ('LB', pnll + gnll - gent)):
                meters[k].update(v.item(), g_clouds.shape[0])
            continue
        r_clouds, gt = _to_eval_scale(outputs['p_prior_samples'][-1], p_clouds, batch, dev, kwargs, util_mode)
        if util_mode == 'generating':
            gen_buf.append(r_clouds)
            ref_buf.append(gt)
            continue
        a = r_clouds.transpose(2, 1).contiguous()
        b = gt.transpose(2, 1).contiguous()
        dl, dr = distChamferCUDA(a, b)
        CD.update((dl.mean(1) + dr.mean(1)).mean().item(), a.shape[0])
        if util_mode == 'predicting':
            F1.update(f_score(a, b).mean().item(), a.shape[0])
    # multi-rank: every rank iterated its own contiguous shard of the dataset (entry._loader -> dist.ShardSampler);
    # meters are combined as (sum, count) and the cloud sets are gathered in rank order = dataset order
    def avg(m):
        sm, cnt = _dist.reduce_sums([m.sum, m.count], dev)
        return sm / max(cnt, 1)
    res = {'inference_sec_per_sample': avg(inf_time)}
    print('Inference time: {} sec/sample'.format(res['inference_sec_per_sample']))
    if util_mode == 'training':
        res.update({k: avg(m) for k, m in meters.items()})
        print('LB: {:.2f} PNLL: {:.2f} GNLL: {:.2f} GENT: {:.2f}'.format(res['LB'], res['PNLL'], res['GNLL'], res['GENT']))
    elif util_mode in ('evaluating', 'predicting'):
        res['CD'] = avg(CD)
        print('CD: {:.6f}'.format(res['CD']))
        if util_mode == 'predicting':
            res['F1'] = avg(F1)
            print('F1: {:.1f}'.format(res['F1']))
    elif util_mode == 'generating':
        # (a rank whose shard is empty still takes part in the gather)
        gen = torch.cat(gen_buf, 0) if gen_buf else torch.zeros((0, 3, n_sampled or 0), device=dev)
        ref = torch.cat(ref_buf, 0) if ref_buf else torch.zeros((0, 3, kwargs.get('cloud_size') or 0), device=dev)
        gen = _dist.gather_rows(gen).transpose(2, 1).contiguous()
        ref = _dist.gather_rows(ref).transpose(2, 1).contiguous()
        bad = torch.isnan(gen).flatten(1).any(1)            # NaN clouds -> random valid duplicates (evaluating.py:237-243)
        if bad.any():
            good = (~bad).nonzero().flatten()
            # CPU generator with a fixed seed: every rank substitutes the same duplicates
            pick = good[torch.randint(len(good), (int(bad.sum()),), generator=torch.Generator().manual_seed(0)).to(good.device)]
            gen[bad] = gen[pick]
        res.update(generation_metrics(gen, ref))
        print('JSD:   \t{:.2f}'.format(1e2 * res['JSD']))
        print('COV-CD:\t{:.1f}'.format(1e2 * res['COV-CD']))
        print('MMD-CD:\t{:.2f}'.format(1e4 * res['MMD-CD']))
        print('1NN-CD:\t{:.1f}'.format(1e2 * res['1NN-CD']))
    return res


def _scores(gg, gt, tt):
    """[COV, MMD, 1-NN accuracy] as one device tensor from the device-resident matrices (ops/metrics.py)."""
    from ...ops.metrics import cd_scores
    return cd_scores(gg, gt, tt)


def _jsd(gen, ref):
    """JSD as a 0-dim device tensor from GPU voxel histograms (ops/metrics.py)."""
    from ...ops.metrics import jsd_from_hists, voxel_hist
    return jsd_from_hists(voxel_hist(gen), voxel_hist(ref))


def generation_metrics(gen, ref):
    """gen, ref (S,N,3) on the GPU (identical on every rank) -> JSD / COV / MMD / 1-NNA
    (evaluating.py:245-253): three fused all-pairs CD matrices (row-sharded across ranks, one collective
    each), then ONE reduction launch for COV / MMD / 1-NNA and a GPU voxel histogram for JSD; the only
    device->host traffic is the four resulting numbers."""
    gg = pairwise_CD(gen, gen)
    tt = pairwise_CD(ref, ref)
    gt = pairwise_CD(gen, ref)
    vals = torch.cat([_jsd(gen, ref).reshape(1).double(), _scores(gg, gt, tt).double()]).tolist()
    return {'JSD': vals[0], 'COV-CD': vals[1], 'MMD-CD': vals[2], '1NN-CD': vals[3]}
