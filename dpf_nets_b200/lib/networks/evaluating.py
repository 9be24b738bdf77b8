"""Evaluation loop with the reference's signature, modes and printed metrics
(lib/networks/evaluating.py:13-321).  'generating' accumulates generated / reference clouds and runs
the all-pairs sweep (three fused Chamfer matrices, row-sharded across ranks) -> JSD, COV, MMD, 1-NNA;
'evaluating' / 'predicting' report per-batch Chamfer (and F1).  Returns the metrics as a dict too."""
from time import time

import numpy as np
import torch

from .utils import AverageMeter, COV, JSD, KNN, MMD, distChamferCUDA, f_score, pairwise_CD


def _denorm(clouds, batch, dev, kwargs):
    """Back to the original scale when the dataset provides it (evaluating.py:118-133)."""
    if not kwargs.get('orig_scale_evaluation') or 'orig_s' not in batch:
        return clouds
    s = batch['orig_s'].to(dev).view(-1, 1, 1)
    c = batch['orig_c'].to(dev).view(-1, 3, 1)
    if kwargs.get('cloud_scale'):
        clouds = clouds / kwargs.get('cloud_scale_scale', 1.0)
    return clouds * s + c


def evaluate(iterator, model, loss_func, **kwargs):
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
        r_clouds = _denorm(outputs['p_prior_samples'][-1], batch, dev, kwargs)
        gt = _denorm(p_clouds, batch, dev, kwargs)
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
    res = {'inference_sec_per_sample': inf_time.avg}
    print('Inference time: {} sec/sample'.format(inf_time.avg))
    if util_mode == 'training':
        res.update({k: m.avg for k, m in meters.items()})
        print('LB: {:.2f} PNLL: {:.2f} GNLL: {:.2f} GENT: {:.2f}'.format(res['LB'], res['PNLL'], res['GNLL'], res['GENT']))
    elif util_mode in ('evaluating', 'predicting'):
        res['CD'] = CD.avg
        print('CD: {:.6f}'.format(CD.avg))
        if util_mode == 'predicting':
            res['F1'] = F1.avg
            print('F1: {:.1f}'.format(F1.avg))
    elif util_mode == 'generating':
        gen = torch.cat(gen_buf, 0).transpose(2, 1).contiguous()
        ref = torch.cat(ref_buf, 0).transpose(2, 1).contiguous()
        bad = torch.isnan(gen).flatten(1).any(1)            # NaN clouds -> random valid duplicates (evaluating.py:237-243)
        if bad.any():
            good = (~bad).nonzero().flatten()
            pick = good[torch.randint(len(good), (int(bad.sum()),), device=good.device)]
            gen[bad] = gen[pick]
        res.update(generation_metrics(gen, ref))
        print('JSD:   \t{:.2f}'.format(1e2 * res['JSD']))
        print('COV-CD:\t{:.1f}'.format(1e2 * res['COV-CD']))
        print('MMD-CD:\t{:.2f}'.format(1e4 * res['MMD-CD']))
        print('1NN-CD:\t{:.1f}'.format(1e2 * res['1NN-CD']))
    return res


def generation_metrics(gen, ref):
    """gen, ref (S,N,3) on the GPU -> JSD / COV / MMD / 1-NNA from three fused all-pairs CD matrices."""
    gg = pairwise_CD(gen, gen)
    tt = pairwise_CD(ref, ref)
    gt = pairwise_CD(gen, ref)
    return {'JSD': JSD(gen.cpu().numpy(), ref.cpu().numpy(), clouds1_flag='gen', clouds2_flag='ref', warning=False),
            'COV-CD': COV(gt), 'MMD-CD': MMD(gt), '1NN-CD': KNN(gg, gt, tt, 1)}
