"""Training epoch loop with the reference's signature and behaviour (lib/networks/training.py:10-87):
NaN guard, stdout meters every `num_workers` iterations, checkpoint dict
{'epoch','iter','model_state','optimizer_state'} every 100*num_workers iterations and at epoch end.
Added: gradient averaging across ranks when torch.distributed is initialised (batch sharding; the decoder
arena's all-reduce overlaps the rest of the backward, dist.GradSync), a NaN guard that all ranks agree on, and
rank-averaged BatchNorm buffers at every checkpoint."""
import os
from sys import stdout
from time import time

import torch

from ... import dist as _dist
from .optimizers import optimizer_state_to_reference
from .utils import AverageMeter, save_model


def train(iterator, model, loss_func, optimizer, scheduler, epoch, iter, **kwargs):
    num_workers = max(1, kwargs.get('num_workers') or 1)
    train_mode = kwargs.get('train_mode')
    model_name = os.path.join(kwargs['path2save'], 'models', 'DPFNets', kwargs.get('model_name'))
    rank, _ = _dist.world()
    meters = {k: AverageMeter() for k in ('time', 'LB', 'PNLL', 'GNLL', 'GENT')}
    model.train()
    torch.set_grad_enabled(True)
    dev = next(model.parameters()).device

    def checkpoint(ep, it):
        _dist.average_buffers(model)      # per-rank BatchNorm running statistics -> their mean (all ranks call this)
        if rank == 0:
            os.makedirs(os.path.dirname(model_name), exist_ok=True)
            save_model({'epoch': ep, 'iter': it, 'model_state': model.state_dict(),
                        'optimizer_state': optimizer_state_to_reference(model, optimizer.state_dict())}, model_name)

    def nan_guard(loss):
        # every rank takes the same decision (a rank that stopped alone would leave the others blocked in the
        # next collective): one MAX all-reduce of the NaN flag when distributed
        if _dist.any_rank(torch.isnan(loss.detach()), dev):
            print('Loss is NaN! Stopping without updating the net...')
            raise SystemExit(1)

    sync = getattr(model, '_dpf_grad_sync', None)      # one set of hooks per model, reused across epochs
    if sync is None:
        sync = _dist.GradSync(model)
        object.__setattr__(model, '_dpf_grad_sync', sync)

    graphed = None
    if kwargs.get('cuda_graph'):     # opt-in (not a reference key): the step as two CUDA graphs, _graphstep.py
        from ._graphstep import GraphedTrainStep
        graphed = getattr(model, '_dpf_graphed_step', None)      # one capture per model, reused across epochs
        if graphed is None or graphed.optimizer is not optimizer:
            graphed = GraphedTrainStep(model, loss_func, optimizer, allreduce=sync.finish)
            graphed.nan_guard = nan_guard
            object.__setattr__(model, '_dpf_graphed_step', graphed)

    end = time()
    for i, batch in enumerate(iterator):
        if iter + i >= len(iterator):
            break
        scheduler(optimizer, epoch, iter + i)
        g_clouds = batch['cloud'].to(dev, non_blocking=True)
        p_clouds = batch['eval_cloud'].to(dev, non_blocking=True)
        n = g_clouds.shape[0]
        if graphed is not None:
            inputs = (g_clouds, p_clouds) + ((batch['image'].to(dev, non_blocking=True),) if train_mode == 'p_rnvp_mc_g_rnvp_vae_ic' else ())
            loss, pnll, gnll, gent = graphed(*inputs)
        else:
            if train_mode == 'p_rnvp_mc_g_rnvp_vae_ic':
                outputs = model(g_clouds, p_clouds, batch['image'].to(dev, non_blocking=True))
            else:
                outputs = model(g_clouds, p_clouds)
            loss, pnll, gnll, gent = loss_func(g_clouds, p_clouds, outputs)
            nan_guard(loss)
        meters['PNLL'].update(pnll.item(), n)
        meters['GNLL'].update(gnll.item(), n)
        meters['GENT'].update(gent.item(), n)
        meters['LB'].update((pnll + gnll - gent).item(), n)
        if graphed is None:
            optimizer.zero_grad()
            loss.backward()
            sync.finish()
            optimizer.step()
        meters['time'].update(time() - end)
        if rank == 0 and (iter + i + 1) % num_workers == 0:
            stdout.write('Epoch: [{0}][{1}/{2}]\tTime {t.val:.3f} ({t.avg:.3f})\tLB {lb.val:.2f} ({lb.avg:.2f})'
                         '\tPNLL {p.val:.2f} ({p.avg:.2f})\tGNLL {g.val:.2f} ({g.avg:.2f})\tGENT {e.val:.2f} ({e.avg:.2f})\n'
                         .format(epoch + 1, iter + i + 1, len(iterator), t=meters['time'], lb=meters['LB'],
                                 p=meters['PNLL'], g=meters['GNLL'], e=meters['GENT']))
            stdout.flush()
        end = time()
        if (iter + i + 1) % (100 * num_workers) == 0:
            checkpoint(epoch, iter + i + 1)
    checkpoint(epoch + 1, 0)
    return meters
