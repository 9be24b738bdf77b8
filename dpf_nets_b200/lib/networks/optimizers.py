"""The reference's custom AMSGrad Adam and cosine LR / beta2 updater (lib/networks/optimizers.py:8-97),
same state keys ('step', 'exp_avg', 'exp_avg_sq', 'max_exp_avg_sq') and the same non-standard
update:  denom = sqrt(v_hat) / sqrt(1 - beta2^t) + eps ;  p -= wd * p + lr * (m / (1 - beta1^t)) / denom
(weight decay NOT scaled by lr).  CUDA tensors go through the multi-tensor kernel (dpf_adam_step_multi, 48 tensors per launch; the
decoder is a single arena tensor), so the per-parameter Python loop of the reference collapses to a
handful of launches."""
import math

import numpy as np
import torch
from torch.optim import Optimizer

from ... import _lib


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        # capturable: the fused launches read {lr, betas, eps, wd, bias corrections} from device memory that a
        # captured H2D copy refreshes from a pinned host table, so step() can live in a CUDA graph (_graphstep.py)
        self.capturable = False
        self._hyper_host = None
        self._hyper_dev = None

    def _hyper_row(self, group, step):
        beta1, beta2 = group['betas']
        return [float(group['lr']), float(beta1), float(beta2), float(group['eps']), float(group['weight_decay']),
                float(1 - beta1 ** step), float(math.sqrt(1 - beta2 ** step)), 0.0]

    def _group_step(self, group):
        steps = {self.state[p]['step'] for p in group['params'] if p.grad is not None and len(self.state[p])}
        if len(steps) != 1:
            raise RuntimeError('capturable Adam needs one shared step count per parameter group, got %s' % sorted(steps))
        return steps.pop()

    def enable_capture(self, device):
        """Allocate the hyper tables (pinned host + device) OUTSIDE any stream capture and switch step() to the
        graph-replayable launches."""
        if self._hyper_host is None:
            self._hyper_host = torch.zeros((len(self.param_groups), 8), dtype=torch.float32).pin_memory()
            self._hyper_dev = torch.zeros((len(self.param_groups), 8), dtype=torch.float32, device=device)
        self.capturable = True

    def prepare_replay(self):
        """Host side of one replayed step: advance the step counters and refresh the pinned hyper table (the
        captured graph copies it to the device before the update kernels run)."""
        for gi, group in enumerate(self.param_groups):
            for p in group['params']:
                if p.grad is not None and len(self.state[p]):
                    self.state[p]['step'] += 1
            self._hyper_host[gi] = torch.tensor(self._hyper_row(group, self._group_step(group)))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            fused = {}      # (device, step) -> parameters that go into one multi-tensor launch
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                    if group['amsgrad']:
                        st['max_exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                if p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous() \
                        and p.grad.dtype == torch.float32:
                    fused.setdefault((p.device, st['step']), []).append(p)
                    continue
                bc1 = 1 - beta1 ** st['step']
                bc2 = math.sqrt(1 - beta2 ** st['step'])
                g = p.grad
                st['exp_avg'].mul_(beta1).add_(g, alpha=1 - beta1)
                st['exp_avg_sq'].mul_(beta2).addcmul_(g, g, value=1 - beta2)
                if group['amsgrad']:
                    torch.max(st['max_exp_avg_sq'], st['exp_avg_sq'], out=st['max_exp_avg_sq'])
                    denom = st['max_exp_avg_sq'].sqrt()
                else:
                    denom = st['exp_avg_sq'].sqrt()
                upd = (st['exp_avg'] / bc1) / (denom / bc2 + group['eps']) * group['lr']
                if group['weight_decay'] != 0:
                    upd = upd + p * group['weight_decay']
                p.sub_(upd)
            if self.capturable and fused:
                if len(fused) != 1:
                    raise RuntimeError('capturable Adam: parameters of a group must share device and step')
                gi = self.param_groups.index(group)
                if self._hyper_host is None:
                    raise RuntimeError('capturable Adam: call enable_capture(device) before the first captured step')
                (device, step), ps = next(iter(fused.items()))
                self._hyper_host[gi] = torch.tensor(self._hyper_row(group, step))
                self._hyper_dev[gi].copy_(self._hyper_host[gi], non_blocking=True)
                self._fused_step(group, device, step, ps, hyper_dev=self._hyper_dev[gi])
                continue
            for (device, step), ps in fused.items():
                self._fused_step(group, device, step, ps)
        return loss

    def _fused_step(self, group, device, step, ps, hyper_dev=None):
        """All CUDA fp32 parameters of one group that share a step count: dpf_adam_step_multi
        (48 tensors per launch) instead of one launch - or ~10 ATen kernels - per parameter."""
        ct = _lib.ctypes
        beta1, beta2 = group['betas']
        n = len(ps)
        arr = ct.c_void_p * n
        states = [self.state[p] for p in ps]
        vmax = arr(*[s['max_exp_avg_sq'].data_ptr() for s in states]) if group['amsgrad'] else None
        if hyper_dev is not None:
            with torch.cuda.device(device):
                _lib.call("dpf_adam_step_multi_dev", n, arr(*[p.data_ptr() for p in ps]), arr(*[p.grad.data_ptr() for p in ps]),
                          arr(*[s['exp_avg'].data_ptr() for s in states]), arr(*[s['exp_avg_sq'].data_ptr() for s in states]),
                          vmax, (ct.c_longlong * n)(*[p.numel() for p in ps]), hyper_dev, device=device)
            return
        with torch.cuda.device(device):
            _lib.call("dpf_adam_step_multi", n, arr(*[p.data_ptr() for p in ps]), arr(*[p.grad.data_ptr() for p in ps]),
                      arr(*[s['exp_avg'].data_ptr() for s in states]), arr(*[s['exp_avg_sq'].data_ptr() for s in states]),
                      vmax, (ct.c_longlong * n)(*[p.numel() for p in ps]), float(group['lr']), float(beta1), float(beta2),
                      float(group['eps']), float(group['weight_decay']), float(1 - beta1 ** step),
                      float(math.sqrt(1 - beta2 ** step)), device=device)


class LRUpdater(object):
    def __init__(self, epoch_length, **kwargs):
        self.epoch_length = epoch_length
        self.cycle_length = kwargs['cycle_length']
        self.min_lr, self.max_lr = kwargs['min_lr'], kwargs['max_lr']
        self.beta1 = kwargs['beta1']
        self.min_beta2, self.max_beta2 = kwargs['min_beta2'], kwargs['max_beta2']

    def __call__(self, optimizer, epoch, iteration):
        pos = ((epoch % self.cycle_length) * self.epoch_length + iteration) / (self.cycle_length * self.epoch_length)
        c = 0.5 * (1.0 + np.cos(np.pi * pos))
        for group in optimizer.param_groups:
            group['lr'] = self.min_lr + (self.max_lr - self.min_lr) * c
            group['betas'] = (self.beta1, self.min_beta2 + (self.max_beta2 - self.min_beta2) * c)


# ---- checkpoint boundary of the optimizer state --------------------------------------------------
# The reference's checkpoints hold one optimizer-state entry per nn.Parameter of ITS modules
# (training.py:75-80, optimizers.py:32-40); here every coupling stack owns ONE arena parameter.  The
# arena's field order equals the reference's parameter registration order (flows.py:25-93), so the two
# layouts convert into each other by slicing / concatenating the per-tensor moments.
def _plan(model):
    """[(parameter, None | [(offset, shape), ...])] in model.parameters() order; a list = arena fields."""
    from ._arena import CouplingStack
    arenas = {id(m.arena): m for m in model.modules() if isinstance(m, CouplingStack)}
    plan = []
    for p in model.parameters():
        m = arenas.get(id(p))
        plan.append((p, None if m is None else [(off, shape) for off, shape in m.layout.param_index.values()]))
    return plan


def optimizer_state_to_reference(model, opt_state):
    """This package's Adam.state_dict() (one group) -> the reference's per-tensor layout."""
    if len(opt_state['param_groups']) != 1:
        raise ValueError('optimizer state conversion supports a single parameter group')
    state, out, idx = opt_state['state'], {}, 0
    for i, (p, fields) in enumerate(_plan(model)):
        st = state.get(i)
        if fields is None:
            if st is not None:
                out[idx] = st
            idx += 1
            continue
        for off, shape in fields:
            if st is not None:
                n = int(np.prod(shape))
                out[idx] = {k: (v.reshape(-1)[off:off + n].view(shape).clone() if torch.is_tensor(v) and v.numel() == p.numel() else v)
                            for k, v in st.items()}
            idx += 1
    group = dict(opt_state['param_groups'][0])
    group['params'] = list(range(idx))
    return {'state': out, 'param_groups': [group]}


def optimizer_state_from_reference(model, ref_state):
    """The reference's per-tensor optimizer state -> this package's layout (arena moments concatenated).
    Returns `ref_state` unchanged when it already has this model's parameter count."""
    plan = _plan(model)
    n_ref = sum(1 if f is None else len(f) for _, f in plan)
    if len(ref_state['param_groups']) != 1:
        raise ValueError('optimizer state conversion supports a single parameter group')
    n_have = len(ref_state['param_groups'][0]['params'])
    if n_have == len(plan) and n_have != n_ref:
        return ref_state
    if n_have != n_ref:
        raise ValueError('optimizer_state has %d parameters; this model has %d (%d in the reference layout): '
                         'the checkpoint belongs to a different architecture' % (n_have, len(plan), n_ref))
    state, out, idx = ref_state['state'], {}, 0
    for i, (p, fields) in enumerate(plan):
        if fields is None:
            if idx in state:
                out[i] = state[idx]
            idx += 1
            continue
        parts = [state.get(idx + j) for j in range(len(fields))]
        idx += len(fields)
        have = [s for s in parts if s is not None]
        if not have:
            continue
        if len(have) != len(parts):
            raise ValueError('optimizer_state covers only part of a coupling stack')
        steps = {int(s['step']) for s in parts}
        if len(steps) != 1:
            raise ValueError('coupling-stack parameters carry different step counts: %s' % sorted(steps))
        merged = {'step': steps.pop()}
        for k in parts[0]:
            if torch.is_tensor(parts[0][k]):
                merged[k] = torch.cat([s[k].reshape(-1) for s in parts]).to(p.device)
        out[i] = merged
    group = dict(ref_state['param_groups'][0])
    group['params'] = list(range(len(plan)))
    return {'state': out, 'param_groups': [group]}
