"""Multi-GPU plumbing for the two sharded paths (one process per GPU, torch.distributed):

* all-pairs Chamfer matrices: every rank holds both cloud sets (`gather_rows` brings the per-rank
  shards of generated / reference clouds together first), computes an interleaved set of rows
  (row r -> rank r % R, which also balances the upper-triangle-only symmetric case), and the row
  blocks are exchanged with ONE collective (sum of disjoint zero-filled blocks == all-gather);
* batch-sharded training: `GradSync` averages gradients across ranks - the decoder's arena gradient
  (the bulk of the bytes, and the FIRST gradient autograd finishes) is all-reduced asynchronously
  while the rest of the backward still runs, everything else goes in flat buckets afterwards.

The compute callback is injectable so that the sharding logic is testable on CPU (gloo).
"""
import torch
import torch.distributed as dist
from torch.utils.data import Sampler


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_rows(n_rows, rank, world_size):
    """(row_start, row_step, n_local) of the interleaved row shard of `rank`."""
    n_local = max(0, (n_rows - rank + world_size - 1) // world_size)
    return rank, world_size, n_local


def sharded_pairwise(compute_rows, S1, S2, device, symmetric=False, group=None):
    """compute_rows(out, row_start, row_step, n_rows, symmetric) fills rows row_start::row_step of the
    zero-initialised (S1,S2) matrix `out` (upper triangle only when symmetric).  Returns the full
    matrix on every rank.  Every rank must hold the SAME cloud sets (see gather_rows)."""
    rank, R = world(group)
    out = torch.zeros((S1, S2), dtype=torch.float32, device=device)
    row_start, row_step, n_local = shard_rows(S1, rank, R)
    if n_local > 0:
        compute_rows(out, row_start, row_step, n_local, symmetric)
    if R > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)   # disjoint row blocks: sum == gather
    if symmetric:
        out = torch.triu(out) + torch.triu(out, 1).t()
    return out


# ---- evaluation: contiguous, un-padded dataset shards + gather ---------------------------------
class ShardSampler(Sampler):
    """Rank r iterates the contiguous index block [r*ceil(S/R), (r+1)*ceil(S/R)) of the dataset in
    order, WITHOUT the duplicate padding of DistributedSampler, so that concatenating the ranks'
    results in rank order (gather_rows) reproduces the single-process iteration order exactly."""

    def __init__(self, n_items, rank=None, world_size=None):
        r, R = world()
        self.rank = r if rank is None else rank
        self.world_size = R if world_size is None else world_size
        per = (n_items + self.world_size - 1) // self.world_size
        self.lo = min(n_items, self.rank * per)
        self.hi = min(n_items, self.lo + per)

    def __iter__(self):
        return iter(range(self.lo, self.hi))

    def __len__(self):
        return self.hi - self.lo


def gather_rows(t, group=None):
    """Concatenate the ranks' tensors along dim 0 in rank order (row counts may differ, trailing
    shape must agree); every rank gets the full tensor.  Identity when not distributed."""
    rank, R = world(group)
    if R == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(R)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(R)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def reduce_sums(values, device, group=None):
    """Sum a list of python floats across ranks (AverageMeter sums / counts)."""
    rank, R = world(group)
    if R == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()


def any_rank(flag, device, group=None):
    """Logical OR of a per-rank python bool / 0-dim tensor across ranks (one MAX all-reduce), so that
    every rank takes the same early-exit decision (NaN guard)."""
    rank, R = world(group)
    f = flag.detach().to(device=device, dtype=torch.float32).reshape(1) if torch.is_tensor(flag) \
        else torch.tensor([1.0 if flag else 0.0], device=device)
    if R > 1:
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)
    return bool(f.item() > 0)


# ---- training: gradient averaging --------------------------------------------------------------
class GradSync:
    """Averages the gradients of `module` across ranks.

    * Parameters with >= `async_numel` elements (the decoder's arena: 3.7 M - 9.9 M of the model's
      4.8 M - 25 M parameters) get a post-accumulate-grad hook that launches their NCCL all-reduce
      as soon as autograd has produced the gradient.  The decoder is the last module of the forward,
      so its gradient is the first one finished and its all-reduce overlaps with the backward of the
      encoder / latent flows / image encoder.
    * The remaining (small) parameters are flattened into buckets of <= `bucket_numel` elements and
      all-reduced by finish() - a fixed number of collectives per step whether or not a parameter
      received a gradient on this rank (missing gradients contribute zeros), so ranks cannot get out
      of step with each other.
    """

    def __init__(self, module, group=None, async_numel=1 << 20, bucket_numel=1 << 23, overlap=True):
        self.group = group
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.big = [p for p in self.params if p.numel() >= async_numel] if overlap else []
        big_ids = {id(p) for p in self.big}
        self.small = [p for p in self.params if id(p) not in big_ids]
        self.bucket_numel = bucket_numel
        self.pending = {}
        self.handles = []
        for p in self.big:
            self.handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _hook(self, p):
        rank, R = world(self.group)
        if R == 1 or (p.is_cuda and torch.cuda.is_current_stream_capturing()):
            return
        self.pending[id(p)] = dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    def finish(self):
        """Call after backward(): all-reduces what the hooks did not, waits, divides by the world size."""
        rank, R = world(self.group)
        if R == 1:
            return
        for p in self.big:
            w = self.pending.pop(id(p), None)
            if w is None:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
            else:
                w.wait()
            p.grad.div_(R)
        bucket, n = [], 0
        for p in self.small:
            if n + p.numel() > self.bucket_numel and bucket:
                self._reduce_bucket(bucket, R)
                bucket, n = [], 0
            bucket.append(p)
            n += p.numel()
        if bucket:
            self._reduce_bucket(bucket, R)

    def _reduce_bucket(self, params, R):
        # one concatenation, one collective, one scaling and one multi-tensor copy back: 4 launches per bucket
        # whatever the number of parameter tensors (the generation model has ~170 small ones, the SVR model ~230)
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(R)
        parts = [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)]
        if hasattr(torch, "_foreach_copy_"):
            torch._foreach_copy_(grads, parts)
        else:
            for g, c in zip(grads, parts):
                g.copy_(c)


def allreduce_arena_grads(module, group=None):
    """Average the gradients of every parameter of `module` across ranks (flat buckets, no overlap);
    kept for callers that do not hold a GradSync."""
    rank, R = world(group)
    if R == 1:
        return
    GradSync(module, group=group, overlap=False).finish()


def average_buffers(module, group=None):
    """BatchNorm running statistics are per rank during training (DDP-style);
    before a checkpoint every rank's floating-point buffers are replaced by the mean over ranks so that
    the saved model does not depend on which rank writes it."""
    rank, R = world(group)
    if R == 1:
        return
    bufs = [b for b in module.buffers() if b.is_floating_point()]
    if not bufs:
        return
    flat = torch.cat([b.reshape(-1) for b in bufs])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(R)
    off = 0
    for b in bufs:
        n = b.numel()
        b.copy_(flat[off:off + n].view_as(b))
        off += n
