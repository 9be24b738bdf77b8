"""Multi-GPU plumbing for the two sharded paths (one process per GPU, torch.distributed):

* all-pairs Chamfer matrices: every rank holds both cloud sets, computes an interleaved set of rows
  (row r -> rank r % R, which also balances the upper-triangle-only symmetric case), and the row
  blocks are exchanged with ONE collective (sum of disjoint zero-filled blocks == all-gather);
* batch-sharded training: gradients of the parameter arena (one tensor per module) are averaged
  with one NCCL all-reduce per arena.

The compute callback is injectable so that the sharding logic is testable on CPU (gloo).
"""
import torch
import torch.distributed as dist


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
    matrix on every rank."""
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


def allreduce_arena_grads(module, group=None):
    """Average the gradients of every parameter of `module` across ranks (one collective per
    parameter tensor; the decoder contributes exactly one: its arena)."""
    rank, R = world(group)
    if R == 1:
        return
    for prm in module.parameters():
        if prm.grad is not None:
            dist.all_reduce(prm.grad, op=dist.ReduceOp.SUM, group=group)
            prm.grad.div_(R)
