"""BatchNorm `num_batches_tracked` bookkeeping for the fused paths: a model forward touches ~35 of these int64 scalars, one
tiny add kernel each; inside `deferred()` they are collected and bumped by ONE multi-tensor add when the outermost context
exits (the values are identical; outside any context `bump` adds immediately, like nn.BatchNorm does)."""
import contextlib

import torch

_pending = None


def bump(counter):
    if _pending is None:
        counter += 1
    else:
        _pending.append(counter)


@contextlib.contextmanager
def deferred():
    global _pending
    if _pending is not None:          # nested: the outermost context flushes
        yield
        return
    _pending = []
    try:
        yield
    finally:
        todo, _pending = _pending, None
        if todo:
            with torch.no_grad():
                torch._foreach_add_(todo, 1)
