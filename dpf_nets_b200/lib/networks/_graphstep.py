"""Whole-model training step as two CUDA graphs (opt-in: config key `cuda_graph`, training.py).

The reference's step (lib/networks/training.py:33-56) is ~600 small kernels on the latent side (FeatureEncoders,
14 latent coupling flows, priors, losses, their autograd) around the point decoder; on a B200 they are
launch-bound (host issue time ~= GPU time, tools/model_step_probe.py).  Shapes are static (drop_last batches), so:

  graph A = model forward + loss + backward      (inputs copied into static buffers first)
  -- host: NaN guard on the loss (training.py:44-47), gradient all-reduce across ranks (eager NCCL) --
  graph B = optimizer step                       (hyper-parameters read from device memory, optimizers.Adam.capturable)

The first `eager_steps` calls run eagerly (they are real training steps: they create the optimizer state and
warm every kernel attribute), the next call captures and replays.  A change of input shape falls back to eager
for that call.
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, loss_func, optimizer, allreduce=None, eager_steps=2):
        self.model, self.loss_func, self.optimizer = model, loss_func, optimizer
        self.allreduce = allreduce
        self.eager_steps = eager_steps
        self.calls = 0
        self.graph_a = self.graph_b = None
        self.static_in = None
        self.static_losses = None
        self.nan_guard = None          # callable(loss tensor) -> None, may raise; runs between the two graphs

    def _forward_backward(self, inputs):
        outputs = self.model(*inputs)
        losses = self.loss_func(inputs[0], inputs[1], outputs)
        return losses

    def _eager(self, inputs):
        losses = self._forward_backward(inputs)
        if self.nan_guard is not None:
            self.nan_guard(losses[0])
        self.optimizer.zero_grad()
        losses[0].backward()
        if self.allreduce is not None:
            self.allreduce()
        self.optimizer.step()
        # detached: a caller that keeps the returned losses would keep this step's autograd graph - and with it the
        # parameters' AccumulateGrad nodes, which are bound to the (legacy default) stream they were created on;
        # the capture's forward would then reuse them and the backward would touch the legacy stream mid-capture
        return tuple(l.detach() for l in losses)

    def _capture(self, inputs):
        self.static_in = [t.clone() for t in inputs]
        self.optimizer.enable_capture(inputs[0].device)   # allocations happen outside the capture
        self.optimizer.zero_grad(set_to_none=True)      # gradients are (re)created inside graph A's memory pool
        torch.cuda.synchronize()
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            losses = self._forward_backward(self.static_in)
            losses[0].backward()
        self.static_losses = tuple(l.detach() for l in losses)
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
            self.optimizer.step()                        # advances the step counters once: the first replay's step

    def __call__(self, *inputs):
        """inputs = (g_clouds, p_clouds[, images]) on the model's device -> (loss, pnll, gnll, gent) tensors."""
        self.calls += 1
        if self.calls <= self.eager_steps:
            return self._eager(inputs)
        first = self.graph_a is None
        if first:
            self._capture(inputs)
        elif any(a.shape != b.shape for a, b in zip(inputs, self.static_in)):
            raise RuntimeError("GraphedTrainStep: input shape changed after capture (use drop_last batches)")
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph_a.replay()
        if self.nan_guard is not None:
            self.nan_guard(self.static_losses[0])
        if self.allreduce is not None:
            self.allreduce()
        if not first:
            self.optimizer.prepare_replay()
        self.graph_b.replay()
        return self.static_losses
