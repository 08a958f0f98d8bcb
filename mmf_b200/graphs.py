"""Whole-step CUDA graph for the small configurations (BASELINE.json configs[0]: MMBT, batch 2, 40 kernel launches of a few
microseconds each - the host cannot issue them as fast as the GPU retires them).

    step = GraphedStep(model, loss_fn, example_batch)      # warm-up, then capture of forward + backward
    loss = step(batch)                                      # copies `batch` into the static inputs, replays the graph

The graph holds the library's kernels exactly as an eager step launches them (same C ABI calls, recorded by stream capture);
nothing is traced or compiled.  What a capture freezes and how it is kept correct:
  * dropout: the host-side (seed, offset) of every keep-bit draw is frozen, so the graph increments a device-resident step
    counter that the generator mixes into its Philox counter (functional.dropout_epoch): every replay draws fresh masks;
  * gradients: captured with `.grad` unset, so the flat gradient buffers are zeroed inside the graph and the parameters'
    `.grad` are views of them afterwards - an optimizer sees fresh gradients after every replay;
  * inputs: static device tensors, refreshed by `copy_` on the replay stream (shapes are fixed at capture).
The reference has no counterpart (its trainer launches eagerly, mmf/trainers/core/training_loop.py:185-213); this is the
B200 answer to SURVEY.md 8a's "at T = R = 36 everything is launch-bound"."""
import torch

from . import functional as F


def _map(fn, x):
    if isinstance(x, dict):
        return {k: _map(fn, v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_map(fn, v) for v in x)
    return fn(x) if torch.is_tensor(x) else x


def _copy_into(dst, src):
    if isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _copy_into(d, s)
    elif torch.is_tensor(dst):
        dst.copy_(src, non_blocking=True)


class GraphedStep:
    def __init__(self, model, loss_fn, example_batch, warmup=3):
        """loss_fn(batch) -> scalar loss; `example_batch`: (nested dict / list of) CUDA tensors with the shapes of every later
        batch.  The model must be in the mode (train / eval) it will be replayed in."""
        self.model, self.loss_fn = model, loss_fn
        self.static = _map(lambda t: t.detach().clone(), example_batch)
        dev = next(t for t in _flatten(self.static) if t.is_cuda).device
        self.epoch = F.dropout_epoch(dev, create=True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):              # lazy engine state, kernel attributes, library work buffers
                model.zero_grad(set_to_none=True)
                self.epoch.add_(1)
                loss_fn(self.static).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        # captured on the warm-up stream: autograd binds every parameter's gradient accumulator to the stream that was current
        # when the accumulator was created, i.e. this one, as long as no autograd graph from an earlier eager step on another
        # stream is still alive (it would keep the old accumulators - and their stream - in use)
        with torch.cuda.graph(self.graph, stream=side):
            self.epoch.add_(1)
            self.loss = loss_fn(self.static)
            self.loss.backward()

    def __call__(self, batch=None):
        if batch is not None:
            _copy_into(self.static, batch)
        self.graph.replay()
        return self.loss


def _flatten(x):
    if isinstance(x, dict):
        for v in x.values():
            yield from _flatten(v)
    elif isinstance(x, (list, tuple)):
        for v in x:
            yield from _flatten(v)
    elif torch.is_tensor(x):
        yield x
