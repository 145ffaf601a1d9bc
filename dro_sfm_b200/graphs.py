"""Whole training step as ONE CUDA graph (second half).

At the training shapes (B = 1-2 per GPU, 40x120 feature maps) a step of the reference model is ~2 000 kernel launches of
a few microseconds each (ResNet encoder, 12-32 GRU sub-steps, the cost calls between them, the loss): eager PyTorch is
bound by Python and launch latency, not by the GPU.  Every operator of this package is capture-safe (no allocation outside
torch's allocator, no host synchronisation, no host-side branching on device data; the loss's second stream is joined by
events), so forward + loss + backward can be recorded once and replayed:

    step = GraphedStep(lambda batch: model(batch)["loss"].sum(), example_batch, model.parameters())
    for batch in loader:
        loss = step(batch)          # copies the batch into the static buffers, replays; gradients are in p.grad
        optimizer.step()

Host-side randomness of the model (the left-right flip of SelfSupModelMF.forward, `random.random() < flip_lr_prob`) is
frozen at capture time: build one GraphedStep per branch and choose on the host.
"""
import torch


def _tree_map(f, x):
    if torch.is_tensor(x):
        return f(x)
    if isinstance(x, dict):
        return {k: _tree_map(f, v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_tree_map(f, v) for v in x)
    return x


def _tree_copy_(dst, src):
    if torch.is_tensor(dst):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _tree_copy_(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _tree_copy_(d, s)


class GraphedStep:
    """fn(batch) -> scalar loss tensor (or a dict with key 'loss'); forward + backward captured in one CUDA graph.

    example_batch: pytree (dict / list / tuple) of CUDA tensors with the shapes and dtypes of every later batch.
    params: the parameters whose .grad the backward pass fills (gradients are OVERWRITTEN by every replay, as after
            zero_grad(); call optimizer.step() after each step)."""

    def __init__(self, fn, example_batch, params, warmup=3):
        self.fn = fn
        self.params = [p for p in params if p.requires_grad]
        self.static_batch = _tree_map(lambda t: t.detach().clone(), example_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):            # cuDNN plans, lazy initialisations, the loss's scratch buffers
                self._zero_grads()
                self._run_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._zero_grads()                              # .grad is re-created inside the graph's private memory pool
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._run_eager()
        self.replays = 0

    def _zero_grads(self):
        for p in self.params:
            p.grad = None

    def _run_eager(self):
        out = self.fn(self.static_batch)
        loss = out["loss"] if isinstance(out, dict) else out
        loss.sum().backward()
        return out

    def __call__(self, batch=None):
        """Replays the step on `batch` (None: on whatever the static buffers hold); returns the static output."""
        if batch is not None:
            _tree_copy_(self.static_batch, batch)
        self.graph.replay()
        self.replays += 1
        return self.static_out
