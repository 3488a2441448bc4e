"""CUDA-graph capture of a whole module-API training step.

The not-MIWAE (MNAR) models train through the `pcvae::` custom ops with autograd and
`torch.optim.Adam` (reference train.py:87-116).  At the reference's own sizes (batch 128, 20 samples)
one step is ~60 short kernel launches and is bound by launch latency and Python, not by the GPU, so in
throughput mode the step is recorded once per batch shape into a CUDA graph and replayed:

    static inputs  <- copy_ of this step's (x, mask, mask_p) and freshly drawn noise   (outside the graph)
    graph replay   =  forward + loss + backward + Adam                                 (one launch)

Only used with `PCVAE_MODE=throughput`: capture needs `Adam(capturable=True)`, whose bias corrections
are fp32 device tensors instead of Python doubles, so parity mode keeps the eager optimiser.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import torch


class NoiseFeed:
    """Stand-in for the models' `noise` mode: hands out pre-allocated static [B, S, L] buffers in call
    order (q branch, p branch, ...) so the draws themselves stay outside the captured region."""

    def __init__(self, fill: Callable[[torch.Tensor], None]):
        self.fill = fill
        self.buffers: List[torch.Tensor] = []
        self.cursor = 0
        self.frozen = False

    def __call__(self, rows: int, samples: int, latent: int, device) -> torch.Tensor:
        if self.cursor == len(self.buffers):
            if self.frozen:
                raise RuntimeError("NoiseFeed: the captured step draws more noise tensors than it was recorded with")
            self.buffers.append(torch.empty(rows, samples, latent, device=device, dtype=torch.float32))
        t = self.buffers[self.cursor]
        if t.shape != (rows, samples, latent):
            raise RuntimeError(f"NoiseFeed: draw {self.cursor} changed shape {tuple(t.shape)} -> {(rows, samples, latent)}")
        self.cursor += 1
        return t

    def refill(self):
        self.cursor = 0
        for t in self.buffers:
            self.fill(t)


class GraphedStep:
    """fn(*static_inputs) -> loss must run forward + loss and return a scalar loss tensor; the optimiser must be
    built with capturable=True and already have taken at least one eager step (state initialised)."""

    def __init__(self, fn: Callable[..., torch.Tensor], optimizer: torch.optim.Optimizer, example_inputs: Sequence[torch.Tensor],
                 feed: NoiseFeed):
        self.fn, self.opt, self.feed = fn, optimizer, feed
        self.static = [torch.empty_like(t) for t in example_inputs]
        self.graph = None
        self.loss = None

    def _load(self, inputs):
        for s, t in zip(self.static, inputs):
            s.copy_(t, non_blocking=True)
        self.feed.refill()

    def capture(self, inputs):
        for s, t in zip(self.static, inputs):
            s.copy_(t, non_blocking=True)
        # one eager pass with the static buffers in place allocates the noise buffers in call order
        # (no backward, no optimizer step: parameters untouched); then fill them for this step
        self.feed.cursor = 0
        with torch.no_grad():
            self.fn(*self.static)
        self.feed.frozen = True
        self.feed.refill()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            loss = self.fn(*self.static)
            loss.backward()
            self.opt.step()
            self.loss = loss.detach()
        self.feed.cursor = 0

    def __call__(self, *inputs) -> torch.Tensor:
        if self.graph is None:
            self.capture(inputs)       # records only; the replay below executes this step
        else:
            self._load(inputs)
        self.graph.replay()
        return self.loss


class GraphedTrainer:
    """Per-batch-shape cache of GraphedStep; shapes seen fewer than `eager_first` times run eagerly (the ragged last
    batch of an epoch, and the very first steps that initialise the optimiser state)."""

    def __init__(self, model, make_fn: Callable[[], Callable[..., torch.Tensor]], optimizer, fill_noise, eager_first: int = 2):
        self.model, self.make_fn, self.opt, self.fill_noise = model, make_fn, optimizer, fill_noise
        self.eager_first = eager_first
        self.seen: Dict[Tuple, int] = {}
        self.graphs: Dict[Tuple, Tuple[GraphedStep, NoiseFeed]] = {}
        self.replays = 0

    def _eager(self, inputs):
        feed = NoiseFeed(self.fill_noise)
        prev = self.model.noise
        self.model.noise = lambda r, s, l, dev: _fresh(feed, r, s, l, dev)
        try:
            self.opt.zero_grad(set_to_none=True)
            loss = self.make_fn()(*inputs)
            loss.backward()
            self.opt.step()
        finally:
            self.model.noise = prev
        return loss.detach()

    def step(self, *inputs) -> torch.Tensor:
        key = tuple((tuple(t.shape), t.dtype) for t in inputs)
        n = self.seen.get(key, 0)
        self.seen[key] = n + 1
        if n < self.eager_first:
            return self._eager(inputs)
        if key not in self.graphs:
            feed = NoiseFeed(self.fill_noise)
            self.graphs[key] = (GraphedStep(self.make_fn(), self.opt, inputs, feed), feed)
        gs, feed = self.graphs[key]
        prev = self.model.noise
        self.model.noise = feed
        try:
            out = gs(*inputs)
        finally:
            self.model.noise = prev
        self.replays += 1
        return out


def _fresh(feed: NoiseFeed, rows, samples, latent, device):
    t = torch.empty(rows, samples, latent, device=device, dtype=torch.float32)
    feed.fill(t)
    return t
