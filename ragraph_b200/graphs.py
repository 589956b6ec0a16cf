"""CUDA-graph replay of a no-grad forward over the ragraph kernels.

The reference's real shapes (Cora-sized node batches, TU graphs of a few dozen nodes) are launch-bound: a
``RAGraph.forward`` is ~25 kernels of a few microseconds each behind ~1 ms of Python and launch overhead
(RAGraph_node/RAGraph.py:39-63 evaluated once per epoch over a fixed graph, RAGraph_graph/RAGraph.py:58-75 once per
graph).  Every op of this package launches on torch's current stream through the C ABI with caller-owned buffers and no
host synchronisation, so a whole forward can be captured once and replayed as ONE graph launch:

    fwd = GraphedForward(model, features, adj)       # warm-up (fills the CSR / store caches), then capture
    out = fwd(new_features)                          # copy into the static input, replay, return the static output

What is captured as a constant: everything that is not a ``dynamic`` tensor argument -- the adjacency (its CSR is built
and cached during warm-up), the library (keys / values / labels of the store at capture time), module parameters.  Call
``recapture()`` after changing any of them.  The returned tensors are the graph's static outputs: they are overwritten
by the next call (clone them to keep them)."""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import Tensor


def _walk(out):
    if isinstance(out, Tensor):
        yield out
    elif isinstance(out, (tuple, list)):
        for o in out:
            yield from _walk(o)


class GraphedForward:
    def __init__(self, fn: Callable, *args, dynamic: Sequence[int] = (0,), warmup: int = 3):
        """``fn(*args)`` is run ``warmup`` times eagerly (no grad) on a side stream, then captured.  ``dynamic`` lists the
        positions of the tensor arguments that change between calls (default: the first -- the node features); they are
        copied into static buffers, every other argument is baked into the graph."""
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedForward needs a CUDA device (no CPU fallback)")
        self.fn = fn
        self.args = list(args)
        self.dynamic = [i for i in dynamic if i < len(args) and isinstance(args[i], Tensor)]
        for i in self.dynamic:
            if not self.args[i].is_cuda:
                raise RuntimeError("GraphedForward: dynamic arguments must be CUDA tensors")
            self.args[i] = self.args[i].detach().clone()          # static input buffer
        self.warmup = max(1, warmup)
        self.graph = None
        self.out = None
        self.recapture()

    def recapture(self) -> None:
        """(re)build the graph: after the library, the adjacency or the module's parameters changed"""
        dev = next((a.device for a in self.args if isinstance(a, Tensor) and a.is_cuda), torch.device("cuda"))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(self.warmup):                           # fills the CSR / store / workspace caches off-graph
                self.fn(*self.args)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self.fn(*self.args)
        if not any(True for _ in _walk(self.out)):
            raise RuntimeError("GraphedForward: fn returned no tensor")

    def __call__(self, *dynamic_values: Tensor):
        if len(dynamic_values) != len(self.dynamic):
            raise RuntimeError(f"GraphedForward: expected {len(self.dynamic)} dynamic argument(s), got {len(dynamic_values)}")
        for i, v in zip(self.dynamic, dynamic_values):
            buf = self.args[i]
            if v.shape != buf.shape or v.dtype != buf.dtype:
                raise RuntimeError(f"GraphedForward: argument {i} must keep shape {tuple(buf.shape)} / dtype {buf.dtype} "
                                   f"(got {tuple(v.shape)} / {v.dtype}); build a new GraphedForward for a new shape")
            if v.data_ptr() != buf.data_ptr():
                buf.copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
