"""CUDA-graph replay of an inference forward (launch-bound small batches).

At the reference's CPU-runnable size (2 clips x 16 frames) the eval forward is ~50 kernel launches of a few
microseconds each: the Python / ctypes dispatch between them, not the GPU, sets the latency.  Every kernel of this
package launches on the current stream, allocates nothing behind the allocator's back and never synchronises, so the
whole forward can be captured once into a CUDA graph and replayed with one launch.

    g = GraphedInference(model, example_batch)      # warm-up + capture
    y = g(batch)                                    # copies the inputs into the static buffers, replays, returns y

`batch` is a tensor, a tuple of tensors, or the AffWild2VA batch dict; shapes must equal the example's.

The captured graph holds raw addresses of the packed / bf16 weight copies in `ops._pack_cache`.  The instance therefore
(1) keeps strong references to every cache value that existed at capture, so a later `ops.clear_caches()` (every
TrainEngine step) cannot hand those buffers back to the allocator under the graph, and (2) records the parameters'
version counters, storage pointers and the cache generation: if the weights were updated in place, re-seated (`load_state_dict`,
`p.data = ...`) or rewritten through the optimizer arena (generation bump), the next call re-captures instead of
replaying stale weights.
"""
import os

import torch

from . import lib as L
from . import ops


def _map(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map(v, fn) for k, v in obj.items()}
    if isinstance(obj, (tuple, list)):
        return type(obj)(_map(v, fn) for v in obj)
    return obj


def _copy_into(dst, src):
    if torch.is_tensor(dst):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k])
    elif isinstance(dst, (tuple, list)):
        for d, s in zip(dst, src):
            _copy_into(d, s)


class GraphedInference:
    def __init__(self, model, example, warmup=3):
        assert not model.training, "graph capture is for eval-mode inference"
        self.model = model
        self.static_in = _map(example, lambda t: t.detach().clone())
        self._call = (lambda x: model(*x)) if isinstance(example, tuple) else model
        self.warmup = warmup
        self.captures = 0
        self._capture()

    def _weights_key(self):
        # per replay: version counters only (~20 us for the AV model); storage pointers are compared as well whenever
        # the tensor list itself is rebuilt (every 64 calls), which also catches `p.data = ...` re-seating
        self._calls = getattr(self, "_calls", 0) + 1
        if self._calls % 64 == 1 or not hasattr(self, "_tensors"):
            self._tensors = list(self.model.parameters()) + list(self.model.buffers())
            self._ptrs = tuple(t.data_ptr() for t in self._tensors)
        return (ops.CACHE_GENERATION, self._ptrs, tuple([t._version for t in self._tensors]))

    def _capture(self):
        model, warmup = self.model, self.warmup
        assert not model.training, "graph capture is for eval-mode inference"
        # warm-up on a side stream: one-time work (cudaFuncSetAttribute, weight packing caches, allocator pools)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                self._call(self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        # kernel -> kernel edges of the captured sequence become programmatic dependencies (csrc/common.cuh): a replay
        # is launch-latency bound, the next kernel's scheduling and prologue overlap the previous kernel's tail
        prev_pdl = L.set_pdl(None)
        if "M3T_PDL" not in os.environ:
            L.set_pdl(True)
        try:
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.static_out = self._call(self.static_in)
        finally:
            L.set_pdl(prev_pdl)
        self._held = [v[2] for v in ops._pack_cache.values()]     # the derived weight copies the graph reads
        self._key = self._weights_key()
        self.captures += 1

    def __call__(self, batch):
        if self._weights_key() != self._key:
            self._capture()
        _copy_into(self.static_in, batch)
        self.graph.replay()
        return self.static_out
