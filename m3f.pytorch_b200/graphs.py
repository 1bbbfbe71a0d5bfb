"""CUDA-graph replay of an inference forward (launch-bound small batches).

At the reference's CPU-runnable size (2 clips x 16 frames) the eval forward is ~50 kernel launches of a few
microseconds each: the Python / ctypes dispatch between them, not the GPU, sets the latency.  Every kernel of this
package launches on the current stream, allocates nothing behind the allocator's back and never synchronises, so the
whole forward can be captured once into a CUDA graph and replayed with one launch.

    g = GraphedInference(model, example_batch)      # warm-up + capture
    y = g(batch)                                    # copies the inputs into the static buffers, replays, returns y

`batch` is a tensor, a tuple of tensors, or the AffWild2VA batch dict; shapes must equal the example's.
"""
import torch


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
        # warm-up on a side stream: one-time work (cudaFuncSetAttribute, weight packing caches, allocator pools)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                self._call(self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self._call(self.static_in)

    def __call__(self, batch):
        _copy_into(self.static_in, batch)
        self.graph.replay()
        return self.static_out
