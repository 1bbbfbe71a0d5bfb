"""Fork / join of the independent branches of the INFERENCE forward onto side CUDA streams.

Why: at evaluation sizes the recurrences set the latency, not the throughput kernels.  A BiGRU layer launch occupies
2 x H/32 x ceil(B/64) SMs (32 of 148 for H = 512 at B <= 64) for T serial steps of ~6 us, and the audio-visual model
runs eight of them back to back (audio 2, visual 2, scorers 1 + 1, fusion 2).  The audio stream does not depend on
the visual one and the two attention scorers do not depend on each other, so they run concurrently on disjoint SMs:
8 -> 5 serial layer launches.  The side branches are forked AFTER the visual convolutions (persistent kernels, one
CTA per SM: an SM held by a recurrence would turn their single wave into two).

Rules that make this safe without `Tensor.record_stream`:
  * a branch starts with `side.wait_stream(current)` and is joined with `current.wait_stream(side)` before its result
    is read, so a block freed on either stream can only be reused by work ordered after its last reader;
  * enabled only without autograd, outside the fp32-parity mode and for B <= 64 clips: then every recurrence is one
    batch slice and all concurrently running cooperative launches together need at most 80 SMs, so each of them is
    fully resident whatever the scheduler does;
  * capturable: inside `torch.cuda.graph` the fork / join become graph dependencies (graphs.GraphedInference).
M3T_STREAMS=0 (or `set_enabled(False)`) keeps everything on the current stream; results are bit-identical.
"""
import os

import torch

_enabled = os.environ.get("M3T_STREAMS", "1") != "0"
_side = {}
MAX_CLIPS = 64
# Training: the fork is allowed only while this flag is set — TrainEngine.capture sets it for small per-GPU shards
# (<= MAX_CLIPS), where the step is latency-bound and is replayed as a CUDA graph: the forks / joins of the forward
# become graph dependencies, and autograd runs each branch's backward on the stream its forward ran on, so the
# recurrences of the audio stream and of one attention scorer overlap the other branch in both directions.
_train_overlap = False


def set_train_overlap(flag):
    global _train_overlap
    prev, _train_overlap = _train_overlap, bool(flag)
    return prev


def set_enabled(flag):
    global _enabled
    prev, _enabled = _enabled, bool(flag)
    return prev


def overlap_ok(x):
    """x: a (B, ...) device tensor of the forward being run."""
    if not _enabled or not torch.is_tensor(x) or not x.is_cuda:
        return False
    if torch.is_grad_enabled() and not _train_overlap:
        return False
    from . import fp32
    return not fp32.enabled() and x.shape[0] <= MAX_CLIPS


def _stream(i, device):
    key = (device.index if device.index is not None else torch.cuda.current_device(), i)
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=device)
    return _side[key]


def run_on_side(i, device, fn, *args):
    """Enqueue fn(*args) on side stream i, ordered after everything enqueued on the current stream so far."""
    cur = torch.cuda.current_stream(device)
    s = _stream(i, device)
    s.wait_stream(cur)
    with torch.cuda.stream(s):
        return fn(*args)


def join(device, *ids):
    cur = torch.cuda.current_stream(device)
    for i in ids:
        cur.wait_stream(_stream(i, device))
