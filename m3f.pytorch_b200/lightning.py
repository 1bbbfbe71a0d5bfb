"""The slice of pytorch_lightning 0.6 that the reference's scripts use (train.py:32-42, eval.py:18-23,
models/model.py:18,27,409,423,437), SURVEY 8(f) N1: `Trainer(...).fit(model)` / `.test(model)`, `LightningModule`
(+ `load_from_checkpoint`), `@data_loader`.  pytorch_lightning is not installed in this image and its 0.6 API is long
gone upstream; `m3t_b200.run` registers this module under that name so the scripts run unchanged.

What `fit` does, in the reference's order: optional sanity validation -> per epoch: DistributedSampler.set_epoch,
training_step -> backward -> gradient all-reduce -> global-norm clip (`gradient_clip_val`) -> optimizer step ->
`on_batch_end`; validation every `check_val_every_n_epoch` epochs (validation_step / validation_end), epoch
schedulers (ReduceLROnPlateau on `val_loss`), best-`val_loss` checkpoint under
`<default_save_path>/lightning_logs/version_N/checkpoints/_ckpt_epoch_E.ckpt`.

B200 design instead of Lightning's: ONE process per GPU.  `distributed_backend='ddp'` joins the process group that
`torchrun` set up; the gradient exchange and the Adam update are `engine.TrainEngine`'s flat-arena all-reduce + fused clip/Adam kernels (generic torch
optimizers — the reference's SGD option — take a flat-buffer all-reduce and `optimizer.step()`).  Started as a plain
`python train.py --gpus 0,1 --distributed` (the reference's README usage; Lightning 0.6 used mp.spawn there), rank 0
re-launches its own command line once per additional GPU with RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* set — the
children run the script from the top like torchrun workers and meet rank 0 in `fit` / `test`.  `'dp'` (Lightning:
replicas in threads of one process) runs on the first listed GPU.  Validation outputs of all ranks are gathered and
`validation_end` / `test_end` run once on rank 0 (Lightning 0.6 ran them per rank on partial data, which lets
ReduceLROnPlateau diverge between ranks); the result is broadcast.
"""
import argparse
import atexit
import functools
import logging
import os
import re
import socket
import subprocess
import sys
import warnings

import torch
import torch.distributed as dist
import torch.nn as nn

log = logging.getLogger("m3t_b200.lightning")

__version__ = "0.6.0+m3t_b200"


# ----------------------------------------------------------------------------------------------- module side
def data_loader(fn):
    """Lazy, memoised dataloader hook; val/test loaders are normalised to a list (Lightning 0.6 semantics)."""
    slot = "_lazy_" + fn.__name__

    @functools.wraps(fn)
    def getter(self):
        if slot not in self.__dict__:
            try:
                value = fn(self)
            except AttributeError as e:     # nn.Module.__getattr__ would turn this into a misleading message
                raise RuntimeError("%s raised AttributeError: %s" % (fn.__name__, e)) from e
            if value is not None and not isinstance(value, list) and fn.__name__ in ("val_dataloader",
                                                                                     "test_dataloader"):
                value = [value]
            self.__dict__[slot] = value
        return self.__dict__[slot]
    return getter


class LightningModule(nn.Module):
    """Hook surface the Trainer calls; subclasses override what they need."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.trainer = None
        self.current_epoch = 0
        self.global_step = 0
        self.on_gpu = False

    # hooks with a default
    def on_batch_start(self, batch):
        return None

    def on_batch_end(self):
        return None

    def on_epoch_start(self):
        return None

    def on_epoch_end(self):
        return None

    def on_save_checkpoint(self, checkpoint):
        return None

    def on_load_checkpoint(self, checkpoint):
        return None

    def train_dataloader(self):
        return None

    def val_dataloader(self):
        return None

    def test_dataloader(self):
        return None

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None):
        """Rebuild the module from a Trainer checkpoint: `hparams` -> Namespace -> cls(hparams) -> load_state_dict."""
        if map_location is None:
            map_location = lambda storage, loc: storage  # noqa: E731
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        if "hparams" not in ckpt:
            raise KeyError("checkpoint %s holds no 'hparams'; build the module and use load_state_dict" %
                           checkpoint_path)
        model = cls(argparse.Namespace(**ckpt["hparams"]))
        model.load_state_dict(ckpt["state_dict"])
        model.on_load_checkpoint(ckpt)
        return model


# ----------------------------------------------------------------------------------------------- helpers
def parse_gpus(gpus):
    """Lightning 0.6: int n -> first n devices; '-1' / -1 -> all; 'a,b' or '2' -> those device INDICES; list as is."""
    if gpus is None or gpus == 0 or gpus == "" or gpus == []:
        return []
    n_dev = torch.cuda.device_count()
    if isinstance(gpus, str):
        gpus = gpus.strip()
        if gpus == "-1":
            return list(range(n_dev))
        return [int(x) for x in gpus.split(",") if x.strip() != ""]
    if isinstance(gpus, int):
        return list(range(n_dev)) if gpus == -1 else list(range(gpus))
    return [int(x) for x in gpus]


def move_to(obj, device):
    """Tensors inside (nested) dict / list / tuple batches go to `device`; strings and numbers stay."""
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: move_to(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(move_to(v, device) for v in obj)
    return obj


def _record_stream(obj, stream):
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream)


class Prefetcher:
    """Iterates a DataLoader one batch ahead of the consumer: the host->device copy of batch i+1 (pinned memory,
    non-blocking, its own CUDA stream) overlaps the training step of batch i — 617 MB of float32 video per 256-clip
    step is ~12 ms of PCIe time that would otherwise sit in front of every step.  The consumer's stream waits on the
    copy's event and the tensors are `record_stream`-ed to it, so the allocator cannot hand their blocks to the next
    copy while the step still reads them.  On the CPU it degenerates to a plain look-ahead."""

    def __init__(self, loader, device):
        self.it = iter(loader)
        self.device = device
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self._ahead = None
        self._issue()

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self._ahead = None
            return
        if self.stream is None:
            self._ahead = (move_to(batch, self.device), None)
            return
        with torch.cuda.stream(self.stream):
            on_dev = move_to(batch, self.device)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self._ahead = (on_dev, ready)

    def __iter__(self):
        while self._ahead is not None:
            batch, ready = self._ahead
            if ready is not None:
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ready)
                _record_stream(batch, cur)
            self._issue()           # the next copy is in flight before the consumer starts on this batch
            yield batch


def prefetch_enabled():
    """Trainer.fit feeds training_step through `Prefetcher` (next batch copied on a side stream while the current
    step runs): 32.6 vs 43.9 ms per 256-clip step on a B200 (tests/bench_trainer.py, profiles/r2_next_session.md).
    M3T_TRAINER_PREFETCH=0 restores the plain copy in front of each step."""
    return os.environ.get("M3T_TRAINER_PREFETCH", "1") == "1"


class SequentialShardSampler(torch.utils.data.Sampler):
    """Evaluation sampler for one-process-per-GPU runs: rank r reads dataset indices r, r+world, ... in order and the
    shards are NOT padded to equal length.  `DistributedSampler` (what the reference's loaders use,
    models/model.py:416-418) repeats samples whenever len(dataset) % world != 0; the gathered validation / test outputs
    would then hold duplicated windows — summed twice by the overlap-add of half-stride windows, concatenated twice on
    tiled tracks, counted twice in val_loss (which drives ReduceLROnPlateau, checkpointing and early stopping)."""

    def __init__(self, dataset, num_replicas=None, rank=None):
        if num_replicas is None:
            num_replicas = dist.get_world_size() if dist.is_available() and dist.is_initialized() else \
                int(os.environ.get("WORLD_SIZE", "1"))
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else \
                int(os.environ.get("RANK", "0"))
        self.n, self.world, self.rank = len(dataset), int(num_replicas), int(rank)

    def __iter__(self):
        return iter(range(self.rank, self.n, self.world))

    def __len__(self):
        return len(range(self.rank, self.n, self.world))

    def set_epoch(self, epoch):
        pass


def _dedup_eval_outputs(outputs):
    """Drop windows that reached rank 0 more than once (a padding sampler handed to the Trainer by user code): a window
    is identified by (vid_name, start_frame).  Only step outputs carrying that bookkeeping are touched."""
    if not (isinstance(outputs, list) and outputs and all(
            isinstance(o, dict) and "vid_names" in o and "start_frames" in o for o in outputs)):
        return outputs
    seen, kept = set(), []
    for o in outputs:
        n = len(o["vid_names"])
        keep = []
        for j in range(n):
            key = (o["vid_names"][j], int(o["start_frames"][j]))
            if key not in seen:
                seen.add(key)
                keep.append(j)
        if len(keep) == n:
            kept.append(o)
        elif keep:
            new = {}
            for k, v in o.items():
                if isinstance(v, torch.Tensor) and v.dim() > 0 and v.size(0) == n:
                    new[k] = v[torch.tensor(keep)]
                elif isinstance(v, (list, tuple)) and len(v) == n:
                    new[k] = type(v)(v[j] for j in keep)
                else:
                    new[k] = v
            kept.append(new)
    return kept


def _scalar(v):
    if torch.is_tensor(v):
        return float(v.detach().float().mean().cpu()) if v.numel() else float("nan")
    try:
        return float(v)
    except (TypeError, ValueError):
        return None


def _flatten_metrics(result):
    """callback metrics = top-level scalars + the 'log' and 'progress_bar' dictionaries of a *_end result."""
    out = {}
    if not isinstance(result, dict):
        return out
    for k, v in result.items():
        if k in ("log", "progress_bar") and isinstance(v, dict):
            for kk, vv in v.items():
                s = _scalar(vv)
                if s is not None:
                    out[kk] = s
        elif not isinstance(v, dict):
            s = _scalar(v)
            if s is not None:
                out[k] = s
    return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


_children = []


def _reap_children(grace=600):
    """Rank 0 waits for the ranks it launched (atexit); after a failure on rank 0 they are stopped at once, since
    they would otherwise sit in a collective until its timeout."""
    for proc in _children:
        try:
            proc.wait(timeout=grace)
        except subprocess.TimeoutExpired:
            proc.terminate()
    del _children[:]


def relaunch_ranks(world):
    """Make this process rank 0 of `world` and start ranks 1..world-1 as copies of its own command line
    (`sys.orig_argv`), one process per GPU, rendezvous on 127.0.0.1."""
    port = _free_port()
    common = dict(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world))
    for rank in range(1, world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), **common)
        _children.append(subprocess.Popen(list(sys.orig_argv), env=env))
    os.environ.update(RANK="0", LOCAL_RANK="0", **common)
    atexit.register(_reap_children)


class EarlyStopping:
    """Keras-style early stopping as Lightning 0.6 shipped it: stop when `monitor` has not improved by more than
    `min_delta` for `patience` consecutive checks.  strict=False ignores a missing metric."""

    def __init__(self, monitor="val_loss", min_delta=0.0, patience=3, verbose=False, mode="min", strict=True):
        if mode not in ("min", "max"):
            raise ValueError("mode must be 'min' or 'max'")
        self.monitor, self.min_delta, self.patience, self.verbose, self.strict = monitor, min_delta, patience, verbose, \
            strict
        self.sign = 1.0 if mode == "min" else -1.0
        self.wait, self.best, self.stopped_epoch = 0, float("inf"), 0

    def on_epoch_end(self, epoch, metrics):
        """True when training should stop."""
        current = metrics.get(self.monitor)
        if current is None:
            if self.strict:
                raise RuntimeError("early stopping is conditioned on %r, which validation_end did not return"
                                   % self.monitor)
            return False
        score = self.sign * current
        if score + self.min_delta < self.best:
            self.best, self.wait = score, 0
            return False
        self.wait += 1
        if self.wait >= self.patience:
            self.stopped_epoch = epoch
            return True
        return False


# ----------------------------------------------------------------------------------------------- trainer
class Trainer:
    def __init__(self, logger=True, checkpoint_callback=True, early_stop_callback=None, default_save_path=None,
                 gradient_clip_val=0, process_position=0, nb_gpu_nodes=1, num_nodes=None, gpus=None,
                 log_gpu_memory=None, show_progress_bar=True, overfit_pct=0.0, track_grad_norm=-1,
                 check_val_every_n_epoch=1, fast_dev_run=False, accumulate_grad_batches=1, max_epochs=1000,
                 min_epochs=1, max_nb_epochs=None, min_nb_epochs=None, train_percent_check=1.0,
                 val_percent_check=1.0, test_percent_check=1.0, val_check_interval=1.0, log_save_interval=100,
                 row_log_interval=10, add_row_log_interval=None, distributed_backend=None, use_amp=False,
                 print_nan_grads=False, weights_summary="full", weights_save_path=None, amp_level="O1",
                 nb_sanity_val_steps=5, num_sanity_val_steps=None, truncated_bptt_steps=None,
                 resume_from_checkpoint=None, max_steps=None, num_processes=1, **unused):
        if unused:
            warnings.warn("Trainer: ignoring arguments %s" % sorted(unused))
        if use_amp:
            raise NotImplementedError("use_amp: the kernels already run bf16 operands with fp32 accumulation")
        if accumulate_grad_batches != 1 or truncated_bptt_steps:
            raise NotImplementedError("accumulate_grad_batches / truncated_bptt_steps are not used by the reference")
        if resume_from_checkpoint:
            raise NotImplementedError("resume_from_checkpoint: the reference restores weights through "
                                      "model.load_from_checkpoint / load_state_dict (train.py:26-30)")
        self.max_epochs = max_nb_epochs if max_nb_epochs is not None else max_epochs
        self.min_epochs = min_nb_epochs if min_nb_epochs is not None else min_epochs
        self.max_steps = max_steps
        # Lightning 0.6: True -> stop on val_loss (patience 3), an error if it is missing; None (what train.py:33
        # passes, and the default) -> the same callback but tolerant of a missing val_loss; False -> disabled
        if early_stop_callback is True:
            self.early_stop_callback = EarlyStopping("val_loss", patience=3, strict=True)
        elif early_stop_callback is None:
            self.early_stop_callback = EarlyStopping("val_loss", patience=3, strict=False)
        elif not early_stop_callback:
            self.early_stop_callback = None
        else:
            self.early_stop_callback = early_stop_callback
        self.gradient_clip_val = float(gradient_clip_val or 0)
        self.default_save_path = default_save_path or os.getcwd()
        self.weights_save_path = weights_save_path
        self.checkpoint_callback = checkpoint_callback
        self.check_val_every_n_epoch = max(int(check_val_every_n_epoch), 1)
        self.nb_sanity_val_steps = nb_sanity_val_steps if num_sanity_val_steps is None else num_sanity_val_steps
        self.fast_dev_run = bool(fast_dev_run)
        self.train_percent_check = train_percent_check
        self.val_percent_check = val_percent_check
        self.test_percent_check = test_percent_check
        self.row_log_interval = add_row_log_interval or row_log_interval
        self.show_progress_bar = show_progress_bar
        self.nb_gpu_nodes = num_nodes if num_nodes is not None else nb_gpu_nodes
        if self.nb_gpu_nodes != 1:
            raise NotImplementedError("one node (8 GPUs over NVSwitch) is the scope; nb_gpu_nodes must be 1")
        self.gpus = gpus
        self.num_processes = int(num_processes)     # CPU ranks (gloo) when gpus is empty: host-logic tests
        self.distributed_backend = distributed_backend
        self.resume_from_checkpoint = resume_from_checkpoint
        self.current_epoch = 0
        self.global_step = 0
        self.callback_metrics = {}
        self.rank, self.world = 0, 1
        self.device = torch.device("cpu")
        self.best_val, self.best_path = None, None
        self.ckpt_dir = None
        self.engine = None
        self.optimizers, self.lr_schedulers, self.plateau = [], [], None

    # ------------------------------------------------------------------ public API
    def fit(self, model):
        self._launch(model, "fit")
        return 1

    def test(self, model=None):
        if model is None:
            raise ValueError("Trainer.test needs the module (the reference passes it, eval.py:23)")
        self._launch(model, "test")

    # ------------------------------------------------------------------ process layout
    def _launch(self, model, mode):
        ids = parse_gpus(self.gpus)
        world = len(ids) if ids else self.num_processes
        under_launcher = int(os.environ.get("WORLD_SIZE", "1")) > 1
        if self.distributed_backend == "ddp" and not under_launcher and world > 1:
            relaunch_ranks(world)
        try:
            self._run(model, mode)
        except BaseException:
            _reap_children(grace=0)
            raise

    def _setup_process(self, model):
        ids = parse_gpus(self.gpus)
        env_world = int(os.environ.get("WORLD_SIZE", "1"))
        use_ddp = self.distributed_backend == "ddp" and env_world > 1
        local = int(os.environ.get("LOCAL_RANK", "0")) if use_ddp else 0
        if ids:
            if not torch.cuda.is_available():
                raise RuntimeError("Trainer(gpus=%r) but no CUDA device is visible" % (self.gpus,))
            if self.distributed_backend != "ddp" and len(ids) > 1:
                warnings.warn("'dp' replicates inside one process; this build runs one process per GPU — using GPU "
                              "%d only (pass --distributed and launch with torchrun for %d GPUs)" % (ids[0], len(ids)))
            dev = ids[local] if local < len(ids) else local
            self.device = torch.device("cuda", dev)
            torch.cuda.set_device(self.device)
        else:
            self.device = torch.device("cpu")
        if use_ddp:
            if not dist.is_initialized():
                dist.init_process_group("nccl" if self.device.type == "cuda" else "gloo")
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        else:
            self.rank, self.world = 0, 1
        model.to(self.device)
        model.trainer = self
        model.on_gpu = self.device.type == "cuda"
        if self.device.type == "cuda" and torch.backends.cudnn.deterministic and \
                os.environ.get("M3T_DETERMINISTIC") is None:
            # the reference's train.py:17 asks cuDNN for deterministic algorithms: the equivalent here is the slotted
            # accumulation mode (bit-reproducible training, ~12 % slower); M3T_DETERMINISTIC=0 keeps the atomics
            from . import raw
            if not raw.DETERMINISTIC:
                raw.set_deterministic(True)
                if self.rank == 0:
                    log.info("torch.backends.cudnn.deterministic is set: deterministic accumulation mode on "
                             "(M3T_DETERMINISTIC=0 to keep the faster atomics)")

    def _run(self, model, mode):
        self._setup_process(model)
        try:
            if mode == "fit":
                self._fit(model)
            else:
                self._test(model)
        except BaseException:
            # A failure on one rank (rank 0's validation_end / checkpoint write, typically) must not enter a barrier
            # the other ranks will never match: tear the group down and stop the ranks this process started, so the
            # real exception surfaces at once instead of after the collective timeout.
            if self.world > 1 and dist.is_initialized():
                try:
                    abort = getattr(dist.distributed_c10d, "_abort_process_group", None)
                    if abort is not None and self.device.type == "cuda":
                        abort()
                    else:
                        dist.destroy_process_group()
                except Exception:
                    pass
            _reap_children(grace=0)
            raise
        else:
            if self.world > 1 and dist.is_initialized():
                dist.barrier()

    # ------------------------------------------------------------------ optimisation plumbing
    def _init_optimizers(self, model):
        conf = model.configure_optimizers()
        if isinstance(conf, torch.optim.Optimizer):
            opts, scheds = [conf], []
        elif isinstance(conf, (list, tuple)) and len(conf) == 2 and isinstance(conf[0], (list, tuple)):
            opts, scheds = list(conf[0]), list(conf[1])
        elif isinstance(conf, (list, tuple)):
            opts, scheds = list(conf), []
        else:
            raise ValueError("configure_optimizers must return an optimizer, a list of them, or ([opts], [scheds])")
        if len(opts) != 1:
            raise NotImplementedError("one optimizer (the reference's configuration)")
        self.optimizers = opts
        self.plateau = None
        self.lr_schedulers = []
        for s in scheds:
            if isinstance(s, torch.optim.lr_scheduler.ReduceLROnPlateau):
                self.plateau = s
            else:
                self.lr_schedulers.append(s)
        opt = opts[0]
        self.engine = None
        if type(opt) is torch.optim.Adam and len(opt.param_groups) == 1 and not opt.param_groups[0].get("amsgrad"):
            # the product step: flat arenas, ONE all-reduce, fused clip + Adam (engine.py / csrc/optim.cu)
            from .engine import TrainEngine
            g = opt.param_groups[0]
            self.engine = TrainEngine(model, lr=g["lr"], weight_decay=g["weight_decay"],
                                      clip=self.gradient_clip_val, betas=tuple(g["betas"]), eps=g["eps"],
                                      overlap_allreduce=False)
            self.engine.all_params = [p for p in g["params"] if p.requires_grad]

    def _generic_step(self, opt):
        params = [p for g in opt.param_groups for p in g["params"] if p.grad is not None]
        if self.world > 1 and params:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / self.world)
            o = 0
            for p in params:
                p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
        if self.gradient_clip_val > 0:
            torch.nn.utils.clip_grad_norm_(params, self.gradient_clip_val)
        opt.step()
        opt.zero_grad(set_to_none=True)

    def _optimizer_step(self):
        opt = self.optimizers[0]
        if self.engine is not None:
            g = opt.param_groups[0]
            self.engine.lr, self.engine.wd = g["lr"], g["weight_decay"]     # schedulers act on the torch optimizer
            self.engine._gather_grads()
            if not self.engine.on_gpu:      # host-side logic tests: the engine's CPU stand-in is a torch Adam
                self.engine.opt.param_groups[0].update(lr=g["lr"], weight_decay=g["weight_decay"])
            self.engine._allreduce_grads()
            self.engine._optimizer_step()
            opt._opt_called = True      # the fused kernel took the step; keeps lr_scheduler's ordering check quiet
        else:
            self._generic_step(opt)

    # ------------------------------------------------------------------ loops
    def _limit(self, n, pct):
        if self.fast_dev_run:
            return min(n, 1)
        return max(int(n * pct), 1) if pct < 1.0 else n

    def _fit(self, model):
        self._init_optimizers(model)
        train_loader = model.train_dataloader()
        val_loaders = model.val_dataloader()
        if train_loader is None:
            raise RuntimeError("fit needs a train_dataloader")
        if self.rank == 0:
            n_par = sum(p.numel() for p in model.parameters())
            log.info("fit: %d parameters (%d trainable), device %s, world %d, %s step", n_par,
                     sum(p.numel() for p in model.parameters() if p.requires_grad), self.device, self.world,
                     "fused flat-arena Adam" if self.engine is not None else type(self.optimizers[0]).__name__)
        if val_loaders and self.nb_sanity_val_steps and not self.fast_dev_run:
            self._evaluate(model, val_loaders, "validation", max_batches=self.nb_sanity_val_steps)
        max_epochs = 1 if self.fast_dev_run else self.max_epochs
        stop = False
        for epoch in range(self.current_epoch, max_epochs):
            self.current_epoch = model.current_epoch = epoch
            sampler = getattr(train_loader, "sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)
            model.train()
            model.on_epoch_start()
            n_batches = self._limit(len(train_loader), self.train_percent_check)
            feed = Prefetcher(train_loader, self.device) if prefetch_enabled() else \
                (move_to(b, self.device) for b in train_loader)
            for bi, batch in enumerate(feed):
                if bi >= n_batches:
                    break
                model.on_batch_start(batch)
                pooled = self.engine is not None and self.engine.on_gpu
                if pooled:      # the step's zero-initialised accumulators come from one pre-cleared pool (raw.StepPool)
                    from . import raw
                    raw.begin_step(self.device)
                try:
                    out = model.training_step(batch, bi)
                    loss = out["loss"] if isinstance(out, dict) else out
                    loss.backward()
                    self._optimizer_step()
                finally:
                    if pooled:
                        raw.end_step(self.device)
                self.global_step += 1
                model.global_step = self.global_step
                model.on_batch_end()
                if self.rank == 0 and self.show_progress_bar and self.global_step % self.row_log_interval == 0:
                    shown = _flatten_metrics({"progress_bar": out.get("progress_bar", {})}) if isinstance(out, dict) \
                        else {}
                    shown.setdefault("loss", _scalar(loss))
                    log.info("epoch %d step %d/%d  %s", epoch, bi + 1, n_batches,
                             "  ".join("%s=%.4f" % kv for kv in sorted(shown.items())))
                if self.max_steps and self.global_step >= self.max_steps:
                    stop = True
                    break
            model.on_epoch_end()
            if val_loaders and (epoch + 1) % self.check_val_every_n_epoch == 0:
                self._evaluate(model, val_loaders, "validation",
                               max_batches=1 if self.fast_dev_run else None, pct=self.val_percent_check)
            for s in self.lr_schedulers:
                s.step()
            if self.plateau is not None and "val_loss" in self.callback_metrics:
                self.plateau.step(self.callback_metrics["val_loss"])
            self._maybe_checkpoint(model, epoch)
            validated = val_loaders and (epoch + 1) % self.check_val_every_n_epoch == 0
            if self.early_stop_callback is not None and validated and epoch >= self.min_epochs - 1 \
                    and self.early_stop_callback.on_epoch_end(epoch, self.callback_metrics):
                if self.rank == 0:
                    log.info("early stopping after epoch %d (%s did not improve for %d checks)", epoch,
                             self.early_stop_callback.monitor, self.early_stop_callback.patience)
                stop = True
            if stop:
                break

    def _evaluate(self, model, loaders, kind, max_batches=None, pct=1.0):
        step = model.validation_step if kind == "validation" else model.test_step
        end = getattr(model, "validation_end" if kind == "validation" else "test_end", None)
        was_training = model.training
        model.eval()
        outputs = []
        with torch.no_grad():
            for di, loader in enumerate(loaders):
                n = self._limit(len(loader), pct)
                if max_batches is not None:
                    n = min(n, max_batches)
                outs = []
                for bi, batch in enumerate(loader):
                    if bi >= n:
                        break
                    batch = move_to(batch, self.device)
                    outs.append(step(batch, bi, di) if len(loaders) > 1 else step(batch, bi))
                outputs.append(outs)
        outputs = outputs[0] if len(outputs) == 1 else outputs
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, move_to(outputs, torch.device("cpu")))
            outputs = _dedup_eval_outputs([o for part in gathered for o in part])
        result = None
        if end is not None:
            if self.rank == 0:
                with torch.no_grad():
                    result = end(outputs)
            if self.world > 1:
                box = [move_to(result, torch.device("cpu")) if self.rank == 0 else None]
                dist.broadcast_object_list(box, src=0)
                result = box[0]
        metrics = _flatten_metrics(result)
        self.callback_metrics.update(metrics)
        if self.rank == 0 and metrics:
            log.info("%s: %s", kind, "  ".join("%s=%.4f" % kv for kv in sorted(metrics.items())))
        model.train(was_training)
        return result

    def _test(self, model):
        loaders = model.test_dataloader()
        if not loaders:
            raise RuntimeError("test needs a test_dataloader")
        return self._evaluate(model, loaders, "test", max_batches=1 if self.fast_dev_run else None,
                              pct=self.test_percent_check)

    # ------------------------------------------------------------------ checkpoints
    def _version_dir(self):
        root = os.path.join(self.weights_save_path or self.default_save_path, "lightning_logs")
        box = [None]
        if self.rank == 0:
            os.makedirs(root, exist_ok=True)
            taken = [int(m.group(1)) for m in (re.fullmatch(r"version_(\d+)", d) for d in os.listdir(root)) if m]
            box[0] = os.path.join(root, "version_%d" % (max(taken) + 1 if taken else 0), "checkpoints")
            os.makedirs(box[0], exist_ok=True)
        if self.world > 1:
            dist.broadcast_object_list(box, src=0)
        return box[0]

    def dump_checkpoint(self, model):
        hp = getattr(model, "hparams", None)
        ckpt = {"epoch": self.current_epoch, "global_step": self.global_step,
                "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "hparams": dict(vars(hp)) if hp is not None and not isinstance(hp, dict) else hp,
                "checkpoint_callback_best": self.best_val,
                "lr_schedulers": [s.state_dict() for s in self.lr_schedulers +
                                  ([self.plateau] if self.plateau is not None else [])]}
        if self.engine is not None and self.engine.params is not None and hasattr(self.engine, "m"):
            ckpt["optimizer_states"] = [{"fused_adam": True, "steps": self.engine.steps,
                                         "exp_avg": self.engine.m.cpu(), "exp_avg_sq": self.engine.v.cpu()}]
        elif self.engine is not None and self.engine.params is not None:
            ckpt["optimizer_states"] = [self.engine.opt.state_dict()]
        else:
            ckpt["optimizer_states"] = [o.state_dict() for o in self.optimizers]
        model.on_save_checkpoint(ckpt)
        return ckpt

    def _maybe_checkpoint(self, model, epoch):
        if not self.checkpoint_callback:
            return
        if self.ckpt_dir is None:
            self.ckpt_dir = self._version_dir()
        val = self.callback_metrics.get("val_loss")
        improved = val is None or self.best_val is None or val < self.best_val
        if not improved:
            return
        if val is not None:
            self.best_val = val
        if self.rank == 0:
            path = os.path.join(self.ckpt_dir, "_ckpt_epoch_%d.ckpt" % epoch)
            torch.save(self.dump_checkpoint(model), path)
            if self.best_path and self.best_path != path and os.path.exists(self.best_path):
                os.remove(self.best_path)       # save_top_k = 1, Lightning's default
            self.best_path = path
            log.info("checkpoint: %s (val_loss=%s)", path, "n/a" if val is None else "%.4f" % val)
