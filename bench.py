#!/usr/bin/env python
"""Headline benchmark: audio-visual M3T training throughput (frames/s) on N B200s (BASELINE.json config 4).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (torchrun for N > 1)
  python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference itself on the host CPU cores

One "step" = forward + ccc_mtl loss + backward + gradient all-reduce (N > 1) + clip(1.0) + Adam on one synthetic
batch of `--clips` clips x 16 frames per GPU of AffWild2VA(modality=audiovisual, fusion_type=attention,
backbone=resnet, split_layer=5).  Prints ONE JSON line (rank 0):
  value / ms_per_step   weak scaling (256 clips per GPU at every N), batch resident in HBM, CUDA events, max over ranks
  e2e                   the same step fed from pinned HOST memory every step (decoded uint8 frames + augmentation rows,
                        the product's input pipeline) with the loss read back;  e2e_f32_clips = float32 clips instead;
                        e2e_trainer_fit = through lightning.Trainer.fit / training_step (N = 1)
  strong                N > 1: BASELINE config 4 AS WRITTEN - global batch 256 split over the N GPUs (32 clips per GPU
                        at N = 8), the whole step replayed as one CUDA graph
  roofline              dominant tensor-core kernel (CUDA events around every launch inside the timed region; traffic
                        from profiles/ncu_dram_traffic.csv);  roofline_hbm = the HBM-bound families next to it
  secondary             N = 1: BASELINE configs 1, 2, 3, 5 (+ the TCN carrier), ms per pass and frames/s
  cpu_baseline          the UNMODIFIED reference (oracle/_ref) on the host cores - a reported baseline, not a target
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels that can be the dominant one, keyed by the
    profiler key, from the committed `ncu --set full` captures at this workload (256 clips x 16 frames per GPU):
    profiles/ncu_dram_traffic.csv  (columns: key, dram_bytes_per_launch, capture)."""
    out = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dram_traffic.csv")) as f:
            for line in f:
                if line.startswith("#") or not line.strip() or line.startswith("key,"):
                    continue
                key, val, cap = [x.strip() for x in line.rsplit(",", 2)]
                out[key] = (float(val), cap)
    except OSError:
        pass
    return out


METRIC = "av_m3t_train_frames_per_sec"
UNIT = "frames/s"
T_FRAMES = 16


def hparams():
    return argparse.Namespace(backbone="resnet", backend="gru", modality="audiovisual", fusion_type="attention",
                              window=T_FRAMES, loss="ccc_mtl", loss_lambda=0.5, num_hidden=512, split_layer=5,
                              num_fc_layers=2, learning_rate=5e-5, optimizer="adam")


def config(clips, world):
    return {"workload": "BASELINE config 4: full audio-visual M3T (Conv3d stem + ResNet-18 trunk + BiGRU audio + "
                        "attention fusion + BiGRU fusion head, ccc_mtl) training step fwd+bwd+allreduce+clip+Adam",
            "per_gpu_clips": clips, "frames_per_clip": T_FRAMES, "global_clips": clips * world,
            "frame": "3x112x112", "parallelism": "dp%d" % world,
            "cache": "inputs larger than L2 (%.0f MB video per step per GPU)" % (clips * 3 * T_FRAMES * 112 * 112 * 4 / 1e6)}


def synth_batch(clips, seed, pin):
    import torch
    g = torch.Generator().manual_seed(seed)
    B, T = clips, T_FRAMES
    b = {
        "video": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8).float(),
        "audio": torch.randn((B, T, 200), generator=g) * 20 - 40,
        "se_features": torch.randn((B, 512, T), generator=g),
        "label_valence": torch.rand((B, T), generator=g) * 2 - 1,
        "label_arousal": torch.rand((B, T), generator=g) * 2 - 1,
        "class_expr": torch.randint(0, 7, (B, T), generator=g),
        "expr_valid": torch.ones((B, T), dtype=torch.bool),
    }
    if pin:
        b = {k: v.pin_memory() for k, v in b.items()}
    return b


def randomise_bn(model, seed):
    """Default init zeroes every residual branch (zero_init_residual); randomise BN affine + running statistics so
    all 20 trunk convolutions carry signal."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.rand(m.bias.shape, generator=g) * 0.4 - 0.2)
            m.running_mean.copy_(torch.rand(m.running_mean.shape, generator=g) * 0.4 - 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)


# ----------------------------------------------------------------------------------------------------------
# clocks (NVML, sampled during the timed region)
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own training step on the host cores (oracle/_ref = the reference's modules vendored
# byte-for-byte by oracle/make_ref.py; falls back to the oracle port when that directory is absent)
# ----------------------------------------------------------------------------------------------------------
REF_SAMPLE_CLIPS = 16      # clips per CPU step: a bounded sample of the 256-clip batch (CPU throughput is batch-insensitive)


def _reference_model():
    """AffWild2VA from the vendored reference (None if oracle/_ref is absent): same hparams as the GPU arm."""
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_root, "models", "model.py")):
        return None
    os.environ["M3T_REFERENCE"] = ref_root
    from oracle import _refload
    backbone = _refload.load("backbone")
    model = _refload.load("model")
    # models/model.py:111 calls the visual stream with 3 arguments, VA_3DResNet.forward takes one: the
    # north-star configuration needs this one-line arity tolerance to run at all; arithmetic is untouched
    if not getattr(backbone.VA_3DResNet, "_m3t_arity", False):
        orig = backbone.VA_3DResNet.forward
        backbone.VA_3DResNet.forward = lambda self, x, *unused: orig(self, x)
        backbone.VA_3DResNet._m3t_arity = True
    hp = _refload.hparams(modality="audiovisual", fusion_type="attention", backbone="resnet", split_layer=5,
                          window=T_FRAMES, loss="ccc_mtl")
    return model.AffWild2VA(hp)


def cpu_reference_steps(steps, warmup, clips):
    """(frames/s, s/step, threads, kind): the reference's training_step + backward + clip(1.0) + Adam(5e-5, wd 1e-4) —
    models/model.py:146-218,388-390, train.py:35 — on all host cores."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(12345)
    batch = synth_batch(clips, 1234, pin=False)
    m = _reference_model()
    if m is not None:
        kind = "reference"
        randomise_bn(m, 7)
        m.train()
        params = [p for p in m.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=5e-5, weight_decay=1e-4)

        def one():
            loss = m.training_step(batch, 0)["loss"]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
    else:
        kind = "port"
        from oracle import ref_torch as R
        from m3t_b200.models.model import AffWild2VA
        hp = hparams()
        mm = AffWild2VA(hp)          # host-side container only: reference-shaped, reference-initialised weights
        randomise_bn(mm, 7)
        sd = {k: v.detach().clone() for k, v in mm.state_dict().items()}
        params = []
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
                params.append(v)
        opt = torch.optim.Adam(params, lr=hp.learning_rate, weight_decay=1e-4)

        def one():
            y = R.affwild2va_forward(batch, sd, hp, train=True)
            loss = R.training_loss(y, batch, hp.loss, hp.loss_lambda)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return clips * T_FRAMES / sec, sec, torch.get_num_threads(), kind


def _cpu_sample_text(kind, clips, warm, steps):
    what = ("the UNMODIFIED reference (oracle/_ref, vendored by oracle/make_ref.py): AffWild2VA.training_step + backward "
            "+ clip_grad_norm_(1.0) + Adam" if kind == "reference" else
            "oracle port of the reference training step (oracle/ref_torch.py; oracle/_ref absent)")
    return "%d clips x %d frames per step (a bounded sample of the 256-clip batch), %d warm-up + %d timed steps of %s" % (
        clips, T_FRAMES, warm, steps, what)


def run_reference(args, rank, world):
    if rank != 0:
        return
    clips = REF_SAMPLE_CLIPS
    fps, sec, cores, kind = cpu_reference_steps(args.steps, args.warmup, clips)
    cfg = config(args.clips, world)
    cfg["reference_arm_sample_clips_per_step"] = clips
    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "impl": "reference",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": _cpu_sample_text(kind, clips, args.warmup, args.steps)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def _hbm_class(key):
    """Which HBM-bound family a profiler key belongs to (None: tensor-core kernel)."""
    if key.startswith("hbm "):
        name = key.split()[1]
        if name.startswith(("bn_", "maxpool_bn", "add_bf16")):
            return "bn_relu_pool_passes"
        return name
    if key.startswith("att_mix"):
        return "att_mix"
    if key.startswith("gru_"):
        return "gru_recurrence"
    return None


def secondary_configs(dev):
    """BASELINE configs 1, 2, 3, 5 on ONE GPU (ms per pass, frames/s; CUDA events, 3 warm-up + 5-10 timed): parity-test
    cases by the bench contract, timed here so that they are driver-visible with the run's clocks record."""
    import torch
    from m3t_b200 import lib
    from m3t_b200.graphs import GraphedInference
    from m3t_b200.models.audio_resnet import AudioResNetTCN
    from m3t_b200.models.backbone import VA_3DResNet
    from m3t_b200.models.model import AffWild2VA
    from m3t_b200.models.vggm import VA_3DVGGM

    def timed(fn, warm=3, iters=10):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.launch_count()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, (lib.launch_count() - n0) // iters

    def hp(**kw):
        d = vars(hparams()).copy()
        d.update(window=32)
        d.update(kw)
        return argparse.Namespace(**d)

    def av_batch(B, T):
        g = torch.Generator().manual_seed(0)
        return {"video": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8).float().to(dev),
                "audio": (torch.randn((B, T, 200), generator=g) * 20 - 40).to(dev),
                "se_features": torch.randn((B, 512, T), generator=g).to(dev)}

    out = []
    torch.manual_seed(12345)
    m = VA_3DResNet(hiddenDim=512, frameLen=16, backend="gru", resnet_ver="v1", nClasses=9, nFCs=2)
    randomise_bn(m, 7)
    m = m.to(dev).eval()
    x = (torch.randint(0, 256, (2, 3, 16, 112, 112)).float().to(dev) - 127.5) / 127.5
    with torch.no_grad():
        ms, nl = timed(lambda: m(x))
        g = GraphedInference(m, x)
        msg, _ = timed(lambda: g(x))
    out.append({"config": 1, "what": "VA_3DResNet eval forward, 2 clips x 16 frames", "ms": round(ms, 4),
                "frames_per_s": round(32 / ms * 1e3, 1), "launches": nl, "ms_cuda_graph_replay": round(msg, 4),
                "frames_per_s_cuda_graph": round(32 / msg * 1e3, 1)})
    del m, g
    m = AudioResNetTCN(dropout=0.0)
    randomise_bn(m, 7)
    m = m.to(dev).train()
    a = (torch.randn(64, 32, 200) * 20 - 40).to(dev)

    def step2():
        m.zero_grad(set_to_none=True)
        m(a).square().mean().backward()

    ms, nl = timed(step2)
    out.append({"config": 2, "what": "audio ResNet over log-Mel windows + TCN head, forward+backward, 64 clips x 32 frames",
                "ms": round(ms, 4), "frames_per_s": round(64 * 32 / ms * 1e3, 1), "launches": nl})
    del m
    # the reference's only TemporalConvNet carrier, training step (fwd + bwd), dropout 0.2 on the device
    m = VA_3DVGGM(inputDim=512, hiddenDim=512, nLayers=2, nClasses=2, frameLen=32, backend="tcn")
    randomise_bn(m, 7)
    m = m.to(dev).train()
    xv = (torch.randint(0, 256, (16, 3, 32, 112, 112)).float().to(dev) - 127.5) / 127.5

    def step_tcn():
        m.zero_grad(set_to_none=True)
        m(xv).square().mean().backward()

    ms, nl = timed(step_tcn, iters=5)
    out.append({"config": "T1", "what": "VA_3DVGGM(backend=tcn) forward+backward, 16 clips x 32 frames",
                "ms": round(ms, 4), "frames_per_s": round(16 * 32 / ms * 1e3, 1), "launches": nl})
    del m, xv
    for name, h in (("v2p_split", hp(backbone="v2p_split", split_layer=3)), ("resnet", hp())):
        torch.manual_seed(12345)
        m = AffWild2VA(h)
        randomise_bn(m, 7)
        m = m.to(dev).eval()
        b = av_batch(32, 32)
        with torch.no_grad():
            ms, nl = timed(lambda: m(b))
        out.append({"config": 3, "what": "AV attention inference, 32 clips x 32 frames, backbone " + name,
                    "ms": round(ms, 4), "frames_per_s": round(1024 / ms * 1e3, 1), "launches": nl})
        del m, b
    for T in (64, 256):
        torch.manual_seed(12345)
        m = AffWild2VA(hp(window=T))
        randomise_bn(m, 7)
        m = m.to(dev).eval()
        b = av_batch(16, T)
        with torch.no_grad():
            ms, nl = timed(lambda: m(b), iters=5)
        out.append({"config": 5, "what": "AV attention eval, 16 clips per GPU x T=%d (batch 128 on 8 GPUs)" % T,
                    "ms": round(ms, 4), "frames_per_s": round(16 * T / ms * 1e3, 1), "launches": nl})
        del m, b
    torch.cuda.empty_cache()
    return out


def trainer_e2e(dev, clips, steps):
    """The same step driven through the reference's call surface: lightning.Trainer.fit -> AffWild2VA.training_step
    (with the reference's .item() syncs) -> backward -> engine; batches come from pinned host memory every step."""
    import torch
    from m3t_b200 import lightning as pl
    from m3t_b200.models.model import AffWild2VA
    hp = hparams()
    hp.scheduler, hp.freeze_enc, hp.test_lr = "none", False, False
    batches = [synth_batch(clips, 100 + i, pin=True) for i in range(2)]

    class _Steps(torch.utils.data.Dataset):
        def __len__(self):
            return steps + 3

        def __getitem__(self, i):
            return batches[i % 2]

    class Timed(AffWild2VA):
        def __init__(self, hparams_):
            super().__init__(hparams_)
            self.events = []

        def on_batch_end(self):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.events.append(e)

        @pl.data_loader
        def train_dataloader(self):
            return torch.utils.data.DataLoader(_Steps(), batch_size=None, pin_memory=False)

        @pl.data_loader
        def val_dataloader(self):
            return None

    torch.manual_seed(12345)
    m = Timed(hp)
    randomise_bn(m, 7)
    tr = pl.Trainer(gradient_clip_val=1.0, max_epochs=1, gpus=str(dev.index), nb_sanity_val_steps=0,
                    checkpoint_callback=False, early_stop_callback=False, show_progress_bar=False,
                    distributed_backend="dp")
    tr.fit(m)
    torch.cuda.synchronize()
    ms = m.events[2].elapsed_time(m.events[-1]) / (len(m.events) - 3)
    h2d = sum(v.numel() * v.element_size() for v in batches[0].values())
    del m, tr
    torch.cuda.empty_cache()
    return {"value": clips * T_FRAMES / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 12, "api": "m3t_b200.lightning.Trainer.fit (prefetching feed) -> "
            "AffWild2VA.training_step -> TrainEngine (1 GPU per process)"}


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from m3t_b200 import lib, raw
    from m3t_b200.engine import TrainEngine
    from m3t_b200.models.model import AffWild2VA

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    hp = hparams()
    torch.manual_seed(12345)
    model = AffWild2VA(hp)
    randomise_bn(model, 7)
    model = model.to(dev).train()
    engine = TrainEngine(model, lr=hp.learning_rate, weight_decay=1e-4, clip=1.0)
    host = synth_batch(args.clips, 1234 + rank, pin=True)
    resident = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    frames_per_step = args.clips * T_FRAMES * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- resident-input timing (value) ----------------
    def dominant_key(summary):
        conv = {k: v for k, v in summary.items() if v["flops_total"] > 0}
        return max(conv, key=lambda k: conv[k]["ms_total"]) if conv else None

    # the last warm-up step runs with events around every tensor-core launch: it names the dominant kernel
    for _ in range(args.warmup - 1):
        engine.step(resident)
    prof_w = raw.KernelProfiler(track_hbm=False)
    raw.set_profiler(prof_w)
    engine.step(resident)
    raw.set_profiler(None)
    torch.cuda.synchronize()
    dom_key = dominant_key(prof_w.summary())
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # inside the timed region only the dominant kernel's launches carry CUDA events (roofline.achieved is measured
    # live, on the launching stream); event pairs around ALL ~230 tensor-core launches of a step cost ~0.8 ms of the
    # headline and around the ~60 BN / pooling passes another 0.5-0.7 ms, so the per-kernel table ("kernels",
    # roofline_hbm) comes from a separate short pass below
    prof = raw.KernelProfiler(track_hbm=False, only={dom_key} if dom_key else set())
    raw.set_profiler(prof)
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = engine.step(resident)
    e1.record()
    barrier()
    raw.set_profiler(None)
    clocks = sampler.result()
    launches = lib.launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = frames_per_step * args.steps / (ms * 1e-3)
    final_loss = float(loss.item())
    dsum = prof.summary()
    hbm_steps = 3
    prof_h = raw.KernelProfiler(track_hbm=True)
    raw.set_profiler(prof_h)
    eh0, eh1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eh0.record()
    for _ in range(hbm_steps):
        engine.step(resident)
    eh1.record()
    torch.cuda.synchronize()
    raw.set_profiler(None)
    hsum, ms_h = prof_h.summary(), eh0.elapsed_time(eh1)
    ksum = hsum                 # every kernel family, from the fully instrumented pass (hbm_steps steps, ms_h)

    if args.no_e2e:
        if rank == 0:
            _emit({"metric": METRIC, "value": value, "ms_per_step": ms / args.steps, "profiling": True,
                   "gpu_launches": int(launches)})
        return
    # ---------------- end-to-end timing (host batch -> device every step, loss read back) ----------------
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(args.warmup + args.steps, dtype=torch.float32).pin_memory()

    def run_e2e(eng, host_batch, n_warm, n_steps):
        """The step through the public API (TrainEngine.step) with HOST buffers: H2D copy of every step's batch (pinned
        memory, side stream, prefetch depth 1) and D2H read of the loss inside the timed region."""
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in host_batch.items()} for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]

        def issue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i % 2])
                for k, v in host_batch.items():
                    bufs[i % 2][k].copy_(v, non_blocking=True)
                ready[i % 2].record(copy_stream)

        for d in done:
            d.record()
        issue_copy(0)
        main = torch.cuda.current_stream()
        for i in range(n_warm + n_steps):
            if i == n_warm:
                barrier()
                e0.record()
            issue_copy(i + 1)
            main.wait_event(ready[i % 2])
            ls = eng.step(bufs[i % 2])
            done[i % 2].record(main)
            loss_host[i:i + 1].copy_(ls.reshape(1), non_blocking=True)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), sum(v.numel() * v.element_size() for v in host_batch.values())

    def u8_batch(clips, host_f32):
        """DECODED uint8 frames (128x128 HWC) + augmentation rows: crop / mirror / cutout / normalise happen inside the
        stem's layout pass on the device (`--device_augment` of the dataset)."""
        from m3t_b200.process.video_input import draw_params
        import random as _random
        import numpy as _np
        _random.seed(1234 + rank)
        _np.random.seed(1234 + rank)
        g8 = torch.Generator().manual_seed(4321 + rank)
        hb = {k: v[:clips] for k, v in host_f32.items() if k != "video"}
        hb = {k: v.clone().pin_memory() for k, v in hb.items()}
        hb["video_u8"] = torch.randint(0, 256, (clips, T_FRAMES, 128, 128, 3), generator=g8,
                                       dtype=torch.uint8).pin_memory()
        hb["video_aug"] = torch.tensor([draw_params(True, bool(i & 1), True, True, 128, 112)
                                        for i in range(clips)], dtype=torch.int32).pin_memory()
        return hb

    ms_u8, h2d_u8 = run_e2e(engine, u8_batch(args.clips, host), args.warmup, args.steps)
    e2e = {"value": frames_per_step * args.steps / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_u8,
           "d2h_bytes_per_step": 4, "ms_per_step": ms_u8 / args.steps,
           "input": "decoded uint8 128x128x3 frames + per-clip crop/mirror/cutout rows from pinned host memory "
                    "(what the dataset hands over with --device_augment; models/dataset.py:46-80 runs on the device)",
           "api": "m3t_b200.engine.TrainEngine.step"}
    e2e_f32 = None
    if not args.no_f32_e2e:
        ms_e2e, h2d = run_e2e(engine, host, args.warmup, args.steps)
        e2e_f32 = {"value": frames_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                   "input": "float32 (B,3,T,112,112) clips as the reference's dataset yields them (models/dataset.py:322)"}

    # ---------------- BASELINE config 4 as written: GLOBAL batch 256 split over the N GPUs (strong scaling) ----------
    strong = None
    if (world > 1 or args.shard_clips) and not args.no_strong and 256 % world == 0:
        sc = args.shard_clips or 256 // world
        sh_host = {k: v[:sc].clone().pin_memory() for k, v in host.items()}
        sh_res = {k: v.to(dev) for k, v in sh_host.items()}
        graphed = not args.no_graph
        if graphed:
            engine.capture(sh_res, warmup=3)
        for _ in range(args.warmup):
            engine.step(sh_res)
        barrier()
        l0 = lib.launch_count()
        e0.record()
        for _ in range(args.steps):
            engine.step(sh_res)
        e1.record()
        barrier()
        ms_s = max_over_ranks(e0.elapsed_time(e1))
        l_s = lib.launch_count() - l0
        u8h = u8_batch(sc, host)
        if graphed:      # the uint8 batch has other keys than the float batch: capture that signature
            engine.capture({k: v.to(dev) for k, v in u8h.items()}, warmup=2)
        ms_su8, h2d_s = run_e2e(engine, u8h, args.warmup, args.steps)
        gclips = sc * world
        strong = {"scaling": "strong", "global_clips": gclips, "per_gpu_clips": sc,
                  "value": gclips * T_FRAMES * args.steps / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s / args.steps,
                  "e2e": {"value": gclips * T_FRAMES * args.steps / (ms_su8 * 1e-3), "unit": UNIT,
                          "h2d_bytes_per_step": h2d_s, "d2h_bytes_per_step": 4, "ms_per_step": ms_su8 / args.steps},
                  "cuda_graph": graphed, "host_launches_per_step": l_s / args.steps,
                  "ideal_ms_per_step": (ms / args.steps) * sc / args.clips,
                  "note": "efficiency vs N=1 = (N=1 ms_per_step at 256 clips) / (N x this ms_per_step); the whole step "
                          "(fwd, loss, bwd, NCCL all-reduce, clip, Adam) is one CUDA-graph launch per replay"}
        engine.graph = None

    def guarded(what, fn):
        """The headline (value / e2e / roofline) is measured by now: a failure in one of the explanatory single-GPU
        legs is reported in its place in the JSON line (and on stderr) instead of costing the whole line."""
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            return {"error": "%s failed: %r" % (what, exc)}

    trainer_leg = None
    if world == 1 and not args.no_trainer:
        del engine, model, resident
        torch.cuda.empty_cache()
        trainer_leg = guarded("e2e_trainer_fit", lambda: trainer_e2e(dev, args.clips, args.steps))
    second = None
    if world == 1 and not args.no_secondary:
        second = guarded("secondary", lambda: secondary_configs(dev))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        if not peak:
            peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
        hbm_peak = peaks.get("hbm_gbs")
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        if not hbm_peak:
            hbm_peak, hbm_src = 6500.0, "fallback (B200_PROFILING.md copy bandwidth)"
        traffic = load_ncu_traffic()
        conv = {k: v for k, v in ksum.items() if v["flops_total"] > 0}
        tot_ms = sum(v["ms_total"] for v in conv.values())
        tot_fl = sum(v["flops_total"] for v in conv.values())
        roof = None
        if dom_key and dom_key in dsum:
            d = dsum[dom_key]        # the dominant kernel's launches inside the headline region
            tr_ = traffic.get(dom_key) if args.clips == 256 else None
            roof = {"bound": "tensor", "kernel": "tcgen05 implicit GEMM: " + dom_key,
                    "achieved": d["tflops"], "peak": peak, "unit": "TFLOP/s", "frac": d["tflops"] / peak,
                    "traffic": tr_[0] if tr_ else None,
                    "traffic_source": ("profiles/ncu_dram_traffic.csv <- " + tr_[1]) if tr_ else None,
                    "peak_source": peak_src,
                    "avg_launch_ms": d["ms_total"] / d["calls"], "launches_timed": d["calls"],
                    "all_conv_kernels": {"tflops": tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms else None,
                                         "share_of_step": tot_ms / ms_h if ms_h else None,
                                         "from": "the fully instrumented pass (%d steps)" % hbm_steps}}
        # HBM-bound families: algorithmic bytes (each operand of a streaming pass once) / event time
        fam = {}
        for k, v in hsum.items():
            c = _hbm_class(k)
            if c is None or v["bytes_total"] <= 0:
                continue
            f = fam.setdefault(c, {"ms_total": 0.0, "bytes_total": 0.0, "calls": 0})
            f["ms_total"] += v["ms_total"]
            f["bytes_total"] += v["bytes_total"]
            f["calls"] += v["calls"]
        hbm_block = None
        if fam:
            dom = max(fam, key=lambda c: fam[c]["ms_total"])
            rows = {c: {"launches_per_step": f["calls"] / hbm_steps, "ms_per_step": round(f["ms_total"] / hbm_steps, 4),
                        "achieved_GBps": round(f["bytes_total"] / (f["ms_total"] * 1e-3) / 1e9, 1),
                        "frac": round(f["bytes_total"] / (f["ms_total"] * 1e-3) / 1e9 / hbm_peak, 4),
                        "share_of_step": round(f["ms_total"] / ms_h, 4)} for c, f in fam.items()}
            d = fam[dom]
            hbm_block = {"bound": "hbm", "kernel": dom, "achieved": d["bytes_total"] / (d["ms_total"] * 1e-3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s",
                         "frac": d["bytes_total"] / (d["ms_total"] * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "peak_source": hbm_src, "families": rows, "steps_timed": hbm_steps,
                         "note": "gru_recurrence is latency-bound (T serial steps), its GB/s is reported, not a target"}
        top = sorted(conv.items(), key=lambda kv: -kv[1]["ms_total"])[:args.top]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(args.clips, world),
            "clips_per_sec": value / T_FRAMES, "loss": final_loss, "clocks": clocks,
            "e2e": e2e, "e2e_f32_clips": e2e_f32, "e2e_trainer_fit": trainer_leg,
            "gpu_launches": int(launches), "roofline": roof, "roofline_hbm": hbm_block, "strong": strong,
            "secondary": second,
            "kernels": [{"key": k, "calls_per_step": v["calls"] / hbm_steps,
                         "ms_per_step": round(v["ms_total"] / hbm_steps, 3),
                         "tflops": round(v["tflops"], 1)} for k, v in top],
            "kernels_from": "a separate fully instrumented pass of %d steps (%.2f ms per step with its ~290 event "
                            "pairs); the headline region times only the dominant kernel" % (hbm_steps, ms_h / hbm_steps),
        }
        if world == 1 and not args.no_cpu_baseline:
            def cpu_leg():
                fps, sec, cores, kind = cpu_reference_steps(3, 1, REF_SAMPLE_CLIPS)
                return {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": _cpu_sample_text(kind, REF_SAMPLE_CLIPS, 1, 3)}
            line["cpu_baseline"] = guarded("cpu_baseline", cpu_leg)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU (x16 frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiling runs)")
    ap.add_argument("--no-f32-e2e", action="store_true", help="skip the float32-clips end-to-end leg")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the global-batch-256 (strong scaling) block")
    ap.add_argument("--shard-clips", type=int, default=0, help="diagnostic: time the strong-scaling shard of this many "
                    "clips on the GPUs given (e.g. 32 = one rank's share at N = 8, without the all-reduce when N = 1)")
    ap.add_argument("--no-graph", action="store_true", help="strong block: eager step instead of the CUDA-graph replay")
    ap.add_argument("--no-trainer", action="store_true", help="N = 1: skip the Trainer.fit end-to-end leg")
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: skip BASELINE configs 1, 2, 3, 5")
    ap.add_argument("--top", type=int, default=8, help="how many tensor-core kernels to list under \"kernels\"")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


def _emit(obj):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else written to fd 1 while the bench runs
    (NCCL's version banner, library chatter) has been redirected to stderr so the line stays alone."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
