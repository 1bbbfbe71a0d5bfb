#!/usr/bin/env python
"""Headline benchmark: audio-visual M3T training throughput (frames/s) on N B200s (BASELINE.json config 4).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (torchrun for N > 1)
  python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference algorithm on the host CPU cores

One "step" = forward + ccc_mtl loss + backward + gradient all-reduce (N > 1) + clip(1.0) + Adam on one synthetic
batch of `--clips` clips x 16 frames per GPU of AffWild2VA(modality=audiovisual, fusion_type=attention,
backbone=resnet, split_layer=5) - weak scaling: the per-GPU shard is fixed, the global batch is clips*N.
Prints ONE JSON line (rank 0).  `value` is measured with the batch resident in HBM; `e2e` copies every step's batch
from pinned host memory (prefetched on a side stream) and reads the loss back; `roofline` is the dominant
tensor-core kernel timed with CUDA events inside the timed region; `cpu_baseline` is the oracle port of the
reference path on the host cores (a reported baseline, not a target).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per call of the kernels that can be the dominant one, from the
# `ncu --set full` capture at this workload (profiles/r1_ncu_full_summary.md, capture prof_r1c; 256 clips x 16 frames per GPU)
NCU_DRAM_TRAFFIC_256 = {
    "wgrad-halo stem 16x56x56": 6.562e9,        # sum over its 2 passes (3 + 2 temporal taps, activation box resident)
    "stem-halo 16x56x56": 3.244e9,
    "stem nd3 16x56x56 c64->64 k5x4x1 s1": 3.248e9,
    "wgrad nd3 16x56x56 c64->64 k5x4x1 s1": 8.788e9,
    "fprop-halo nd2 1x28x28 c64->64 k1x3x3 s1": 0.772e9,
    "dgrad-halo nd2 1x28x28 c64->64 k1x3x3 s1": 0.772e9,
    "wgrad-halo nd2 1x28x28 c64->64 k1x3x3 s1": 0.826e9,
    "fprop-halo nd2 1x14x14 c128->128 k1x3x3 s1": 0.3615e9,
    "dgrad-halo nd2 1x14x14 c128->128 k1x3x3 s1": 0.3615e9,
}

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
# CPU arm: the oracle port of the reference training step (fp32, PyTorch CPU kernels, all host threads)
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, clips):
    import torch
    from oracle import ref_torch as R
    from m3t_b200.models.model import AffWild2VA
    torch.set_num_threads(os.cpu_count() or 1)
    hp = hparams()
    torch.manual_seed(12345)
    m = AffWild2VA(hp)          # host-side container only: gives reference-shaped, reference-initialised weights
    randomise_bn(m, 7)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
            params.append(v)
    opt = torch.optim.Adam(params, lr=hp.learning_rate, weight_decay=1e-4)
    batch = synth_batch(clips, 1234, pin=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        y = R.affwild2va_forward(batch, sd, hp, train=True)
        loss = R.training_loss(y, batch, hp.loss, hp.loss_lambda)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return clips * T_FRAMES / sec, sec, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    clips = 8
    fps, sec, cores = cpu_reference_steps(args.steps, args.warmup, clips)
    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config(args.clips, world), "impl": "reference",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d clips x %d frames per step (same model, loss, clip, Adam); the reference is "
                                   "pure Python/PyTorch and cannot travel, so its oracle port is timed" % (clips, T_FRAMES)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
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

    # ---------------- resident-input timing (value) ----------------
    for _ in range(args.warmup):
        engine.step(resident)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    prof = raw.KernelProfiler()
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
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = frames_per_step * args.steps / (ms * 1e-3)
    final_loss = float(loss.item())
    ksum = prof.summary()

    # ---------------- end-to-end timing (host batch -> device every step, loss read back) ----------------
    if args.no_e2e:
        if rank == 0:
            _emit({"metric": METRIC, "value": value, "ms_per_step": ms / args.steps, "profiling": True,
                   "gpu_launches": int(launches)})
        return
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(args.warmup + args.steps, dtype=torch.float32).pin_memory()

    def run_e2e(host):
        """The step through the public API with HOST buffers: H2D copy of every step's batch (pinned memory, side
        stream, prefetch depth 1) and D2H read of the loss inside the timed region."""
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]

        def issue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[i % 2])
                for k, v in host.items():
                    bufs[i % 2][k].copy_(v, non_blocking=True)
                ready[i % 2].record(copy_stream)

        for d in done:
            d.record()
        issue_copy(0)
        main = torch.cuda.current_stream()
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                barrier()
                e0.record()
            issue_copy(i + 1)
            main.wait_event(ready[i % 2])
            loss = engine.step(bufs[i % 2])
            done[i % 2].record(main)
            loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), sum(v.numel() * v.element_size() for v in host.values())

    ms_e2e, h2d = run_e2e(host)
    e2e_value = frames_per_step * args.steps / (ms_e2e * 1e-3)
    # same step fed with DECODED uint8 frames (128x128 HWC) + augmentation rows: crop / mirror / cutout / normalise
    # happen inside the stem's layout pass on the device, 3.1x fewer host->device bytes
    e2e_u8 = None
    if not args.no_u8:
        from m3t_b200.process.video_input import draw_params
        import random as _random
        import numpy as _np
        _random.seed(1234 + rank)
        _np.random.seed(1234 + rank)
        g8 = torch.Generator().manual_seed(4321 + rank)
        host_u8 = {k: v for k, v in host.items() if k != "video"}
        host_u8["video_u8"] = torch.randint(0, 256, (args.clips, T_FRAMES, 128, 128, 3), generator=g8,
                                            dtype=torch.uint8).pin_memory()
        host_u8["video_aug"] = torch.tensor([draw_params(True, bool(i & 1), True, True, 128, 112)
                                             for i in range(args.clips)], dtype=torch.int32).pin_memory()
        ms_u8, h2d_u8 = run_e2e(host_u8)
        e2e_u8 = {"value": frames_per_step * args.steps / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_u8,
                  "d2h_bytes_per_step": 4, "ms_per_step": ms_u8 / args.steps,
                  "input": "uint8 128x128x3 decoded frames + crop/mirror/cutout rows (models/dataset.py:46-80 on device)"}

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
        conv = {k: v for k, v in ksum.items() if v["flops_total"] > 0}
        tot_ms = sum(v["ms_total"] for v in conv.values())
        tot_fl = sum(v["flops_total"] for v in conv.values())
        dom_key = max(conv, key=lambda k: conv[k]["ms_total"]) if conv else None
        roof = None
        if dom_key:
            d = conv[dom_key]
            roof = {"bound": "tensor", "kernel": "tcgen05 implicit GEMM: " + dom_key,
                    "achieved": d["tflops"], "peak": peak, "unit": "TFLOP/s", "frac": d["tflops"] / peak,
                    "traffic": NCU_DRAM_TRAFFIC_256.get(dom_key) if args.clips == 256 else None,
                    "peak_source": peak_src,
                    "avg_launch_ms": d["ms_total"] / d["calls"], "launches_timed": d["calls"],
                    "all_conv_kernels": {"tflops": tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms else None,
                                         "share_of_step": tot_ms / ms if ms else None}}
        top = sorted(conv.items(), key=lambda kv: -kv[1]["ms_total"])[:args.top]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(args.clips, world),
            "clips_per_sec": value / T_FRAMES, "loss": final_loss, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "e2e_u8_frames": e2e_u8,
            "gpu_launches": int(launches), "roofline": roof,
            "kernels": [{"key": k, "calls": v["calls"], "ms_total": round(v["ms_total"], 3),
                         "tflops": round(v["tflops"], 1)} for k, v in top],
        }
        if world == 1 and not args.no_cpu_baseline:
            clips = 8
            fps, sec, cores = cpu_reference_steps(3, 1, clips)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d clips x %d frames per step, 1 warm-up + 3 timed steps of the same "
                                              "training step through oracle/ref_torch.py" % (clips, T_FRAMES)}
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
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-u8", action="store_true", help="skip the uint8-frames end-to-end leg")
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
