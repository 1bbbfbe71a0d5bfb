"""Helpers shared by the golden-fixture tests."""
import argparse
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    if "inputs_gen" in fx:      # inputs too large to commit: regenerate from (kind, shape, seed)
        fx["inputs"] = {}
        for k, d in fx["inputs_gen"].items():
            g = torch.Generator().manual_seed(d["seed"])
            if d["kind"] == "randn_relu":
                fx["inputs"][k] = torch.randn(d["shape"], generator=g).relu_()
            elif d["kind"] == "randint_u8":
                fx["inputs"][k] = torch.randint(0, 256, d["shape"], generator=g, dtype=torch.uint8)
            else:
                raise KeyError(d["kind"])
    if "inputs_recipe" in fx:   # fixtures at BASELINE sizes (oracle/make_golden_sizes.py): everything is re-drawn
        fx["inputs"] = from_recipe(fx["inputs_recipe"])
    return fx


def from_recipe(r):
    g = torch.Generator().manual_seed(r["seed"])
    B, T = r["B"], r["T"]
    if r["kind"] == "video_u8":
        return {"video_u8": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8)}
    if r["kind"] == "av_batch":     # draw order of oracle/make_golden_sizes.py::av_batch
        return {
            "video_u8": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8),
            "audio": torch.randn((B, T, 200), generator=g) * 20 - 40,
            "se_features": torch.randn((B, 512, T), generator=g),
            "label_valence": torch.rand((B, T), generator=g) * 2 - 1,
            "label_arousal": torch.rand((B, T), generator=g) * 2 - 1,
            "class_expr": torch.randint(0, 7, (B, T), generator=g),
            "expr_valid": torch.ones((B, T), dtype=torch.bool),
        }
    raise KeyError(r["kind"])


def rel_err(y, ref):
    """Range-normalised max error."""
    y = y.detach().float().cpu()
    ref = ref.detach().float().cpu()
    return float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def grad_err(g, packed):
    """Compare a gradient tensor with a packed golden gradient (full tensor or strided sample + norm)."""
    g = g.detach().float().cpu().contiguous().view(-1)
    if "full" in packed:
        ref = packed["full"]
        got = g
    else:
        ref = packed["sample"]
        got = g[::packed["stride"]][: ref.numel()]
    scale = max(float(ref.abs().max()), 1e-12)
    e = float((got - ref).abs().max()) / scale
    n = abs(float(g.norm()) - packed["norm"]) / max(packed["norm"], 1e-12)
    return max(e, n)


def hparams_ns(d):
    return argparse.Namespace(**d)


def ref_batch(inputs, device="cpu"):
    b = {}
    for k, v in inputs.items():
        if k == "video_u8":
            b["video"] = v.float().to(device)
        else:
            b[k] = v.to(device)
    return b
