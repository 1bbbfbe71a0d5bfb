"""ORACLE support (test infrastructure): golden fixtures AT BASELINE.json's configuration sizes, from the UNMODIFIED
reference modules (VERDICT r1 "Next round" item 1).

Run on the build box only (needs /root/reference):   python -m oracle.make_golden_sizes [name ...]

Inputs and weights are seeded recipes (nothing large is committed): `inputs_recipe` = (kind, B, T, seed) re-drawn by
tests/golden_util.load; weights = oracle.ref_torch.synth_state_dict(spec, seed, **synth_kw).  Stored per fixture: the
reference's fp32 output `out`, the bf16-storage-emulating oracle's output `out_emu` (so the GPU box does not have to
run the CPU oracle at these sizes), and for train fixtures the loss / packed gradients.

  cfg1_va3dresnet_eval[_hard]     BASELINE config 1: VA_3DResNet(frameLen 16, gru, v1, 9 classes, 2 FCs).eval() on
                                  (2,3,16,112,112)  [SURVEY 8(d) "parity anchor"]
  cfg3_av_{resnet,v2psplit}_eval  config 3: AV attention inference, 32 clips x 32 frames (1024 frames), both backbones
  cfg4_av_resnet_eval_256x16[_hard]  config 4's batch (256 clips x 16 frames = 4096 frames) in eval mode, run clip-chunked
                                  through the reference (eval has no cross-clip op)
  vggm_tcn_{eval,train}           VA_3DVGGM(backend='tcn') — the reference's only TemporalConvNet carrier
                                  (models/backbone.py:107-111,139-141); train with the Dropout modules set to p=0
  av_v2psplit_attention_train     the AV model model.py runs as-is, training_step + backward (8 clips x 16 frames)

`_hard` = SURVEY 8(d)'s BatchNorm recipe exactly as written (gamma ~ U(0.5,1.5) on every BN, bn2 included:
synth_kw bn2_gain=1.0), i.e. the un-softened, high-gain trunk the bench itself runs.
"""
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402
from oracle import ref_torch as R  # noqa: E402
from oracle.make_golden import pack_grad, save, spec_of  # noqa: E402

CHUNK = 16


def av_batch(B, T, seed):
    """Same draw order as oracle/make_golden.py::av_batch and tests/golden_util.py (kind 'av_batch')."""
    g = torch.Generator().manual_seed(seed)
    return {
        "video_u8": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8),
        "audio": torch.randn((B, T, 200), generator=g) * 20 - 40,
        "se_features": torch.randn((B, 512, T), generator=g),
        "label_valence": torch.rand((B, T), generator=g) * 2 - 1,
        "label_arousal": torch.rand((B, T), generator=g) * 2 - 1,
        "class_expr": torch.randint(0, 7, (B, T), generator=g),
        "expr_valid": torch.ones((B, T), dtype=torch.bool),
    }


def to_ref_batch(b, lo=None, hi=None):
    r = {k: (v if lo is None else v[lo:hi]) for k, v in b.items()}
    r["video"] = r.pop("video_u8").float()
    return r


def load_synth(module, seed, **kw):
    spec = spec_of(module)
    module.load_state_dict(R.synth_state_dict(spec, seed, **kw), strict=True)
    return spec


def chunked(fn, B):
    outs = []
    for lo in range(0, B, CHUNK):
        outs.append(fn(lo, min(B, lo + CHUNK)))
    return torch.cat(outs)


def gen_cfg1(hard):
    backbone = _refload.load("backbone")
    ctor = dict(hiddenDim=512, frameLen=16, backend="gru", resnet_ver="v1", nClasses=9, nFCs=2)
    kw = dict(bn2_gain=1.0) if hard else {}
    m = backbone.VA_3DResNet(**ctor)
    spec = load_synth(m, 21, **kw)
    m.eval()
    rec = dict(kind="video_u8", B=2, T=16, seed=31)
    v = torch.randint(0, 256, (2, 3, 16, 112, 112), generator=torch.Generator().manual_seed(31), dtype=torch.uint8)
    x = (v.float() - 127.5) / 127.5
    with torch.no_grad():
        out = m(x)
        sd = R.synth_state_dict(spec, 21, **kw)
        with R.bf16_emulation():
            emu = R.va_3dresnet(x, sd, 16)
    save("cfg1_va3dresnet_eval" + ("_hard" if hard else ""),
         dict(kind="VA_3DResNet", ctor=ctor, mode="eval", seed=21, spec=spec, synth_kw=kw, inputs_recipe=rec, out=out,
              out_emu=emu))


def _av(name, B, T, seed_w, seed_in, backbone_name, split_layer, hard, with_emu=True):
    backbone = _refload.load("backbone")
    model = _refload.load("model")
    hp = _refload.hparams(modality="audiovisual", fusion_type="attention", backbone=backbone_name,
                          split_layer=split_layer, window=T, loss="ccc_mtl")
    kw = dict(bn2_gain=1.0) if hard else {}
    orig_fwd = backbone.VA_3DResNet.forward
    backbone.VA_3DResNet.forward = lambda self, x, *unused: orig_fwd(self, x)   # SURVEY F4 (3-argument call)
    try:
        m = model.AffWild2VA(hp)
        spec = load_synth(m, seed_w, **kw)
        m.eval()
        b = av_batch(B, T, seed_in)
        t0 = time.time()
        with torch.no_grad():
            out = chunked(lambda lo, hi: m(to_ref_batch(b, lo, hi)), B)
        t_ref = time.time() - t0
        emu = None
        if with_emu:
            sd = R.synth_state_dict(spec, seed_w, **kw)
            with torch.no_grad(), R.bf16_emulation():
                emu = chunked(lambda lo, hi: R.affwild2va_forward(to_ref_batch(b, lo, hi), sd, hp), B)
    finally:
        backbone.VA_3DResNet.forward = orig_fwd
    print("  reference: %.1f s for %d frames on %d threads" % (t_ref, B * T, torch.get_num_threads()))
    save(name, dict(kind="AffWild2VA", hparams=vars(hp), mode="eval", seed=seed_w, spec=spec, synth_kw=kw,
                    inputs_recipe=dict(kind="av_batch", B=B, T=T, seed=seed_in), out=out, out_emu=emu,
                    labels={k: b[k] for k in ("label_valence", "label_arousal")}))


def gen_vggm_tcn():
    backbone = _refload.load("backbone")
    ctor = dict(inputDim=512, hiddenDim=512, nLayers=2, nClasses=2, frameLen=16, backend="tcn")
    for mode in ("eval", "train"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = backbone.VA_3DVGGM(**ctor)
        spec = load_synth(m, 23)
        m.train(mode == "train")
        for d in m.modules():
            if isinstance(d, torch.nn.Dropout):
                d.p = 0.0                       # Philox masks cannot be matched
        B, T = (2, 8) if mode == "eval" else (8, 16)   # train: BN statistics over >= 128 samples everywhere
        rec = dict(kind="video_u8", B=B, T=T, seed=33)
        v = torch.randint(0, 256, (B, 3, T, 112, 112), generator=torch.Generator().manual_seed(33), dtype=torch.uint8)
        x = (v.float() - 127.5) / 127.5
        sd = R.synth_state_dict(spec, 23)
        fx = dict(kind="VA_3DVGGM", ctor=ctor, mode=mode, seed=23, spec=spec, synth_kw={}, inputs_recipe=rec,
                  dropout_p=0.0)
        if mode == "eval":
            with torch.no_grad():
                fx["out"] = m(x)
                with R.bf16_emulation():
                    fx["out_emu"] = R.va_3dvggm(x, sd, "tcn")
        else:
            out = m(x)
            cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(34))
            (out * cot).sum().backward()
            fx.update(out=out.detach(), cot=cot,
                      grads={"param." + n: pack_grad(p.grad) for n, p in m.named_parameters() if p.grad is not None})
            with torch.no_grad(), R.bf16_emulation():
                fx["out_emu"] = R.va_3dvggm(x, sd, "tcn", train=True)
        save("vggm_tcn_" + mode, fx)


def gen_v2psplit_train():
    model = _refload.load("model")
    hp = _refload.hparams(modality="audiovisual", fusion_type="attention", backbone="v2p_split", split_layer=3,
                          window=16, loss="ccc_mtl")
    m = model.AffWild2VA(hp)
    spec = load_synth(m, 24)
    m.train()
    b = av_batch(8, 16, 35)      # 128 frames: the 1x1-spatial conv5 BatchNorms see 128 samples
    res = m.training_step(to_ref_batch(b), 0)
    loss = res["loss"]
    loss.backward()
    grads = {"param." + n: pack_grad(p.grad) for n, p in m.named_parameters() if p.grad is not None}
    with torch.no_grad():
        out_tr = m(to_ref_batch(b))
    save("av_v2psplit_attention_train", dict(kind="AffWild2VA", hparams=vars(hp), mode="train", seed=24, spec=spec,
                                             synth_kw={}, inputs_recipe=dict(kind="av_batch", B=8, T=16, seed=35),
                                             out=out_tr, loss=float(loss.detach()), grads=grads))


GENERATORS = {
    "cfg1_va3dresnet_eval": lambda: gen_cfg1(False),
    "cfg1_va3dresnet_eval_hard": lambda: gen_cfg1(True),
    "cfg3_av_resnet_eval": lambda: _av("cfg3_av_resnet_eval", 32, 32, 25, 36, "resnet", 5, False),
    "cfg3_av_resnet_eval_hard": lambda: _av("cfg3_av_resnet_eval_hard", 32, 32, 25, 36, "resnet", 5, True),
    "cfg3_av_v2psplit_eval": lambda: _av("cfg3_av_v2psplit_eval", 32, 32, 26, 37, "v2p_split", 3, False),
    "cfg4_av_resnet_eval_256x16": lambda: _av("cfg4_av_resnet_eval_256x16", 256, 16, 27, 38, "resnet", 5, False),
    "cfg4_av_resnet_eval_256x16_hard": lambda: _av("cfg4_av_resnet_eval_256x16_hard", 256, 16, 27, 38, "resnet", 5,
                                                   True),
    "vggm_tcn": gen_vggm_tcn,
    "av_v2psplit_attention_train": gen_v2psplit_train,
}


def main(argv):
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    for name in (argv or list(GENERATORS)):
        t0 = time.time()
        GENERATORS[name]()
        print("  [%s: %.1f s]" % (name, time.time() - t0))


if __name__ == "__main__":
    main(sys.argv[1:])
