"""ORACLE support (test infrastructure): generate tests/golden/*.pt from the UNMODIFIED reference modules.

Run on the build box only (needs /root/reference):   python -m oracle.make_golden
Each fixture holds: the constructor recipe, the state_dict *spec* (key -> shape) and seed (weights are
re-synthesised with oracle.ref_torch.synth_state_dict, so no weights are committed), small inputs, the
reference's outputs, and - for train-mode fixtures - the reference's parameter / input gradients for the scalar
loss  sum(out * cot)  (large gradients are stored as a strided sample plus their L2 norm).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402
from oracle.ref_torch import synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FULL_GRAD_MAX = 4096
SAMPLE = 1024


def spec_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def load_synth(module, seed):
    spec = spec_of(module)
    sd = synth_state_dict(spec, seed)
    module.load_state_dict(sd, strict=True)
    return spec


def pack_grad(g):
    g = g.detach().float().contiguous().view(-1)
    if g.numel() <= FULL_GRAD_MAX:
        return {"full": g.clone(), "norm": float(g.norm())}
    stride = g.numel() // SAMPLE
    return {"sample": g[::stride].clone(), "stride": stride, "norm": float(g.norm())}


def run_with_grads(module, fwd, inputs_requiring_grad, cot_seed):
    out = fwd()
    g = torch.Generator().manual_seed(cot_seed)
    cot = torch.randn(out.shape, generator=g)
    (out * cot).sum().backward()
    grads = {"param." + n: pack_grad(p.grad) for n, p in module.named_parameters() if p.grad is not None}
    for n, t in inputs_requiring_grad.items():
        grads["input." + n] = pack_grad(t.grad)
    return out.detach(), cot, grads


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print("%-34s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def materialise(gen):
    """Inputs too large to commit are stored as (kind, shape, seed) recipes (tests/golden_util.py mirrors this)."""
    out = {}
    for k, d in gen.items():
        g = torch.Generator().manual_seed(d["seed"])
        if d["kind"] == "randn_relu":
            out[k] = torch.randn(d["shape"], generator=g).relu_()
        elif d["kind"] == "randint_u8":
            out[k] = torch.randint(0, 256, d["shape"], generator=g, dtype=torch.uint8)
        else:
            raise KeyError(d["kind"])
    return out


def video(B, T, seed):
    return torch.randint(0, 256, (B, 3, T, 112, 112), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    rnn = _refload.load("rnn")
    tcn = _refload.load("tcn")
    att = _refload.load("att_fusion")
    resnet = _refload.load("resnet")
    backbone = _refload.load("backbone")
    model = _refload.load("model")
    utils = _refload.load("utils")

    # ---- GRU module (audio stream shape, scorer shape) ----
    for name, ctor, xs in (("gru_audio", dict(input_size=200, hidden_size=256, num_layers=2, num_classes=9, num_fcs=2),
                            (3, 5, 200)),
                           ("gru_scorer", dict(input_size=512, hidden_size=128, num_layers=1, num_classes=1, num_fcs=1),
                            (2, 6, 512)),
                           ("gru_nohead", dict(input_size=512, hidden_size=512, num_layers=2, num_classes=-1),
                            (2, 4, 512))):
        m = rnn.GRU(**ctor)
        spec = load_synth(m, 11)
        x = rnd(xs, 1).requires_grad_(True)
        out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 2)
        save(name, dict(kind="GRU", ctor=ctor, seed=11, spec=spec, inputs={"x": x.detach()}, out=out, cot=cot,
                        grads=grads))

    # ---- AttFusion ----
    m = att.AttFusion([512, 512], 128)
    spec = load_synth(m, 12)
    xa = rnd((2, 6, 512), 3).requires_grad_(True)
    xv = rnd((2, 6, 512), 4).requires_grad_(True)
    out, cot, grads = run_with_grads(m, lambda: m(xa, xv), {"x_a": xa, "x_v": xv}, 5)
    save("attfusion", dict(kind="AttFusion", ctor=dict(input_dim=[512, 512], hidden_dim=128), seed=12, spec=spec,
                           inputs={"x_a": xa.detach(), "x_v": xv.detach()}, out=out, cot=cot, grads=grads))

    # ---- TemporalConvNet (dropout 0 so train == eval; Philox masks cannot be matched) ----
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = tcn.TemporalConvNet(512, [512, 512], 3, dropout=0.0)
    spec = load_synth(m, 13)
    x = rnd((2, 512, 12), 6).requires_grad_(True)
    out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 7)
    save("tcn", dict(kind="TemporalConvNet", ctor=dict(num_inputs=512, num_channels=[512, 512], kernel_size=3,
                                                       dropout=0.0), seed=13, spec=spec,
                     inputs={"x": x.detach()}, out=out, cot=cot, grads=grads))

    # ---- ResNet trunk, eval and train ----
    for mode in ("eval", "train"):
        m = resnet.ResNet(resnet.BasicBlock, [2, 2, 2, 2], 512, zero_init_residual=True, agg_mode="ap",
                          fmap_out_size=3)
        spec = load_synth(m, 14)
        m.train(mode == "train")
        gen = {"x": dict(kind="randn_relu", shape=(16, 64, 28, 28), seed=8)}
        x = materialise(gen)["x"].requires_grad_(True)
        out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 9)
        save("resnet_trunk_" + mode, dict(kind="ResNet", mode=mode, seed=14, spec=spec, inputs_gen=gen,
                                          out=out, cot=cot, grads=grads))

    # ---- VA_3DResNet (config-1 model at reduced T), eval forward and train forward+backward ----
    ctor = dict(hiddenDim=512, frameLen=4, backend="gru", resnet_ver="v1", nClasses=9, nFCs=2)
    m = backbone.VA_3DResNet(**ctor)
    spec = load_synth(m, 15)
    m.eval()
    v = video(1, 4, 10)
    with torch.no_grad():
        out = m((v.float() - 127.5) / 127.5)
    save("va3dresnet_eval", dict(kind="VA_3DResNet", ctor=ctor, mode="eval", seed=15, spec=spec,
                                 inputs={"video_u8": v}, out=out))
    ctor = dict(hiddenDim=512, frameLen=8, backend="gru", resnet_ver="v1", nClasses=9, nFCs=2)
    m = backbone.VA_3DResNet(**ctor)
    spec = load_synth(m, 15)
    m.train()
    gen = {"video_u8": dict(kind="randint_u8", shape=(2, 3, 8, 112, 112), seed=11)}
    v = materialise(gen)["video_u8"]
    x = ((v.float() - 127.5) / 127.5)
    out, cot, grads = run_with_grads(m, lambda: m(x), {}, 12)
    save("va3dresnet_train", dict(kind="VA_3DResNet", ctor=ctor, mode="train", seed=15, spec=spec,
                                  inputs_gen=gen, out=out, cot=cot, grads=grads))

    # ---- VA_3DVGGM_Split (the AV visual stream model.py really runs), eval ----
    ctor = dict(hiddenDim=512, frameLen=4, backend="gru", split_layer=3, nClasses=-1, nFCs=2, use_mtl=True)
    m = backbone.VA_3DVGGM_Split(**ctor)
    spec = load_synth(m, 16)
    m.eval()
    v = video(2, 4, 13)
    se = rnd((2, 512, 4), 14)
    with torch.no_grad():
        out = m((v.float() - 127.5) / 127.5, se, se)
    save("vggm_split3_eval", dict(kind="VA_3DVGGM_Split", ctor=ctor, mode="eval", seed=16, spec=spec,
                                  inputs={"video_u8": v, "se_features": se}, out=out))

    # ---- AffWild2VA, audiovisual + attention fusion ----
    def av_batch(B, T, seed):
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

    def to_ref_batch(b):
        r = dict(b)
        r["video"] = r.pop("video_u8").float()
        return r

    # (a) north-star model: --backbone resnet --split_layer 5 (needs the forward-arity tolerance)
    hp = _refload.hparams(modality="audiovisual", fusion_type="attention", backbone="resnet", split_layer=5, window=4,
                          loss="ccc_mtl")
    orig_fwd = backbone.VA_3DResNet.forward
    backbone.VA_3DResNet.forward = lambda self, x, *unused: orig_fwd(self, x)   # F4: model.py:111 passes 3 args
    try:
        m = model.AffWild2VA(hp)
        spec = load_synth(m, 17)
        m.eval()
        b = av_batch(2, 4, 15)
        with torch.no_grad():
            out = m(to_ref_batch(b))
        save("av_resnet_attention_eval", dict(kind="AffWild2VA", hparams=vars(hp), mode="eval", seed=17, spec=spec,
                                              inputs=b, out=out))
        # training_step on 2 clips x 8 frames (16 frames: BatchNorm statistics over >= 256 samples everywhere)
        hp_tr = _refload.hparams(modality="audiovisual", fusion_type="attention", backbone="resnet", split_layer=5,
                                 window=8, loss="ccc_mtl")
        m = model.AffWild2VA(hp_tr)
        spec = load_synth(m, 17)
        m.train()
        b = av_batch(2, 8, 19)
        res = m.training_step(to_ref_batch(b), 0)
        loss = res["loss"]
        loss.backward()
        grads = {"param." + n: pack_grad(p.grad) for n, p in m.named_parameters() if p.grad is not None}
        with torch.no_grad():
            out_tr = m(to_ref_batch(b))
        save("av_resnet_attention_train", dict(kind="AffWild2VA", hparams=vars(hp_tr), mode="train", seed=17,
                                               spec=spec, inputs=b, out=out_tr, loss=float(loss.detach()),
                                               grads=grads))
    finally:
        backbone.VA_3DResNet.forward = orig_fwd

    # (b) the AV model model.py runs as-is: --backbone v2p_split (split_layer 3)
    hp = _refload.hparams(modality="audiovisual", fusion_type="attention", backbone="v2p_split", split_layer=3,
                          window=4, loss="ccc_mtl")
    m = model.AffWild2VA(hp)
    spec = load_synth(m, 18)
    m.eval()
    b = av_batch(2, 4, 16)
    with torch.no_grad():
        out = m(to_ref_batch(b))
    save("av_v2psplit_attention_eval", dict(kind="AffWild2VA", hparams=vars(hp), mode="eval", seed=18, spec=spec,
                                            inputs=b, out=out))

    # ---- CCC ----
    a, c = rnd((3, 40), 20), rnd((3, 40), 21) * 0.5 + 0.1
    save("ccc", dict(kind="ccc", inputs={"r1": a, "r2": c},
                     out=torch.stack([utils.concordance_cc2(a[i], c[i], "none").squeeze() for i in range(3)])))


if __name__ == "__main__":
    main()
