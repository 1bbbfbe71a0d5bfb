"""ORACLE support (test infrastructure): golden fixtures for the rest of the reference's model zoo,
from the UNMODIFIED reference modules.  Run on the build box only:   python -m oracle.make_golden_zoo [name ...]

  resnetv2_trunk_{eval,train}   ResNetV2(BasicBlockV2, [2,2,2,2], 512, agg_mode='ap')      models/resnet.py:127-251
  va3dresnet_v2_eval            VA_3DResNet(resnet_ver='v2') — the constructor's default     models/backbone.py:336-337
  attencdec_{eval,train}        AttEncDec (Attention + Decoder, fusion_type 'att_dec')       models/rnn.py:84-165
  cbam_{eval,train}             CBAM(128) (ChannelGate + SpatialGate)                        models/cbam.py:1-112
  resnet_cbam_train             ResNet(BasicBlock, [1,1,1,1], use_cbam=True)                  models/resnet.py:32-35,48-49
  vggface_{eval,train}          VGGFace (13 conv + ReLU, ceil-mode pooling, fc1)             models/vggface.py:7-50
  densenet_{eval,train}         DenseNet52_3D(392, agg_mode='ap') on (4,64,8,28,28)          models/densenet.py:5-93
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402
from oracle import ref_torch as R  # noqa: E402
from oracle.make_golden import materialise, rnd, run_with_grads, save, spec_of, video  # noqa: E402


def load_synth(module, seed, **kw):
    spec = spec_of(module)
    module.load_state_dict(R.synth_state_dict(spec, seed, **kw), strict=True)
    return spec


def gen_resnetv2():
    resnet = _refload.load("resnet")
    for mode in ("eval", "train"):
        m = resnet.ResNetV2(resnet.BasicBlockV2, [2, 2, 2, 2], 512, zero_init_residual=False, agg_mode="ap",
                            fmap_out_size=3)
        spec = load_synth(m, 51)
        m.train(mode == "train")
        gen = {"x": dict(kind="randn_relu", shape=(16, 64, 28, 28), seed=52)}
        x = materialise(gen)["x"].requires_grad_(True)
        out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 53)
        save("resnetv2_trunk_" + mode, dict(kind="ResNetV2", mode=mode, seed=51, spec=spec, inputs_gen=gen, out=out,
                                            cot=cot, grads=grads))


def gen_va3dresnet_v2():
    backbone = _refload.load("backbone")
    ctor = dict(hiddenDim=512, frameLen=4, backend="gru", resnet_ver="v2", nClasses=9, nFCs=2)
    m = backbone.VA_3DResNet(**ctor)
    spec = load_synth(m, 54)
    m.eval()
    v = video(2, 4, 55)
    with torch.no_grad():
        out = m((v.float() - 127.5) / 127.5)
    save("va3dresnet_v2_eval", dict(kind="VA_3DResNet", ctor=ctor, mode="eval", seed=54, spec=spec,
                                    inputs={"video_u8": v}, out=out))


def gen_attencdec():
    rnn = _refload.load("rnn")
    m = rnn.AttEncDec()
    spec = load_synth(m, 56)
    x = rnd((3, 10, 1024), 57).requires_grad_(True)
    m.eval()
    with torch.no_grad():
        out_eval = m(x)
    save("attencdec_eval", dict(kind="AttEncDec", mode="eval", seed=56, spec=spec, inputs={"x": x.detach()},
                                out=out_eval))
    m.train()
    out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 58)
    save("attencdec_train", dict(kind="AttEncDec", mode="train", seed=56, spec=spec, inputs={"x": x.detach()}, out=out,
                                 cot=cot, grads=grads))


def gen_cbam():
    cbam = _refload.load("cbam")
    resnet = _refload.load("resnet")
    for mode in ("eval", "train"):
        m = cbam.CBAM(128)
        spec = load_synth(m, 59)
        m.train(mode == "train")
        gen = {"x": dict(kind="randn_relu", shape=(4, 128, 14, 14), seed=60)}
        x = materialise(gen)["x"].requires_grad_(True)
        out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 61)
        save("cbam_" + mode, dict(kind="CBAM", ctor=dict(gate_channels=128), mode=mode, seed=59, spec=spec,
                                  inputs_gen=gen, out=out, cot=cot, grads=grads))
    m = resnet.ResNet(resnet.BasicBlock, [1, 1, 1, 1], 512, zero_init_residual=True, agg_mode="ap", fmap_out_size=3,
                      use_cbam=True)
    spec = load_synth(m, 62)
    m.train()
    gen = {"x": dict(kind="randn_relu", shape=(16, 64, 28, 28), seed=63)}
    x = materialise(gen)["x"].requires_grad_(True)
    out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 64)
    save("resnet_cbam_train", dict(kind="ResNetCBAM", mode="train", seed=62, spec=spec, inputs_gen=gen, out=out,
                                   cot=cot, grads=grads))


def gen_vggface():
    vgg = _refload.load("vggface")
    for mode in ("eval", "train"):
        m = vgg.VGGFace()
        spec = load_synth(m, 65)
        m.train(mode == "train")
        m.dropout.p = 0.0                      # Philox masks cannot be matched; the dropout kernel has its own case
        v = video(2, 1, 66)[:, :, 0]           # (2,3,112,112) uint8
        x = (v.float() - 127.5) / 127.5
        fx = dict(kind="VGGFace", mode=mode, seed=65, spec=spec, inputs={"image_u8": v}, dropout_p=0.0)
        if mode == "eval":
            with torch.no_grad():
                fx["out"] = m(x)
        else:
            out, cot, grads = run_with_grads(m, lambda: m(x), {}, 67)
            fx.update(out=out, cot=cot, grads=grads)
        save("vggface_" + mode, fx)


def gen_densenet():
    dn = _refload.load("densenet")
    for mode in ("eval", "train"):
        m = dn.DenseNet52_3D(392, agg_mode="ap", fmap_out_size=3)
        spec = load_synth(m, 68)
        m.train(mode == "train")
        gen = {"x": dict(kind="randn_relu", shape=(4, 64, 8, 28, 28), seed=69)}     # 32 frames: 288 samples in the 3x3 stage
        x = materialise(gen)["x"].requires_grad_(True)
        out, cot, grads = run_with_grads(m, lambda: m(x), {"x": x}, 70)
        save("densenet_" + mode, dict(kind="DenseNet52_3D", mode=mode, seed=68, spec=spec, inputs_gen=gen, out=out,
                                      cot=cot, grads=grads))


GENERATORS = {"resnetv2": gen_resnetv2, "va3dresnet_v2": gen_va3dresnet_v2, "attencdec": gen_attencdec,
              "cbam": gen_cbam, "vggface": gen_vggface, "densenet": gen_densenet}


def main(argv):
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    for name in (argv or list(GENERATORS)):
        t0 = time.time()
        GENERATORS[name]()
        print("  [%s: %.1f s]" % (name, time.time() - t0))


if __name__ == "__main__":
    main(sys.argv[1:])
