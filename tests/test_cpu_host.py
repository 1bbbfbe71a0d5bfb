"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/m3t_b200.h declares (no
compute calls), the module mirrors carry the reference's state_dict contract, the product never imports the oracle,
and the data-parallel gradient reduction is correct under gloo with world_size 2."""
import argparse
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_symbols_exported():
    from m3t_b200 import lib as L
    h = L.load()
    syms = L.header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(h, s)]
    assert not missing, missing
    assert h.m3t_abi_version() == 1


def test_header_cites_reference():
    txt = open(os.path.join(ROOT, "include", "m3t_b200.h")).read()
    assert len(re.findall(r"models/[a-z_]+\.py:\d+", txt)) >= 15


@pytest.mark.parametrize("name", ["gru_audio", "gru_scorer", "gru_nohead", "attfusion", "tcn", "resnet_trunk_eval",
                                  "va3dresnet_eval", "vggm_split3_eval", "av_resnet_attention_eval",
                                  "av_v2psplit_attention_eval"])
def test_state_dict_contract(name):
    """Same keys and shapes as the reference module the fixture was generated from (strict load both ways)."""
    from oracle.ref_torch import synth_state_dict
    from tests.golden_util import load
    fx = load(name)
    kind = fx["kind"]
    if kind == "GRU":
        from m3t_b200.models.rnn import GRU
        m = GRU(**fx["ctor"])
    elif kind == "AttFusion":
        from m3t_b200.models.att_fusion import AttFusion
        m = AttFusion(**fx["ctor"])
    elif kind == "TemporalConvNet":
        from m3t_b200.models.tcn import TemporalConvNet
        m = TemporalConvNet(**fx["ctor"])
    elif kind == "ResNet":
        from m3t_b200.models.resnet import BasicBlock, ResNet
        m = ResNet(BasicBlock, [2, 2, 2, 2], 512, zero_init_residual=True, agg_mode="ap", fmap_out_size=3)
    elif kind == "VA_3DResNet":
        from m3t_b200.models.backbone import VA_3DResNet
        m = VA_3DResNet(**fx["ctor"])
    elif kind == "VA_3DVGGM_Split":
        from m3t_b200.models.vggm import VA_3DVGGM_Split
        m = VA_3DVGGM_Split(**fx["ctor"])
    else:
        from m3t_b200.models.model import AffWild2VA
        m = AffWild2VA(argparse.Namespace(**fx["hparams"]))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == fx["spec"]
    m.load_state_dict(synth_state_dict(fx["spec"], fx["seed"]), strict=True)


def test_reference_init_distributions():
    """zero_init_residual zeroes bn2.weight; GRU biases are zero and W_hh blocks orthogonal (models/rnn.py:57-69)."""
    from m3t_b200.models.backbone import VA_3DResNet
    m = VA_3DResNet(nClasses=9, nFCs=2, frameLen=4, resnet_ver="v1")
    assert float(m.resnet.layer1[0].bn2.weight.abs().max()) == 0.0
    assert float(m.resnet.layer1[0].bn1.weight.min()) == 1.0
    w = m.gru.gru.weight_hh_l0[:512]
    assert torch.allclose(w @ w.t(), torch.eye(512), atol=1e-4)
    assert float(m.gru.gru.bias_ih_l0.abs().max()) == 0.0
    c = m.c3d[0].weight
    assert abs(float(c.std()) - (2.0 / (5 * 7 * 7 * 64)) ** 0.5) < 2e-3


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "m3f.pytorch_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn


def test_ops_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from m3t_b200.models.rnn import GRU
    m = GRU(200, 256, 1, 2)
    with pytest.raises(Exception):
        m(torch.randn(1, 3, 200))


def test_stem_index_maps():
    from m3t_b200 import ops
    idx = ops.stem_s2d_index("cpu")
    valid = idx[idx >= 0]
    assert idx.numel() == 80 * 16 and valid.numel() == 3 * 5 * 7 * 7 and valid.unique().numel() == valid.numel()
    v = ops.vggm_s2d_index("cpu")
    vv = v[v >= 0]
    assert v.numel() == 12 * 16 and vv.numel() == 81 and vv.unique().numel() == 81


def test_logmel_oracle_vs_torchaudio():
    ta = pytest.importorskip("torchaudio")
    import numpy as np
    from oracle import melspec as OM
    from m3t_b200.process.extract_melspec import mel_filterbank
    assert np.abs(mel_filterbank() - OM.mel_filters()).max() < 1e-6
    fps = 30.0
    hop = int(1 / 3 * 1 / fps * 16000)
    y = torch.randn(32000, generator=torch.Generator().manual_seed(0)) * 0.1
    ms = ta.transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=hop, f_min=0, f_max=8000, n_mels=40,
                                      power=2.0, norm="slaney", mel_scale="slaney", center=True, pad_mode="constant")
    ref = ta.transforms.AmplitudeToDB("power", top_db=80)(ms(y)).t().numpy()
    assert np.abs(ref - OM.logmel(y.numpy(), fps)).max() < 1e-3


def test_data_parallel_gradient_allreduce_gloo():
    """world_size 2 on CPU/gloo: after TrainEngine's reduction every rank holds the mean of the per-rank gradients."""
    script = os.path.join(ROOT, "tests", "dp_gloo_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script], cwd=ROOT, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("DP_OK") == 2, r.stdout[-2000:]


def test_stride2_dgrad_parity_decomposition_host_logic():
    """ops._parity_taps (which filter taps / paddings each output parity of a stride-2 data gradient uses) replayed with
    plain CPU convolutions: the four scattered sub-convolutions must equal conv_transpose2d."""
    import torch.nn.functional as F
    from m3t_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (H, K, pad) in ((28, 3, 1), (7, 3, 1), (14, 1, 0), (7, 1, 0), (12, 7, 3)):
        Cin, Cout, N = 3, 4, 2
        P = (H + 2 * pad - K) // 2 + 1
        w = torch.randn((Cout, Cin, K, K), generator=g)
        dy = torch.randn((N, Cout, P, P), generator=g)
        ref = F.conv_transpose2d(dy, w, stride=2, padding=pad, output_padding=H - ((P - 1) * 2 - 2 * pad + K))
        # flipped dgrad pack [Cin][taps (flipped)][Cout], as m3t_pack_filter lays it out
        wd = w.permute(1, 2, 3, 0).reshape(Cin, K * K, Cout).flip(1)
        dx = torch.zeros((N, Cin, H, H))
        for ph in range(2):
            for pw in range(2):
                tp = ops._parity_taps(K, pad, ph, pw, "cpu")
                Ha, Wa = (H - ph + 1) // 2, (H - pw + 1) // 2
                if tp is None or Ha == 0 or Wa == 0:
                    continue
                idx, nh, nw, plh, plw = tp
                phh, pwh = Ha - P - plh + nh - 1, Wa - P - plw + nw - 1
                assert min(phh, pwh, plh, plw) >= 0
                wsub = wd.index_select(1, idx).reshape(Cin, nh, nw, Cout).permute(0, 3, 1, 2)   # [Cin][Cout][nh][nw]
                sub = F.conv2d(F.pad(dy, (plw, pwh, plh, phh)), wsub)
                assert sub.shape[-2:] == (Ha, Wa)
                dx[:, :, ph::2, pw::2] = sub
        assert (dx - ref).abs().max() < 1e-4, (H, K, pad, float((dx - ref).abs().max()))


def test_split_operand_arithmetic_host_model():
    """The arithmetic the fp32-parity mode relies on, modelled on the CPU: x = hi + lo with bf16 pieces, and
    a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi with exact fp32 products - relative error ~2^-16 (three terms) and ~2^-23
    (three pieces, six terms), against float64."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(64, 2048, generator=g)
    b = torch.randn(32, 2048, generator=g)
    ref = a.double() @ b.double().t()

    def pieces(x, n):
        out, r = [], x.clone()
        for _ in range(n):
            p = r.bfloat16().float()
            out.append(p)
            r = r - p
        return out

    a2, b2 = pieces(a, 2), pieces(b, 2)
    assert (a - a2[0] - a2[1]).abs().max() <= a.abs().max() * 2.0 ** -16
    three = sum((x.double() @ y.double().t()) for x, y in ((a2[0], b2[1]), (a2[1], b2[0]), (a2[0], b2[0])))
    a3, b3 = pieces(a, 3), pieces(b, 3)
    six = sum((a3[i].double() @ b3[j].double().t()) for i, j in ((0, 2), (1, 1), (2, 0), (0, 1), (1, 0), (0, 0)))
    scale = ref.abs().max()
    assert (three - ref).abs().max() / scale < 2.0 ** -14
    assert (six - ref).abs().max() / scale < 2.0 ** -21


def test_dropout_oracle_statistics():
    """oracle/dropout.py (the restatement m3t_dropout_bf16 is checked against bit for bit on the GPU): keep rate,
    determinism in (seed, index), independence of neighbours and of seeds, inverted-dropout scaling."""
    import numpy as np
    from oracle import dropout as D
    n = 1 << 20
    for p in (0.2, 0.5):
        m = D.keep_mask(99, n, p)
        assert abs(m.mean() - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n)
        assert np.array_equal(m, D.keep_mask(99, n, p)) and np.array_equal(m[:1000], D.keep_mask(99, 1000, p))
        assert abs(np.corrcoef(m[:-1], m[1:])[0, 1]) < 5e-3 and abs(np.corrcoef(m, D.keep_mask(100, n, p))[0, 1]) < 5e-3
    assert D.keep_mask(5, 4096, 0.0).all()
    x = np.full(4096, 2.0, dtype=np.float32)
    y = D.dropout(x, 0.2, 7)
    assert set(np.unique(y)) == {0.0, 2.5}
    assert D.uniform_u32(7, 4).tolist() == [1674306020, 72105175, 3868737664, 2503666544]   # splitmix64 in C (gcc)


def test_extract_melspec_task_list_and_error_contract(tmp_path, monkeypatch, capsys):
    """process/extract_melspec.py keeps the reference script's contract: tasks for videos of >= 15 fps only, `_left` /
    `_right` crops share the source wav, existing outputs are skipped (1), failures are reported and counted (-1)."""
    from m3t_b200.process import extract_melspec as E
    monkeypatch.chdir(tmp_path)
    os.makedirs("splits")
    with open("splits/frames_fps.csv", "w") as f:
        f.write("a,100,30.0\nb_left,50,25.0\nc,80,12.0\n")
    tasks = E.build_tasks("wavs", "out")
    assert tasks == [(30.0, os.path.join("wavs", "a.wav"), os.path.join("out", "a.npy")),
                     (25.0, os.path.join("wavs", "b_left.wav"), os.path.join("out", "b_left.npy"))]
    os.makedirs("out")
    open(os.path.join("out", "a.npy"), "w").close()
    assert E.extract_melspec(tasks[0]) == 1
    assert E.extract_melspec(tasks[1]) == -1                       # wavs/b.wav does not exist
    assert "wavs/b.wav" in capsys.readouterr().out.replace(os.sep, "/")
    assert E.main(["wavs", "out"]) == 0
    assert "progress: 2/2" in capsys.readouterr().out


@pytest.mark.skipif(not os.path.isdir("/root/reference/process"), reason="reference tree not present")
def test_checkpoint_surgery_scripts_meet_the_key_contract(tmp_path):
    """The reference's two checkpoint-surgery scripts, run unmodified on state_dicts of THIS build's modules, produce
    files this build's next-stage models load (they pin the state_dict key contract):
    export_pretrained_ckpts.py  VGG-M `visual.v2p.*` (VoxCeleb pre-training, fc backend) -> `visual.shared.*` +
                                `visual.{v,a}_private.*` of the split model;
    merge_av_checkpoints.py     audio stream + visual stream -> the `--fusion_checkpoint` of the audio-visual model."""
    import subprocess
    import sys
    from m3t_b200.models.model import AffWild2VA
    from m3t_b200.models.vggm import VA_3DVGGM
    ref = "/root/reference/process"

    def hp(**kw):
        d = dict(backbone="v2p_split", backend="gru", modality="visual", fusion_type="attention", window=8,
                 loss="ccc_mtl", loss_lambda=0.5, num_hidden=512, split_layer=3, num_fc_layers=2, learning_rate=5e-5,
                 optimizer="adam")
        d.update(kw)
        return argparse.Namespace(**d)

    torch.manual_seed(0)
    pre = VA_3DVGGM(frameLen=8, backend="fc", nClasses=1000)
    torch.save({"state_dict": {"visual." + k: v for k, v in pre.state_dict().items()}}, tmp_path / "vox.ckpt")
    r = subprocess.run([sys.executable, os.path.join(ref, "export_pretrained_ckpts.py"), "vox.ckpt"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    video_sd = torch.load(tmp_path / "video_checkpoint.pt")["state_dict"]
    visual = AffWild2VA(hp())
    own = visual.state_dict()
    assert set(video_sd) <= set(own) and all(video_sd[k].shape == own[k].shape for k in video_sd)
    res = visual.load_state_dict(video_sd, strict=False)
    assert not res.unexpected_keys
    assert all(k.split(".")[1] in ("gru_v", "gru_a") for k in res.missing_keys), res.missing_keys   # only the heads
    assert torch.equal(visual.visual.shared[0].weight, pre.v2p[0].weight)
    assert torch.equal(visual.visual.a_private[0].weight, pre.v2p[12].weight)

    audio = AffWild2VA(hp(modality="audio"))
    torch.save({"state_dict": audio.state_dict()}, tmp_path / "audio.ckpt")
    torch.save({"state_dict": visual.state_dict()}, tmp_path / "video.ckpt")
    r = subprocess.run([sys.executable, os.path.join(ref, "merge_av_checkpoints.py"), "audio.ckpt", "video.ckpt"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    fused = torch.load(tmp_path / "fused_av.pt")["state_dict"]
    av = AffWild2VA(hp(modality="audiovisual"))
    own = av.state_dict()
    assert set(fused) <= set(own) and all(fused[k].shape == own[k].shape for k in fused)
    res = av.load_state_dict(fused, strict=False)          # train.py:26-28
    assert not res.unexpected_keys
    assert {k.split(".")[0] for k in res.missing_keys} == {"proj_v", "att_fuse", "fusion"}
    assert torch.equal(av.audio.gru.weight_hh_l0, audio.audio.gru.weight_hh_l0)


def test_gru_cluster_index_maps():
    """Host model of gru_fwd_cluster_kernel's index arithmetic (csrc/gru.cu): the DSMEM pull writes every element of
    the 16 x H hidden-state tile exactly once from the right (peer, row, column); the 256 threads own every (row, unit)
    of the CTA's 16 x 64 slice exactly once; the ldmatrix row addresses of a warp cover its 8 units of each gate; the
    shared-memory budget fits for every hidden size."""
    import numpy as np
    for H in (128, 256, 512):
        ncta, ldk = H // 64, H + 8
        smem = (3 * 64 + 16) * ldk * 2 + 2 * 16 * 64 * 2
        assert smem + 1024 <= 227 * 1024, (H, smem)
        # pull: vector v -> (peer, row, c8); destination element offset in the h tile
        hits = np.zeros((16, H), dtype=int)
        src = {}
        nvec = 16 * ncta * 8
        for tid in range(256):
            for v0 in range(tid, nvec, 4 * 256):
                for u in range(4):
                    v = v0 + u * 256
                    if v < nvec:
                        peer, rem = v >> 7, v & 127
                        row, c8 = rem >> 3, rem & 7
                        hits[row, peer * 64 + c8 * 8: peer * 64 + c8 * 8 + 8] += 1
                        src[(row, peer * 64 + c8 * 8)] = (peer, row * 64 + c8 * 8)      # (rank, element in its slice)
        assert (hits == 1).all()
        assert all(col // 64 == peer and off == row * 64 + col % 64 for (row, col), (peer, off) in src.items())
    owned = np.zeros((16, 64), dtype=int)
    for tid in range(256):
        warp, lane = tid >> 5, tid & 31
        j0 = 8 * warp + (lane & 3) * 2
        for rs in range(2):
            b = (lane >> 2) + rs * 8
            owned[b, j0:j0 + 2] += 1
    assert (owned == 1).all()
    for warp in range(8):
        rows_rz = {(lane >> 4) * 64 + 8 * warp + (lane & 7) for lane in range(32)}
        rows_n = {2 * 64 + 8 * warp + (lane & 7) for lane in range(32)}
        assert rows_rz == {g * 64 + 8 * warp + i for g in (0, 1) for i in range(8)}
        assert rows_n == {128 + 8 * warp + i for i in range(8)}
        # accumulator columns of m16n8: thread (lane) holds units (lane & 3) * 2 + {0, 1} of the warp's 8
        assert {8 * warp + (lane & 3) * 2 + q for lane in range(32) for q in range(2)} == set(range(8 * warp, 8 * warp + 8))


def test_gru_bwd_cluster_index_maps():
    """Host model of gru_bwd_cluster_kernel's index arithmetic (csrc/gru.cu): the 256 threads own every (row, unit) of
    the CTA's 16 x 32 slice once; the W_hh^T slice load fills every (k, gate, unit) of the [H][96] tile once from the
    right source element; the warps' n-tiles cover all H columns of the partial once; the reduce-scatter reads every
    (rank, row, own column) once, low ranks in the even lane and high ranks in the odd lane of a pair, and each thread
    ends with the two units it owns; the shared-memory budget fits for every cluster size."""
    import numpy as np
    for CS in (4, 8, 16):
        H, K3, bld = CS * 32, CS * 96, 3 * 32 + 8
        smem = (H + 2 * 16) * bld * 2 + 3 * 16 * (H + 8) * 4 + 6 * 256 * 8     # W, 2 A tiles, 3 partials, slots
        assert smem + 1024 <= 227 * 1024, (CS, smem)
        own = np.zeros((16, 32), dtype=int)
        for tid in range(256):
            row, jl = tid >> 4, (tid & 15) * 2
            own[row, jl:jl + 2] += 1
        assert (own == 1).all()
        for js in (0, CS - 1):
            filled = np.zeros((H, 96), dtype=int)
            for i in range(H * 12):
                k, rem = divmod(i, 12)
                g, v = rem >> 2, rem & 3
                src0 = k * K3 + g * H + js * 32 + v * 8          # element of W_hh^T[dir] = [H k][3H (g, j)]
                for e in range(8):
                    col = g * 32 + v * 8 + e
                    filled[k, col] += 1
                    assert (src0 + e) % K3 == g * H + js * 32 + col % 32 and (src0 + e) // K3 == k
            assert (filled == 1).all()
        ntw = CS // 2
        cols = np.zeros(H, dtype=int)
        for warp in range(8):
            for nt in range(ntw):
                for lane in range(32):
                    c = (warp * ntw + nt) * 8 + (lane & 3) * 2
                    if lane >> 2 == 0:
                        cols[c:c + 2] += 1
        assert (cols == 1).all()
        js = CS - 1
        reads = np.zeros((CS, 16, 32), dtype=int)
        for tid in range(256):
            row, hsel, quad = tid >> 4, tid & 1, (tid & 15) >> 1
            for s2 in range(CS // 2):
                reads[hsel * (CS // 2) + s2, row, quad * 4: quad * 4 + 4] += 1
            mine = {quad * 4 + (2 if hsel else 0), quad * 4 + (3 if hsel else 1)}
            assert mine == {(tid & 15) * 2, (tid & 15) * 2 + 1}
        assert (reads == 1).all()


def test_pack_tiling_host_functions():
    """m3t_pack_entry_tiles (host arithmetic of the one-launch weight re-pack): tiles of 32 output channels x
    min(64, 288 / taps) input channels x all taps cover every element once and fit the kernel's shared-memory tile;
    a plain copy is cut into 2048-value tiles."""
    import ctypes
    from m3t_b200 import lib
    L = lib.load()
    L.m3t_pack_entry_tiles.restype = ctypes.c_longlong
    for Cout, Cin, taps in ((64, 64, 9), (128, 64, 9), (512, 512, 9), (96, 16, 27), (1536, 512, 1), (3072, 200, 1),
                            (2, 1024, 1), (64, 3, 25), (33, 7, 4)):
        tci = max(1, min(64, 288 // taps))
        want = -(-Cout // 32) * -(-Cin // tci)
        assert L.m3t_pack_entry_tiles(Cout, Cin, taps, 0) == want
        assert L.m3t_pack_entry_tiles(Cout, Cin, taps, 1) == want
        rp = tci * taps
        pitch = (rp + 1) & ~1
        pitch += 2 if pitch % 4 == 0 else 0
        assert pitch <= 292 and (pitch // 2) % 2 == 1          # odd number of 4-byte words per row: conflict-free columns
        assert want * 32 * tci >= Cout * Cin
    assert L.m3t_pack_entry_tiles(1536, 1, 1, -1) == 1
    assert L.m3t_pack_entry_tiles(4096, 1, 1, -1) == 2
    assert L.m3t_pack_entry_tiles(0, 1, 1, 0) < 0


def test_engine_arena_admits_late_gradients():
    """ADVICE r1 (engine.py arena membership): a parameter that first receives a gradient after the arena was laid
    out joins it (with the moments of the others preserved) instead of being silently skipped."""
    import torch.nn as nn
    from m3t_b200.engine import TrainEngine

    class Two(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(4, 4), nn.Linear(4, 4)
            self.use_b = False

        def forward(self, batch):
            y = self.a(batch["x"])
            return self.b(y) if self.use_b else y

        def compute_loss(self, y, batch, sync_free=False):
            return (y ** 2).mean(), {}

    torch.manual_seed(0)
    m = Two()
    ref = Two()
    ref.load_state_dict(m.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-4)
    eng = TrainEngine(m, lr=1e-2, weight_decay=1e-4, clip=None)
    batch = {"x": torch.randn(8, 4)}
    for step in range(4):
        m.use_b = ref.use_b = step >= 2
        eng.step(batch)
        opt.zero_grad(set_to_none=True)
        ref.compute_loss(ref(batch), batch)[0].backward()
        opt.step()
        assert len(eng.params) == (4 if step >= 2 else 2)
    for (k, p), q in zip(m.state_dict().items(), ref.state_dict().values()):
        assert torch.allclose(p, q, atol=1e-6), k


def test_torch_library_registration():
    """north star: the kernels are torch custom ops.  Every op of custom_ops.OP_NAMES is registered under torch.ops.m3t
    with a schema and a fake (meta) implementation whose output shapes / dtypes match the C-ABI wrappers'."""
    import m3t_b200.custom_ops as C
    for n in C.OP_NAMES:
        assert hasattr(torch.ops.m3t, n), n
        assert str(getattr(torch.ops.m3t, n).default._schema).startswith("m3t::" + n)
    bf = dict(dtype=torch.bfloat16, device="meta")
    f32 = dict(dtype=torch.float32, device="meta")
    A, B = torch.empty(128, 64, **bf), torch.empty(32, 64, **bf)
    assert torch.ops.m3t.gemm(A, B, False, False, True, None, None, None, False).shape == (128, 32)
    assert torch.ops.m3t.gemm(A, torch.empty(128, 48, **bf), True, True, False, None, None, None, False).shape == (64, 48)
    x = torch.empty(4, 10, 512, **bf)
    y = torch.ops.m3t.linear(x, torch.empty(9, 512, **f32), torch.empty(9, **f32), False, True)
    assert y.shape == (4, 10, 9) and y.dtype == torch.float32
    s = torch.empty(4, 10, 1, **f32)
    assert torch.ops.m3t.att_mix(x, x, s, s).shape == x.shape
    H = 128
    prm = [torch.empty(3 * H, 512, **f32), torch.empty(3 * H, H, **f32), torch.empty(3 * H, **f32),
           torch.empty(3 * H, **f32)] * 2
    assert torch.ops.m3t.gru_layer(x, *prm).shape == (4, 10, 2 * H)
    from m3t_b200 import raw
    geom = raw.conv_geom(2, 5, 1, 28, 28, 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), (0, 1, 1), (1, 1, 1))
    out = torch.ops.m3t.conv_fprop(torch.empty(5, 28, 28, 64, **bf), torch.empty(128, 576, **bf), geom, None, None, None,
                                   False)
    assert out.shape == (5, 1, 14, 14, 128)
    assert torch.ops.m3t.conv_wgrad(torch.empty(5, 28, 28, 64, **bf), out, geom).shape == (128, 576)
    assert torch.ops.m3t.logmel(torch.empty(16000, **f32), 30.0, True, 80.0).shape == (1 + 16000 // 177, 40)
