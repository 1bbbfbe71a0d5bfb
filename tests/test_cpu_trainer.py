"""CPU tests of the callers in front of / behind the hot path: the Aff-Wild2 window dataset against
the reference's own class on a synthetic tree, the Lightning-0.6 surface (`Trainer`, `LightningModule`,
`data_loader`) on a small CPU module, the world-size-2 gloo run, and the script launcher's import redirection."""
import argparse
import glob
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import _refload
from tests import synth_affwild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------- dataset
def _same(a, b, path=""):
    if torch.is_tensor(a) or torch.is_tensor(b):
        assert torch.is_tensor(a) and torch.is_tensor(b), path
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), path
    elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), path
    else:
        assert a == b, (path, a, b)


def _seed(s):
    random.seed(s)
    np.random.seed(s)


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
@pytest.mark.parametrize("modality,input_size,balance,release",
                         [("audiovisual", 128, False, "vipl"), ("visual", 256, False, "vipl"),
                          ("audio", 128, False, "vipl"), ("audiovisual", 128, True, "vipl"),
                          ("audio", 128, True, "vipl"), ("visual", 112, False, "ibug")])
def test_dataset_matches_reference_class(tmp_path, monkeypatch, modality, input_size, balance, release):
    """Every sample of every split equals the reference dataset's bit for bit under the same seeds: window choice,
    crop / mirror / cutout draws, missing-frame rule, edge padding of the last window, feature padding, masks."""
    root = str(tmp_path / "data")
    synth_affwild.build(root, str(tmp_path), input_size=input_size, release=release)
    monkeypatch.chdir(tmp_path)
    ref_ds = _refload.load("dataset").AffWild2SequenceDataset
    from m3t_b200.models.dataset import AffWild2SequenceDataset
    for split, stride in (("train", 1), ("val", 1), ("val", 2), ("test", 2)):
        sets = []
        for cls in (ref_ds, AffWild2SequenceDataset):
            for f in glob.glob(str(tmp_path / "*.pkl")):       # each class must do its own window scan
                os.remove(f)
            _seed(11)
            sets.append(cls(split, root, 8, 3, True, release, input_size, modality, balance, stride))
        r, m = sets
        assert len(r) == len(m) and r.sample_src == m.sample_src
        if split == "train":
            assert {k: list(v) for k, v in r.avail_windows.items()} == {k: list(v) for k, v in m.avail_windows.items()}
        for i in range(len(r)):
            _seed(100 + i)
            a = r[i]
            _seed(100 + i)
            b = m[i]
            assert a.keys() == b.keys()
            for k in a:
                _same(a[k], b[k], "%s[%d].%s" % (split, i, k))


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
def test_dataset_u8_mode_draws_the_same_augmentation(tmp_path, monkeypatch):
    """emit_u8: the uint8 frames + parameter row, pushed through the oracle's clip assembly, give the float clip of
    the reference's load_video (128-pixel tracks: no resize involved)."""
    from oracle.video_input import assemble_clip
    root = str(tmp_path / "data")
    synth_affwild.build(root, str(tmp_path), input_size=128)
    monkeypatch.chdir(tmp_path)
    ref_ds = _refload.load("dataset").AffWild2SequenceDataset
    from m3t_b200.models.dataset import AffWild2SequenceDataset
    _seed(3)
    r = ref_ds("train", root, 8, 2, True, "vipl", 128, "visual")
    _seed(3)
    m = AffWild2SequenceDataset("train", root, 8, 2, True, "vipl", 128, "visual", emit_u8=True)
    for i in range(len(r)):
        _seed(50 + i)
        a = r[i]
        _seed(50 + i)
        b = m[i]
        row = [int(x) for x in b["video_aug"]]
        clip = assemble_clip(b["video_u8"].numpy(), *row[:7])
        assert a["start"] == b["start"] and np.array_equal(a["video"].numpy(), clip), i


def test_one_runs_and_load_audio(tmp_path):
    from m3t_b200.models.dataset import load_audio, one_runs
    assert one_runs(np.array([1, 1, 0, 1, 0, 0, 1, 1, 1])).tolist() == [[0, 2], [3, 4], [6, 9]]
    assert one_runs(np.zeros(4)).tolist() == []
    mel = np.arange(10 * 40, dtype=np.float32).reshape(10, 40)
    np.save(tmp_path / "m.npy", mel)
    out = load_audio(str(tmp_path / "m.npy"), 1, 4)            # frames 1..4 -> mel rows 3..7, 6..10, 9..13, 12..16
    assert out.shape == (4, 200)
    assert np.array_equal(out[0], mel[3:8].reshape(-1))
    assert np.array_equal(out[2, :40], mel[9]) and not out[2, 40:].any() and not out[3].any()


# ---------------------------------------------------------------------------------------------- Lightning surface
def _toy_hparams(**over):
    d = dict(lr=1e-2, optimizer="adam", scheduler="none", n=64, batch_size=8, distributed=False, seed=5)
    d.update(over)
    return argparse.Namespace(**d)


def _toy_module():
    from m3t_b200 import lightning as pl

    class Toy(pl.LightningModule):
        def __init__(self, hparams):
            super().__init__()
            self.hparams = hparams
            torch.manual_seed(hparams.seed)
            self.net = nn.Sequential(nn.Linear(6, 16), nn.Tanh(), nn.Linear(16, 2))
            self.unused = nn.Linear(3, 3)           # never reached by autograd, like resnet.fc on the hot path
            g = torch.Generator().manual_seed(77)
            self.x = torch.randn(hparams.n, 6, generator=g)
            self.t = self.x[:, :2] * 0.5 - self.x[:, 2:4]
            self.ended, self.batch_ends = [], 0

        def forward(self, batch):
            return self.net(batch["x"])

        def compute_loss(self, y, batch, sync_free=False):
            return ((y - batch["t"]) ** 2).mean(), {}

        def training_step(self, batch, batch_idx):
            loss, _ = self.compute_loss(self(batch), batch)
            return {"loss": loss, "progress_bar": {"loss": loss}, "log": {"loss": loss}}

        def on_batch_end(self):
            self.batch_ends += 1

        def validation_step(self, batch, batch_idx):
            return {"se": ((self(batch) - batch["t"]) ** 2).sum(dim=1), "idx": batch["idx"]}

        def validation_end(self, outputs):
            se = torch.cat([o["se"] for o in outputs])
            self.ended.append(sorted(int(i) for o in outputs for i in o["idx"]))
            return {"val_loss": se.mean(), "log": {"val_rmse": se.mean().sqrt()}}

        def test_step(self, batch, batch_idx):
            return self.validation_step(batch, batch_idx)

        def test_end(self, outputs):
            torch.save(torch.cat([o["se"] for o in outputs]), "toy_test.pt")
            return {}

        def configure_optimizers(self):
            hp = self.hparams
            if hp.optimizer == "adam":
                opt = torch.optim.Adam(self.parameters(), lr=hp.lr, weight_decay=1e-4)
            else:
                opt = torch.optim.SGD(self.parameters(), lr=hp.lr, momentum=0.9, weight_decay=5e-4)
            if hp.scheduler == "exp":
                return [opt], [torch.optim.lr_scheduler.ExponentialLR(opt, 0.5)]
            if hp.scheduler == "plateau":
                return [opt], [torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.5, patience=0)]
            return opt

        def _loader(self, shuffle):
            ds = [{"x": self.x[i], "t": self.t[i], "idx": i} for i in range(self.hparams.n)]
            sampler = None
            if self.hparams.distributed:
                sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=False)
            return torch.utils.data.DataLoader(ds, batch_size=self.hparams.batch_size, sampler=sampler, shuffle=False)

        @pl.data_loader
        def train_dataloader(self):
            return self._loader(True)

        @pl.data_loader
        def val_dataloader(self):
            return self._loader(False)

        @pl.data_loader
        def test_dataloader(self):
            return self.val_dataloader()

    return Toy


def _manual_fit(Toy, hp, epochs, clip):
    """The same optimisation written out with stock torch: Adam(L2 1e-4) / SGD, global-norm clip, epoch scheduler."""
    m = Toy(hp)
    conf = m.configure_optimizers()
    opt, scheds = (conf[0][0], conf[1]) if isinstance(conf, list) or isinstance(conf, tuple) else (conf, [])
    for _ in range(epochs):
        for s in range(0, hp.n, hp.batch_size):
            b = {"x": m.x[s:s + hp.batch_size], "t": m.t[s:s + hp.batch_size]}
            opt.zero_grad()
            m.compute_loss(m(b), b)[0].backward()
            torch.nn.utils.clip_grad_norm_([p for p in m.parameters() if p.grad is not None], clip)
            opt.step()
        for sc in scheds:
            sc.step()
    return m


@pytest.mark.parametrize("optimizer,scheduler", [("adam", "none"), ("adam", "exp"), ("sgd", "exp")])
def test_trainer_fit_equals_handwritten_loop(tmp_path, monkeypatch, optimizer, scheduler):
    from m3t_b200.lightning import Trainer
    monkeypatch.chdir(tmp_path)
    Toy = _toy_module()
    hp = _toy_hparams(optimizer=optimizer, scheduler=scheduler)
    m = Toy(hp)
    tr = Trainer(early_stop_callback=None, check_val_every_n_epoch=1, gradient_clip_val=0.05,
                 default_save_path=str(tmp_path), max_epochs=3, gpus=None, nb_gpu_nodes=1, distributed_backend="dp")
    assert tr.fit(m) == 1
    assert (tr.engine is not None) == (optimizer == "adam")
    ref = _manual_fit(Toy, hp, 3, 0.05)
    for (k, a), b in zip(m.state_dict().items(), ref.state_dict().values()):
        assert torch.allclose(a, b, atol=2e-6, rtol=1e-5), (k, (a - b).abs().max())
    assert m.batch_ends == 3 * (hp.n // hp.batch_size) and tr.global_step == m.batch_ends
    # sanity pass (5 batches) + one validation per epoch, each over the whole loader
    assert len(m.ended) == 4 and len(m.ended[0]) == 5 * hp.batch_size and m.ended[1] == list(range(hp.n))
    assert "val_loss" in tr.callback_metrics and "val_rmse" in tr.callback_metrics
    ck = glob.glob(str(tmp_path / "lightning_logs" / "version_0" / "checkpoints" / "_ckpt_epoch_*.ckpt"))
    assert len(ck) == 1                                         # save_top_k = 1: only the best epoch stays
    back = Toy.load_from_checkpoint(ck[0])
    assert vars(back.hparams) == vars(hp)
    saved = torch.load(ck[0], weights_only=False)
    assert saved["epoch"] == int(ck[0].rsplit("_", 1)[1].split(".")[0]) and saved["optimizer_states"]
    tr.test(m)
    assert torch.load(tmp_path / "toy_test.pt").shape == (hp.n,)


def test_trainer_plateau_scheduler_sees_val_loss(tmp_path, monkeypatch):
    """ReduceLROnPlateau is stepped with validation_end's `val_loss`; the fused step picks the new lr up."""
    from m3t_b200.lightning import Trainer
    monkeypatch.chdir(tmp_path)
    Toy = _toy_module()
    m = Toy(_toy_hparams(scheduler="plateau", lr=0.5))         # lr far too high: val_loss stops improving at once
    tr = Trainer(gradient_clip_val=0, default_save_path=str(tmp_path), max_epochs=6, nb_sanity_val_steps=0,
                 show_progress_bar=False, early_stop_callback=False)
    tr.fit(m)
    assert tr.current_epoch == 5
    lr_now = tr.optimizers[0].param_groups[0]["lr"]
    assert lr_now < 0.5 and tr.engine.lr in (lr_now, lr_now * 2)   # engine.lr is refreshed at each step
    assert tr.plateau is not None and tr.lr_schedulers == []


def test_data_loader_decorator_and_gpu_parsing():
    from m3t_b200 import lightning as pl

    class M(pl.LightningModule):
        calls = 0

        @pl.data_loader
        def val_dataloader(self):
            M.calls += 1
            return "loader"

        @pl.data_loader
        def train_dataloader(self):
            return "train"

        @pl.data_loader
        def test_dataloader(self):
            return self.missing_attribute

    m = M()
    assert m.val_dataloader() == ["loader"] and m.val_dataloader() == ["loader"] and M.calls == 1
    assert m.train_dataloader() == "train"
    with pytest.raises(RuntimeError, match="AttributeError"):
        m.test_dataloader()
    assert pl.parse_gpus("2") == [2] and pl.parse_gpus("0, 3") == [0, 3] and pl.parse_gpus(2) == [0, 1]
    assert pl.parse_gpus(None) == [] and pl.parse_gpus([1]) == [1]
    with pytest.raises(NotImplementedError):
        pl.Trainer(nb_gpu_nodes=2)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        if torch.cuda.is_available():
            raise RuntimeError("no CUDA device (skipped: a GPU is present)")
        pl.Trainer(gpus="0").fit(M())


def test_trainer_ddp_gloo_world2(tmp_path):
    """Two ranks (gloo): sharded batches + gradient mean == the single-process run on the full batches; validation
    outputs of both ranks reach rank 0's validation_end; one checkpoint, written by rank 0."""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tests", "trainer_gloo_worker.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0 and r.stdout.count("TRAINER_DDP_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


# ---------------------------------------------------------------------------------------------- launcher
def test_run_redirects_the_scripts_imports(tmp_path):
    """`python -m m3t_b200.run script.py`: pytorch_lightning / models / matplotlib resolve to this build."""
    script = tmp_path / "probe_script.py"
    script.write_text(
        "import matplotlib\nmatplotlib.use('Agg')\n"
        "from pytorch_lightning import Trainer\nimport pytorch_lightning as pl\n"
        "from models.model import AffWild2VA\nfrom models.dataset import AffWild2SequenceDataset\n"
        "from models.rnn import GRU\n"
        "import sys\nassert __name__ == '__main__' and sys.argv[1:] == ['--flag', '3'], sys.argv\n"
        "assert issubclass(AffWild2VA, pl.LightningModule) and Trainer.__module__ == 'm3t_b200.lightning'\n"
        "print('REDIRECT_OK', AffWild2VA.__module__)\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "m3t_b200.run", str(script), "--flag", "3"], capture_output=True,
                       text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0 and "REDIRECT_OK m3t_b200.models.model" in r.stdout, r.stdout + r.stderr[-2000:]


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
@pytest.mark.parametrize("script", ["train.py", "eval.py"])
def test_reference_scripts_parse_their_command_line_unchanged(tmp_path, script):
    """The reference's own train.py / eval.py, unmodified, import and build their full argument parser on this
    implementation (everything up to the first kernel launch, which needs the GPU: tests/gpu_cases.py runs a fit)."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "m3t_b200.run", os.path.join(_refload.REFERENCE_ROOT, script), "--help"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    for opt in ("--gpus", "--checkpoint", "--backbone", "--fusion_type", "--windows_per_epoch", "--test_on_val",
                "--max_nb_epochs", "--distributed"):
        assert opt in r.stdout, opt


def test_model_hooks_configure_optimizers_and_args():
    """argparse surface == the reference's defaults (oracle/_refload.hparams lists them), scheduler / freeze_enc
    branches of configure_optimizers."""
    from m3t_b200.models.model import AffWild2VA
    parser = AffWild2VA.add_model_specific_args(argparse.ArgumentParser(add_help=False))
    got = vars(parser.parse_args([]))
    want = vars(_refload.hparams())
    for k in ("gpus", "nodes", "seed", "fusion_checkpoint", "checkpoint"):     # train.py's own arguments
        want.pop(k)
    assert got.pop("device_augment") is False
    want["dataset_path"] = got["dataset_path"]          # _refload.hparams leaves the site-specific default empty
    assert got == want
    if _refload.available():                            # and against the reference's own parser, default by default
        ref_parser = _refload.load("model").AffWild2VA.add_model_specific_args(argparse.ArgumentParser(add_help=False))
        assert vars(ref_parser.parse_args([])) == got
    hp = argparse.Namespace(**dict(got, modality="audiovisual", backbone="resnet", fusion_type="attention",
                                   split_layer=5, window=4, freeze_enc=True, scheduler="plateau"))
    m = AffWild2VA(hp)
    conf = m.configure_optimizers()
    opt, sched = conf[0][0], conf[1][0]
    assert isinstance(sched, torch.optim.lr_scheduler.ReduceLROnPlateau) and type(opt) is torch.optim.Adam
    trainable = {n.split(".")[0] for n, p in m.named_parameters() if p.requires_grad}
    assert trainable == {"fusion", "proj_v", "att_fuse"}
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == \
        sum(p.numel() for n, p in m.named_parameters() if n.split(".")[0] in trainable)
    hp2 = argparse.Namespace(**dict(vars(hp), freeze_enc=False, scheduler="cyclic", optimizer="sgd"))
    m2 = AffWild2VA(hp2)
    assert type(m2.configure_optimizers()) is torch.optim.SGD and hasattr(m2, "cyclic_scheduler")
    m2.on_batch_end()
    hp3 = argparse.Namespace(**dict(vars(hp), freeze_enc=False, test_lr=True))
    m3 = AffWild2VA(hp3)
    assert type(m3.configure_optimizers()) is torch.optim.Adam and m3.lr_test.get_lr()[0] > hp3.learning_rate


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
def test_validation_end_matches_reference_hook(tmp_path, monkeypatch):
    """validation_end on window outputs in the reference's format (stride = window: concatenation branch, host only):
    same val_loss / CCC / MSE and the same predictions_val.pt as the reference's hook run on a stub self."""
    monkeypatch.chdir(tmp_path)
    ref_model = _refload.load("model")
    from m3t_b200.models.model import AffWild2VA
    g = torch.Generator().manual_seed(9)
    outputs, W = [], 8
    for vids in (("a", "b"), ("b", "a")):
        out = {k: [] for k in ("v_gt", "a_gt", "v_pred", "a_pred")}
        out["vid_names"], starts = list(vids), []
        for j, v in enumerate(vids):
            n = W if len(outputs) == 0 else 5
            for k in out:
                if k != "vid_names":
                    out[k].append(torch.rand(n, generator=g) * 2 - 1)
            starts.append(0 if len(outputs) == 0 else W)
        out["v_gt"][0][1] = -5.0                      # an unannotated frame: excluded from the metrics
        out["start_frames"] = torch.tensor(starts)
        outputs.append(out)
    hp = _refload.hparams(test_on_val=False, window=W)
    stub = argparse.Namespace(hparams=hp)
    want = ref_model.AffWild2VA.validation_end(stub, outputs)
    want_file = torch.load("predictions_val.pt")
    os.remove("predictions_val.pt")
    got = AffWild2VA.validation_end(argparse.Namespace(hparams=hp, _per_video=lambda *a, **k:
                                                       AffWild2VA._per_video(stub, *a, **k)), outputs)
    got_file = torch.load("predictions_val.pt")
    assert torch.allclose(got["val_loss"], want["val_loss"], atol=1e-6)
    for k in want["log"]:
        assert torch.allclose(torch.as_tensor(got["log"][k]), torch.as_tensor(want["log"][k]), atol=1e-6), k
    for k in want_file:
        assert want_file[k].keys() == got_file[k].keys()
        for v in want_file[k]:
            assert torch.equal(want_file[k][v], got_file[k][v]), (k, v)


def test_fit_on_synthetic_tree_with_stub_forward(tmp_path, monkeypatch):
    """Everything around the kernels, on the CPU: Aff-Wild2 tree -> dataset -> collate -> training_step (ccc_mtl loss,
    expression accuracy) -> fused-step stand-in -> validation_step / validation_end -> plateau scheduler ->
    checkpoint -> load_from_checkpoint.  Only `forward` is replaced (a Linear on the audio features): the real one
    launches sm_100a kernels (tests/gpu_cases.py::case_trainer_scripts runs it on the GPU)."""
    from m3t_b200.lightning import Trainer
    from m3t_b200.models.model import AffWild2VA
    synth_affwild.build(str(tmp_path / "data"), str(tmp_path), input_size=128)
    monkeypatch.chdir(tmp_path)

    class Stubbed(AffWild2VA):
        def __init__(self, hparams):
            super().__init__(hparams)
            self.stub = nn.Linear(200, 9)

        def forward(self, batch):
            return self.stub(batch["audio"])

    parser = AffWild2VA.add_model_specific_args(argparse.ArgumentParser(add_help=False))
    hp = parser.parse_args(["--modality", "audio", "--window", "8", "--windows_per_epoch", "4", "--batch_size", "4",
                            "--dataset_path", str(tmp_path / "data"), "--release", "vipl", "--input_size", "128",
                            "--workers", "0", "--checkpoint_path", str(tmp_path), "--max_nb_epochs", "2",
                            "--learning_rate", "1e-2"])
    _seed(1)
    torch.manual_seed(1)
    m = Stubbed(hp)
    w0 = m.stub.weight.detach().clone()
    tr = Trainer(early_stop_callback=None, check_val_every_n_epoch=1, gradient_clip_val=1.0,
                 default_save_path=hp.checkpoint_path, max_epochs=hp.max_nb_epochs, gpus=None, nb_gpu_nodes=1,
                 distributed_backend="dp")
    tr.fit(m)
    assert tr.engine is not None and tr.plateau is not None
    assert tr.global_step == 2 * 2                   # 2 train videos x 4 windows / batch 4, 2 epochs
    assert not torch.equal(w0, m.stub.weight) and torch.isfinite(m.stub.weight).all()
    assert all(p.grad is None for p in m.audio.parameters())       # the unused stream stays outside the arena
    for k in ("val_loss", "val_ccc_v", "val_ccc_a", "val_mse_v", "val_mse_a"):
        assert np.isfinite(tr.callback_metrics[k]), k
    pred = torch.load("predictions_val.pt")
    assert sorted(pred["valence_pred"]) == ["vidC"]                # vidD is < 15 fps: dropped for audio-only
    assert len(pred["valence_pred"]["vidC"]) == 27 == len(pred["arousal_gt"]["vidC"])
    ck = glob.glob(str(tmp_path / "lightning_logs" / "version_0" / "checkpoints" / "*.ckpt"))
    assert len(ck) == 1
    back = Stubbed.load_from_checkpoint(ck[0])
    assert back.hparams.window == 8 and back.state_dict().keys() == m.state_dict().keys()
    # half-stride evaluation needs the device kernel: refused on the CPU, not silently emulated
    hp.test_on_val = False
    with pytest.raises(RuntimeError, match="m3t_overlap_add_f32"):
        tr.test(m)


@pytest.mark.skipif(not _refload.available() or torch.cuda.is_available(), reason="reference tree + no GPU")
def test_reference_smoothing_script_resolves_its_imports(tmp_path):
    """get_smoothed_ccc.py, unmodified, through the launcher: `models.utils.smooth_predictions / concordance_cc2_np`
    resolve to this build and the script reaches the device filter, which refuses to run without a GPU (its numerics
    are the `smooth_predictions_api` / `postproc_eval` GPU cases)."""
    track = torch.linspace(-1, 1, 50)
    torch.save({k: {"vid": track.clone()} for k in ("valence_gt", "arousal_gt", "valence_pred", "arousal_pred")},
               tmp_path / "predictions_val.pt")
    r = subprocess.run([sys.executable, "-m", "m3t_b200.run", os.path.join(_refload.REFERENCE_ROOT,
                                                                           "get_smoothed_ccc.py")],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, PYTHONPATH=ROOT),
                       cwd=str(tmp_path))
    assert r.returncode != 0 and "m3t_wiener1d_f64" in r.stderr, r.stderr[-1500:]


def test_plain_python_start_relaunches_one_process_per_rank(tmp_path):
    """`python script.py` with distributed_backend='ddp' and 2 ranks (the reference's README start, without torchrun):
    rank 0 re-launches its own command line as rank 1; two processes in total, identical parameters, fit followed by
    test in the same script does not launch again."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "scripts", "relaunch_script.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    started = sorted(open(f).read() for f in glob.glob(str(tmp_path / "started_*")))
    assert started == ["1", "none"], started          # the launched copy saw RANK=1; the original had no RANK yet
    ranks = [open(tmp_path / ("rank_%d" % i)).read().split() for i in (0, 1)]
    assert ranks[0][0] == ranks[1][0] == "2" and ranks[0][1] == ranks[1][1], ranks
    assert (tmp_path / "toy_test.pt").exists()


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
@pytest.mark.parametrize("loss,expr", [("ccc_mtl", "some"), ("ccc_mtl", "none"), ("mse_mtl", "some"), ("ccc", "some")])
def test_step_hooks_match_reference_methods(loss, expr):
    """training_step / validation_step / test_step of the task module against the reference's own methods, both fed the
    same network output (`forward` replaced by a constant): loss, every logged term, expression accuracy, and the
    per-clip lists handed to validation_end / test_end."""
    ref_cls = _refload.load("model").AffWild2VA
    from m3t_b200.models.model import AffWild2VA
    hp = _refload.hparams(modality="audio", loss=loss, window=6)
    g = torch.Generator().manual_seed(4)
    B, T, C = 3, 6, 9 if "mtl" in loss else 2
    y_hat = torch.randn(B, T, C, generator=g)
    batch = {"audio": torch.zeros(B, T, 200),
             "label_valence": torch.rand(B, T, generator=g) * 2 - 1, "label_arousal": torch.rand(B, T, generator=g) * 2 - 1,
             "class_expr": torch.randint(0, 7, (B, T), generator=g),
             "expr_valid": (torch.rand(B, T, generator=g) > 0.4) if expr == "some" else torch.zeros(B, T, dtype=torch.bool),
             "length": torch.tensor([6, 4, 1]), "vid_name": ["a", "a", "b"], "start": torch.tensor([0, 6, 0])}
    torch.manual_seed(0)
    ref, mine = ref_cls(hp), AffWild2VA(hp)
    for m in (ref, mine):
        m.forward = lambda b: y_hat.clone().requires_grad_(True)
    want, got = ref.training_step(batch, 0), mine.training_step(batch, 0)
    assert want.keys() == got.keys()
    assert torch.allclose(want["loss"], got["loss"], atol=1e-6)
    for part in ("progress_bar", "log"):
        assert want[part].keys() == got[part].keys(), part
        for k in want[part]:
            assert abs(float(want[part][k]) - float(got[part][k])) < 1e-6, (part, k)
    got["loss"].backward()                                   # the returned loss carries the graph
    for name in ("validation_step", "test_step"):
        for flag in (False, True):
            hp.test_on_val = flag
            w, o = getattr(ref, name)(batch, 0), getattr(mine, name)(batch, 0)
            assert w.keys() == o.keys(), (name, flag)
            for k in w:
                if k == "vid_names":
                    assert w[k] == o[k]
                elif k == "start_frames":
                    assert torch.equal(w[k], o[k])
                else:
                    assert len(w[k]) == len(o[k]) and all(torch.equal(a, b) for a, b in zip(w[k], o[k])), (name, k)


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
@pytest.mark.parametrize("optimizer,scheduler,freeze", [("adam", "plateau", False), ("adam", "exp", False),
                                                        ("sgd", "cyclic", False), ("adam", "plateau", True)])
def test_configure_optimizers_matches_reference(optimizer, scheduler, freeze):
    """Same optimiser class and hyper-parameters, same scheduler class and settings, same set of trainable parameters
    as the reference's configure_optimizers (models/model.py:375-407)."""
    ref_cls = _refload.load("model").AffWild2VA
    from m3t_b200.models.model import AffWild2VA
    hp = _refload.hparams(modality="audiovisual", backbone="resnet", fusion_type="attention", split_layer=5, window=4,
                          optimizer=optimizer, scheduler=scheduler, freeze_enc=freeze, learning_rate=3e-4)

    def unpack(conf):
        return (conf[0][0], conf[1][0]) if isinstance(conf, (list, tuple)) else (conf, None)

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ref = ref_cls(hp)
            r_opt, r_sch = unpack(ref.configure_optimizers())
        except TypeError:            # ReduceLROnPlateau(verbose=...) no longer exists in this torch: compare what can be
            ref, r_opt, r_sch = None, None, None
        mine = AffWild2VA(hp)
        m_opt, m_sch = unpack(mine.configure_optimizers())
    if scheduler == "cyclic":
        r_sch, m_sch = (getattr(ref, "cyclic_scheduler", None) if ref is not None else None), mine.cyclic_scheduler
    if ref is None:
        assert scheduler == "plateau" and isinstance(m_sch, torch.optim.lr_scheduler.ReduceLROnPlateau)
        assert (m_sch.factor, m_sch.patience, m_sch.min_lrs) == (hp.decay_factor, 3, [1e-6])
        ref = ref_cls(hp)
        if freeze:                   # the freezing happens before the optimiser is built: still comparable
            try:
                ref.configure_optimizers()
            except TypeError:
                pass
    else:
        assert type(r_opt) is type(m_opt) and type(r_sch) is type(m_sch)
        keys = ("lr", "weight_decay", "betas", "eps", "momentum", "nesterov", "dampening")
        a, b = r_opt.param_groups[0], m_opt.param_groups[0]
        assert {k: a[k] for k in keys if k in a} == {k: b[k] for k in keys if k in b}
        for k in ("gamma", "factor", "patience", "min_lrs", "base_lrs", "max_lrs", "total_size", "step_ratio"):
            if hasattr(r_sch, k):
                assert getattr(r_sch, k) == getattr(m_sch, k), k
    assert [n for n, p in ref.named_parameters() if p.requires_grad] == \
        [n for n, p in mine.named_parameters() if p.requires_grad]
    assert sum(p.numel() for p in m_opt.param_groups[0]["params"]) == \
        sum(p.numel() for p in mine.parameters() if p.requires_grad)


def test_early_stopping_as_lightning_06(tmp_path, monkeypatch):
    """`early_stop_callback=None` (what train.py passes) is Lightning 0.6's default callback: val_loss, patience 3,
    tolerant of a missing metric; False disables it; True insists on the metric."""
    from m3t_b200 import lightning as pl
    monkeypatch.chdir(tmp_path)
    Toy = _toy_module()

    class Stuck(Toy):
        def validation_end(self, outputs):
            self.ended.append(None)
            return {"val_loss": torch.tensor(1.0 + 0.01 * len(self.ended))}       # never improves after the first

    m = Stuck(_toy_hparams())
    tr = pl.Trainer(early_stop_callback=None, max_epochs=20, nb_sanity_val_steps=0, default_save_path=str(tmp_path),
                    show_progress_bar=False)
    tr.fit(m)
    assert tr.current_epoch == 3 and tr.early_stop_callback.stopped_epoch == 3      # best at 0, then 3 bad checks
    assert len(glob.glob(str(tmp_path / "lightning_logs" / "version_0" / "checkpoints" / "*.ckpt"))) == 1

    class Silent(Toy):
        def validation_end(self, outputs):
            return {}

    tr = pl.Trainer(early_stop_callback=None, max_epochs=5, nb_sanity_val_steps=0, default_save_path=str(tmp_path),
                    show_progress_bar=False, checkpoint_callback=False)
    tr.fit(Silent(_toy_hparams()))
    assert tr.current_epoch == 4                                                     # non-strict: keeps going
    with pytest.raises(RuntimeError, match="val_loss"):
        pl.Trainer(early_stop_callback=True, max_epochs=5, nb_sanity_val_steps=0, default_save_path=str(tmp_path),
                   show_progress_bar=False, checkpoint_callback=False).fit(Silent(_toy_hparams()))
    es = pl.EarlyStopping("acc", min_delta=0.1, patience=2, mode="max")
    assert [es.on_epoch_end(i, {"acc": a}) for i, a in enumerate((0.5, 0.55, 0.7, 0.75, 0.6))] == \
        [False, False, False, False, True]


def test_trainer_loop_controls(tmp_path, monkeypatch):
    """check_val_every_n_epoch, fast_dev_run, train_percent_check and max_steps bound the loops as in Lightning 0.6."""
    from m3t_b200.lightning import Trainer
    monkeypatch.chdir(tmp_path)
    Toy = _toy_module()
    common = dict(default_save_path=str(tmp_path), show_progress_bar=False, nb_sanity_val_steps=0,
                  early_stop_callback=False, checkpoint_callback=False)
    m = Toy(_toy_hparams())
    Trainer(max_epochs=4, check_val_every_n_epoch=2, **common).fit(m)
    assert len(m.ended) == 2 and m.batch_ends == 4 * 8
    m = Toy(_toy_hparams())
    tr = Trainer(max_epochs=4, fast_dev_run=True, **common)
    tr.fit(m)
    assert m.batch_ends == 1 and len(m.ended) == 1 and len(m.ended[0]) == 8 and tr.current_epoch == 0
    m = Toy(_toy_hparams())
    Trainer(max_epochs=2, train_percent_check=0.5, val_percent_check=0.25, **common).fit(m)
    assert m.batch_ends == 2 * 4 and all(len(e) == 2 * 8 for e in m.ended)
    m = Toy(_toy_hparams())
    tr = Trainer(max_epochs=10, max_steps=11, **common)
    tr.fit(m)
    assert tr.global_step == 11 and tr.current_epoch == 1
    with pytest.raises(NotImplementedError):
        Trainer(resume_from_checkpoint="x.ckpt")
    with pytest.raises(NotImplementedError):
        Trainer(use_amp=True)


def test_prefetcher_iteration_and_opt_in(tmp_path, monkeypatch):
    """Prefetcher yields the loader's batches in order, one ahead, and is empty for an empty loader; with
    M3T_TRAINER_PREFETCH=1 a fit gives the same parameters as the plain feed."""
    from m3t_b200 import lightning as pl
    pulled = []

    class Loader:
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            for i in range(self.n):
                pulled.append(i)
                yield {"x": torch.full((2,), float(i)), "name": "b%d" % i}

    seen = []
    for b in pl.Prefetcher(Loader(4), torch.device("cpu")):
        seen.append((int(b["x"][0]), b["name"], len(pulled)))
    assert seen == [(0, "b0", 2), (1, "b1", 3), (2, "b2", 4), (3, "b3", 4)]       # always one batch ahead
    assert list(pl.Prefetcher(Loader(0), torch.device("cpu"))) == []
    monkeypatch.chdir(tmp_path)
    Toy = _toy_module()
    kw = dict(default_save_path=str(tmp_path), show_progress_bar=False, nb_sanity_val_steps=0, max_epochs=2,
              early_stop_callback=False, checkpoint_callback=False, gradient_clip_val=0.05)
    a = Toy(_toy_hparams())
    pl.Trainer(**kw).fit(a)
    monkeypatch.setenv("M3T_TRAINER_PREFETCH", "1")
    b = Toy(_toy_hparams())
    pl.Trainer(**kw).fit(b)
    assert all(torch.equal(p, q) for p, q in zip(a.state_dict().values(), b.state_dict().values()))


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
@pytest.mark.parametrize("trial", [0, 1, 2, 3])
def test_dataset_fuzz_vs_reference(tmp_path, monkeypatch, trial):
    """Random trees (video lengths, missing frames, unannotated labels, frame rates, 128 / 256-pixel tracks, window
    4 / 8 / 16, plain / noisy-balanced windows): every sample bit-identical to the reference class, and where the
    reference raises (e.g. a one-frame last window past the end of a short feature file) this one raises the same
    exception type."""
    ref_ds = _refload.load("dataset").AffWild2SequenceDataset
    from m3t_b200.models.dataset import AffWild2SequenceDataset
    rng = np.random.default_rng(trial)
    monkeypatch.chdir(tmp_path)
    vids = {}
    for i, split in enumerate(["train", "train", "val", "val", "test"]):
        n = int(rng.integers(20, 70))
        vids["v%d" % i] = dict(split=split, frames=n, fps=float(rng.choice([12.0, 25.0, 30.0])),
                               expr=bool(rng.integers(0, 2)), missing=[int(x) for x in rng.integers(0, n, 3)],
                               bad=[int(x) for x in rng.integers(0, n, 4)])
    size = int(rng.choice([128, 256]))
    root = str(tmp_path / "data")
    synth_affwild.build(root, str(tmp_path), videos=vids, input_size=size, seed=trial)
    W = int(rng.choice([4, 8, 16]))
    compared = 0
    for modality in ("audiovisual", "audio"):
        for split, stride in (("train", 1), ("val", 2), ("test", 2)):
            sets, errs = [], []
            for cls in (ref_ds, AffWild2SequenceDataset):
                for f in glob.glob(str(tmp_path / "*.pkl")):
                    os.remove(f)
                _seed(trial)
                try:
                    sets.append(cls(split, root, W, 2, True, "vipl", size, modality, bool(trial % 2), stride))
                    errs.append(None)
                except Exception as e:  # noqa: BLE001
                    sets.append(None)
                    errs.append(type(e).__name__)
            assert errs[0] == errs[1], (modality, split, errs)
            if errs[0]:
                continue
            r, m = sets
            assert r.sample_src == m.sample_src
            for i in range(len(r)):
                got = []
                for ds in (r, m):
                    _seed(7 * trial + i)
                    try:
                        got.append(ds[i])
                    except Exception as e:  # noqa: BLE001
                        got.append(type(e).__name__)
                a, b = got
                if isinstance(a, str) or isinstance(b, str):
                    assert a == b, (modality, split, i, a, b)
                    continue
                assert a.keys() == b.keys()
                for k in a:
                    _same(a[k], b[k], "%s %s[%d].%s" % (modality, split, i, k))
                compared += 1
    assert compared > 20


def test_eval_shards_cover_every_window_once():
    """ADVICE r1 (lightning.py distributed evaluation): evaluation shards must not repeat samples when
    len(dataset) % world != 0 (DistributedSampler pads), and gathered outputs are de-duplicated by
    (vid_name, start_frame) if a padding sampler is used anyway."""
    from m3t_b200.lightning import SequentialShardSampler, _dedup_eval_outputs
    for n in (1, 7, 8, 13):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s = SequentialShardSampler(list(range(n)), num_replicas=world, rank=r)
                idx = list(s)
                assert len(idx) == len(s)
                seen += idx
            assert sorted(seen) == list(range(n)), (n, world)
    # a padded pair of shards: 5 windows over 2 ranks -> rank 1 repeats window 0
    def out(ids):
        return {"vid_names": ["v%d" % (i // 3) for i in ids], "start_frames": torch.tensor([(i % 3) * 8 for i in ids]),
                "v_pred": [torch.full((4,), float(i)) for i in ids], "tag": "x"}
    gathered = [out([0, 2]), out([4]), out([1, 3]), out([0])]
    kept = _dedup_eval_outputs(gathered)
    ids = [int(t[0]) for o in kept for t in o["v_pred"]]
    assert ids == [0, 2, 4, 1, 3]
    mixed = _dedup_eval_outputs([out([0, 1]), out([1, 2])])
    assert [int(t[0]) for o in mixed for t in o["v_pred"]] == [0, 1, 2]
    assert mixed[1]["start_frames"].tolist() == [16] and mixed[1]["vid_names"] == ["v0"] and mixed[1]["tag"] == "x"
    assert _dedup_eval_outputs([{"loss": 1.0}]) == [{"loss": 1.0}]
