"""Task module (reference: models/model.py `AffWild2VA`): stream construction, forward, losses, the Lightning-0.6
hooks (training / validation / test steps and their `*_end` aggregation, optimiser + scheduler configuration,
dataloaders, argparse surface), so the reference's train.py / eval.py drive it unchanged through `m3t_b200.run`
(SURVEY 8(f) N1).  The base class is `m3t_b200.lightning.LightningModule` (pytorch_lightning 0.6 is not installable).

Every configuration of the reference's constructor is built: modality in {audio, visual, audiovisual}, backbone in
{resnet, v2p, v2p_split, densenet, vggface}, fusion_type in {concat, attention, att_dec}; the hot path (BASELINE
configs) is resnet / v2p_split + attention.
"""
import os
import sys
from argparse import ArgumentParser

import torch
import torch.nn as nn
from torch.nn import functional as F
from torch.utils.data import DataLoader

from .. import fp32, ops, streams
from .. import lightning as pl
from .att_fusion import AttFusion
from .backbone import VA_3DResNet
from .rnn import GRU, AttEncDec
from .utils import concordance_cc2, mse

LR_TEST_MAX_LR = 0.01
LR_TEST_STEPS = 1096 * 3

_Base = pl.LightningModule


def _hp(hparams, name, default):
    """Callers outside train.py (tests, bench) pass a partial Namespace; absent switches read as the argparse default."""
    return getattr(hparams, name, default)


class AffWild2VA(_Base):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        hp = hparams
        use_mtl = 'mtl' in hp.loss
        fc_outputs = 7 + 2 if use_mtl else 2
        rnn_fc_classes = -1 if hp.modality == 'audiovisual' else fc_outputs
        if 'visual' in hp.modality:
            if hp.backbone == 'resnet':
                self.visual = VA_3DResNet(hiddenDim=hp.num_hidden, frameLen=hp.window, backend=hp.backend,
                                          resnet_ver='v1', nClasses=rnn_fc_classes, nFCs=hp.num_fc_layers)
            elif hp.backbone in ('v2p', 'v2p_split'):
                from .vggm import VA_3DVGGM, VA_3DVGGM_Split
                if hp.backbone == 'v2p':
                    self.visual = VA_3DVGGM(hiddenDim=hp.num_hidden, frameLen=hp.window, backend=hp.backend,
                                            nClasses=rnn_fc_classes, nFCs=hp.num_fc_layers)
                else:
                    self.visual = VA_3DVGGM_Split(hiddenDim=hp.num_hidden, frameLen=hp.window, backend=hp.backend,
                                                  split_layer=hp.split_layer, nClasses=rnn_fc_classes,
                                                  nFCs=hp.num_fc_layers, use_mtl=use_mtl)
            elif hp.backbone == 'vggface':
                from .backbone import VA_VGGFace
                self.visual = VA_VGGFace(hiddenDim=hp.num_hidden, frameLen=hp.window, backend=hp.backend,
                                         nClasses=rnn_fc_classes, nFCs=hp.num_fc_layers)
            elif hp.backbone == 'densenet':
                from .backbone import VA_3DDenseNet
                self.visual = VA_3DDenseNet(hiddenDim=hp.num_hidden, frameLen=hp.window, backend=hp.backend,
                                            nClasses=rnn_fc_classes, nFCs=hp.num_fc_layers)
            else:
                raise ValueError("unknown backbone %r" % hp.backbone)
        if 'audio' in hp.modality:
            self.audio = GRU(200, 256, 2, rnn_fc_classes, hp.num_fc_layers)
        if hp.modality == 'audiovisual':
            self.proj_v = nn.Linear(hp.num_hidden * (2 if hp.split_layer == 5 else 4), 512)
            if hp.fusion_type == 'attention':
                self.att_fuse = AttFusion([512, 512], 128)
                self.fusion = GRU(512, hp.num_hidden, 2, fc_outputs, hp.num_fc_layers)
            elif hp.fusion_type == 'concat':
                self.fusion = GRU(512 * 2, hp.num_hidden, 2, fc_outputs, hp.num_fc_layers)
            elif hp.fusion_type == 'att_dec':
                self.fusion = AttEncDec()
            else:
                raise ValueError("unknown fusion_type %r" % hp.fusion_type)
        self.history = {'lr': [], 'loss': []}

    # ------------------------------------------------------------------ forward
    def _visual(self, batch, after_features=None):
        hp = self.hparams
        if hp.backbone in ('vggface', 'densenet'):
            return self.visual.forward_bf16(batch['video'], normalise=True, after_features=after_features)
        if hp.backbone == 'resnet':
            # normalisation (video - 127.5) / 127.5 (reference :106) is folded into the stem's input pass
            if 'video_u8' in batch:
                # decoded uint8 HWC frames + augmentation parameters instead of the float32 clip tensor: crop, mirror
                # and cutout (models/dataset.py:46-80) happen inside the same pass (process/video_input.py)
                video = ops.RawClips(batch['video_u8'], batch['video_aug'])
            else:
                video = batch['video']
            return self.visual.forward_bf16(video, normalise=True, after_features=after_features)
        return self.visual.forward_bf16(batch['video'], batch['se_features'], batch['se_features'], normalise=True,
                                        after_features=after_features)

    def forward(self, batch):
        hp = self.hparams
        if hp.modality == 'audio':
            return ops.as_f32(self.audio.forward_bf16(batch['audio']))
        if hp.modality == 'visual':
            return ops.as_f32(self._visual(batch))
        audio = batch['audio']
        if streams.overlap_ok(audio):
            # inference: the audio BiGRU runs on a side stream next to the visual stream's recurrent head; it is
            # forked by the backbone right after its convolutions (streams.py)
            box = []
            v = self._visual(batch, lambda: box.append(streams.run_on_side(0, audio.device, self.audio.forward_bf16,
                                                                           audio)))
            a = box[0] if box else self.audio.forward_bf16(audio)
            v = ops.linear(v, self.proj_v.weight, self.proj_v.bias)
            if box:
                streams.join(audio.device, 0)
        else:
            a = self.audio.forward_bf16(audio)
            v = self._visual(batch)
            if fp32.enabled():
                v = fp32.linear(v, self.proj_v.weight, self.proj_v.bias)
            else:
                v = ops.linear(v, self.proj_v.weight, self.proj_v.bias)
        if hp.fusion_type == 'att_dec':
            # reference :119-125: teacher forcing when the batch carries 'valence' / 'arousal' tracks
            f = torch.cat((ops.as_bf16(a), ops.as_bf16(v)), dim=-1)
            if 'arousal' in batch:
                return self.fusion(f, torch.stack((batch['valence'], batch['arousal']), dim=-1))
            return self.fusion(f)
        if hp.fusion_type == 'concat':
            f = torch.cat((a, v), dim=-1)
        else:
            f = self.att_fuse.forward_bf16(a, v)
        return ops.as_f32(self.fusion.forward_bf16(f))

    # ------------------------------------------------------------------ losses (reference :132-144)
    def ccc_loss(self, y_hat, y):
        return 1 - concordance_cc2(y_hat.view(-1), y.view(-1), 'none').squeeze()

    def ce_loss(self, y_hat, y, mask):
        loss = F.cross_entropy(y_hat.view(-1, y_hat.size(-1)), y.view(-1), reduction='none')
        return (loss * mask.view(-1).float()).mean()

    def mse_loss(self, y_hat, y):
        return F.mse_loss(y_hat, y)

    def compute_loss(self, y_hat, batch, sync_free=False):
        """The differentiable part of training_step (reference :146-182).  sync_free=True skips the `.item()` host
        synchronisation of the reference's `valid_expr > 0` test (the masked CE of an all-invalid batch is 0)."""
        hp = self.hparams
        if 'mtl' in hp.loss:
            v_hat, a_hat = y_hat[..., 7], y_hat[..., -1]
        else:
            v_hat, a_hat = y_hat[..., -2], y_hat[..., -1]
        v, a = batch['label_valence'], batch['label_arousal']
        if sync_free and 'ccc' in hp.loss and y_hat.is_cuda and os.environ.get("M3T_FUSED_LOSS", "1") == "1":
            # the engine's step: both CCC terms, the masked cross-entropy and dL/dy_hat in ONE launch (m3t_av_loss)
            mtl = 'mtl' in hp.loss
            out = ops.AVLossFn.apply(y_hat, v, a, batch['class_expr'] if mtl else None,
                                     batch['expr_valid'] if mtl else None, 7 if mtl else -2, -1, 7 if mtl else 0,
                                     hp.loss_lambda, 0.8)
            logs = {'loss_v': out[1], 'loss_a': out[2], 'loss': out[0]}
            if mtl:
                logs['loss_expr'] = out[3]
            return out[0], logs
        if 'mse' in hp.loss:
            loss_v, loss_a = self.mse_loss(v_hat, v), self.mse_loss(a_hat, a)
        else:
            assert 'ccc' in hp.loss, 'invalid loss specification'
            loss_v, loss_a = self.ccc_loss(v_hat, v), self.ccc_loss(a_hat, a)
        loss = hp.loss_lambda * loss_v + (1 - hp.loss_lambda) * loss_a
        logs = {'loss_v': loss_v, 'loss_a': loss_a}
        if 'mtl' in hp.loss:
            mask = batch['expr_valid']
            if sync_free or int(mask.view(-1).long().sum().item()) > 0:
                loss_expr = self.ce_loss(y_hat[..., :7], batch['class_expr'], mask)
                loss = loss + 0.8 * loss_expr
                logs['loss_expr'] = loss_expr
        logs['loss'] = loss
        return loss, logs


    # ------------------------------------------------------------------ Lightning hooks (reference :146-373)
    def training_step(self, batch, batch_idx):
        y_hat = self.forward(batch)
        loss, logs = self.compute_loss(y_hat, batch)
        progress = dict(logs)
        if 'loss_expr' in logs:
            # expression accuracy over the frames that carry a label (reference :178-180)
            keep = batch['expr_valid'].view(-1)
            hit = (y_hat[..., :7].argmax(dim=-1).view(-1)[keep] == batch['class_expr'].view(-1)[keep]).sum().item()
            progress['acc_expr'] = hit / int(keep.long().sum().item())
        if _hp(self.hparams, 'test_lr', False):
            self._lr_range_test_record(loss, batch_idx)
        return {'loss': loss, 'progress_bar': progress, 'log': dict(logs)}

    def _lr_range_test_record(self, loss, batch_idx):
        """Learning-rate range test bookkeeping (reference :198-209): exponentially smoothed loss against the
        BatchExponentialLR schedule; plots and exits after LR_TEST_STEPS batches."""
        from .lr_finder import plot_lr
        hist = self.history
        if len(hist['lr']) == LR_TEST_STEPS:
            plot_lr(hist)
            print('Saved LR-loss plot.')
            sys.exit(0)
        hist['lr'].append(self.lr_test.get_lr()[0])
        value = loss.item()
        hist['loss'].append(value if batch_idx == 0 else 0.05 * value + 0.95 * hist['loss'][-1])

    def on_batch_end(self):
        if _hp(self.hparams, 'test_lr', False):
            self.lr_test.step()
        if _hp(self.hparams, 'scheduler', 'plateau') == 'cyclic':
            self.cyclic_scheduler.step()

    @staticmethod
    def _valid_prefixes(track, lens):
        return [track[i, :int(lens[i])] for i in range(track.size(0))]

    def validation_step(self, batch, batch_idx):
        """Per clip: the first `length` frames of prediction and ground truth, plus the bookkeeping that lets
        validation_end put the windows back on their videos (reference :220-242).  One device->host copy per batch."""
        lens = batch['length'].cpu()
        va_hat = self.forward(batch)[..., -2:].cpu()
        return {'v_gt': self._valid_prefixes(batch['label_valence'].cpu(), lens),
                'a_gt': self._valid_prefixes(batch['label_arousal'].cpu(), lens),
                'v_pred': self._valid_prefixes(va_hat[..., 0], lens),
                'a_pred': self._valid_prefixes(va_hat[..., 1], lens),
                'vid_names': batch['vid_name'], 'start_frames': batch['start'].cpu()}

    def test_step(self, batch, batch_idx):
        if _hp(self.hparams, 'test_on_val', False):
            return self.validation_step(batch, batch_idx)
        lens = batch['length'].cpu()
        va_hat = self.forward(batch)[..., -2:].cpu()
        return {'v_pred': self._valid_prefixes(va_hat[..., 0], lens),
                'a_pred': self._valid_prefixes(va_hat[..., 1], lens),
                'vid_names': batch['vid_name'], 'start_frames': batch['start'].cpu()}

    def _per_video(self, outputs, keys, overlapped):
        """Windows -> per-video tracks.  keys: the per-clip lists to carry (e.g. v_gt, a_gt, v_pred, a_pred).
        overlapped=False: windows tile the video, tracks are their concatenation in start order (reference :274-278);
        overlapped=True: half-stride windows are summed on their frames and every frame from window//2 on is halved
        (reference :279-297, :352-366) — `m3t_overlap_add_f32`, all videos in one launch."""
        names, vid_of, starts, segs = [], {}, [], []
        for out in outputs:
            for j, name in enumerate(out['vid_names']):
                vid_of.setdefault(name, len(vid_of))
                names.append(name)
                starts.append(int(out['start_frames'][j]))
                segs.append(torch.stack([out[k][j].float() for k in keys], dim=-1))      # (len, C)
        order = sorted(range(len(segs)), key=lambda i: (vid_of[names[i]], starts[i]))
        videos = list(vid_of)
        if not overlapped:
            tracks = {k: {} for k in keys}
            for v in videos:
                whole = torch.cat([segs[i] for i in order if names[i] == v])
                for c, k in enumerate(keys):
                    tracks[k][v] = whole[:, c].contiguous()
            return tracks
        from ..process import postproc
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("overlap-add of half-stride windows runs on the device (m3t_overlap_add_f32); "
                               "the module is on %s" % dev)
        lens = [segs[i].size(0) for i in order]
        stacked = torch.nn.utils.rnn.pad_sequence([segs[i] for i in order], batch_first=True)
        if stacked.size(1) < self.hparams.window:
            stacked = F.pad(stacked, (0, 0, 0, self.hparams.window - stacked.size(1)))
        t = postproc.overlap_add(stacked.to(dev).contiguous(), [starts[i] for i in order],
                                 [vid_of[names[i]] for i in order], lens, self.hparams.window, len(videos))
        per_video = [x.cpu() for x in t.split()]
        return {k: {v: per_video[n][:, c].contiguous() for n, v in enumerate(videos)} for c, k in enumerate(keys)}

    def validation_end(self, outputs):
        """Global CCC / MSE over every valid frame, `val_loss` = 1 - mean CCC, and predictions_val.pt with the
        per-video tracks (reference :244-318)."""
        flat = {k: torch.cat([torch.cat(o[k]) for o in outputs]) for k in ('v_gt', 'a_gt', 'v_pred', 'a_pred')}
        ok = (flat['v_gt'].abs() <= 1) & (flat['a_gt'].abs() <= 1)
        ccc_v = concordance_cc2(flat['v_gt'][ok], flat['v_pred'][ok])
        ccc_a = concordance_cc2(flat['a_gt'][ok], flat['a_pred'][ok])
        mse_v = mse(flat['v_pred'][ok], flat['v_gt'][ok])
        mse_a = mse(flat['a_pred'][ok], flat['a_gt'][ok])
        val_loss = 1 - 0.5 * (ccc_v + ccc_a)
        tracks = self._per_video(outputs, ('v_gt', 'a_gt', 'v_pred', 'a_pred'),
                                 overlapped=_hp(self.hparams, 'test_on_val', False))
        torch.save({'valence_gt': tracks['v_gt'], 'arousal_gt': tracks['a_gt'],
                    'valence_pred': tracks['v_pred'], 'arousal_pred': tracks['a_pred']}, 'predictions_val.pt')
        bar = {'val_ccc_v': ccc_v, 'val_ccc_a': ccc_a}
        return {'val_loss': val_loss, 'progress_bar': bar,
                'log': dict(bar, val_mse_v=mse_v, val_mse_a=mse_a, val_loss=val_loss)}

    def test_end(self, outputs):
        """predictions_test.pt: per-video overlap-added valence / arousal tracks (reference :340-373)."""
        if _hp(self.hparams, 'test_on_val', False):
            return self.validation_end(outputs)
        tracks = self._per_video(outputs, ('v_pred', 'a_pred'), overlapped=True)
        torch.save({'valence_pred': tracks['v_pred'], 'arousal_pred': tracks['a_pred']}, 'predictions_test.pt')
        return {}

    # ------------------------------------------------------------------ optimiser (reference :375-407)
    def configure_optimizers(self):
        hp = self.hparams
        if _hp(hp, 'freeze_enc', False):
            # only the fusion stage trains: projection, attention scorers, fusion GRU
            trainable = [self.fusion, self.proj_v] + ([self.att_fuse] if hp.fusion_type == 'attention' else [])
            self.requires_grad_(False)
            for part in trainable:
                part.requires_grad_(True)
        params = [p for p in self.parameters() if p.requires_grad]
        if hp.optimizer == 'adam':
            opt = torch.optim.Adam(params, lr=hp.learning_rate, weight_decay=1e-4)
        elif hp.optimizer == 'sgd':
            opt = torch.optim.SGD(params, lr=hp.learning_rate, momentum=0.9, weight_decay=5e-4)
        else:
            raise ValueError("optimizer must be 'adam' or 'sgd', got %r" % (hp.optimizer,))
        if _hp(hp, 'test_lr', False):
            from .lr_finder import BatchExponentialLR
            self.lr_test = BatchExponentialLR(opt, LR_TEST_MAX_LR, LR_TEST_STEPS)
            return opt
        kind = _hp(hp, 'scheduler', None)
        if kind == 'cyclic':
            self.cyclic_scheduler = torch.optim.lr_scheduler.CyclicLR(
                opt, hp.min_lr, hp.learning_rate, step_size_up=5000, cycle_momentum=hp.optimizer == 'sgd')
            return opt
        if kind == 'exp':
            return [opt], [torch.optim.lr_scheduler.ExponentialLR(opt, hp.decay_factor)]
        if kind == 'plateau':
            return [opt], [torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=hp.decay_factor, patience=3,
                                                                      min_lr=1e-6)]
        return opt

    # ------------------------------------------------------------------ data (reference :409-446)
    def _loader(self, split, inv_test_stride=1):
        from .dataset import AffWild2SequenceDataset
        hp = self.hparams
        if hp.mode != 'video':
            raise NotImplementedError("only mode='video' exists in the reference")
        on_device_aug = _hp(hp, 'device_augment', False) and hp.backbone == 'resnet' and 'visual' in hp.modality
        ds = AffWild2SequenceDataset(split, hp.dataset_path, hp.window, hp.windows_per_epoch, hp.cutout, hp.release,
                                     hp.input_size, hp.modality, hp.resample, inv_test_stride, emit_u8=on_device_aug)
        if hp.distributed:
            # training: the reference's DistributedSampler; evaluation: an un-padded, in-order shard per rank, so the
            # outputs gathered on rank 0 hold every window exactly once (pl.SequentialShardSampler)
            sampler = torch.utils.data.distributed.DistributedSampler(ds) if split == 'train' else \
                pl.SequentialShardSampler(ds)
            return DataLoader(ds, batch_size=hp.batch_size, num_workers=hp.workers, pin_memory=True, sampler=sampler)
        return DataLoader(ds, batch_size=hp.batch_size, shuffle=split == 'train', num_workers=hp.workers,
                          pin_memory=True)

    @pl.data_loader
    def train_dataloader(self):
        return self._loader('train')

    @pl.data_loader
    def val_dataloader(self):
        return self._loader('val', 2 if self.hparams.test_on_val else 1)

    @pl.data_loader
    def test_dataloader(self):
        if self.hparams.test_on_val:
            return self.val_dataloader()
        return self._loader('test', 2)

    @staticmethod
    def add_model_specific_args(parent_parser):
        """The reference's command line (models/model.py:448-493), plus --device_augment (crop / mirror / cutout /
        normalise inside the stem's input kernel from uint8 frames; resnet backbone)."""
        parser = ArgumentParser(parents=[parent_parser])
        flag = dict(action='store_true', default=False)
        for name, kw in (
                ('--backbone', dict(default='v2p_split', type=str)), ('--backend', dict(default='gru', type=str)),
                ('--modality', dict(default='visual', type=str)), ('--fusion_type', dict(default='concat', type=str)),
                ('--freeze_enc', flag), ('--resample', flag), ('--mode', dict(default='video', type=str)),
                ('--window', dict(default=32, type=int)), ('--windows_per_epoch', dict(default=200, type=int)),
                ('--learning_rate', dict(default=5e-5, type=float)), ('--min_lr', dict(default=1e-8, type=float)),
                ('--decay_factor', dict(default=0.5, type=float)), ('--batch_size', dict(default=96, type=int)),
                ('--optimizer', dict(default='adam', type=str)), ('--scheduler', dict(default='plateau', type=str)),
                ('--test_lr', flag), ('--test_on_val', flag), ('--loss', dict(default='ccc_mtl', type=str)),
                ('--loss_lambda', dict(default=0.5, type=float)), ('--num_hidden', dict(default=512, type=int)),
                ('--split_layer', dict(default=3, type=int)), ('--num_fc_layers', dict(default=2, type=int)),
                ('--cutout', flag), ('--distributed', flag),
                ('--dataset_path', dict(default='/.data/zhangyuanhang/Aff-Wild2', type=str)),
                ('--release', dict(default='vipl', type=str)), ('--input_size', dict(default=256, type=int)),
                ('--checkpoint_path', dict(default='.', type=str)), ('--workers', dict(default=8, type=int)),
                ('--max_nb_epochs', dict(default=30, type=int)), ('--device_augment', flag)):
            parser.add_argument(name, **kw)
        return parser
