"""Task module (reference: models/model.py `AffWild2VA`): stream construction, forward, losses, training_step,
optimiser.  It keeps the LightningModule hook names but does not need pytorch_lightning (absent in this image; if it
is importable the class derives from it so the reference's train.py/eval.py can drive it).

Only the hot-path configurations are built: modality in {audio, visual, audiovisual}, backbone in {resnet,
v2p_split(*)}, fusion_type in {concat, attention}.   (*) when models.vggm is available.
"""
import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import fp32, ops
from .att_fusion import AttFusion
from .backbone import VA_3DResNet
from .rnn import GRU
from .utils import concordance_cc2

try:  # pragma: no cover - not installed here
    import pytorch_lightning as _pl
    _Base = _pl.LightningModule
except Exception:  # noqa: BLE001
    _Base = nn.Module


class AffWild2VA(_Base):
    def __init__(self, hparams):
        super().__init__()
        try:
            self.hparams = hparams
        except AttributeError:  # newer Lightning makes hparams read-only
            self.save_hyperparameters(vars(hparams))
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
            else:
                raise NotImplementedError("backbone %r is outside the hot path (SURVEY.md section 2)" % hp.backbone)
        if 'audio' in hp.modality:
            self.audio = GRU(200, 256, 2, rnn_fc_classes, hp.num_fc_layers)
        if hp.modality == 'audiovisual':
            self.proj_v = nn.Linear(hp.num_hidden * (2 if hp.split_layer == 5 else 4), 512)
            if hp.fusion_type == 'attention':
                self.att_fuse = AttFusion([512, 512], 128)
                self.fusion = GRU(512, hp.num_hidden, 2, fc_outputs, hp.num_fc_layers)
            elif hp.fusion_type == 'concat':
                self.fusion = GRU(512 * 2, hp.num_hidden, 2, fc_outputs, hp.num_fc_layers)
            else:
                raise NotImplementedError("fusion_type %r is outside the hot path" % hp.fusion_type)
        self.history = {'lr': [], 'loss': []}

    # ------------------------------------------------------------------ forward
    def _visual(self, batch):
        hp = self.hparams
        if hp.backbone == 'resnet':
            # normalisation (video - 127.5) / 127.5 (reference :106) is folded into the stem's input pass
            if 'video_u8' in batch:
                # decoded uint8 HWC frames + augmentation parameters instead of the float32 clip tensor: crop, mirror
                # and cutout (models/dataset.py:46-80) happen inside the same pass (process/video_input.py)
                return self.visual.forward_bf16(ops.RawClips(batch['video_u8'], batch['video_aug']), normalise=True)
            return self.visual.forward_bf16(batch['video'], normalise=True)
        return self.visual.forward_bf16(batch['video'], batch['se_features'], batch['se_features'], normalise=True)

    def forward(self, batch):
        hp = self.hparams
        if hp.modality == 'audio':
            return ops.as_f32(self.audio.forward_bf16(batch['audio']))
        if hp.modality == 'visual':
            return ops.as_f32(self._visual(batch))
        a = self.audio.forward_bf16(batch['audio'])
        v = self._visual(batch)
        if fp32.enabled():
            v = fp32.linear(v, self.proj_v.weight, self.proj_v.bias)
        else:
            v = ops.linear(v, self.proj_v.weight, self.proj_v.bias)
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

    def training_step(self, batch, batch_idx):
        y_hat = self.forward(batch)
        loss, logs = self.compute_loss(y_hat, batch)
        return {'loss': loss, 'progress_bar': dict(logs), 'log': dict(logs)}

    def validation_step(self, batch, batch_idx):
        y_hat = self.forward(batch).cpu()
        return {'valence_hat': y_hat[..., -2], 'arousal_hat': y_hat[..., -1]}

    # ------------------------------------------------------------------ optimiser (reference :375-407)
    def configure_optimizers(self):
        hp = self.hparams
        if hp.optimizer == 'adam':
            opt = torch.optim.Adam(self.parameters(), lr=hp.learning_rate, weight_decay=1e-4)
        else:
            opt = torch.optim.SGD(self.parameters(), lr=hp.learning_rate, momentum=0.9, weight_decay=5e-4)
        return opt
