"""Test driver with the import and call surface of the reference's train.py / eval.py (which cannot travel to the GPU
box): `pytorch_lightning.Trainer`, `models.model.AffWild2VA`, the module's own argument parser, `load_state_dict` from
a Trainer checkpoint, `fit` or `test`.  Run through `python -m m3t_b200.run tests/scripts/fit_script.py ...`."""
import logging
import random
from argparse import ArgumentParser

import numpy as np
import torch
from pytorch_lightning import Trainer

from models.model import AffWild2VA

logging.basicConfig(level=logging.INFO)

if __name__ == '__main__':
    own = ArgumentParser(add_help=False)
    own.add_argument('--gpus', type=str, default='0')
    own.add_argument('--nodes', type=int, default=1)
    own.add_argument('--seed', type=int, default=12345)
    own.add_argument('--checkpoint', type=str, default='')
    own.add_argument('--evaluate', action='store_true', default=False)
    hparams = AffWild2VA.add_model_specific_args(own).parse_args()
    for seeder in (random.seed, np.random.seed, torch.manual_seed, torch.cuda.manual_seed):
        seeder(hparams.seed)
    model = AffWild2VA(hparams)
    backend = 'ddp' if hparams.distributed else 'dp'
    if hparams.evaluate:
        state = torch.load(hparams.checkpoint, map_location=lambda storage, loc: storage, weights_only=False)
        model.load_state_dict(state['state_dict'])
        Trainer(gpus=hparams.gpus, nb_gpu_nodes=hparams.nodes, distributed_backend=backend).test(model)
    else:
        if hparams.checkpoint:
            model = model.load_from_checkpoint(hparams.checkpoint)
        Trainer(early_stop_callback=None, check_val_every_n_epoch=1, gradient_clip_val=1.0,
                default_save_path=hparams.checkpoint_path, max_epochs=hparams.max_nb_epochs, gpus=hparams.gpus,
                nb_gpu_nodes=hparams.nodes, distributed_backend=backend).fit(model)
