"""Host side of the on-device input pipeline: the random draws of the reference's `load_video`
(models/dataset.py:46-80) and `sequence_cutout` (:16-31) turned into the int32 parameter row the fused kernel
`m3t_video_augment_prep_s2d_w4` consumes, in the reference's draw order so that seeding `random` / `numpy.random`
the same way selects the same crop window and cutout hole.

    frames  uint8 [B, T, input_size, input_size, 3]   decoded frames (cv2 order), pinned host or device
    params  int32 [B, 8] = {crop_x, crop_y, flip, cut_y1, cut_y2, cut_x1, cut_x2, 0}
    batch['video_u8'], batch['video_aug'] = frames.cuda(), params.cuda()   ->  AffWild2VA.forward

Only the no-resize case of the reference (input_size 128: crop 112 x 112) is on the device; frames from 256-pixel
tracks must be resized by the decoder stage first.
"""
import random

import numpy as np
import torch


def draw_params(is_training=False, mirror_augment=False, crop_augment=False, cutout_augment=False, input_size=128,
                out_size=112):
    """One clip's parameter row; consumes `random` and `np.random` exactly as load_video + sequence_cutout do."""
    crop_x = crop_y = 0
    if crop_augment:
        if is_training:
            crop_x = random.randint(0, input_size // 8)
            crop_y = random.randint(0, input_size // 8)
        else:
            crop_x = crop_y = input_size // 16
        if input_size * 7 // 8 != out_size:
            raise NotImplementedError("resize branch of load_video (input_size > 128) is not on the device")
    elif input_size != out_size:
        raise ValueError("without crop_augment the frames must already be %d pixels" % out_size)
    flip = 1 if (mirror_augment and is_training) else 0
    cy1 = cy2 = cx1 = cx2 = 0
    if cutout_augment and is_training:
        h = w = out_size
        length = h // 2
        y = np.random.randint(h)
        x = np.random.randint(w)
        cy1, cy2 = int(np.clip(y - length, 0, h)), int(np.clip(y + length, 0, h))
        cx1, cx2 = int(np.clip(x - length, 0, w)), int(np.clip(x + length, 0, w))
    return [crop_x, crop_y, flip, cy1, cy2, cx1, cx2, 0]


def make_batch_inputs(frames_u8, rows, device="cuda"):
    """frames_u8: uint8 tensor [B,T,Hs,Ws,3]; rows: B parameter rows -> (video_u8, video_aug) device tensors."""
    params = torch.tensor(rows, dtype=torch.int32)
    return frames_u8.to(device, non_blocking=True).contiguous(), params.to(device)
