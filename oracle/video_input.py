"""ORACLE (test infrastructure): CPU restatement of the reference's clip assembly from decoded frames,
models/dataset.py:46-80 `load_video` (crop -> mirror -> stack -> THWC->CTHW float32 -> cutout) and :16-31
`sequence_cutout`, with the random draws passed in explicitly.  Pinned against the reference's own load_video run
on JPEG files written for the purpose (tests/golden/video_input.pt, oracle/make_golden_video_input.py)."""
import numpy as np


def assemble_clip(frames, crop_x, crop_y, flip, cy1, cy2, cx1, cx2, out_size=112, fill_value=127.5):
    """frames: uint8 [T, Hs, Ws, 3] decoded images -> float32 [3, T, out_size, out_size] (values 0..255)."""
    out = []
    for img in frames:
        img = img[crop_y:crop_y + out_size, crop_x:crop_x + out_size]
        if flip:
            img = img[:, ::-1]
        out.append(img)
    seq = np.stack(out).transpose(3, 0, 1, 2).astype(np.float32)
    if cy2 > cy1 and cx2 > cx1:
        seq[:, :, cy1:cy2, cx1:cx2] = fill_value
    return seq
