"""Generate tests/golden/logmel_torchaudio.pt: log-Mel features of seeded waveforms computed by torchaudio
(MelSpectrogram(n_fft 512, win 400, hop int(16000/(3 fps)), 40 Slaney mels, norm 'slaney', power 2, centred) +
AmplitudeToDB('power', top_db 80)) - an independent published implementation of the librosa call the reference makes
(process/extract_melspec.py:13-20; librosa itself is not installed).  Both centre-padding conventions
librosa has used are recorded ('constant' since 0.10, 'reflect' before).
Run on the build box:  python -m oracle.make_golden_logmel
"""
import math
import os

import torch
import torchaudio as ta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    g = torch.Generator().manual_seed(5)
    out = []
    for fps, n, mode in ((30.0, 16000 + 321, "constant"), (25.0, 16000, "reflect")):
        t = torch.arange(n) / 16000.0
        y = 0.3 * torch.sin(2 * math.pi * 440 * t) + 0.05 * torch.randn(n, generator=g) * (t > 0.4) + 1e-4
        hop = int(1 / 3 * 1 / fps * 16000)
        ms = ta.transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=hop, f_min=0, f_max=8000,
                                          n_mels=40, power=2.0, norm="slaney", mel_scale="slaney", center=True,
                                          pad_mode=mode)
        db = ta.transforms.AmplitudeToDB("power", top_db=80)(ms(y)).t().contiguous()
        out.append({"fps": fps, "pad_mode": mode, "wave": y, "logmel_db": db})
    path = os.path.join(ROOT, "tests", "golden", "logmel_torchaudio.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), [tuple(o["logmel_db"].shape) for o in out])


if __name__ == "__main__":
    main()
