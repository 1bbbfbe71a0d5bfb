"""Times the stem's layout pass (m3t_video_prep_s2d_w4) at bench size: python tests/bench_video_prep.py  (M3T_VIDEO_PREP=0|1)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import raw
v = torch.randint(0, 256, (256, 3, 16, 112, 112), dtype=torch.uint8).float().cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = raw.video_prep_s2d_w4(v, True)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
gb = (v.numel() * 4 + out.numel() * 2) / 1e9
print("M3T_VIDEO_PREP=%s  ms %.3f (best %.3f)  %.0f GB/s" % (os.environ.get("M3T_VIDEO_PREP", "1"), sorted(ts)[len(ts) // 2],
                                                            min(ts), gb / (sorted(ts)[len(ts) // 2] * 1e-3)))
