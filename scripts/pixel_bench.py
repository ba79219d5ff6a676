"""Pixel stage alone on BASELINE config 3 frames (RMCV_BGR_STRIP=1: the strip kernel): average launch duration over back-to-back calls (CUDA events)."""
import sys, statistics, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H, B = 1280, 1024, 1024
frames = np.stack([synth.make_frame(s, W, H, 10) for s in range(16)] * (B // 16))
c = rb.Context(max_width=W, max_height=H, max_batch=B)
d = c.device_buffer(frames.nbytes); m = c.device_buffer(B * H * W); d.upload(frames)
ms = []
REP = 4
for i in range(8):
    c.timer_start()
    for _ in range(REP): c.extract_color_batch(d.ptr, W, H, B, rb.CAMP_BLUE, 80, m.ptr)
    t = c.timer_stop() / REP
    if i >= 3: ms.append(t)
t = statistics.median(ms)
print("bgr pixel stage: %.4f ms per 1024 frames  %.0f GB/s (4 B/px)" % (t, B * H * W * 4 / (t * 1e-3) / 1e9))
