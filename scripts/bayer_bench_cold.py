"""GPU aid: BASELINE config 2 (64 x 1440x1080 raw Bayer -> mask) measured COLD: four distinct 199 MB working sets are rotated so
that nothing is served by the 126 MB L2 between repetitions."""
import sys, statistics, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H, B = 1440, 1080, 64
raw = np.stack([synth.bgr_to_bayer(synth.make_frame(s, W, H, 10), synth.BAYER_BG) for s in range(8)] * 8)
c = rb.Context(max_width=W, max_height=H, max_batch=B)
NR = 4
ds = [c.device_buffer(raw.nbytes) for _ in range(NR)]; ms_ = [c.device_buffer(B * H * W) for _ in range(NR)]
for k, d in enumerate(ds):
    d.upload(np.roll(raw, k, axis=0))
t_all = []
REP = 20
for i in range(10):
    c.timer_start()
    for r in range(REP):
        c.bayer_extract_color_batch(ds[r % NR].ptr, W, H, B, synth.BAYER_BG, rb.CAMP_BLUE, 80, ms_[r % NR].ptr)
    t = c.timer_stop() / REP
    if i >= 3:
        t_all.append(t)
t = statistics.median(t_all)
print("bayer pixel stage COLD: %.4f ms per launch  %.0f GB/s (2 B/px)  frac %.3f of 6535" % (t, B * H * W * 2 / (t * 1e-3) / 1e9, B * H * W * 2 / (t * 1e-3) / 1e9 / 6535.4))
