"""Full detection from raw Bayer frames (BASELINE config 2 material at the size of config 3): frames/s and stage times, two
calls in flight.  RMCV_BAYER_GENERIC=1 selects the generic shared-memory Bayer kernel instead of the strip kernel."""
import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H, B = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1280, 1024, 1024)
raw = np.stack([synth.bgr_to_bayer(synth.make_frame(s, W, H, synth.plates_for_seed(s)), synth.BAYER_BG) for s in range(32)] * (B // 32))
import os
c = rb.Context(max_width=W, max_height=H, max_batch=B, chunk_frames=int(os.environ.get("CHUNK", "0")))
d = c.device_buffer(raw.nbytes); m = c.device_buffer(B * H * W); d.upload(raw)
p = rb.default_params()
for _ in range(3):
    c.bayer_detect_batch(d.ptr, W, H, B, synth.BAYER_BG, p, m.ptr); res = c.fetch_results()
c.profile(True); c.profile_read(reset=True)
steps = 8
c.timer_start()
c.bayer_detect_batch(d.ptr, W, H, B, synth.BAYER_BG, p, m.ptr)
for _ in range(1, steps):
    c.bayer_detect_batch(d.ptr, W, H, B, synth.BAYER_BG, p, m.ptr)
    res = c.fetch_results()
res = c.fetch_results()
ms = c.timer_stop()
prof = c.profile_read(reset=True)
print("bayer full detect %dx%d x%d: %.0f frames/s, %.3f ms/step, stages %s, blobs/frame %.1f" % (
    W, H, B, B * steps / (ms * 1e-3), ms / steps, {k: round(v[0] / steps, 3) for k, v in prof.items()}, res.total_blobs / B))
