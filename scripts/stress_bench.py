"""BASELINE config 4: 4096x3072 stress frames (250 plates -> ~500 light blobs), stage times and frames/s."""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H, B = 4096, 3072, int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = np.stack([synth.make_stress_frame(s, W, H, 250) for s in range(4)] * (B // 4))
c = rb.Context(max_width=W, max_height=H, max_batch=B, max_blobs_per_frame=1024, max_armours_per_frame=4096)
d = c.device_buffer(frames.nbytes); m = c.device_buffer(B * H * W); d.upload(frames)
p = rb.default_params()
for _ in range(3):
    c.detect_batch(d.ptr, W, H, B, p, m.ptr); res = c.fetch_results()
c.profile(True); c.profile_read(reset=True)
steps = 6
c.timer_start()
c.detect_batch(d.ptr, W, H, B, p, m.ptr)
for _ in range(1, steps):
    c.detect_batch(d.ptr, W, H, B, p, m.ptr); res = c.fetch_results()
res = c.fetch_results()
ms = c.timer_stop()
prof = c.profile_read(reset=True)
print("stress %dx%d x%d: %.0f frames/s, %.3f ms/call, stages %s, chunk %d, contours/frame %.0f armours/frame %.0f" % (
    W, H, B, B * steps / (ms * 1e-3), ms / steps, {k: round(v[0] / steps, 3) for k, v in prof.items()}, c.chunk_frames,
    res.total_contours / B, res.total_armours / B))
