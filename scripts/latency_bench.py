"""GPU aid: batch-1 latency (BASELINE config 5) of rmcv_detect_batch + rmcv_fetch_results on a device-resident frame."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H = 1280, 1024
ctx = rb.Context(max_width=W, max_height=H, max_batch=1)
frames = [synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in range(16)]
bufs = []
for f in frames:
    b = ctx.device_buffer(f.nbytes); b.upload(f); bufs.append(b)
d_mask = ctx.device_buffer(H * W)
p = rb.default_params()
lat = []
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for i in range(N):
    t0 = time.perf_counter()
    ctx.detect_batch(bufs[i % 16].ptr, W, H, 1, p, d_mask.ptr)
    ctx.fetch_results()
    if i >= N // 6:
        lat.append(1e6 * (time.perf_counter() - t0))
lat.sort()
ctx.profile(True); ctx.profile_read(reset=True)
for i in range(200):
    ctx.detect_batch(bufs[i % 16].ptr, W, H, 1, p, d_mask.ptr); ctx.fetch_results()
prof = ctx.profile_read(reset=True)
print("p50 %.1f us  p99 %.1f us  min %.1f" % (lat[len(lat) // 2], lat[int(len(lat) * 0.99)], lat[0]),
      "| stage us:", {k: round(1e3 * v[0] / 200, 1) for k, v in prof.items()})
