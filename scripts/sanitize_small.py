"""GPU aid: a short tour of the small-chunk (latency-mode) and large-chunk paths for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
n = 0
for (W, H, B, chunk) in ((1280, 1024, 1, 0), (1280, 1024, 5, 2), (640, 480, 16, 0), (333, 77, 7, 1), (1280, 1024, 40, 0), (2048, 1536, 2, 0)):
    frames = np.stack([synth.make_frame(100 + s, W, H, synth.plates_for_seed(100 + s), blue=(s % 2 == 0)) for s in range(B)])
    for target in (rb.CAMP_BLUE, rb.CAMP_RED):
        prm = rb.default_params(target=target)
        with rb.Context(max_width=W, max_height=H, max_batch=B, chunk_frames=chunk) as c:
            res = c.detect_batch_host(frames, prm, np.empty((B, H, W), np.uint8))
            buf = c.device_buffer(frames.nbytes); buf.upload(frames)
            dm = c.device_buffer(B * H * W)
            for _ in range(2):
                c.detect_batch(buf.ptr, W, H, B, prm, dm.ptr); c.detect_batch(buf.ptr, W, H, B, prm, dm.ptr)
                r1 = c.fetch_results(); r2 = c.fetch_results()
            n += sum(res.frames[f].n_armours for f in range(B))
            c.get_label_map(0, W, H); c.get_contours(0)
print("sanitize tour done, armours:", n)
