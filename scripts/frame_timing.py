"""GPU debug aid: CUDA-event time of every stage of the detection path (one chunk at a time, no overlap)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H, B = 1280, 1024, int(os.environ.get("FT_BATCH", "592"))
ctx = rb.Context(max_width=W, max_height=H, max_batch=B)
frames = np.stack([synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in range(min(B, 64))])
frames = np.concatenate([frames] * (-(-B // len(frames))))[:B]
d_in = ctx.device_buffer(frames.nbytes); d_in.upload(frames)
d_mask = ctx.device_buffer(B * H * W)
for i in range(3):
    ctx.detect_batch(d_in.ptr, W, H, B, rb.default_params(), d_mask.ptr)
    res = ctx.fetch_results()
ctx.profile(True); ctx.profile_read(reset=True)
N = 5
for i in range(N):
    ctx.detect_batch(d_in.ptr, W, H, B, rb.default_params(), d_mask.ptr)
    res = ctx.fetch_results()
prof = ctx.profile_read(reset=True)
tot = sum(v[0] for v in prof.values())
print("contours/frame", res.total_contours / B, "chunk", ctx.chunk_frames)
print("[stage us per 1024 frames]", {k: round(1e3 * v[0] / N / B * 1024, 1) for k, v in prof.items()}, "| total", round(1e3 * tot / N / B * 1024, 1))
