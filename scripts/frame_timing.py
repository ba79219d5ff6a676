"""GPU debug aid: per-phase cycle counts of the frame kernel (RMCV_FRAME_TIMING=1)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RMCV_FRAME_TIMING"] = "1"
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H, B = 1280, 1024, 128
ctx = rb.Context(max_width=W, max_height=H, max_batch=B)
frames = np.stack([synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in range(B)])
d_in = ctx.device_buffer(frames.nbytes); d_in.upload(frames)
for i in range(3):
    ctx.detect_batch(d_in.ptr, W, H, B, rb.default_params(), None)
    res = ctx.fetch_results()
print("contours/frame", res.total_contours / B)
