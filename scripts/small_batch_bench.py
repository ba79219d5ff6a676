"""GPU aid: throughput of small device-resident batches (1280x1024) with two calls in flight, for the chained-launch trade-off."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H = 1280, 1024
out = {}
for B in (1, 4, 16):
    ctx = rb.Context(max_width=W, max_height=H, max_batch=B)
    fr = np.stack([synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in range(B)])
    buf = ctx.device_buffer(fr.nbytes); buf.upload(fr)
    d_mask = ctx.device_buffer(B * H * W)
    p = rb.default_params()
    for depth in (1, 2):
        for _ in range(50):
            ctx.detect_batch(buf.ptr, W, H, B, p, d_mask.ptr); ctx.fetch_results()
        n = 600; t0 = time.perf_counter(); inflight = 0
        for i in range(n):
            ctx.detect_batch(buf.ptr, W, H, B, p, d_mask.ptr); inflight += 1
            if inflight >= depth: ctx.fetch_results(); inflight -= 1
        while inflight: ctx.fetch_results(); inflight -= 1
        dt = time.perf_counter() - t0
        out["b%d_d%d" % (B, depth)] = round(n * B / dt)
    del ctx
print("small batches, frames/s (batch_depth):", out)
