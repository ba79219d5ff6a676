"""GPU: sweeps the launch parameters of the pixel kernel (env overrides read by launch_pixel_stage) and of the frame
kernel, printing achieved algorithmic GB/s / frames per second.  Tuning aid; not part of the test suite."""
import itertools
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb  # noqa: E402
from rmcv_b200 import synth  # noqa: E402

W, H, B = 1280, 1024, 512
ctx = rb.Context(max_width=W, max_height=H, max_batch=B)
frames = np.stack([synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in range(32)])
frames = np.concatenate([frames] * (B // 32))
d_in = ctx.device_buffer(frames.nbytes); d_mask = ctx.device_buffer(B * W * H)
d_in.upload(frames)
prm = rb.default_params()


def time_pixel(reps=5):
    ts = []
    for i in range(reps + 2):
        ctx.timer_start()
        ctx.extract_color_batch(d_in.ptr, W, H, B, 1, 80, d_mask.ptr)
        ms = ctx.timer_stop()
        if i >= 2:
            ts.append(ms)
    return statistics.median(ts)


def time_detect(reps=5):
    ts = []
    for i in range(reps + 2):
        ctx.timer_start()
        ctx.detect_batch(d_in.ptr, W, H, B, prm, d_mask.ptr)
        ctx.fetch_results()
        ms = ctx.timer_stop()
        if i >= 2:
            ts.append(ms)
    return statistics.median(ts)


out = []
base = time_pixel()
print(f"default pixel: {base:.3f} ms  {B * W * H * 4 / base / 1e6:.0f} GB/s", flush=True)
for BH, RC, S, NT in itertools.product((16, 32, 64), (4, 6, 8, 12), (2, 3, 4), (128, 160, 256, 320, 480)):
    os.environ.update(RMCV_PIX_BH=str(BH), RMCV_PIX_RC=str(RC), RMCV_PIX_S=str(S), RMCV_PIX_NT=str(NT))
    try:
        ms = time_pixel(3)
    except rb.RmcvError as e:
        print("skip", BH, RC, S, NT, str(e)[:60]); ctx.sync() if False else None
        continue
    gbs = B * W * H * 4 / ms / 1e6
    out.append((gbs, BH, RC, S, NT))
    print(f"BH={BH} RC={RC} S={S} NT={NT}: {ms:.3f} ms {gbs:.0f} GB/s", flush=True)
out.sort(reverse=True)
print("TOP", out[:8])
for k in ("RMCV_PIX_BH", "RMCV_PIX_RC", "RMCV_PIX_S", "RMCV_PIX_NT"):
    os.environ.pop(k, None)
g, BH, RC, S, NT = out[0]
os.environ.update(RMCV_PIX_BH=str(BH), RMCV_PIX_RC=str(RC), RMCV_PIX_S=str(S), RMCV_PIX_NT=str(NT))
for rs in (2048, 4096, 8192):
    os.environ["RMCV_FRAME_RS"] = str(rs)
    ms = time_detect()
    print(f"detect with best pixel params, Rs={rs}: {ms:.3f} ms  {B / ms * 1e3:.0f} frames/s", flush=True)
json.dump(out[:20], open(os.path.join(ROOT, "gpurun_out", "sweep_pixel.json"), "w"))
