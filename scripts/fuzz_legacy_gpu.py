"""GPU aid: the legacy rows (rmcv_min_area_rects, rmcv_match_lightblobs with and without the ellipse) on the external
contours of drawn-shape and noise masks against cv2.  usage: fuzz_legacy_gpu.py [cases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O
from oracle import cv_restate as R
from tests import _compare as CMP
from tests.test_gpu_legacy import rect_equal, rect_identical

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
args = (1.5, 80.0, 70.0, 10.0, 99999.0)
n_rect = ties = bad = n_verdict = n_box = 0
with rb.Context(max_width=1280, max_height=1024, max_batch=1) as ctx:
    for n in range(cases):
        W, H = int(rng.integers(100, 900)), int(rng.integers(80, 600))
        m = synth.shape_mask(rng, W, H) if n % 2 == 0 else (cv2.GaussianBlur((rng.random((H, W)) < 0.3).astype(np.float32), (0, 0), 2.0) > 0.33)
        m = R.close3x3(m)
        cs, _ = cv2.findContours(m.astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        cs = [c.reshape(-1, 2).astype(np.int32) for c in cs if len(c) >= 6][:400]
        if not cs:
            continue
        for c, g in zip(cs, ctx.min_area_rects(cs)):
            ref = cv2.minAreaRect(c.reshape(-1, 1, 2))
            n_rect += 1
            if not rect_identical(g, ref):
                bad += 1
                print("MISMATCH minAreaRect", g, ref)
            ties += not (abs(g[2] - ref[1][0]) <= 2e-3 * max(1, ref[1][0]) and abs(g[4] - ref[2]) <= 0.02)
        for fe in (True, False):
            for c, (ok, box) in zip(cs, ctx.match_lightblobs(cs, *args, fit_ellipse=fe)):
                rok, rbox = O.match_lightblob(c, *args, fit_ellipse=fe)
                v = O.classify_contour(c, 70.0, (1.5, 80.0), (10.0, 99999.0))
                fragile = v.status != 0 and (v.ellipse.w < 2.0 or CMP.near_blob_gate(v, CMP.oracle_params()) or not np.isfinite(v.ellipse.w))
                if not fe and rok is not None and rbox is not None:   # ratio taken from minAreaRect: a tie can move it across the gate
                    pass
                n_verdict += 1
                if ok != rok and not fragile:
                    bad += 1
                    print("MISMATCH verdict", fe, ok, rok, v.ellipse, len(c))
                if ok and rok:
                    n_box += 1
                    if fe and not (abs(box[0] - rbox.cx) <= 0.5 and abs(box[1] - rbox.cy) <= 0.5):
                        bad += 1; print("MISMATCH box", box, rbox)
                    if not fe and not rect_equal(box, ((rbox.cx, rbox.cy), (rbox.w, rbox.h), rbox.angle)):
                        bad += 1; print("MISMATCH rect box", box, rbox)
print("fuzz_legacy: %d cases, minAreaRect %d (%d equal-area ties), verdicts %d, boxes %d, %d mismatches" % (cases, n_rect, ties, n_verdict, n_box, bad))
sys.exit(1 if bad else 0)
