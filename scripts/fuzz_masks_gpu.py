"""GPU aid: closed noise masks of many sizes and densities (hundreds to thousands of components, holes, nesting, concavities)
and, with FUZZ_SHAPES=1, drawn shapes (rings in rings, C-shapes holding components in their concavity, spirals, combs)
through the whole path; external contours (count, first pixel, size, ordered points of a sample) and the label map against
cv2.  usage: fuzz_masks_gpu.py [cases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import cv_restate as R
from oracle import rm_oracle as O
from tests import _compare as CMP

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
ran = 0
total_contours = 0
rep = CMP.Report()
FULL = os.environ.get("FUZZ_FULL")   # also every contour's statistics, fit, blob and armour through tests/_compare.py
for n in range(cases):
    big = os.environ.get("FUZZ_BIG")        # frames above 2 Mpx: the label kernel works on global arrays
    W = int(rng.integers(1500, 2600)) if big else int(rng.integers(16, 900))
    H = int(rng.integers(1400, 2000)) if big else int(rng.integers(8, 500))
    dens = rng.uniform(0.05, 0.5)
    m = rng.random((H, W)) < dens
    kind = 1 if big else int(rng.integers(0, 3))
    if os.environ.get("FUZZ_SHAPES"):
        kind = 3
        W = int(rng.integers(200, 1400)); H = int(rng.integers(150, 1100))
        m = synth.shape_mask(rng, W, H)
        dens = 0.0
    if kind == 1:
        m = cv2.GaussianBlur(m.astype(np.float32), (0, 0), float(rng.uniform(3.0, 8.0) if big else rng.uniform(1.0, 3.0))) > dens * rng.uniform(0.8, 1.1)
    elif kind == 2:
        m = cv2.dilate(m.astype(np.uint8), np.ones((2, 2), np.uint8)).astype(bool) & (rng.random((H, W)) < 0.9)
    m = R.close3x3(m)
    frame = np.zeros((H, W, 3), np.uint8); frame[..., 0] = np.where(m, 200, 0)
    contours, _ = cv2.findContours(m.astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if len(contours) > 8000:
        continue
    prm = rb.default_params(area_range=(0.0, 1e300))
    with rb.Context(max_width=W, max_height=H, max_batch=1, max_blobs_per_frame=8192, max_armours_per_frame=16384) as c:
        try:
            res = c.detect_batch_host(frame[None], prm)
        except rb.RmcvError as e:
            if "capacity" in str(e):
                continue
            raise
        ran += 1
        total_contours += len(contours)
        det = c.frame_detections(res, 0)
        got = [(tuple(ci.first), ci.n_points) for ci in det.contours]
        want = [((int(p[0][0][0]), int(p[0][0][1])), len(p)) for p in contours]
        ok = got == want
        if ok and len(contours):
            pts = c.get_contours(0)
            for k in range(0, len(contours), max(1, len(contours) // 16)):
                ok = ok and np.array_equal(pts[k], contours[k].reshape(-1, 2))
        if ok and FULL:
            p = CMP.oracle_params(dict(area_range=(10.0, 99999.0)))
            prm2 = rb.default_params()
            try:
                res2 = c.detect_batch_host(frame[None], prm2)
                ref = O.detect_frame(frame)
                rep.merge(CMP.compare_frame(c.frame_detections(res2, 0), ref, p, where="case %d" % n))
            except AssertionError as e:
                ok = False
                print("COMPARE", str(e)[:300])
        if not ok:
            bad += 1
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.save(os.path.join(ROOT, "gpurun_out", "fuzz_mask_fail_%d.npy" % bad), np.packbits(m, axis=1))
            print("MISMATCH", dict(W=W, H=H, kind=kind, dens=round(dens, 3)), len(got), len(want),
                  sorted(set(want) - set(got))[:3], sorted(set(got) - set(want))[:3])
print("fuzz_masks: %d cases, %d compared (%d contours), %d mismatches" % (cases, ran, total_contours, bad))
if FULL:
    print("compared:", {k: getattr(rep, k) for k in ("frames", "contours", "fitted", "direct", "fallback", "rng_band", "near_gate", "degenerate", "blobs", "armours")},
          "worst centre / axis / angle / vertex:", rep.worst_centre, rep.worst_axis_rel, rep.worst_angle, rep.worst_vertex)
sys.exit(1 if bad else 0)
