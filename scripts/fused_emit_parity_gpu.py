"""GPU aid: the opt-in fused pixel+emit kernel (RMCV_FUSED_EMIT=1, read once per process) against the oracle: 24 frames of
1280x1024 so that the fixed-geometry band kernel is the one launched; every frame through the full comparator, label maps
and ordered contour points included.  Exit code 0 = parity."""
import os, sys
os.environ["RMCV_FUSED_EMIT"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O
from tests import _compare as CMP

B, W, H = 24, 1280, 1024
frames = np.stack([synth.make_frame(900 + s, W, H, synth.plates_for_seed(900 + s), blue=(s % 2 == 0)) for s in range(B)])
bad = 0
rep = CMP.Report()
for target in (rb.CAMP_BLUE, rb.CAMP_RED):
    prm = rb.default_params(target=target)
    with rb.Context(max_width=W, max_height=H, max_batch=B) as c:
        masks = np.empty((B, H, W), np.uint8)
        l0 = c.kernel_launches()
        res = c.detect_batch_host(frames, prm, masks)
        launches = c.kernel_launches() - l0
        assert launches == 5, "expected pixel+emit, label, contour, fit, order = 5 launches for one chunk, got %d" % launches
        for f in range(B):
            ref = O.detect_frame(frames[f], target=target)
            try:
                assert np.array_equal(masks[f], ref.binary), "mask differs"
                rep.merge(CMP.compare_frame(c.frame_detections(res, f), ref, CMP.oracle_params(dict(target=int(target))), where="frame %d" % f))
                assert np.array_equal(c.get_label_map(f, W, H), O.blob_label_map(ref.binary, ref.contours)), "label map differs"
                for k, pts in enumerate(c.get_contours(f)):
                    assert np.array_equal(pts, ref.contours[k]), "contour %d points differ" % k
            except AssertionError as e:
                bad += 1
                print("MISMATCH", f, str(e)[:300])
print("fused pixel+emit: %d frames x 2 camps, %d mismatches; compared %s" % (B, bad, {k: getattr(rep, k) for k in ("frames", "contours", "fitted", "blobs", "armours")}))
sys.exit(1 if bad else 0)
