"""GPU aid: large frames with hundreds of light blobs (BASELINE config 4 and odd sizes around it) through the whole path,
every frame through the full comparator of tests/_compare.py.  usage: fuzz_stress_gpu.py [cases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O
from tests import _compare as CMP

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
rep = CMP.Report()
bad = 0
for n in range(cases):
    W, H = (4096, 3072) if n % 3 == 0 else (int(rng.integers(1500, 4097)), int(rng.integers(1100, 3073)))
    plates = int(rng.integers(60, 320))
    B = int(rng.choice([1, 2, 3]))
    frames = np.stack([synth.make_stress_frame(int(rng.integers(0, 1 << 30)), W, H, plates) for _ in range(B)])
    if n % 4 == 1:
        frames = np.ascontiguousarray(frames[:, ::-1])
    prm = rb.default_params()
    with rb.Context(max_width=W, max_height=H, max_batch=B, max_blobs_per_frame=2048, max_armours_per_frame=8192) as c:
        mask = np.empty((B, H, W), np.uint8)
        res = c.detect_batch_host(frames, prm, mask)
        for f in range(B):
            ref = O.detect_frame(frames[f])
            try:
                assert np.array_equal(mask[f], ref.binary), "mask differs"
                rep.merge(CMP.compare_frame(c.frame_detections(res, f), ref, CMP.oracle_params(), where="case %d frame %d" % (n, f)))
                lab = c.get_label_map(f, W, H)
                assert np.array_equal(lab, O.blob_label_map(ref.binary, ref.contours)), "label map differs"
            except AssertionError as e:
                bad += 1
                print("MISMATCH", dict(W=W, H=H, B=B, f=f, plates=plates), str(e)[:300])
print("fuzz_stress: %d cases, %d mismatches" % (cases, bad))
print("compared:", {k: getattr(rep, k) for k in ("frames", "contours", "fitted", "direct", "fallback", "rng_band", "near_gate", "degenerate", "blobs", "armours")},
      "worst centre / axis / angle / vertex:", rep.worst_centre, rep.worst_axis_rel, rep.worst_angle, rep.worst_vertex)
sys.exit(1 if bad else 0)
