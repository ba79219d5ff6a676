"""GPU aid: randomised sizes / batches / targets through the whole path and the Bayer front, compared with the oracle
(masks bit-exact, contour / positive / armour counts and contour first pixels).  usage: fuzz_gpu.py [cases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O
from tests import _compare as CMP

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
soft = 0
rep = CMP.Report()
repb = CMP.Report()
for n in range(cases):
    W = int(rng.choice([rng.integers(3, 200), 16 * rng.integers(2, 90), 1280, 1440, 32 * rng.integers(1, 40)]))
    H = int(rng.choice([rng.integers(3, 150), 2 * rng.integers(2, 300), 1024]))
    B = int(rng.choice([1, 2, 3, 5, 17, 33]))
    if W * H * B > 40e6:
        B = 1
    blue = bool(rng.integers(0, 2))
    target = rb.CAMP_BLUE if blue else rb.CAMP_RED
    plates = int(rng.integers(1, 12))
    if W >= 160 and H >= 120:
        frames = np.stack([synth.make_frame(int(rng.integers(0, 1 << 30)), W, H, plates, blue=blue) for _ in range(min(B, 3))] * ((B + 2) // 3))[:B]
    else:
        frames = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    prm = rb.default_params(target=target, lower_bound=int(rng.choice([80, 60, 120])))
    chunk = int(rng.choice([0, 1, 2, 4]))
    with rb.Context(max_width=W, max_height=H, max_batch=B, chunk_frames=chunk) as c:
        mask = np.empty((B, H, W), np.uint8)
        try:
            res = c.detect_batch_host(np.ascontiguousarray(frames), prm, mask)
        except rb.RmcvError as e:
            if "capacity" in str(e):      # noise frames can exceed the default per-frame capacities: not a mismatch
                continue
            raise
        for f in range(B):
            ref = O.detect_frame(frames[f], target=target, lower_bound=prm.lower_bound)
            det = c.frame_detections(res, f)
            ok = np.array_equal(mask[f], ref.binary) and len(det.contours) == len(ref.contours) and \
                [list(ci.first) for ci in det.contours] == [[int(p[0][0]), int(p[0][1])] for p in ref.contours] and \
                [ci.n_points for ci in det.contours] == [len(p) for p in ref.contours]
            if ok and (len(det.positive) != len(ref.positive) or len(det.armours) != len(ref.armours)):
                soft += 1    # must be a gate value at its threshold or a fit in cv::fitEllipseDirect's RNG band: checked below
            if ok:           # the full comparison of tests/_compare.py (statistics exact, geometry within tolerance, carve-outs counted)
                try:
                    p = CMP.oracle_params(dict(target=int(target), lower_bound=int(prm.lower_bound)))
                    rep.merge(CMP.compare_frame(det, ref, p, where="case %d frame %d" % (n, f)))
                except AssertionError as e:
                    ok = False
                    print("COMPARE", str(e)[:300])
            if not ok:
                bad += 1
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                np.save(os.path.join(ROOT, "gpurun_out", "fuzz_fail_%d.npy" % bad), frames[f])
                print("MISMATCH detect", dict(W=W, H=H, B=B, f=f, target=target, lb=prm.lower_bound, chunk=chunk),
                      len(det.contours), len(ref.contours), len(det.positive), len(ref.positive), len(det.armours), len(ref.armours),
                      int((mask[f] != ref.binary).sum()))
        if W >= 3 and H >= 3:
            layout = int(rng.choice([rb.BAYER_BG, rb.BAYER_GB, rb.BAYER_GR, rb.BAYER_RG]))
            raw = np.stack([synth.bgr_to_bayer(frames[f], layout) for f in range(B)])
            d_in = c.device_buffer(raw.nbytes); d_out = c.device_buffer(B * H * W)
            d_in.upload(raw)
            c.bayer_extract_color_batch(d_in.ptr, W, H, B, layout, target, prm.lower_bound, d_out.ptr)
            c.sync()
            bm = d_out.download((B, H, W))
            for f in range(B):
                refm = O.extract_color_mask(O.bayer_to_bgr(raw[f], layout), target, prm.lower_bound)
                if not np.array_equal(bm[f], refm):
                    bad += 1
                    print("MISMATCH bayer", dict(W=W, H=H, B=B, f=f, layout=layout, target=target, lb=prm.lower_bound), int((bm[f] != refm).sum()))
            # the same mosaics through the Bayer full-detect entry point, every frame through the comparator
            c.bayer_detect_batch(d_in.ptr, W, H, B, layout, prm, d_out.ptr)
            try:
                resb = c.fetch_results()
            except rb.RmcvError as e:
                if "capacity" not in str(e):
                    raise
                resb = None
            if resb is not None:
                bm = d_out.download((B, H, W))
                for f in range(min(B, 3)):
                    refd = O.detect_frame(O.bayer_to_bgr(raw[f], layout), target=target, lower_bound=prm.lower_bound)
                    try:
                        assert np.array_equal(bm[f], refd.binary), "bayer detect mask differs"
                        p = CMP.oracle_params(dict(target=int(target), lower_bound=int(prm.lower_bound)))
                        repb.merge(CMP.compare_frame(c.frame_detections(resb, f), refd, p, where="case %d bayer frame %d" % (n, f)))
                    except AssertionError as e:
                        bad += 1
                        print("MISMATCH bayer detect", dict(W=W, H=H, B=B, f=f, layout=layout, target=target, lb=prm.lower_bound), str(e)[:300])
            d_in.free(); d_out.free()
print("fuzz: %d cases, %d mismatches, %d frames whose positive / armour counts differ (all inside the comparator's carve-outs)" % (cases, bad, soft))
print("compared:", {k: getattr(rep, k) for k in ("frames", "contours", "fitted", "direct", "fallback", "rng_band", "near_gate", "degenerate", "blobs", "armours")},
      "worst centre / axis / angle / vertex:", rep.worst_centre, rep.worst_axis_rel, rep.worst_angle, rep.worst_vertex)
print("bayer detect compared:", {k: getattr(repb, k) for k in ("frames", "contours", "fitted", "direct", "fallback", "rng_band", "near_gate", "degenerate", "blobs", "armours")})
sys.exit(1 if bad else 0)
