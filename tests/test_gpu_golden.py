"""The CUDA path against the committed golden fixtures (tests/golden/*.json, made by scripts/make_golden.py from the cv2
oracle): mask bytes, contour order / first pixel / size / area / ordered points, light blobs and armours of the hot path,
the Bayer front, and the next rows f1 (pose), f2 (icon crop) on the same frames.  Nothing here calls the oracle."""
import glob
import json
import os
import zlib

import numpy as np
import pytest

import rmcv_b200 as rb
from rmcv_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_gpu_matches_golden(path):
    rec = json.load(open(path))
    case = rec["case"]
    W, H = case["width"], case["height"]
    img = synth.make_frame(case["seed"], W, H, case["plates"], blue=case["blue"])
    assert zlib.crc32(img.tobytes()) == rec["frame_crc32"], "synthetic generator drifted"
    prm = rb.default_params(target=case["target"])
    with rb.Context(max_width=W, max_height=H, max_batch=1) as c:
        mask = np.empty((1, H, W), np.uint8)
        res = c.detect_batch_host(img[None], prm, mask)
        assert zlib.crc32(mask[0].tobytes()) == rec["mask_crc32"]
        assert int((mask[0] > 0).sum()) == rec["mask_foreground"]
        det = c.frame_detections(res, 0)
        contours = c.get_contours(0)
        assert len(det.contours) == len(rec["contours"]) == len(contours)
        for k, (ci, pts, g) in enumerate(zip(det.contours, contours, rec["contours"])):
            assert list(ci.first) == g["first"] and ci.n_points == g["n"] and ci.area2 == g["area2"], f"contour {k}"
            assert list(ci.bbox) == g["bbox"] and ci.status == g["status"], f"contour {k}"
            assert zlib.crc32(np.ascontiguousarray(pts, np.int32).tobytes()) == g["points_crc32"], f"contour {k} points"
            if g["ellipse"] is not None and ci.fit_branch >= 0 and not (0.7e-10 <= abs(ci.det0) <= 1e-10 * (1 + 1e-6)):
                e, r = ci.ellipse, g["ellipse"]
                assert max(abs(e[0] - r[0]), abs(e[1] - r[1])) <= 1e-3 and abs(e[2] - r[2]) <= 1e-5 * max(1, r[2]) + 1e-4 \
                    and abs(e[3] - r[3]) <= 1e-5 * max(1, r[3]) + 1e-4, f"contour {k} ellipse {e} vs {r}"
        assert len(det.positive) == len(rec["positive"])
        for b, g in zip(det.positive, rec["positive"]):
            assert abs(b.angle - g["angle"]) <= 1e-3 and np.abs(np.array(b.vertices) - np.array(g["vertices"])).max() <= 2e-3
        assert len(det.armours) == len(rec["armours"])
        poses = c.solve_pnp(det.armours, rec["camera"]["matrix"], rec["camera"]["dist"], tuple(rec["camera"]["exact_size"]))
        d = c.device_buffer(img.nbytes)
        d.upload(img)
        icons, _, _ = c.icon_batch(d.ptr, W, H, det.armours)
        d.free()
        for k, (a, g) in enumerate(zip(det.armours, rec["armours"])):
            assert (a.i, a.j) == (g["i"], g["j"]) and list(a.bounding_box) == g["bounding_box"], f"armour {k}"
            assert np.abs(np.array(a.vertices) - np.array(g["vertices"])).max() <= 2e-3
            rvec, tvec, _, ok = poses[k]
            assert ok and np.abs(rvec - g["rvec"]).max() <= 1e-6 and (np.abs(tvec - g["tvec"]) / np.abs(g["tvec"]).max()).max() <= 1e-6
            if np.abs(np.array(a.icon) - np.array(g["icon"])).max() == 0:   # same float vertices -> same bytes
                assert zlib.crc32(icons[k].tobytes()) == g["icon20_crc32"], f"armour {k} icon"
        # Bayer front on the same frame
        raw = synth.bgr_to_bayer(img, synth.BAYER_BG)
        d_in = c.device_buffer(raw.nbytes); d_out = c.device_buffer(H * W)
        d_in.upload(raw)
        c.bayer_extract_color_batch(d_in.ptr, W, H, 1, synth.BAYER_BG, case["target"], 80, d_out.ptr)
        c.sync()
        assert zlib.crc32(d_out.download((H, W)).tobytes()) == rec["bayer_bg_mask_crc32"]
        d_in.free(); d_out.free()
