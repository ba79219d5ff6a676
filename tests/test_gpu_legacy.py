"""GPU parity of the legacy rows of SURVEY §8 (a6 rm::MatchLightBlob / rm::FindLightBlobs incl. cv::minAreaRect and the
bounding-rect camp vote, a7 rm::LightBlobOverlap) through the C ABI against the cv2 oracle."""
import cv2
import numpy as np
import pytest

import rmcv_b200 as rb
from oracle import rm_oracle as O
from rmcv_b200 import synth
from tests import _compare as CMP

pytestmark = pytest.mark.gpu


def R_det0(contour) -> float:
    """|det M| of the first direct-fit attempt of cv::fitEllipseDirect (oracle/cv_restate.py), to recognise its RNG band."""
    from oracle import cv_restate as R
    try:
        return float(R.fit_ellipse_direct(np.asarray(contour, np.int32))["det0"])
    except Exception:
        return 1.0


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=1280, max_height=1024, max_batch=2) as c:
        yield c


def rect_equal(got, ref, tol=2e-3):
    """The SAME rectangle as cv2.minAreaRect: OpenCV's float32 rotating calipers are restated literally (tie rules
    included), so there is no "another minimum-area rectangle" allowance any more — centre within 2e-3 px, sizes and
    angle to float rounding."""
    (cx, cy, w, h, a), ((rx, ry), (rw, rh), ra) = got, ref
    return max(abs(cx - rx), abs(cy - ry)) <= tol and abs(w - rw) <= 1e-5 * max(1, rw) and abs(h - rh) <= 1e-5 * max(1, rh) \
        and abs(a - ra) <= 1e-4


def rect_identical(got, ref):
    (rx, ry), (rw, rh), ra = ref
    return np.asarray(got, np.float32).tobytes() == np.asarray([rx, ry, rw, rh, ra], np.float32).tobytes()


def test_min_area_rect_matches_cv2(ctx):
    """cv::minAreaRect (src/objdetect.cpp:16) bit for bit: synthetic-frame contours, noise blobs, drawn shapes, thin lines,
    single points, collinear points — zero ties, zero differing floats."""
    from oracle import cv_restate as R
    n = 0
    pools = []
    for seed in range(4):
        fr = O.detect_frame(synth.make_frame(400 + seed, 1280, 1024, synth.plates_for_seed(400 + seed)))
        pools.append(fr.contours)
    rng = np.random.default_rng(8)
    for k in range(12):
        W, H = int(rng.integers(100, 700)), int(rng.integers(80, 500))
        m = synth.shape_mask(rng, W, H) if k % 3 == 0 else (cv2.GaussianBlur((rng.random((H, W)) < 0.3).astype(np.float32), (0, 0),
                                                                             float(rng.uniform(1.2, 4.0))) > 0.33)
        cs, _ = cv2.findContours(R.close3x3(m).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        pools.append([c.reshape(-1, 2).astype(np.int32) for c in cs][:600])
    pools.append([np.array([[5, 5]], np.int32), np.array([[5, 5], [9, 5]], np.int32), np.array([[3, 3], [4, 4], [5, 5], [6, 6]], np.int32),
                  np.array([[0, 0], [0, 7], [0, 3]], np.int32), np.array([[2, 9], [2, 9], [2, 9]], np.int32)])
    for cs in pools:
        got = ctx.min_area_rects(cs)
        for c, g in zip(cs, got):
            ref = cv2.minAreaRect(c.reshape(-1, 1, 2).astype(np.int32))
            assert rect_identical(g, ref), f"minAreaRect {g} vs cv2 {ref} ({len(c)} points)"
            n += 1
    assert n > 1500
    # hand-checkable: an axis-aligned 20x100 block -> ((100, 20), -90) in OpenCV 4.13
    ys, xs = np.mgrid[10:110, 30:50]
    block = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    (cx, cy, w, h, a), = ctx.min_area_rects([block])
    assert (round(w), round(h), round(a)) == (99, 19, -90) and abs(cx - 39.5) < 1e-3 and abs(cy - 59.5) < 1e-3


@pytest.mark.parametrize("fit_ellipse", [True, False])
def test_match_and_find_lightblobs_legacy(ctx, fit_ellipse):
    args = (1.5, 80.0, 70.0, 10.0, 99999.0)
    nmatch = 0
    for seed, blue in ((410, True), (411, False), (412, True)):
        img = synth.make_frame(seed, 1280, 1024, synth.plates_for_seed(seed), blue=blue)
        contours, _ = O.extract_color(img, rb.CAMP_BLUE if blue else rb.CAMP_RED, 80)
        got = ctx.match_lightblobs(contours, *args, fit_ellipse=fit_ellipse)
        keep = []
        for c, (ok, box) in zip(contours, got):
            rok, rbox = O.match_lightblob(c, *args, fit_ellipse=fit_ellipse)
            v = O.classify_contour(c, 70.0, (1.5, 80.0), (10.0, 99999.0))
            fragile = v.status != 0 and (v.ellipse.w < 2.0 or CMP.near_blob_gate(v, CMP.oracle_params()))
            if not fragile:
                assert ok == rok, f"seed {seed}: MatchLightBlob verdict {ok} != {rok}"
            if ok and rok:
                nmatch += 1
                if fit_ellipse:
                    band = 0.7e-10 <= abs(R_det0(c)) <= 1e-10 * (1 + 1e-6)     # cv::fitEllipseDirect's own RNG band
                    tolc = 0.5 if band else 2e-3
                    assert abs(box[0] - rbox.cx) <= tolc and abs(box[1] - rbox.cy) <= tolc
                else:
                    assert rect_identical(box, ((rbox.cx, rbox.cy), (rbox.w, rbox.h), rbox.angle))
            keep.append(ok == rok)
        if all(keep):
            blobs = ctx.find_lightblobs_legacy(contours, *args, source=img, fit_ellipse=fit_ellipse)
            ref = O.find_lightblobs_legacy(contours, *args, source=img, fit_ellipse=fit_ellipse)
            assert [b.target for b in blobs] == [b.target for b in ref], "camp vote differs"
            assert all(b.target == (rb.CAMP_BLUE if blue else rb.CAMP_RED) for b in blobs)
            for b, r, c in zip(blobs, ref, [c for c, (ok, _) in zip(contours, got) if ok]):
                band = 0.7e-10 <= abs(R_det0(c)) <= 1e-10 * (1 + 1e-6)
                tolc = 0.5 if (band and fit_ellipse) else 2e-3
                assert abs(b.center[0] - r.center[0]) <= tolc and abs(b.center[1] - r.center[1]) <= tolc
                if not (band and fit_ellipse):
                    assert np.abs(np.asarray(b.vertices) - np.asarray(r.vertices)).max() <= 2e-3
    assert nmatch > 30
    # a non-3-channel source yields nothing (src/objdetect.cpp:35)
    assert ctx.find_lightblobs_legacy(contours, *args, source=img[..., 0], fit_ellipse=fit_ellipse) == []


def test_guide_light_vote(ctx):
    """Green bars: G > B and G > R over the bounding rect -> CAMP_GUIDELIGHT."""
    img = np.zeros((200, 300, 3), np.uint8)
    cv2.ellipse(img, ((100, 100), (12, 70), 5), (40, 255, 60), -1)
    cv2.ellipse(img, ((200, 100), (12, 70), -5), (40, 255, 60), -1)
    contours, _ = O.extract_color(img, rb.CAMP_GUIDELIGHT, 80)
    assert len(contours) == 2
    blobs = ctx.find_lightblobs_legacy(contours, 1.5, 80.0, 70.0, 10.0, 99999.0, source=img)
    ref = O.find_lightblobs_legacy(contours, 1.5, 80.0, 70.0, 10.0, 99999.0, source=img)
    assert [b.target for b in blobs] == [b.target for b in ref] == [rb.CAMP_GUIDELIGHT, rb.CAMP_GUIDELIGHT]


def test_lightblob_overlap(ctx):
    fr = O.detect_frame(synth.make_frame(420, 1280, 1024, 12))
    blobs = sorted(fr.positive, key=lambda b: b.center[0])
    cb = [rb.LightBlob(b.angle, b.target, b.center, b.vertices, b.size) for b in blobs]
    n = len(blobs)
    seen = set()
    for left in range(-1, n):
        for right in range(left, n + 2):
            got = ctx.lightblob_overlap(cb, left, right)
            ref = O.lightblob_overlap(blobs, left, right)
            assert got == ref, f"LightBlobOverlap({left},{right}) = {got}, oracle {ref}"
            seen.add(ref)
    assert seen == {True, False}


def test_solve_pnp_matches_cv2(ctx):
    """f1 (next row): rm::solve_PnP per armour (src/mobility.cpp:166-190, executable/main.cpp:183-192) on the GPU
    against cv2.solvePnP(SOLVEPNP_IPPE_SQUARE) with the reference's camera, plus the camera -> world transform."""
    frames = np.stack([synth.make_frame(s, 1280, 1024, 10) for s in (430, 431)])
    res = ctx.detect_batch_host(frames, rb.default_params())
    M = np.array([[0.0007941130268316332, 0.009683274185178004, -0.9999528006788897, -27.25811584661768],
                  [0.9989588796104363, 0.04560298009571095, 0.001234930707386894, -51.46996511920027],
                  [0.04561278583864914, -0.9989127101040636, -0.009636978810429797, 77.11760876626687],
                  [0.0, 0.0, 0.0, 1.0]])   # h_gripper2camera of executable/main.cpp:18-22
    n = 0
    for f in range(2):
        arm = ctx.frame_detections(res, f).armours
        poses = ctx.solve_pnp(arm, O.MAIN_CAMMAT, O.MAIN_DISCOF, (27.0, 27.0), cam2world=M)
        assert len(poses) == len(arm) > 0
        for a, (rvec, tvec, pos, ok) in zip(arm, poses):
            rr, rt = O.solve_pnp(a.vertices)
            assert ok
            assert np.abs(rvec - rr).max() <= 1e-8 and (np.abs(tvec - rt) / np.abs(rt).max()).max() <= 1e-8
            assert np.abs(pos - O.camera_to_world(rt, M)).max() <= 1e-6 * max(1.0, np.abs(rt).max())
            n += 1
    assert n >= 10
    assert ctx.solve_pnp([], O.MAIN_CAMMAT, O.MAIN_DISCOF) == []
    with pytest.raises(rb.RmcvError):
        ctx.solve_pnp(arm[:1], O.MAIN_CAMMAT, O.MAIN_DISCOF, (27.0, 20.0))


def test_fused_poses_in_detect_results(ctx):
    """f1 fused: with rmcv_set_camera every detect call also solves the pose of every armour it finds; poses[k] belongs
    to armours[k] and equals the standalone rmcv_solve_pnp / cv2 result.  rmcv_clear_camera turns it off."""
    frames = np.stack([synth.make_frame(s, 1280, 1024, 10 + s % 9) for s in range(440, 442)])
    M = np.array([[0.0, 0.0, -1.0, -27.0], [1.0, 0.0, 0.0, -51.0], [0.0, -1.0, 0.0, 77.0], [0.0, 0.0, 0.0, 1.0]])
    res = ctx.detect_batch_host(frames, rb.default_params())
    assert not res.poses
    ctx.set_camera(O.MAIN_CAMMAT, O.MAIN_DISCOF, (27.0, 27.0), cam2world=M)
    try:
        res = ctx.detect_batch_host(frames, rb.default_params())
        assert res.poses
        n = 0
        for f in range(len(frames)):
            det = ctx.frame_detections(res, f)
            assert len(det.poses) == len(det.armours)
            alone = ctx.solve_pnp(det.armours, O.MAIN_CAMMAT, O.MAIN_DISCOF, (27.0, 27.0), cam2world=M)
            for a, (rvec, tvec, pos, ok), (rvec2, tvec2, pos2, ok2) in zip(det.armours, det.poses, alone):
                assert ok and ok2
                assert np.array_equal(rvec, rvec2) and np.array_equal(tvec, tvec2) and np.array_equal(pos, pos2)
                rr, rt = O.solve_pnp(a.vertices)
                assert np.abs(rvec - rr).max() <= 1e-8 and (np.abs(tvec - rt) / np.abs(rt).max()).max() <= 1e-8
                n += 1
        assert n >= 20
        with pytest.raises(rb.RmcvError):
            ctx.set_camera(O.MAIN_CAMMAT, O.MAIN_DISCOF, (27.0, 20.0))
    finally:
        ctx.clear_camera()
    res = ctx.detect_batch_host(frames, rb.default_params())
    assert not res.poses


@pytest.mark.parametrize("bits,mirror,flip", [(8, False, False), (8, True, False), (10, False, True), (12, True, True), (12, False, False)])
def test_camera_frontend_variants(ctx, bits, mirror, flip):
    """f4 (next row): 8/10/12-bit mosaics with mirror / flip (hardware/src/daheng.cpp:91-187) -> front-end kernel ->
    Bayer pixel kernel, against the oracle's restatement of ProcessData followed by extract_color."""
    W, H = 1280, 1024
    scene = synth.make_frame(440 + bits, W, H, 8)
    layout = rb.BAYER_GB if mirror else rb.BAYER_BG           # what daheng::capture passes (daheng.cpp:81)
    # the camera delivers the mosaic BEFORE mirror/flip: build it so that ProcessData ends at `scene`'s orientation
    pre = scene[::-1] if flip else scene
    pre = pre[:, ::-1] if mirror else pre
    sensor_layout = ctx.frontend_layout(layout, W, H, mirror, False)   # layout of the unmirrored sensor data
    raw8 = synth.bgr_to_bayer(np.ascontiguousarray(pre), ctx.frontend_layout(sensor_layout, W, H, False, flip))
    rng = np.random.default_rng(bits)
    if bits > 8:
        sh = 4 if bits == 12 else 2
        raw = (raw8.astype(np.uint16) << sh) | rng.integers(0, 1 << sh, raw8.shape, dtype=np.uint16)
    else:
        raw = raw8
    ref_bgr = O.daheng_process(raw, bits, layout, flip, mirror)
    ref_mask = O.extract_color_mask(ref_bgr, rb.CAMP_BLUE, 80)
    assert ref_mask.any()
    d_raw = ctx.device_buffer(raw.nbytes); d_r8 = ctx.device_buffer(W * H); d_mask = ctx.device_buffer(W * H)
    try:
        d_raw.upload(np.ascontiguousarray(raw))
        ctx.raw_frontend_batch(d_raw.ptr, W, H, 1, bits, mirror, flip, d_r8.ptr)
        lay = ctx.frontend_layout(layout, W, H, False, flip)
        ctx.bayer_extract_color_batch(d_r8.ptr, W, H, 1, lay, rb.CAMP_BLUE, 80, d_mask.ptr)
        ctx.sync()
        mask = d_mask.download((H, W))
    finally:
        d_raw.free(); d_r8.free(); d_mask.free()
    bad = np.argwhere(mask != ref_mask)
    assert bad.size == 0, f"bits {bits} mirror {mirror} flip {flip}: {len(bad)} mask bytes differ, first {bad[0]}"
