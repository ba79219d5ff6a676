"""GPU parity of the whole path (extract_color -> filter_lightblobs -> filter_armours, executable/main.cpp:172-176)
through the C ABI against the cv2 oracle: masks / contour discovery / contour statistics bit-exact, ellipse, light-blob
and armour geometry within the tolerances of tests/_compare.py."""
import json
import os

import numpy as np
import pytest

import rmcv_b200 as rb
from oracle import rm_oracle as O
from rmcv_b200 import synth
from tests import _compare as CMP

pytestmark = pytest.mark.gpu
PRM = CMP.oracle_params()
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def oracle_kwargs(p):
    return dict(target=p["target"], lower_bound=p["lower_bound"], tilt_max=p["tilt_max"], ratio_range=p["ratio_range"],
                area_range=p["area_range"], angle_difference_max=p["angle_difference_max"], shear_max=p["shear_max"],
                lenght_ratio_max=p["lenght_ratio_max"])


def c_params(p):
    return rb.default_params(target=p["target"], lower_bound=p["lower_bound"], tilt_max=p["tilt_max"], ratio_range=p["ratio_range"],
                             area_range=p["area_range"], angle_difference_max=p["angle_difference_max"], shear_max=p["shear_max"],
                             lenght_ratio_max=p["lenght_ratio_max"])


def detect_and_compare(ctx, frames, p=PRM, check_points=True, check_labels=True, what=""):
    B, H, W, _ = frames.shape
    masks = np.empty((B, H, W), np.uint8)
    res = ctx.detect_batch_host(frames, c_params(p), masks)
    total = CMP.Report()
    for f in range(B):
        ref = O.detect_frame(frames[f], **oracle_kwargs(p))
        bad = np.argwhere(masks[f] != ref.binary)
        assert bad.size == 0, f"{what} frame {f}: {len(bad)} mask bytes differ"
        det = ctx.frame_detections(res, f)
        total.merge(CMP.compare_frame(det, ref, p, where=f"{what} frame {f}"))
        if check_points:
            try:
                for k, rc in enumerate(ref.contours):
                    pts = ctx.get_contour(f, k)
                    assert np.array_equal(pts, rc), f"{what} frame {f} contour {k}: ordered points differ"
            except rb.RmcvError as e:
                assert e.status == rb.abi.RMCV_ERR_STATE
        if check_labels:
            try:
                lab = ctx.get_label_map(f, W, H)
                assert np.array_equal(lab, O.blob_label_map(ref.binary, ref.contours)), f"{what} frame {f}: label map (blob pixel sets) differs"
            except rb.RmcvError as e:
                assert e.status == rb.abi.RMCV_ERR_STATE
    return total


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=1280, max_height=1024, max_batch=16) as c:
        yield c


def mask_to_bgr(mask: np.ndarray) -> np.ndarray:
    """BGR frame whose blue-minus-red difference is 200 on the mask (the 3x3 close still applies on both sides)."""
    img = np.zeros(mask.shape + (3,), np.uint8)
    img[..., 0] = np.where(mask, 200, 0)
    return img


def test_config1_single_frame(ctx):
    """BASELINE config 1: seed 1, 8 plates, blue."""
    frame = synth.make_frame(1, 1280, 1024, 8, blue=True)
    rep = detect_and_compare(ctx, frame[None], what="config1")
    assert rep.contours >= 16 and rep.blobs >= 16 and rep.armours >= 8


def test_synthetic_batch_blue_and_red(ctx):
    seeds = list(range(10, 22))
    frames = np.stack([synth.make_frame(s, 1280, 1024, synth.plates_for_seed(s), blue=True) for s in seeds])
    rep = detect_and_compare(ctx, frames, what="blue batch")
    assert rep.direct > 0 and rep.fallback > 0, "both fitEllipseDirect branches must be exercised"
    red = np.stack([synth.make_frame(s, 1280, 1024, synth.plates_for_seed(s), blue=False) for s in seeds[:4]])
    p = CMP.oracle_params(dict(target=rb.CAMP_RED))
    rep2 = detect_and_compare(ctx, red, p, what="red batch")
    rep.merge(rep2)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity_report.json"), "w") as fh:
        json.dump({k: v for k, v in rep.__dict__.items()}, fh, indent=1)
    assert rep.rng_band <= 0.05 * max(rep.fitted, 1)


MICRO = {
    "square3": (np.pad(np.ones((3, 3), bool), 3), 8, 8),            # 3x3 square -> 8 points, area 4 -> area2 8
    "line5": (np.pad(np.ones((5, 1), bool), 3), 8, 0),              # 1x5 line -> 8 points (revisits), area 0
    "pixel": (np.pad(np.ones((1, 1), bool), 3), 1, 0),
    "diag3": (np.pad(np.eye(3, dtype=bool), 3), 4, 0),
    "bar2x8": (np.pad(np.ones((2, 8), bool), 3), 16, 14),           # n=16, area 7
}


@pytest.mark.parametrize("name", sorted(MICRO))
def test_micro_masks_known_answers(ctx, name):
    """Hand-checkable masks.  They are built so that the 3x3 close leaves them unchanged."""
    mask, n, area2 = MICRO[name]
    frame = mask_to_bgr(mask)
    ref = O.detect_frame(frame)
    assert np.array_equal(ref.binary > 0, mask), "micro mask changed by the close; pick another"
    p = c_params(PRM)
    res = ctx.detect_batch_host(frame[None], p)
    det = ctx.frame_detections(res, 0)
    assert len(det.contours) == 1
    assert det.contours[0].n_points == n and det.contours[0].area2 == area2
    assert np.array_equal(ctx.get_contour(0, 0), ref.contours[0])


def test_ring_with_nested_dot_and_corner_pixels(ctx):
    m = np.zeros((24, 24), bool)
    yy, xx = np.mgrid[0:24, 0:24]
    r2 = (yy - 12) ** 2 + (xx - 12) ** 2
    m[(r2 <= 81) & (r2 >= 36)] = True   # ring
    m[11:14, 11:14] = True              # nested dot (must not be reported)
    m[0, 0] = True; m[23, 23] = True    # corner pixels
    frame = mask_to_bgr(m)
    ref = O.detect_frame(frame)
    assert len(ref.contours) == 3
    rep = detect_and_compare(ctx, frame[None], what="ring")
    assert rep.contours == 3


@pytest.mark.parametrize("seed", range(24))
def test_random_masks(ctx, seed):
    """Random masks of random density incl. holes, nesting, border contact, thin parts (findContours semantics)."""
    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.integers(8, 90)), int(rng.integers(8, 140))
    dens = rng.uniform(0.15, 0.75)
    m = rng.random((H, W)) < dens
    if seed % 3 == 1:   # blobby: smooth the noise so that components are large and have holes
        import cv2
        m = cv2.GaussianBlur(m.astype(np.float32), (0, 0), 2.0) > dens * 0.9
    # wide area range so that small blobs are fitted too (degenerate fits are compared only when finite)
    p = CMP.oracle_params(dict(area_range=(10.0, 99999.0)))
    detect_and_compare(ctx, mask_to_bgr(m)[None], p, what=f"random {seed} {W}x{H}")


def test_concavity_next_to_holed_components(ctx):
    """Regression (found by scripts/fuzz_gpu.py): a small component in a concavity of a larger one, where the concavity
    opens across the edge of the bounding box of the frame's holed components — the background gap just inside the box
    must inherit "outer" from the gap below it that lies outside the box, or the small component is taken for a nested
    one and dropped.  The mask (a closed 832x37 noise mask with 52 holes, 324 external contours) is a committed fixture."""
    path = os.path.join(os.path.dirname(__file__), "golden", "mask_concavity_next_to_holes_37x832.npy")
    m = np.unpackbits(np.load(path), axis=1)[:, :832].astype(bool)
    p = CMP.oracle_params(dict(area_range=(10.0, 99999.0)))
    rep = detect_and_compare(ctx, mask_to_bgr(m)[None], p, check_points=False, what="concavity fixture")
    assert rep.contours == 324


@pytest.mark.parametrize("seed", range(6))
def test_wide_noise_masks_with_many_holes(ctx, seed):
    """Closed noise masks, wide and flat (hundreds of components and dozens of holes per frame): every concavity /
    hole / bounding-box configuration of the gap labelling shows up somewhere."""
    rng = np.random.default_rng(7000 + seed)
    H, W = int(rng.integers(24, 60)), int(rng.integers(500, 1000))
    from oracle import cv_restate as R
    m = R.close3x3(rng.random((H, W)) < rng.uniform(0.3, 0.45))
    p = CMP.oracle_params(dict(area_range=(10.0, 99999.0)))
    with rb.Context(max_width=W, max_height=H, max_batch=1, max_blobs_per_frame=2048) as c:
        detect_and_compare(c, mask_to_bgr(m)[None], p, check_points=False, what=f"noise {seed} {W}x{H}")


@pytest.mark.parametrize("seed", range(8))
def test_drawn_shapes(ctx, seed):
    """Rings in rings, arcs holding other shapes in their concavity, spirals, combs (synth.shape_mask): the external
    contours, their order and the blob pixel sets against cv2."""
    rng = np.random.default_rng(9100 + seed)
    W, H = int(rng.integers(200, 900)), int(rng.integers(150, 600))
    from oracle import cv_restate as R
    m = R.close3x3(synth.shape_mask(rng, W, H))
    p = CMP.oracle_params(dict(area_range=(10.0, 99999.0)))
    with rb.Context(max_width=W, max_height=H, max_batch=1, max_blobs_per_frame=4096) as c:
        detect_and_compare(c, mask_to_bgr(m)[None], p, check_points=(seed < 3), what=f"shapes {seed} {W}x{H}")


def test_upright_symmetric_bars(ctx):
    """Exactly mirror-symmetric upright / level bars (rectangles, ellipses, diamonds of many sizes): the xy coefficient of
    the fitted conic is zero and cv::fitEllipseNoDirect leaves the angle of an unswapped box at 0 — the fallback branch of a
    singular direct fit must report 0, not -90 (found by scripts/fuzz_gpu.py: the blob flipped from positive to negative)."""
    import cv2
    img = np.zeros((1024, 1280), np.uint8)
    x = 12
    for k, (w, h) in enumerate([(20, 93), (19, 92), (10, 60), (11, 61), (24, 120), (18, 88), (21, 95), (16, 75), (22, 101), (14, 66),
                                 (20, 94), (20, 92), (23, 110), (17, 80), (12, 57)]):
        for row, kind in enumerate(("rect", "ellipse", "diamond")):
            y = 40 + row * 300
            if kind == "rect":
                img[y:y + h, x:x + w] = 255
            elif kind == "ellipse":
                cv2.ellipse(img, (x + w // 2, y + h // 2), (w // 2, h // 2), 0, 0, 360, 255, -1)
            else:
                cv2.fillConvexPoly(img, np.array([[x + w // 2, y], [x + 2 * (w // 2), y + h // 2], [x + w // 2, y + 2 * (h // 2)], [x, y + h // 2]], np.int32), 255)
        img[940:940 + w, x:x + min(h, 70)] = 255      # level bars
        x += max(w, min(h, 70)) + 14
    rep = detect_and_compare(ctx, mask_to_bgr(img > 0)[None], what="upright symmetric bars")
    assert rep.fallback > 0 and rep.direct > 0


def test_rng_band_blob_whose_jittered_fit_is_far_away():
    """A 17-row noise frame (committed fixture, found by scripts/fuzz_gpu.py) with a thin ragged blob whose direct fit is
    singular inside cv::fitEllipseDirect's RNG band: with the oracle's seed the reference returns a jittered direct fit that
    is 3 px away from its own fallback; the GPU must then equal the fallback exactly (tests/_compare.py)."""
    fr = np.load(os.path.join(os.path.dirname(__file__), "golden", "rng_band_thin_blob.npz"))["frame"]
    p = CMP.oracle_params(dict(target=0, lower_bound=120))
    with rb.Context(max_width=fr.shape[1], max_height=fr.shape[0], max_batch=1) as c:
        rep = detect_and_compare(c, fr[None], p, what="rng band fixture")
    assert rep.rng_band >= 1


def test_large_capacities(ctx):
    """Capacities far above the defaults (8192 blobs, 16384 armours per frame): the wide order kernel cannot stage that
    many light blobs in shared memory and reads them in place."""
    rng = np.random.default_rng(77)
    from oracle import cv_restate as R
    m = R.close3x3(rng.random((120, 400)) < 0.35)
    p = CMP.oracle_params(dict(area_range=(10.0, 99999.0)))
    with rb.Context(max_width=400, max_height=120, max_batch=2, max_blobs_per_frame=8192, max_armours_per_frame=16384) as c:
        detect_and_compare(c, np.stack([mask_to_bgr(m), mask_to_bgr(m[::-1].copy())]), p, check_points=False, what="large capacities")


def test_nested_levels(ctx):
    """Component inside a hole inside a component inside a hole ...: only the outermost is external."""
    m = np.zeros((60, 60), bool)
    for k, r in enumerate(range(28, 2, -4)):
        m[30 - r:30 + r, 30 - r:30 + r] = (k % 2 == 0)
    detect_and_compare(ctx, mask_to_bgr(m)[None], what="nested squares")
    # spiral / comb shapes whose background is connected to the border through long corridors
    c = np.zeros((40, 64), bool)
    c[2:38, 2:62] = True
    for x in range(6, 60, 6):
        c[2:30, x:x + 2] = False
    detect_and_compare(ctx, mask_to_bgr(c)[None], what="comb")


def test_multi_chunk_batch_and_device_api():
    seeds = list(range(40, 51))
    frames = np.stack([synth.make_frame(s, 640, 480, 5, blue=True) for s in seeds])
    with rb.Context(max_width=640, max_height=480, max_batch=16, chunk_frames=3) as c:
        rep = detect_and_compare(c, frames, check_points=True, what="multi-chunk host")
        # device-resident entry point gives identical results
        B, H, W, _ = frames.shape
        d_in = c.device_buffer(frames.nbytes); d_mask = c.device_buffer(B * H * W)
        d_in.upload(frames)
        c.detect_batch(d_in.ptr, W, H, B, c_params(PRM), d_mask.ptr)
        res = c.fetch_results()
        masks = d_mask.download((B, H, W))
        for f in range(B):
            ref = O.detect_frame(frames[f])
            assert np.array_equal(masks[f], ref.binary)
            CMP.compare_frame(c.frame_detections(res, f), ref, PRM, where=f"device api frame {f}")
        d_in.free(); d_mask.free()


def test_standalone_filter_lightblobs_and_armours(ctx):
    """rm::filter_lightblobs / rm::filter_armours on caller-supplied inputs (the reference's stage boundaries)."""
    frame = synth.make_frame(3, 1280, 1024, 14)
    ref = O.detect_frame(frame)
    pos, neg = rb.filter_lightblobs(ref.contours, 70, (1.5, 80), (10, 99999), rb.CAMP_BLUE)
    assert len(pos) == len(ref.positive) and len(neg) == len(ref.negative)
    infos, _ = ctx.filter_lightblobs_raw(ref.contours, c_params(PRM))
    pinfo = [i for i in infos if i.status == rb.CONTOUR_POSITIVE]
    for b, rbk, inf in zip(pos, ref.positive, pinfo):
        tol = CMP.LOOSE["vert"] if CMP.in_rng_band(inf.det0) else CMP.TOL_VERT
        assert float(np.max(np.abs(b.vertices - rbk.vertices))) <= tol
    for inf, v, rc in zip(infos, ref.verdicts, ref.contours):
        assert inf.n_points == v.n and inf.area2 == int(round(2 * v.area)) and inf.first == (int(rc[0][0]), int(rc[0][1]))
    for a, b in zip(neg, ref.negative):
        assert np.array_equal(a, b)
    # armours from the ORACLE's light blobs: identical inputs -> gates and geometry must agree to fp32 rounding
    opos = [rb.LightBlob(b.angle, b.target, b.center, b.vertices, b.size) for b in ref.positive]
    arm = rb.filter_armours(opos, 12, 22, 0.4, rb.CAMP_BLUE)
    assert [(a.i, a.j) for a in arm] == [(a.i, a.j) for a in ref.armours]
    for a, ra in zip(arm, ref.armours):
        assert float(np.max(np.abs(a.icon - ra.icon))) <= 1e-4
        assert float(np.max(np.abs(a.vertices - ra.vertices))) <= 1e-4
        assert a.bounding_box == ra.bounding_box
    # lightblob ctor from oracle ellipses: exact
    boxes = [(v.ellipse.cx, v.ellipse.cy, v.ellipse.w, v.ellipse.h, v.ellipse.angle) for v in ref.verdicts if v.status == 1]
    made = ctx.make_lightblobs(boxes, rb.CAMP_BLUE)
    for b, rbk in zip(made, ref.positive):
        assert np.array_equal(b.vertices, rbk.vertices) and b.angle == np.float32(rbk.angle) and b.size == rbk.size
    # degenerate inputs of the reference (src/objdetect.cpp:120, :64)
    assert rb.filter_armours(opos[:1], 12, 22, 0.4, rb.CAMP_BLUE) == []
    assert rb.filter_lightblobs([], 70, (1.5, 80), (10, 99999), rb.CAMP_BLUE) == ([], [])
    p2, n2 = rb.filter_lightblobs([np.array([[1, 1], [2, 1], [2, 2]], np.int32)], 70, (1.5, 80), (10, 99999), rb.CAMP_BLUE)
    assert p2 == [] and n2 == []


def test_rm_mirror_call_site(ctx):
    """The reference's call site (executable/main.cpp:172-176) written against the mirror API."""
    frame = synth.make_frame(8, 1280, 1024, 6)
    contours, binary = rb.extract_color(frame, rb.CAMP_BLUE, 80)
    positive, negative = rb.filter_lightblobs(contours, 70, (1.5, 80), (10, 99999), rb.CAMP_BLUE)
    armours = rb.filter_armours(positive, 12, 22, 0.4, rb.CAMP_BLUE)
    ref = O.detect_frame(frame)
    assert np.array_equal(binary, ref.binary)
    assert len(contours) == len(ref.contours) and all(np.array_equal(a, b) for a, b in zip(contours, ref.contours))
    assert len(positive) == len(ref.positive) and len(negative) == len(ref.negative)
    assert [(a.i, a.j) for a in armours] == [(a.i, a.j) for a in ref.armours]


def test_capacity_overflow_is_reported():
    frame = synth.make_frame(1, 640, 480, 6)
    with rb.Context(max_width=640, max_height=480, max_batch=1, max_blobs_per_frame=4) as c:
        with pytest.raises(rb.RmcvError) as ei:
            c.detect_batch_host(frame[None], rb.default_params())
        assert ei.value.status == rb.abi.RMCV_ERR_CAPACITY
    with rb.Context(max_width=640, max_height=480, max_batch=1, max_runs_per_frame=16) as c:
        with pytest.raises(rb.RmcvError) as ei:
            c.detect_batch_host(frame[None], rb.default_params())
        assert ei.value.status == rb.abi.RMCV_ERR_CAPACITY


def test_hole_plane_invariant_and_repeatability(ctx):
    """Frames with holes (annulus) processed twice give identical results (the hole plane is restored to zero)."""
    frames = np.stack([synth.make_frame(s, 1280, 1024, 6) for s in (60, 61)])
    a = ctx.detect_batch_host(frames, rb.default_params())
    first = [(ctx.frame_detections(a, f)) for f in range(2)]
    b = ctx.detect_batch_host(frames, rb.default_params())
    second = [(ctx.frame_detections(b, f)) for f in range(2)]
    for x, y in zip(first, second):
        assert [c.__dict__ for c in x.contours] == [c.__dict__ for c in y.contours]
        assert len(x.armours) == len(y.armours)


def test_empty_full_and_degenerate_frames(ctx):
    """Edge cases of the path: no foreground at all, everything foreground, one-pixel-high / one-pixel-wide frames."""
    p = CMP.oracle_params(dict(area_range=(0.0, 1e12)))
    empty = np.zeros((1, 64, 96, 3), np.uint8)
    res = ctx.detect_batch_host(empty, c_params(p))
    assert res.total_contours == 0 and res.total_blobs == 0 and res.total_armours == 0
    full = np.zeros((1, 40, 72, 3), np.uint8); full[..., 0] = 255
    detect_and_compare(ctx, full, p, what="all foreground")
    for shape in ((1, 50), (50, 1), (2, 2), (1, 1), (3, 200)):
        rng = np.random.default_rng(shape[0] * 1000 + shape[1])
        m = rng.random(shape) < 0.6
        detect_and_compare(ctx, mask_to_bgr(m)[None], p, what=f"thin {shape}")


def test_all_targets_full_path(ctx):
    """CAMP_RED and CAMP_GUIDELIGHT through the whole path (channel pairs of src/imgproc.cpp:56-65)."""
    import cv2
    img = synth.make_frame(70, 1280, 1024, 8, blue=False)
    detect_and_compare(ctx, img[None], CMP.oracle_params(dict(target=rb.CAMP_RED)), what="red")
    g = np.zeros((300, 400, 3), np.uint8)
    for k, (x, a) in enumerate(((80, 5), (160, -4), (260, 8), (330, 0))):
        cv2.ellipse(g, ((x, 150), (14, 90), a), (30, 250, 40), -1)
    rep = detect_and_compare(ctx, g[None], CMP.oracle_params(dict(target=rb.CAMP_GUIDELIGHT)), what="guide light")
    assert rep.blobs == 4 and rep.armours >= 1


def test_pitched_device_frames_full_path(ctx):
    """cv::Mat-style row pitch and frame stride larger than the payload, device-resident entry point."""
    W, H, B = 300, 200, 3
    frames = np.stack([synth.make_frame(80 + s, W, H, 3) for s in range(B)])
    pitch, fstride = W * 3 + 52, (W * 3 + 52) * H + 4096
    buf = np.zeros(B * fstride, np.uint8)
    for f in range(B):
        for y in range(H):
            o = f * fstride + y * pitch
            buf[o:o + W * 3] = frames[f, y].ravel()
    mp, ms = W + 20, (W + 20) * H + 512
    d_in = ctx.device_buffer(buf.nbytes); d_mask = ctx.device_buffer(B * ms)
    d_in.upload(buf)
    p = c_params(PRM)
    import ctypes
    ctx._check(ctx.lib.rmcv_detect_batch(ctx.h, d_in.ptr, pitch, fstride, W, H, B, ctypes.byref(p), d_mask.ptr, mp, ms), "rmcv_detect_batch")
    res = ctx.fetch_results()
    masks = d_mask.download((B * ms,))
    for f in range(B):
        ref = O.detect_frame(frames[f])
        got = np.stack([masks[f * ms + y * mp: f * ms + y * mp + W] for y in range(H)])
        assert np.array_equal(got, ref.binary)
        CMP.compare_frame(ctx.frame_detections(res, f), ref, PRM, where=f"pitched frame {f}")
    d_in.free(); d_mask.free()


def test_long_contour_fallback_centre_beyond_2_pow_24():
    """cv::fitEllipseNoDirect sums its centre as a float Point2f point by point (the fallback every contour of >~ 200 points
    takes): once n * x reaches 2^24 that sum rounds in contour order.  The order-free path returns the exactly rounded
    centre and says so (RMCV_FIT_FALLBACK_LONG); below the bound the result is the reference's, float for float."""
    W, H = 4096, 1200
    frame = np.zeros((H, W, 3), np.uint8)
    # a long thin serpentine near the right border: ~6000 contour points at x ~ 3900  ->  n * x ~ 2.3e7 > 2^24
    for k, y in enumerate(range(100, 1100, 40)):
        frame[y:y + 6, 3700:4090, 0] = 230
        x = 4084 if k % 2 == 0 else 3700
        frame[y:y + 46, x:x + 6, 0] = 230
    # the same shape near the origin stays below the bound
    small = np.zeros_like(frame)
    small[:, :400] = frame[:, 3696:4096]
    prm = CMP.oracle_params(dict(area_range=(10.0, 1e9)))
    with rb.Context(max_width=W, max_height=H, max_batch=1) as c:
        for img, expect_long in ((frame, True), (small, False)):
            masks = np.empty((1, H, W), np.uint8)
            res = c.detect_batch_host(img[None], c_params(prm), masks)
            ref = O.detect_frame(img, **oracle_kwargs(prm))
            det = c.frame_detections(res, 0)
            assert np.array_equal(masks[0], ref.binary) and len(det.contours) == len(ref.contours) == 1
            ci = det.contours[0]
            assert ci.n_points == len(ref.contours[0]) > 5000
            sx = int(ref.contours[0][:, 0].sum())
            assert (sx >= 1 << 24) == expect_long
            assert ci.fit_branch == (rb.abi.FIT_FALLBACK_LONG if expect_long else rb.abi.FIT_FALLBACK)
            rep = CMP.compare_frame(det, ref, prm, where="long contour")
            assert rep.long_fallback == (1 if expect_long else 0)
            if not expect_long:
                e = ref.verdicts[0].ellipse
                assert np.float32(ci.ellipse).tobytes() == np.float32([e.cx, e.cy, e.w, e.h, e.angle]).tobytes()


def test_small_calls_are_ordered_against_the_pixel_stream():
    """Calls of a few frames run on a slot stream of their own.  The ordering the header promises must still hold
    without any host synchronisation in between: an async upload on the pixel stream before a call is seen by it, an async
    overwrite of its input after it waits for it, an async mask download after it copies the finished mask."""
    W, H = 1280, 1024
    with rb.Context(max_width=W, max_height=H, max_batch=1) as c:
        prm = rb.default_params(target=rb.CAMP_BLUE)
        hs = [c.pinned((H, W, 3)) for _ in range(3)]
        frames = [synth.make_frame(7100 + s, W, H, synth.plates_for_seed(7100 + s), blue=True) for s in range(3)]
        for hbuf, f in zip(hs, frames):
            hbuf.array[:] = f
        d_frame = c.device_buffer(frames[0].nbytes)
        d_masks = [c.device_buffer(H * W) for _ in range(3)]
        h_masks = [c.pinned((H, W)) for _ in range(3)]
        for rep in range(4):   # several rounds: slots and result sets rotate
            for k in range(3):  # upload k -> detect k -> download mask k, nothing but stream order in between
                c._check(c.lib.rmcv_memcpy_h2d(c.h, d_frame.ptr, hs[k].array.ctypes.data, frames[k].nbytes), "h2d")
                c.detect_batch(d_frame.ptr, W, H, 1, prm, d_masks[k].ptr)
                c._check(c.lib.rmcv_memcpy_d2h(c.h, h_masks[k].array.ctypes.data, d_masks[k].ptr, H * W), "d2h")
            res = [c.fetch_results() for _ in range(3)]
            c.sync()
            for k in range(3):
                ref = O.detect_frame(frames[k], target=rb.CAMP_BLUE)
                assert np.array_equal(h_masks[k].array, ref.binary), "round %d frame %d: mask copied before it was complete" % (rep, k)
                det = c.frame_detections(res[k], 0)
                assert len(det.contours) == len(ref.contours) and len(det.armours) == len(ref.armours), (rep, k)
                h_masks[k].array[:] = 0


def test_call_of_several_small_chunks_is_complete_when_fetched():
    """A call cut into chunks of a few frames runs every chunk as a chain on its slot's stream (three slots, three streams);
    rmcv_fetch_results must wait for all of them, not only for the stream of the last chunk.  Four one-frame chunks land on
    streams 0, 1, 2, 0: frames 0 and 3 are black (their chains are over at once), frames 1 and 2 are crowded, and two different
    batches alternate so that a record left over from the previous call is noticed."""
    W, H, B = 1280, 1024, 4
    prm = rb.default_params(target=rb.CAMP_BLUE)
    black = np.zeros((H, W, 3), np.uint8)
    batches = [np.stack([black, synth.make_frame(7300 + 50 * k, W, H, 40, blue=True), synth.make_frame(7301 + 50 * k, W, H, 40, blue=True), black])
               for k in range(2)]

    def digest(res):
        return [(res.frames[f].n_contours, res.frames[f].n_positive, res.frames[f].n_armours,
                 [bytes(res.contours[res.frames[f].contour_offset + i]) for i in range(res.frames[f].n_contours)]) for f in range(B)]

    with rb.Context(max_width=W, max_height=H, max_batch=B, chunk_frames=1) as c:
        # the host entry point synchronises every stream before it returns
        want = [digest(c.detect_batch_host(fr, prm, np.empty((B, H, W), np.uint8))) for fr in batches]
        assert want[0] != want[1] and want[0][1][0] > 40
        bufs = []
        for fr in batches:
            b = c.device_buffer(fr.nbytes); b.upload(fr); bufs.append(b)
        dm = c.device_buffer(B * H * W)
        for it in range(60):
            k = it & 1
            c.detect_batch(bufs[k].ptr, W, H, B, prm, dm.ptr)
            assert digest(c.fetch_results()) == want[k], "iteration %d: results fetched before every chunk had finished" % it
