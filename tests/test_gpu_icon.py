"""f2 (next row, first half): the icon crop that feeds the SVM — rm::affine_correction + flatten_image per armour on the GPU
(rmcv_icon_batch) against the same OpenCV calls (cv2.getAffineTransform / warpAffine / resize).  Bit-exact."""
import os

import numpy as np
import pytest

import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O

pytestmark = pytest.mark.gpu
SEED_OFFSET = int(os.environ.get("RMCV_TEST_SEED", "0"))   # other random cases: RMCV_TEST_SEED=n pytest -m gpu ...


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=1280, max_height=1024, max_batch=1) as c:
        yield c


def armour_with_icon(icon):
    z = ((0.0, 0.0),) * 4
    return rb.Armour(tuple((float(x), float(y)) for x, y in icon), z, (0.0, 0.0, 0.0, 0.0), 0, 1, (0.0,) * 6)


def check(ctx, frame, armours, out_size=(20, 20), what=""):
    H, W, _ = frame.shape
    d = ctx.device_buffer(frame.nbytes)
    try:
        d.upload(np.ascontiguousarray(frame))
        icons, rows, clamped = ctx.icon_batch(d.ptr, W, H, armours, out_size)
    finally:
        d.free()
    assert len(clamped) == len(armours)
    for k, a in enumerate(armours):
        ref, v = O.affine_correction(frame, np.array(a.icon, np.float32), out_size)
        assert np.array_equal(np.array(clamped[k].icon, np.float32), v), f"{what} armour {k}: clamped vertices"
        bad = np.argwhere(icons[k] != ref)
        assert bad.size == 0, f"{what} armour {k}: {len(bad)} icon bytes differ, first {bad[0]}, icon {a.icon}"
        assert np.array_equal(rows[k], O.flatten_image(ref)[0]), f"{what} armour {k}: float row"
    return icons


def test_icons_of_detected_armours(ctx):
    """The armours the detector finds on synthetic frames (executable/main.cpp:178-181)."""
    n = 0
    for seed in (5, 6, 7):
        frame = synth.make_frame(seed, 1280, 1024, 12)
        res = ctx.detect_batch_host(frame[None], rb.default_params())
        arm = ctx.frame_detections(res, 0).armours
        assert len(arm) >= 5
        check(ctx, frame, arm, what=f"seed {seed}")
        n += len(arm)
    assert n >= 20


def test_random_quadrilaterals_clamping_and_degenerate_boxes(ctx):
    """Random icon quadrilaterals on a random image: tilted, partly outside the frame (the vertices are clamped in place),
    one pixel wide, collinear (singular affine system), and boxes that already have the output size."""
    rng = np.random.default_rng(9 + SEED_OFFSET)
    frame = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    arms = []
    for _ in range(40):
        cx, cy = rng.uniform(-20, 420), rng.uniform(-20, 320)
        w, h = rng.uniform(4, 120), rng.uniform(4, 90)
        ang = rng.uniform(-0.6, 0.6)
        c, s = np.cos(ang), np.sin(ang)
        base = np.array([[-w / 2, h / 2], [-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2]])   # icon order: 0 = left-down, 1 = left-up, 2 = right-up
        arms.append(armour_with_icon(base @ np.array([[c, s], [-s, c]]) + [cx, cy]))
    arms.append(armour_with_icon([(10, 50), (10, 10), (10.2, 10), (10.2, 50)]))       # one pixel wide
    arms.append(armour_with_icon([(5, 5), (50, 50), (100, 100), (150, 150)]))           # collinear
    arms.append(armour_with_icon([(100, 119), (100, 100), (119, 100), (119, 119)]))     # 20 x 20 box: resize copies
    arms.append(armour_with_icon([(-50, -50), (-40, -60), (-30, -50), (-40, -40)]))     # entirely outside: clamps to a corner
    arms.append(armour_with_icon([(0.5, 1.5), (2.5, 0.5), (3.5, 2.5), (1.5, 3.5)]))     # half-integer vertices: round half to even
    check(ctx, frame, arms, what="random")
    check(ctx, frame, arms[:10], out_size=(32, 16), what="random 32x16")


def test_arguments(ctx):
    d = ctx.device_buffer(64 * 64 * 3)
    icons, rows, clamped = ctx.icon_batch(d.ptr, 64, 64, [])
    assert icons.shape == (0, 20, 20, 3) and clamped == []
    with pytest.raises(rb.RmcvError):
        ctx.icon_batch(d.ptr, 64, 64, [armour_with_icon([(1, 1)] * 4)], out_size=(0, 20))
    d.free()


def train_linear_svm(rng, n_per_class=12, classes=(0, 1, 2, 3, 4, 5, 6)):
    """A cv::ml::SVM like executable/svm/optimizer.cpp:16-21 trains (C_SVC, LINEAR), on random 20x20x3 'icons'."""
    import cv2
    X, y = [], []
    for c in classes:
        proto = rng.integers(0, 256, 1200).astype(np.float32)
        for _ in range(n_per_class):
            X.append(np.clip(proto + rng.normal(0, 40, 1200), 0, 255).astype(np.float32)); y.append(c)
    svm = cv2.ml.SVM_create()
    svm.setType(cv2.ml.SVM_C_SVC); svm.setKernel(cv2.ml.SVM_LINEAR)
    svm.setTermCriteria((cv2.TERM_CRITERIA_MAX_ITER + cv2.TERM_CRITERIA_EPS, 1000, 1e-3))
    svm.train(np.array(X), cv2.ml.ROW_SAMPLE, np.array(y, np.int32))
    return svm, np.array(X)


def test_svm_predict_matches_cv2(ctx):
    """f2, second half: cv::ml::SVM::predict (C_SVC, LINEAR, one-vs-one vote) on the GPU against cv2 on training rows,
    random rows and rows near the decision boundaries (averages of two classes)."""
    rng = np.random.default_rng(4 + SEED_OFFSET)
    classes = (1, 2, 3, 4, 5, 6, 9)
    svm, X = train_linear_svm(rng, classes=classes)
    model = rb.SvmModel.from_cv2(svm, classes)
    rows = np.concatenate([X, rng.integers(0, 256, (64, 1200)).astype(np.float32),
                           ((X[rng.integers(0, len(X), 200)] + X[rng.integers(0, len(X), 200)]) / 2).astype(np.float32)])
    ref = svm.predict(rows)[1].reshape(-1).astype(np.int32)
    got = ctx.svm_predict(model, rows)
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} of {len(ref)} labels differ"
    assert len(set(ref.tolist())) >= 5
    assert len(ctx.svm_predict(model, np.zeros((0, 1200), np.float32))) == 0


def test_identify_batch_chain(ctx):
    """executable/main.cpp:178-181: icon crop + identity in one call equals oracle crop + cv2 predict."""
    rng = np.random.default_rng(8 + SEED_OFFSET)
    svm, _ = train_linear_svm(rng)
    model = rb.SvmModel.from_cv2(svm, range(7))
    frame = synth.make_frame(11, 1280, 1024, 14)
    res = ctx.detect_batch_host(frame[None], rb.default_params())
    arm = ctx.frame_detections(res, 0).armours
    d = ctx.device_buffer(frame.nbytes)
    try:
        d.upload(frame)
        got = ctx.identify_batch(d.ptr, 1280, 1024, arm, model)
    finally:
        d.free()
    ref = [int(svm.predict(O.flatten_image(O.affine_correction(frame, np.array(a.icon, np.float32))[0]))[1][0, 0]) for a in arm]
    assert got.tolist() == ref and len(ref) >= 8
    with pytest.raises(rb.RmcvError):
        ctx.identify_batch(0, 1280, 1024, arm, model)
