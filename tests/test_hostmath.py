"""CPU: the per-blob / per-pair device arithmetic (rmcv_b200/csrc/blob_math.cuh) compiled for the host by
tests/hostmath (test infrastructure, never part of the product library) against the cv2 oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import rm_oracle as O
from rmcv_b200 import _abi as A
from rmcv_b200 import synth

LIB = os.path.join(os.path.dirname(__file__), "hostmath", "libhostmath.so")


@pytest.fixture(scope="module")
def hm():
    return C.CDLL(LIB)


def fit(hm, c):
    xy = np.ascontiguousarray(c, np.int32)
    box, det = A.RotatedRect(), C.c_double()
    br = hm.hm_fit_points(xy.ctypes.data_as(C.c_void_p), len(xy), C.byref(box), C.byref(det))
    return br, box, det.value


def test_ellipse_lightblob_armour_math_matches_oracle(hm):
    P = A.Params(1, 80, 70, 1.5, 80, 10, 99999, 12, 22, 0.4)
    branches = {1: 0, 2: 0}
    n_arm = 0
    for seed in range(8):
        img = synth.make_frame(seed, 1280, 1024, synth.plates_for_seed(seed))
        fr = O.detect_frame(img)
        blobs = []
        for c, v in zip(fr.contours, fr.verdicts):
            if v.status == 0:
                continue
            br, box, det = fit(hm, c)
            branches[br] += 1
            e = v.ellipse
            if not (0.7e-10 <= det <= 1.0e-10 * (1 + 1e-6)):
                assert max(abs(box.cx - e.cx), abs(box.cy - e.cy)) <= 1e-3
                assert max(abs(box.w - e.w) / e.w, abs(box.h - e.h) / e.h) <= 1e-5
                if e.h / e.w > 1.0001:
                    assert abs(((box.angle - e.angle) + 90) % 180 - 90) <= 1e-3
            rr = A.RotatedRect(e.cx, e.cy, e.w, e.h, e.angle)
            assert hm.hm_blob_gates(C.byref(rr), C.byref(P)) == v.status
            if v.status == 1:
                ob = O.make_lightblob(e, 1)
                lb = A.LightBlob()
                hm.hm_make_lightblob(C.byref(rr), 1, C.byref(lb))
                vv = np.array([[lb.vertices[i][0], lb.vertices[i][1]] for i in range(4)], np.float32)
                assert np.array_equal(vv, ob.vertices) and lb.angle == np.float32(ob.angle)
                assert (lb.size[0], lb.size[1]) == (np.float32(ob.size[0]), np.float32(ob.size[1]))
                blobs.append((ob, lb))
        oa = {(a.i, a.j): a for a in fr.armours}
        for i in range(len(blobs)):
            for j in range(i + 1, len(blobs)):
                g = (C.c_float * 6)()
                ok = hm.hm_pair_gates(C.byref(blobs[i][1]), C.byref(blobs[j][1]), C.byref(P), g)
                assert bool(ok) == ((i, j) in oa)
                og = O.pair_gates(blobs[i][0], blobs[j][0])
                assert all(np.float32(g[k]) == np.float32(og[k]) for k in range(6))
                if ok:
                    n_arm += 1
                    ar = A.Armour()
                    hm.hm_make_armour(C.byref(blobs[i][1]), C.byref(blobs[j][1]), C.byref(ar))
                    a = oa[(i, j)]
                    ic = np.array([[ar.icon[k][0], ar.icon[k][1]] for k in range(4)], np.float32)
                    ve = np.array([[ar.vertices[k][0], ar.vertices[k][1]] for k in range(4)], np.float32)
                    assert np.array_equal(ic, a.icon) and np.array_equal(ve, a.vertices) and tuple(ar.bounding_box) == a.bounding_box
    assert branches[1] > 50 and branches[2] > 5 and n_arm > 50


def _outline(w, h, x0=100, y0=200):
    pts = [(x0 + i, y0) for i in range(w)] + [(x0 + w - 1, y0 + j) for j in range(1, h)] + \
          [(x0 + w - 1 - i, y0 + h - 1) for i in range(1, w)] + [(x0, y0 + h - 1 - j) for j in range(1, h - 1)]
    return np.array(pts, np.int32)


@pytest.mark.parametrize("wh", [(20, 93), (93, 20), (19, 92), (10, 60), (60, 10), (11, 61), (24, 120), (120, 24), (7, 7), (30, 31)])
def test_mirror_symmetric_outlines_both_fit_branches(hm, wh):
    """Axis-aligned, exactly mirror-symmetric contours: the conic's xy coefficient is zero, so both fits take their
    special branches (cv::fitEllipseDirect: theta from the sign of a - c; cv::fitEllipseNoDirect: t = g1 - g0 and an angle
    that is only assigned when the axes are swapped, i.e. 0 for an upright ellipse).  Some of the sizes are singular for
    the direct fit (|det| < 1e-10) and go through the fallback."""
    import cv2
    c = _outline(*wh)
    br, box, det = fit(hm, c)
    cv2.setRNGSeed(0)
    (cx, cy), (w, h), ang = cv2.fitEllipseDirect(c.reshape(-1, 1, 2))
    if 0.7e-10 <= det <= 1.0e-10 * (1 + 1e-6):
        pytest.skip("RNG band")
    assert abs(box.cx - cx) <= 1e-3 and abs(box.cy - cy) <= 1e-3
    assert abs(box.w - w) <= 1e-5 * w + 1e-4 and abs(box.h - h) <= 1e-5 * h + 1e-4
    if h / w > 1.0001:
        assert abs(((box.angle - ang) + 90) % 180 - 90) <= 1e-3, (br, box.angle, ang)   # modulo 180 like tests/_compare.py


def test_extend_cord_special_cases(hm):
    """Vertical / horizontal cords take the exact branches of rm::utils::ExtendCord (src/core.cpp:298-331)."""
    def blob(cx, verts):
        b = A.LightBlob()
        b.angle, b.target = 90.0, 1
        b.center[0], b.center[1] = cx, 50.0
        for i, (x, y) in enumerate(verts):
            b.vertices[i][0], b.vertices[i][1] = x, y
        b.size[0], b.size[1] = 10.0, 40.0
        return b
    L = blob(10.0, [(5, 70), (5, 30), (15, 30), (15, 70)])
    Rr = blob(110.0, [(105, 70), (105, 30), (115, 30), (115, 70)])
    ar = A.Armour()
    hm.hm_make_armour(C.byref(L), C.byref(Rr), C.byref(ar))
    ol = O.LightBlob(90.0, 1, (10.0, 50.0), np.array([(5, 70), (5, 30), (15, 30), (15, 70)], np.float32), (10.0, 40.0))
    orr = O.LightBlob(90.0, 1, (110.0, 50.0), np.array([(105, 70), (105, 30), (115, 30), (115, 70)], np.float32), (10.0, 40.0))
    a = O.make_armour(ol, orr)
    ic = np.array([[ar.icon[k][0], ar.icon[k][1]] for k in range(4)], np.float32)
    assert np.array_equal(ic, a.icon) and tuple(ar.bounding_box) == a.bounding_box


def fit_int(hm, c, ox, oy, P):
    xy = np.ascontiguousarray(c, np.int32)
    box, det, status = A.RotatedRect(), C.c_double(), C.c_int()
    br = hm.hm_fit_points_int(xy.ctypes.data_as(C.c_void_p), len(xy), int(ox), int(oy), C.byref(P), C.byref(box), C.byref(det),
                              C.byref(status))
    return br, box, det.value, status.value


def test_integer_sum_route_matches_oracle_and_point_route(hm):
    """The kernels accumulate exact integer sums about a per-component origin (the root run) and shift them to the
    mean in double; this must agree with the oracle for any origin inside or near the blob."""
    P = A.Params(1, 80, 70, 1.5, 80, 10, 99999, 12, 22, 0.4)
    nfit = nband = 0
    worst = [0.0, 0.0, 0.0]
    for seed in range(8, 14):
        img = synth.make_frame(seed, 1280, 1024, synth.plates_for_seed(seed))
        fr = O.detect_frame(img)
        for c, v in zip(fr.contours, fr.verdicts):
            xs, ys = np.asarray(c)[:, 0], np.asarray(c)[:, 1]
            for ox, oy in ((xs.min(), ys.min()), (xs.max(), ys.max()), (int(c[0][0]), int(c[0][1])), (xs.min() - 700, ys.max() + 700)):
                br, box, det, status = fit_int(hm, c, ox, oy, P)
                assert status == v.status
                if v.status == 0:
                    continue
                e = v.ellipse
                if 0.7e-10 <= det <= 1.0e-10 * (1 + 1e-6):
                    nband += 1
                    continue
                nfit += 1
                worst[0] = max(worst[0], abs(box.cx - e.cx), abs(box.cy - e.cy))
                worst[1] = max(worst[1], abs(box.w - e.w) / e.w, abs(box.h - e.h) / e.h)
                if e.h / e.w > 1.0001:
                    worst[2] = max(worst[2], abs(((box.angle - e.angle) + 90) % 180 - 90))
    assert nfit > 400
    assert worst[0] <= 1e-3 and worst[1] <= 1e-5 and worst[2] <= 1e-3, worst


def test_solve_pnp_ippe_square_matches_cv2(hm):
    """rm::solve_PnP (src/mobility.cpp:166-190): the host build of pnp_math.cuh against cv2.solvePnP(IPPE_SQUARE) on
    armours of synthetic frames and on random quadrilaterals, with the camera of executable/main.cpp:8-17."""
    K = np.ascontiguousarray(O.MAIN_CAMMAT.ravel())
    dist = np.ascontiguousarray(O.MAIN_DISCOF)

    def mine(pts, size=(27.0, 27.0), roi=(0.0, 0.0)):
        p = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(4, 2))
        rv, tv = np.zeros(3), np.zeros(3)
        ok = hm.hm_solve_pnp(p.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p),
                             C.c_float(size[0]), C.c_float(size[1]), C.c_float(roi[0]), C.c_float(roi[1]),
                             rv.ctypes.data_as(C.c_void_p), tv.ctypes.data_as(C.c_void_p))
        return ok, rv, tv

    worst_r = worst_t = 0.0
    n = 0
    quads = []
    for seed in (500, 501):
        fr = O.detect_frame(synth.make_frame(seed, 1280, 1024, 10))
        quads += [(a.vertices, (27.0, 27.0), (0.0, 0.0)) for a in fr.armours]
    rng = np.random.default_rng(5)
    for _ in range(200):
        cx, cy, s = rng.uniform(100, 1180), rng.uniform(100, 900), rng.uniform(15, 200)
        q = np.array([[cx - s, cy - s], [cx - s, cy + s], [cx + s, cy + s], [cx + s, cy - s]], np.float32)
        q += rng.uniform(-0.25 * s, 0.25 * s, (4, 2)).astype(np.float32)
        quads.append((q, (float(rng.uniform(5, 60)), float(rng.uniform(5, 60))) if _ % 3 else (27.0, 27.0),
                      (float(rng.integers(0, 50)), float(rng.integers(0, 50)))))
    for q, size, roi in quads:
        if size[0] != size[1]:
            size = (size[0], size[0])          # IPPE_SQUARE asserts a square object
        rv, tv = O.solve_pnp(q, exact_size=size, roi=roi)
        ok, r2, t2 = mine(q, size, roi)
        assert ok == 1
        worst_r = max(worst_r, float(np.abs(rv - r2).max()))
        worst_t = max(worst_t, float((np.abs(tv - t2) / np.abs(tv).max()).max()))
        n += 1
    assert n > 200
    assert worst_r <= 1e-8 and worst_t <= 1e-8, (worst_r, worst_t)


def test_min_area_rect_literal_calipers_bit_equal_to_cv2():
    """a6: cv::minAreaRect (src/objdetect.cpp:16) as legacy.cu evaluates it — gift-wrapped hull put into cv::convexHull's
    vertex order, OpenCV's float32 rotating calipers restated literally (calipers.cuh) — compiled for the host and
    compared with cv2.minAreaRect: every float identical, no equal-area ties left."""
    import cv2

    class RR(C.Structure):
        _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("w", C.c_float), ("h", C.c_float), ("angle", C.c_float)]
    lib = C.CDLL(LIB)
    rng = np.random.default_rng(3)
    tot = 0
    extra = [np.array([[5, 5]], np.int32), np.array([[5, 5], [9, 5]], np.int32), np.array([[3, 3], [4, 4], [5, 5]], np.int32),
             np.array([[0, 0], [0, 7], [0, 3]], np.int32), np.array([[2, 9], [2, 9]], np.int32),
             np.array([[0, 0], [10, 0], [10, 10], [0, 10]], np.int32), np.array([[0, 0], [10, 0], [10, 10], [0, 10]][::-1], np.int32)]
    for t in range(40):
        W, H = int(rng.integers(60, 500)), int(rng.integers(60, 400))
        m = (cv2.GaussianBlur((rng.random((H, W)) < 0.3).astype(np.float32), (0, 0), float(rng.uniform(1.2, 5))) > 0.33).astype(np.uint8) * 255
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        for c in [c.reshape(-1, 2) for c in cs] + (extra if t == 0 else []):
            xy = np.ascontiguousarray(c, np.int32)
            b = RR()
            lib.hm_min_area_rect(xy.ctypes.data_as(C.POINTER(C.c_int32)), len(xy), C.byref(b))
            r = cv2.minAreaRect(xy.reshape(-1, 1, 2))
            got = np.array([b.cx, b.cy, b.w, b.h, b.angle], np.float32)
            ref = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
            assert got.tobytes() == ref.tobytes(), (len(xy), got, ref)
            tot += 1
    assert tot > 5000
