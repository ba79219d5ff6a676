"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A one-for-one restatement of rmcv's per-frame detection hot path in Python,
calling the *same OpenCV functions* the reference calls (cv2 4.13.0 here; the
reference pins `opencv4 >= 4.8.0#21`, vcpkg.json:26-33).  The reference's C++
cannot be built in this image (no OpenCV C++ headers/libs), so every rm::
function is restated line by line and the OpenCV calls are made through cv2.

PINNING: /root/reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF RUN HERE: oracle/_ref/librmcv_ref.so is the reference's own
src/core.cpp, src/objdetect.cpp, src/imgproc.cpp, src/mobility.cpp compiled
unmodified against a types-only OpenCV stub whose cv:: calls are served by
this image's cv2 (oracle/Makefile, oracle/ref_bridge.py).  tests/test_ref_pin.py
holds every rm:: restatement below bit-equal to that build (1e5 random boxes,
1e5 pairs, whole synthetic frames, legacy/next rows), and tests/golden/*.json
are that build's outputs ("source": "_ref").  What stays unpinnable: the
OpenCV version (cv2 4.13.0 here vs the pinned >= 4.8.0) and the closed Daheng
SDK behind the Bayer front (row a0).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Citations are relative to /root/reference/.
C++ arithmetic rules restated here (verified with g++ 13.3):
  * `abs(float)`  -> float overload (fabsf)
  * `atan2/sin/cos/pow/sqrt/round/fmax` on floats -> double overloads, rounded
    to float on assignment
  * `-` on CV_8U Mats saturates; inRange is inclusive; MORPH_CLOSE ignores the border.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import cv2
import numpy as np

f32 = np.float32

CAMP_RED, CAMP_BLUE, CAMP_GUIDELIGHT, CAMP_NEUTRAL = 0, 1, 2, -1  # include/core.h:20-23
CV_PI_F32 = float(f32(math.pi))  # static_cast<float>(CV_PI)

#: seed installed before every fitEllipseDirect call: its singular-matrix retry draws from
#: cv::theRNG(), which makes the reference itself call-order dependent.
FIT_RNG_SEED = 0


# --------------------------------------------------------------------------- types
@dataclass
class RotatedRect:  # cv::RotatedRect as returned by cv2: ((cx,cy),(w,h),angle), all float32
    cx: float
    cy: float
    w: float
    h: float
    angle: float

    @staticmethod
    def from_cv(r) -> "RotatedRect":
        (cx, cy), (w, h), a = r
        return RotatedRect(float(f32(cx)), float(f32(cy)), float(f32(w)), float(f32(h)), float(f32(a)))

    def to_cv(self):
        return ((self.cx, self.cy), (self.w, self.h), self.angle)


@dataclass
class LightBlob:  # rm::lightblob, include/core.h:89-99
    angle: float
    target: int
    center: Tuple[float, float]
    vertices: np.ndarray  # 4x2 float32: left-down, left-up, right-up, right-down
    size: Tuple[float, float]  # (width=min, height=max)


@dataclass
class Armour:  # rm::armour public geometry, include/core.h:110-112
    icon: np.ndarray  # 4x2 float32
    vertices: np.ndarray  # 4x2 float32
    bounding_box: Tuple[float, float, float, float]  # x, y, w, h
    i: int = -1  # pair indices into the positive list (not in the reference; kept for tests)
    j: int = -1
    gates: Tuple[float, ...] = field(default_factory=tuple)


def range_contains(lo, hi, v) -> bool:
    """rm::range<T>::contains, include/core.h:40-43 (inclusive; NaN -> False)."""
    return bool(v >= lo and v <= hi)


# --------------------------------------------------------------------------- a1 pixel stage
def channel_pair(target: int) -> Tuple[int, int]:
    """src/imgproc.cpp:56-65: which two planes are subtracted."""
    if target == CAMP_GUIDELIGHT:
        return 1, 2
    return (0, 2) if target == CAMP_BLUE else (2, 0)


def extract_color_mask(image: np.ndarray, target: int, lower_bound: int) -> np.ndarray:
    """src/imgproc.cpp:52-69 — split, saturating difference, inRange, 3x3 close."""
    channels = cv2.split(image)  # :53
    a, b = channel_pair(target)
    gray = cv2.subtract(channels[a], channels[b])  # :58/:63  (CV_8U MatExpr '-' saturates)
    binary = cv2.inRange(gray, int(lower_bound), 255)  # :59/:64
    kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))  # :68
    binary = cv2.morphologyEx(binary, cv2.MORPH_CLOSE, kernel)  # :69
    return binary


def find_external_contours(binary: np.ndarray) -> List[np.ndarray]:
    """src/imgproc.cpp:71-72 — findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE)."""
    contours, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    return [np.ascontiguousarray(c.reshape(-1, 2)) for c in contours]


def extract_color(image: np.ndarray, target: int, lower_bound: int):
    """rm::extract_color, include/imgproc.h:29, src/imgproc.cpp:50-75 -> (contours, binary)."""
    binary = extract_color_mask(image, target, lower_bound)
    return find_external_contours(binary), binary


# --------------------------------------------------------------------------- a0 Bayer front (stand-in)
_BAYER_CODE = {  # Daheng DX_PIXEL_COLOR_FILTER -> OpenCV code giving BGR output
    4: cv2.COLOR_BayerBGGR2BGR if hasattr(cv2, "COLOR_BayerBGGR2BGR") else cv2.COLOR_BayerRG2BGR,  # BAYERBG
    2: cv2.COLOR_BayerGBRG2BGR if hasattr(cv2, "COLOR_BayerGBRG2BGR") else cv2.COLOR_BayerGR2BGR,  # BAYERGB
    3: cv2.COLOR_BayerGRBG2BGR if hasattr(cv2, "COLOR_BayerGRBG2BGR") else cv2.COLOR_BayerGB2BGR,  # BAYERGR
    1: cv2.COLOR_BayerRGGB2BGR if hasattr(cv2, "COLOR_BayerRGGB2BGR") else cv2.COLOR_BayerBG2BGR,  # BAYERRG
}


def bayer_to_bgr(raw: np.ndarray, layout: int = 4) -> np.ndarray:
    """Stand-in for the closed DxRaw8toRGB24(RAW2RGB_NEIGHBOUR) (hardware/src/daheng.cpp:143-148):
    OpenCV bilinear demosaic.  PARITY UNPINNED — the Daheng SDK has no source or binary here."""
    return cv2.cvtColor(raw, _BAYER_CODE[layout])


# --------------------------------------------------------------------------- a3 lightblob ctor
def reorder_vertices(box: RotatedRect) -> np.ndarray:
    """rm::utils::reorder_vertices, src/core.cpp:265-283 (RotatedRect::points + stable y sort)."""
    temp = cv2.boxPoints(box.to_cv()).astype(np.float32)  # RotatedRect::points, :268
    order = np.argsort(temp[:, 1], kind="stable")  # std::sort on 4 items = insertion sort, stable; :271-274
    temp = temp[order]
    swap_up = temp[0, 0] < temp[1, 0]
    swap_down = temp[2, 0] < temp[3, 0]
    out = np.empty((4, 2), np.float32)
    out[0] = temp[2] if swap_down else temp[3]  # left down
    out[1] = temp[0] if swap_up else temp[1]  # left up
    out[2] = temp[1] if swap_up else temp[0]  # right up
    out[3] = temp[3] if swap_down else temp[2]  # right down
    return out


def make_lightblob(box: RotatedRect, target: int) -> LightBlob:
    """rm::lightblob::lightblob, src/core.cpp:9-19."""
    ang = f32(box.angle)
    angle = ang - f32(90) if ang > 90 else ang + f32(90)
    verts = reorder_vertices(box)
    size = (float(min(f32(box.h), f32(box.w))), float(max(f32(box.h), f32(box.w))))
    return LightBlob(float(angle), int(target), (float(f32(box.cx)), float(f32(box.cy))), verts, size)


# --------------------------------------------------------------------------- a2 filter_lightblobs
def fit_ellipse_direct(contour: np.ndarray) -> RotatedRect:
    cv2.setRNGSeed(FIT_RNG_SEED)
    return RotatedRect.from_cv(cv2.fitEllipseDirect(contour.reshape(-1, 1, 2).astype(np.int32)))


@dataclass
class ContourVerdict:  # what filter_lightblobs decides per contour (kept for tests)
    n: int
    area: float
    status: int  # 0 skipped (:64), 1 positive, 2 negative
    ellipse: RotatedRect | None = None
    ratio: float = float("nan")
    tilt: float = float("nan")  # |angle-90|


STATUS_SKIPPED, STATUS_POSITIVE, STATUS_NEGATIVE = 0, 1, 2


def classify_contour(contour: np.ndarray, tilt_max, ratio_range, area_range) -> ContourVerdict:
    """Loop body of rm::filter_lightblobs, src/objdetect.cpp:62-84."""
    n = int(contour.shape[0])
    area = float(cv2.contourArea(contour.reshape(-1, 1, 2).astype(np.int32))) if n > 0 else 0.0
    if n < 6 or not range_contains(float(area_range[0]), float(area_range[1]), area):  # :64
        return ContourVerdict(n, area, STATUS_SKIPPED)
    e = fit_ellipse_direct(contour)  # :68   (minAreaRect at :69 is dead code)
    negative = False
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = f32(max(f32(e.w), f32(e.h))) / f32(min(f32(e.w), f32(e.h)))  # :71-73
    if not range_contains(f32(ratio_range[0]), f32(ratio_range[1]), ratio):  # :74
        negative = True
    ang = f32(e.angle)
    angle = ang - f32(90) if ang > 90 else ang + f32(90)  # :78
    tilt = abs(f32(angle - f32(90)))  # :79  abs(float) -> fabsf
    if tilt > f32(tilt_max):
        negative = True
    return ContourVerdict(n, area, STATUS_NEGATIVE if negative else STATUS_POSITIVE, e, float(ratio), float(tilt))


def filter_lightblobs(contours: Sequence[np.ndarray], tilt_max, ratio_range, area_range, enemy):
    """rm::filter_lightblobs, include/objdetect.h:47-49, src/objdetect.cpp:55-87 -> (positive, negative)."""
    positive: List[LightBlob] = []
    negative: List[np.ndarray] = []
    for c in contours:
        v = classify_contour(c, tilt_max, ratio_range, area_range)
        if v.status == STATUS_SKIPPED:
            continue
        if v.status == STATUS_NEGATIVE:
            negative.append(c)  # :82
        else:
            positive.append(make_lightblob(v.ellipse, enemy))  # :83
    return positive, negative


# --------------------------------------------------------------------------- a5 armour ctor helpers
def point_distance(p1, p2) -> np.float32:
    """rm::utils::PointDistance(Point2f), src/core.cpp:285-288: float diffs, pow/sqrt in double."""
    dx = float(f32(p1[0]) - f32(p2[0]))
    dy = float(f32(p1[1]) - f32(p2[1]))
    return f32(math.sqrt(dx * dx + dy * dy))


def extend_cord(pt1, pt2, delta_len) -> Tuple[np.ndarray, np.ndarray]:
    """rm::utils::ExtendCord, src/core.cpp:295-380."""
    p1x, p1y, p2x, p2y = f32(pt1[0]), f32(pt1[1]), f32(pt2[0]), f32(pt2[1])
    d = f32(delta_len)
    d1 = np.zeros(2, np.float32)
    d2 = np.zeros(2, np.float32)
    if p1x == p2x:
        d1[0] = p1x
        d2[0] = p1x
        if p1y > p2y:
            d1[1] = p1y + d
            d2[1] = p2y - d
        else:
            d1[1] = p1y - d
            d2[1] = p2y + d
    elif p1y == p2y:
        d1[1] = p1y
        d2[1] = p1y
        if p1x > p2x:
            d1[0] = p1x + d
            d2[0] = p2x - d
        else:
            d1[0] = p1x - d
            d2[0] = p2x + d
    else:
        k = f32(p1y - p2y) / f32(p1x - p2x)
        theta = f32(math.atan2(float(abs(f32(p1y - p2y))), float(abs(f32(p1x - p2x)))))
        zoom_y = f32(math.sin(float(theta)) * float(d))
        zoom_x = f32(math.cos(float(theta)) * float(d))
        if k > 0:
            if p1x > p2x:
                d1[:] = (p1x + zoom_x, p1y + zoom_y)
                d2[:] = (p2x - zoom_x, p2y - zoom_y)
            else:
                d1[:] = (p1x - zoom_x, p1y - zoom_y)
                d2[:] = (p2x + zoom_x, p2y + zoom_y)
        else:
            if p1x < p2x:
                d1[:] = (p1x - zoom_x, p1y + zoom_y)
                d2[:] = (p2x + zoom_x, p2y - zoom_y)
            else:
                d1[:] = (p1x + zoom_x, p1y - zoom_y)
                d2[:] = (p2x - zoom_x, p2y + zoom_y)
    return d1, d2


def line_center(p1, p2) -> np.ndarray:
    """rm::utils::LineCenter, src/core.cpp:401-404."""
    return np.array([f32(p1[0]) / f32(2) + f32(p2[0]) / f32(2), f32(p1[1]) / f32(2) + f32(p2[1]) / f32(2)], np.float32)


def calc_perspective(inp: np.ndarray, out_ratio=1.0) -> np.ndarray:
    """rm::utils::CalcPerspective, src/core.cpp:382-399."""
    left = point_distance(inp[0], inp[1])
    right = point_distance(inp[2], inp[3])
    max_h = f32(max(float(left), float(right)))  # fmax in double
    size_w = f32(max_h * f32(out_ratio))
    size_h = max_h
    c = line_center(line_center(inp[0], inp[1]), line_center(inp[2], inp[3]))
    out = np.empty((4, 2), np.float32)
    out[0] = (c[0] - size_w / f32(2), c[1] - size_h / f32(2))
    out[1] = (c[0] - size_w / f32(2), c[1] + size_h / f32(2))
    out[2] = (c[0] + size_w / f32(2), c[1] + size_h / f32(2))
    out[3] = (c[0] + size_w / f32(2), c[1] - size_h / f32(2))
    return out


def make_armour(b0: LightBlob, b1: LightBlob) -> Armour:
    """rm::armour::armour, src/core.cpp:21-49 (geometry only; the KalmanFilter member is omitted)."""
    blobs = [b0, b1]
    if f32(b1.center[0]) < f32(b0.center[0]):  # std::sort of 2 items by center.x, :26-30
        blobs = [b1, b0]
    L, R = blobs
    v = np.empty((4, 2), np.float32)
    v[0], v[1], v[2], v[3] = L.vertices[3], L.vertices[2], R.vertices[1], R.vertices[0]  # :32-37
    dl = point_distance(v[0], v[1])
    dr = point_distance(v[2], v[3])
    off_l = f32(_c_round(float(f32(f32(dl / f32(0.5)) - dl) / f32(2))))  # :41
    off_r = f32(_c_round(float(f32(f32(dr / f32(0.5)) - dr) / f32(2))))  # :42
    icon = np.empty((4, 2), np.float32)
    icon[0], icon[1] = extend_cord(v[0], v[1], off_l)  # :43
    icon[3], icon[2] = extend_cord(v[3], v[2], off_r)  # :44
    x, y, w, h = cv2.boundingRect(icon.reshape(-1, 1, 2))  # :46  (Rect -> Rect2f)
    verts = calc_perspective(v)  # :48
    return Armour(icon, verts, (float(x), float(y), float(w), float(h)))


def _c_round(x: float) -> float:
    """C round(): half away from zero."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


# --------------------------------------------------------------------------- a4 filter_armours
def pair_gates(bi: LightBlob, bj: LightBlob):
    """Gate quantities of rm::filter_armours for one pair, src/objdetect.cpp:131-159.
    Returns (angle_difference, shear_i, shear_j, length_ratio, dy, dx, hsum)."""
    ai, aj = f32(bi.angle), f32(bj.angle)
    angle_difference = abs(f32(ai - aj))  # :131
    y = abs(f32(f32(bi.center[1]) - f32(bj.center[1])))  # :135
    x = abs(f32(f32(bi.center[0]) - f32(bj.center[0])))  # :136
    rect_angle = f32(math.atan2(float(y), float(x)) * 180.0 / CV_PI_F32)  # :137 (double, rounded once)

    def shear(a):
        if a > 90:
            return abs(f32(abs(f32(a - rect_angle)) - f32(90)))  # :138-140
        return abs(f32(abs(f32(f32(f32(180) - a) - rect_angle)) - f32(90)))

    shear_i, shear_j = shear(ai), shear(aj)
    hi, hj = f32(bi.size[1]), f32(bj.size[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = f32(min(hi, hj)) / f32(max(hi, hj))  # :149
    hsum = f32(hi + hj)
    return (float(angle_difference), float(shear_i), float(shear_j), float(ratio), float(y), float(x), float(hsum))


def pair_passes(g, angle_difference_max, shear_max, lenght_ratio_max) -> bool:
    ad, si, sj, ratio, dy, dx, hsum = (f32(v) for v in g)
    if ad > f32(angle_difference_max):  # :132
        return False
    if si > f32(shear_max) or sj > f32(shear_max):  # :144
        return False
    if ratio < f32(lenght_ratio_max):  # :150
        return False
    if dy > f32(hsum / f32(2)):  # :153-154
        return False
    if dx > f32(hsum * f32(2)):  # :157-158
        return False
    return True


def filter_armours(lightblobs: Sequence[LightBlob], angle_difference_max, shear_max, lenght_ratio_max, enemy) -> List[Armour]:
    """rm::filter_armours, include/objdetect.h:70-71, src/objdetect.cpp:114-166."""
    armours: List[Armour] = []
    n = len(lightblobs)
    if n < 2:
        return armours
    for i in range(n - 1):
        if lightblobs[i].target != enemy:
            continue
        for j in range(i + 1, n):
            if lightblobs[j].target != enemy:
                continue
            g = pair_gates(lightblobs[i], lightblobs[j])
            if not pair_passes(g, angle_difference_max, shear_max, lenght_ratio_max):
                continue
            a = make_armour(lightblobs[i], lightblobs[j])  # :161
            a.i, a.j, a.gates = i, j, g
            armours.append(a)
    return armours


# --------------------------------------------------------------------------- a6 legacy
def match_lightblob(contour: np.ndarray, min_ratio, max_ratio, tilt_angle, min_area, max_area, fit_ellipse=True):
    """rm::MatchLightBlob, src/objdetect.cpp:9-28 -> (ok, box)."""
    n = int(contour.shape[0])
    c = contour.reshape(-1, 1, 2).astype(np.int32)
    if n < 6:
        return False, None
    area = cv2.contourArea(c)
    if area < f32(min_area) or area > f32(max_area):  # :12 (double vs float compare)
        return False, None
    ellipse = fit_ellipse_direct(contour)  # :15
    box = ellipse if fit_ellipse else RotatedRect.from_cv(cv2.minAreaRect(c))  # :16
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = f32(max(f32(box.w), f32(box.h))) / f32(min(f32(box.w), f32(box.h)))  # :19
    if ratio > f32(max_ratio) or ratio < f32(min_ratio):
        return False, None
    ang = f32(ellipse.angle)
    angle = ang - f32(90) if ang > 90 else ang + f32(90)  # :23
    if abs(f32(angle - f32(90))) > f32(tilt_angle):
        return False, None
    return True, box


def find_lightblobs_legacy(contours, min_ratio, max_ratio, tilt_angle, min_area, max_area, source, fit_ellipse=True):
    """rm::FindLightBlobs, src/objdetect.cpp:30-53 (camp from the bbox mean of the source image)."""
    out: List[LightBlob] = []
    if source.ndim != 3 or source.shape[2] != 3:
        return out
    for c in contours:
        ok, box = match_lightblob(c, min_ratio, max_ratio, tilt_angle, min_area, max_area, fit_ellipse)
        if not ok:
            continue
        x, y, w, h = cv2.boundingRect(c.reshape(-1, 1, 2).astype(np.int32))
        m = cv2.mean(source[y:y + h, x:x + w])  # :43
        if m[1] > m[0] and m[1] > m[2]:
            out.append(make_lightblob(box, CAMP_GUIDELIGHT))
        else:
            out.append(make_lightblob(box, CAMP_BLUE if m[0] > m[2] else CAMP_RED))
    return out


def lightblob_overlap(blobs: Sequence[LightBlob], left: int, right: int) -> bool:
    """rm::LightBlobOverlap, src/objdetect.cpp:89-112.  The reference's bound check admits
    right == size() (one past the end, UB in C++); the restatement rejects it instead."""
    if left < 0 or right >= len(blobs) or right - left < 2:
        return False
    if blobs[left].target != blobs[right].target:
        return False
    lower_y = min(min(blobs[left].vertices[1][1], blobs[left].vertices[2][1]),
                  min(blobs[right].vertices[1][1], blobs[right].vertices[2][1]))
    upper_y = max(max(blobs[left].vertices[0][1], blobs[left].vertices[3][1]),
                  max(blobs[right].vertices[0][1], blobs[right].vertices[3][1]))
    for i in range(left, right):
        if blobs[i].target != blobs[left].target:
            continue
        cx, cy = blobs[i].center
        if blobs[left].center[0] < cx < blobs[right].center[0] and lower_y < cy < upper_y:
            return True
    return False


# --------------------------------------------------------------------------- f4: camera front-end variants (next row)
def daheng_process(raw: np.ndarray, bits: int, layout: int, flip: bool, mirror: bool) -> np.ndarray:
    """hardware/src/daheng.cpp:91-187 (ProcessData) with the declared stand-ins for the closed SDK calls:
    DxImageMirror(HORIZONTAL_MIRROR) = fliplr of the raw mosaic; DxRaw16toRaw8(DX_BIT_2_9 / DX_BIT_4_11) = bits 2..9 /
    4..11; DxRaw8toRGB24(RAW2RGB_NEIGHBOUR, layout, flip) = cv2 bilinear demosaic with that layout, then flipud.
    `layout` is what the caller passes to DxRaw8toRGB24, i.e. the layout of the mosaic AFTER the mirror (daheng.cpp:81)."""
    r = np.asarray(raw)
    if mirror:
        r = r[:, ::-1]
    if bits > 8:
        r = ((r.astype(np.uint16) >> (4 if bits == 12 else 2)) & 0xFF).astype(np.uint8)
    bgr = bayer_to_bgr(np.ascontiguousarray(r, np.uint8), layout)
    return np.ascontiguousarray(bgr[::-1]) if flip else bgr


# --------------------------------------------------------------------------- f1: rm::solve_PnP (next row)
#: camera intrinsics of the reference's only caller, executable/main.cpp:8-14 (float literals stored into double Mats)
MAIN_CAMMAT = np.array([[f32(1782.672144409928), 0.0, f32(598.8983414505224)],
                        [0.0, f32(1783.860175007369), f32(523.4209809658056)],
                        [0.0, 0.0, 1.0]], np.float64)
MAIN_DISCOF = np.array([f32(-0.03436366268485048), f32(0.1953669264956857), f32(0.0001485060439399386),
                        f32(-0.003814875777013483), f32(-0.3181808766352414)], np.float64)


def solve_pnp(points_image, camera_matrix=MAIN_CAMMAT, dist_coeffs=MAIN_DISCOF, exact_size=(27.0, 27.0), roi=(0, 0)):
    """rm::solve_PnP, src/mobility.cpp:166-190: cv::solvePnP(SOLVEPNP_IPPE_SQUARE) on the four armour vertices
    (points 1, 2, 3, 0 against the canonical square).  Returns (rvec[3], tvec[3]) as float64."""
    w, h = f32(exact_size[0]), f32(exact_size[1])
    obj = np.array([[-w / f32(2), h / f32(2), 0], [w / f32(2), h / f32(2), 0],
                    [w / f32(2), -h / f32(2), 0], [-w / f32(2), -h / f32(2), 0]], np.float32)
    p = np.asarray(points_image, np.float32).reshape(4, 2)
    off = np.array([f32(roi[0]), f32(roi[1])], np.float32)
    coord = np.stack([p[1] + off, p[2] + off, p[3] + off, p[0] + off]).astype(np.float32)
    ok, rvec, tvec = cv2.solvePnP(obj, coord, np.asarray(camera_matrix, np.float64), np.asarray(dist_coeffs, np.float64),
                                  flags=cv2.SOLVEPNP_IPPE_SQUARE)
    assert ok
    return rvec.ravel().astype(np.float64), tvec.ravel().astype(np.float64)


def camera_to_world(tvec, cam2world):
    """executable/main.cpp:186-192: world = M * [tvec; 1] for a 4x4 homogeneous transform M."""
    v = np.asarray(cam2world, np.float64).reshape(4, 4) @ np.array([tvec[0], tvec[1], tvec[2], 1.0])
    return v[:3]


# --------------------------------------------------------------------------- whole path + derived oracles
@dataclass
class FrameResult:
    binary: np.ndarray
    contours: List[np.ndarray]
    verdicts: List[ContourVerdict]
    positive: List[LightBlob]
    negative: List[np.ndarray]
    armours: List[Armour]


def detect_frame(image: np.ndarray, target=CAMP_BLUE, lower_bound=80, tilt_max=70.0, ratio_range=(1.5, 80.0),
                 area_range=(10.0, 99999.0), angle_difference_max=12.0, shear_max=22.0, lenght_ratio_max=0.4) -> FrameResult:
    """executable/main.cpp:172-176 — the three calls back to back, with per-contour verdicts kept."""
    contours, binary = extract_color(image, target, lower_bound)
    verdicts = [classify_contour(c, tilt_max, ratio_range, area_range) for c in contours]
    positive, negative = [], []
    for c, v in zip(contours, verdicts):
        if v.status == STATUS_POSITIVE:
            positive.append(make_lightblob(v.ellipse, target))
        elif v.status == STATUS_NEGATIVE:
            negative.append(c)
    armours = filter_armours(positive, angle_difference_max, shear_max, lenght_ratio_max, target)
    return FrameResult(binary, contours, verdicts, positive, negative, armours)


def blob_label_map(binary: np.ndarray, contours: Sequence[np.ndarray]) -> np.ndarray:
    """Oracle for 'blob pixel sets' (SURVEY §8c): int32 H×W map, value = index of the external
    contour whose component owns the pixel, -1 elsewhere (background and nested components)."""
    _, labels = cv2.connectedComponents(binary, connectivity=8, ltype=cv2.CV_32S)
    lut = np.full(int(labels.max()) + 1, -1, np.int32)
    for k, c in enumerate(contours):
        x, y = int(c[0][0]), int(c[0][1])
        lut[labels[y, x]] = k
    lut[0] = -1
    return lut[labels]


# --------------------------------------------------------------------------- f2 icon crop (next row, first half)
def affine_correction(source: np.ndarray, vertices: np.ndarray, out_size=(20, 20)):
    """rm::affine_correction, src/imgproc.cpp:9-35, through the same OpenCV calls.  Returns (calibration, vertices clamped
    in place like the reference does)."""
    v = np.asarray(vertices, np.float32).copy()
    rows, cols = source.shape[:2]
    v[:, 0] = np.maximum(np.float32(0), np.minimum(v[:, 0], np.float32(cols) - np.float32(1)))   # :11-15
    v[:, 1] = np.maximum(np.float32(0), np.minimum(v[:, 1], np.float32(rows) - np.float32(1)))
    pts = np.rint(v).astype(np.int32)            # std::vector<cv::Point>(Point2f...): cvRound
    x, y, w, h = cv2.boundingRect(pts.reshape(-1, 1, 2))   # :17
    src = np.float32([[v[1, 0] - np.float32(x), v[1, 1] - np.float32(y)], [v[2, 0] - np.float32(x), v[2, 1] - np.float32(y)],
                      [v[0, 0] - np.float32(x), v[0, 1] - np.float32(y)]])
    dst = np.float32([[0, 0], [w, 0], [0, h]])
    warp = cv2.getAffineTransform(src, dst)      # :28
    roi = source[y:y + h, x:x + w]
    calibration = cv2.warpAffine(roi, warp, (w, h))   # :31 (an empty dsize means the source size)
    calibration = cv2.resize(calibration, tuple(out_size))   # :32
    return calibration, v


def flatten_image(image: np.ndarray) -> np.ndarray:
    """rm::utils::flatten_image(input, CV_32FC1), src/core.cpp:202-216 (no resize)."""
    return image.reshape(1, -1).astype(np.float32)


# --------------------------------------------------------------------------- f3 tracking (next row)
class TrackedArmour:
    """The tracking side of rm::armour (include/core.h:103-122): cv::KalmanFilter(6, 6, 0, CV_64F) observer, measurement,
    identity_history, and the fields the tracking loop reads.  Restates src/core.cpp:51-161 through cv2.KalmanFilter."""

    def __init__(self, bounding_box, position, identity, timestamp):
        self.bounding_box = tuple(np.float32(v) for v in bounding_box)
        self.position = np.asarray(position, np.float64).copy()
        self.identity = int(identity)
        self.timestamp = int(timestamp)
        self.lost_count = 0
        self.identity_history = {}
        self.observer = cv2.KalmanFilter(6, 6, 0, cv2.CV_64F)   # src/core.cpp:21
        self.measurement = np.zeros((6, 1), np.float64)
        self.initialized = False

    def reset(self, process_noise, measurement_noise, error):   # src/core.cpp:51-69
        k = self.observer
        k.measurementMatrix = np.eye(6)
        k.processNoiseCov = np.eye(6) * process_noise
        k.measurementNoiseCov = np.eye(6) * measurement_noise
        k.errorCovPost = np.eye(6) * error
        self.measurement = np.zeros((6, 1), np.float64)
        F = np.eye(6)
        F[0, 3] = F[1, 4] = F[2, 5] = 1.0
        k.transitionMatrix = F
        self.initialized = False

    def _set_dt(self, dt):
        F = self.observer.transitionMatrix.copy()
        F[0, 3] = F[1, 4] = F[2, 5] = dt
        self.observer.transitionMatrix = F

    def update_observation(self, obs: "TrackedArmour", tick_frequency):   # src/core.cpp:71-106
        self.identity_history[obs.identity] = self.identity_history.get(obs.identity, 0) + 1
        if self.initialized:
            with np.errstate(divide="ignore", invalid="ignore"):
                dt = np.float64(obs.timestamp - self.timestamp) / np.float64(tick_frequency)
                self._set_dt(dt)
                self.observer.predict()
                self.measurement[3:6, 0] = (obs.position - self.measurement[0:3, 0]) / dt
            self.measurement[0:3, 0] = obs.position
            self.observer.correct(self.measurement)
        else:
            self.measurement[0:3, 0] = obs.position
            self.observer.correct(self.measurement)
            self.initialized = True
        self.timestamp = obs.timestamp

    def update_time(self, new_timestamp, tick_frequency):   # src/core.cpp:108-121
        if not self.initialized:
            return
        self._set_dt(np.float64(new_timestamp - self.timestamp) / np.float64(tick_frequency))
        self.observer.predict()

    def identity_max(self):   # src/core.cpp:123-143
        total = sum(np.exp(np.float64(c)) for c in self.identity_history.values())
        best, best_id = 0.0, -1
        for ident in sorted(self.identity_history):   # std::map iterates in key order
            prob = np.exp(np.float64(self.identity_history[ident])) / total
            if prob > best:
                best, best_id = prob, ident
        return best_id, best

    def max_iou(self, armours):   # src/core.cpp:145-161
        index, best = -1, np.float32(0)
        for i, a in enumerate(armours):
            v = rect_iou(self.bounding_box, a.bounding_box)
            if v > best:
                best, index = v, i
        return index, best


def rect_iou(a, b) -> np.float32:
    """intersection.area() / (a.area() + b.area() - intersection.area()) on cv::Rect2f (src/core.cpp:151-154).  The
    intersection restates cv::Rect_::operator&= of OpenCV 4.5+ (modules/core/include/opencv2/core/types.hpp); cv2 has no
    binding for it, so this part of the oracle is not pinned by cv2 itself."""
    f = np.float32
    a = [f(v) for v in a]; b = [f(v) for v in b]
    iw = ih = f(0)
    if not (a[2] <= 0 or a[3] <= 0 or b[2] <= 0 or b[3] <= 0):
        xmin, xmax = (a, b) if a[0] < b[0] else (b, a)
        ymin, ymax = (a, b) if a[1] < b[1] else (b, a)
        apart = (xmin[0] < 0 and f(xmin[0] + xmin[2]) < xmax[0]) or (ymin[1] < 0 and f(ymin[1] + ymin[3]) < ymax[1])
        if not apart:
            iw = min(f(xmin[2] - f(xmax[0] - xmin[0])), xmax[2])
            ih = min(f(ymin[3] - f(ymax[1] - ymin[1])), ymax[3])
            if iw <= 0 or ih <= 0:
                iw = ih = f(0)
    inter = f(iw * ih)
    with np.errstate(divide="ignore", invalid="ignore"):
        return f(inter / f(f(f(a[2] * a[3]) + f(b[2] * b[3])) - inter))


def tracking_step(tracking: list, armours: list, tick_frequency, noise=(5e-5, 0.5, 0.05)) -> list:
    """One iteration of the tracking thread's loop, executable/main.cpp:60-85, including its erase-inside-a-for-loop
    (the track behind an erased one is skipped).  `armours`: TrackedArmour observations of one frame; they are reset() like
    process_function does (main.cpp:195) when they enter the list."""
    armours = list(armours)
    if not armours:
        return tracking
    for a in armours:
        a.reset(*noise)
    if not tracking:
        return armours
    i = 0
    while i < len(tracking):
        index, iou = tracking[i].max_iou(armours)
        if iou > 0.5:
            tracking[i].update_observation(armours[index], tick_frequency)
            del armours[index]
        else:
            lost = tracking[i].lost_count
            tracking[i].lost_count += 1
            if lost > 25:
                del tracking[i]
            else:
                tracking[i].update_time(tracking[i].timestamp, tick_frequency)
        i += 1
    tracking.extend(armours)
    return tracking
