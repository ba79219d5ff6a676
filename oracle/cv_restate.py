"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

numpy restatements of the OpenCV internals that rmcv's hot path leans on, in the
*order-free* form the CUDA kernels use.  They are the
executable specification of the kernels and are themselves pinned against cv2
in tests/test_oracle.py:

  close3x3_bits        morphologyEx(MORPH_CLOSE, 3x3 rect)           (A.1)   src/imgproc.cpp:68-69
  ARC_LUT / contour_stats   findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) as a local
                       3x3 arc rule: point multiset, contour.size(), contourArea (A.2-A.4)   src/imgproc.cpp:72
  fit_ellipse_direct   cv::fitEllipseDirect incl. its fallback to fitEllipseNoDirect (A.6)   src/objdetect.cpp:68
  box_points           cv::RotatedRect::points (A.9)                 src/core.cpp:268
  bounding_rect_f      cv::boundingRect(vector<Point2f>) (A.10)      src/core.cpp:46
  bayer_bilinear_bgr   cv2.cvtColor(COLOR_Bayer*2BGR) integer rule (A.7) — stand-in for DxRaw8toRGB24
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

f32 = np.float32

# direction codes 0..7 = E, NE, N, NW, W, SW, S, SE (image coordinates, y down)
DIRS = ((1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1))


# --------------------------------------------------------------------------- morphology on bit rows
def close3x3(t: np.ndarray) -> np.ndarray:
    """3x3 close on a bool H×W array: dilate with 0 outside, then erode with 1 outside (A.1)."""
    H, W = t.shape
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = t
    d = np.zeros((H, W), bool)
    for dy in range(3):
        for dx in range(3):
            d |= p[dy:dy + H, dx:dx + W]
    q = np.ones((H + 2, W + 2), bool)
    q[1:-1, 1:-1] = d
    e = np.ones((H, W), bool)
    for dy in range(3):
        for dx in range(3):
            e &= q[dy:dy + H, dx:dx + W]
    return e


def threshold_bits(image: np.ndarray, a: int, b: int, lower_bound: int) -> np.ndarray:
    """inRange(sat_u8(ch[a]-ch[b]), lb, 255) as a bool array."""
    d = image[..., a].astype(np.int32) - image[..., b].astype(np.int32)
    d = np.clip(d, 0, 255)
    return (d >= int(lower_bound)) & (d <= 255)


# --------------------------------------------------------------------------- arc LUT (A.3)
def build_arc_lut() -> np.ndarray:
    """LUT[code] for the 8-neighbour foreground code (bit k = neighbour in direction k is foreground).

    Entry layout (uint32): bits 0-2 = number of arcs m (0..4); arc i (i<m) occupies 5 bits at 3+5*i:
    low 2 bits = index (dir/2) of a 4-neighbour inside the arc (the background pixel whose
    outer/hole status decides whether the arc is emitted), next 3 bits = direction of the edge
    target q (foreground neighbour at the counter-clockwise end of the arc).
    Bit 31 set = isolated pixel (one point, no edge; outer test uses any 4-neighbour, index 0).
    """
    lut = np.zeros(256, np.uint32)
    for code in range(256):
        fg = [(code >> k) & 1 for k in range(8)]
        if code == 0:
            lut[code] = (1 << 31) | 1  # one arc, test dir E (idx 0), no edge
            continue
        arcs = []
        # walk the cyclic sequence starting right after some foreground neighbour
        start = next(k for k in range(8) if fg[k])
        k = (start + 1) % 8
        steps = 0
        while steps < 8:
            if not fg[k]:
                run = []
                while not fg[k]:
                    run.append(k)
                    k = (k + 1) % 8
                    steps += 1
                four = [d for d in run if d % 2 == 0]
                if four:
                    arcs.append((four[0] // 2, k))  # k is now the first fg after the run
            else:
                k = (k + 1) % 8
                steps += 1
        v = len(arcs)
        for i, (t4, q) in enumerate(arcs):
            v |= (t4 | (q << 2)) << (3 + 5 * i)
        lut[code] = v
    return lut


ARC_LUT = build_arc_lut()


def neighbour_codes(fg: np.ndarray) -> np.ndarray:
    """uint8 code per pixel: bit k set iff the neighbour in direction k is foreground (outside = 0)."""
    H, W = fg.shape
    p = np.zeros((H + 2, W + 2), np.uint8)
    p[1:-1, 1:-1] = fg
    code = np.zeros((H, W), np.uint8)
    for k, (dx, dy) in enumerate(DIRS):
        code |= (p[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] << k).astype(np.uint8)
    return code


def label8(fg: np.ndarray) -> Tuple[np.ndarray, int]:
    from scipy import ndimage
    return ndimage.label(fg, structure=np.ones((3, 3), int))


def outer_background(fg: np.ndarray) -> np.ndarray:
    """bool map of the background 4-connected to a virtual frame around the image (A.4)."""
    from scipy import ndimage
    H, W = fg.shape
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = fg
    lab, _ = ndimage.label(~p)  # 4-connectivity default
    return (lab == lab[0, 0])[1:-1, 1:-1]


def contour_points(fg: np.ndarray):
    """Order-free contour emission.  Returns (labels, px, py, qdx, qdy, lab_of_point) where each entry is
    one contour point (with multiplicity) of an external component and (qdx,qdy) its edge step (0,0 = none)."""
    fg = fg.astype(bool)
    H, W = fg.shape
    labels, _ = label8(fg)
    outer = outer_background(fg)
    po = np.ones((H + 2, W + 2), bool)  # outside the image counts as outer background
    po[1:-1, 1:-1] = outer
    code = neighbour_codes(fg)
    ys, xs = np.nonzero(fg & (code != 255))
    ent = ARC_LUT[code[ys, xs]]
    PX, PY, QX, QY, LB = [], [], [], [], []
    for i in range(4):
        has = (ent & 7) > i
        if not has.any():
            break
        a = (ent >> (3 + 5 * i)) & 31
        t4 = (a & 3) * 2
        q = (a >> 2) & 7
        iso = (ent >> 31) & 1
        dirs = np.array(DIRS)
        tx = xs + dirs[t4, 0]
        ty = ys + dirs[t4, 1]
        ok = has & po[ty + 1, tx + 1]
        qd = dirs[q] * (1 - iso)[:, None]
        PX.append(xs[ok]); PY.append(ys[ok]); QX.append(qd[ok, 0]); QY.append(qd[ok, 1]); LB.append(labels[ys[ok], xs[ok]])
    if not PX:
        z = np.zeros(0, np.int64)
        return labels, z, z, z, z, z
    return labels, np.concatenate(PX), np.concatenate(PY), np.concatenate(QX), np.concatenate(QY), np.concatenate(LB)


def contour_stats(fg: np.ndarray) -> List[Dict]:
    """Per external component, in cv2 contour order (reverse raster order of the first pixel):
    n (contour.size()), area2 (= 2*contourArea, exact integer), first pixel, bbox, points multiset."""
    labels, px, py, qx, qy, lb = contour_points(fg)
    out = []
    H, W = fg.shape
    for L in np.unique(lb):
        m = lb == L
        x, y, dx, dy = px[m], py[m], qx[m], qy[m]
        cross = int(np.sum(x * (y + dy) - (x + dx) * y))
        ys_, xs_ = np.nonzero(labels == L)
        first = int(np.min(ys_ * W + xs_))
        out.append(dict(label=int(L), n=int(m.sum()), area2=abs(cross), first=(first % W, first // W),
                        bbox=(int(xs_.min()), int(ys_.min()), int(xs_.max() - xs_.min() + 1), int(ys_.max() - ys_.min() + 1)),
                        points=np.stack([x, y], 1)))
    out.sort(key=lambda d: -(d["first"][1] * W + d["first"][0]))
    return out


# --------------------------------------------------------------------------- fitEllipse (A.6)
def _moments(dx: np.ndarray, dy: np.ndarray) -> Dict[str, float]:
    """The 14 sums the fits need (plus n), from centred, *unscaled* double coordinates."""
    return dict(
        n=float(len(dx)),
        x=dx.sum(), y=dy.sum(),
        xx=(dx * dx).sum(), xy=(dx * dy).sum(), yy=(dy * dy).sum(),
        xxx=(dx ** 3).sum(), xxy=(dx * dx * dy).sum(), xyy=(dx * dy * dy).sum(), yyy=(dy ** 3).sum(),
        xxxx=(dx ** 4).sum(), xxxy=(dx ** 3 * dy).sum(), xxyy=(dx * dx * dy * dy).sum(), xyyy=(dx * dy ** 3).sum(), yyyy=(dy ** 4).sum(),
    )


def _scaled(m: Dict[str, float], scale: float) -> Dict[str, float]:
    out = {}
    for k, v in m.items():
        out[k] = v if k == "n" else v * scale ** len(k)
    return out


def direct_fit_from_moments(m: Dict[str, float], scale: float, cx: float, cy: float):
    """Halir-Flusser direct fit as cv::fitEllipseDirect evaluates it, from the scaled moment sums.
    Returns (det, box or None)."""
    n = m["n"]
    s = _scaled(m, scale)
    # DM = A^T A / n with A rows [x^2, xy, y^2, x, y, 1]
    DM = np.array([
        [s["xxxx"], s["xxxy"], s["xxyy"], s["xxx"], s["xxy"], s["xx"]],
        [s["xxxy"], s["xxyy"], s["xyyy"], s["xxy"], s["xyy"], s["xy"]],
        [s["xxyy"], s["xyyy"], s["yyyy"], s["xyy"], s["yyy"], s["yy"]],
        [s["xxx"], s["xxy"], s["xyy"], s["xx"], s["xy"], s["x"]],
        [s["xxy"], s["xyy"], s["yyy"], s["xy"], s["yy"], s["y"]],
        [s["xx"], s["xy"], s["yy"], s["x"], s["y"], n],
    ]) / n
    S1 = DM[0:3, 0:3]
    S2 = DM[0:3, 3:6]
    S3 = DM[3:6, 3:6]
    # TM = -adj(S3) * S2^T  (so that T = TM/Ts), Ts = det(S3)
    Ts = np.linalg.det(S3)
    adj = np.linalg.inv(S3) * Ts
    TM = -adj @ S2.T
    Mp = S1 + (S2 @ TM) / Ts
    M = np.array([Mp[2] / 2.0, -Mp[1], Mp[0] / 2.0])
    det = abs(np.linalg.det(M))
    if not (det > 1.0e-10):
        return det, None
    w, v = np.linalg.eig(M)
    v = np.real(v)
    cond = 4.0 * v[0, :] * v[2, :] - v[1, :] ** 2
    i = int(np.argmax(cond))
    pv = v[:, i].copy()
    norm = math.sqrt(float(pv @ pv))
    sg = 1
    for c in pv:
        sg *= (-1 if c < 0.0 else 1)
    if sg <= 0:
        norm = -norm
    pv = pv / norm
    Q = (TM @ pv) / Ts
    a_, b_, c_ = pv
    u1 = c_ * Q[0] * Q[0] - b_ * Q[0] * Q[1] + a_ * Q[1] * Q[1] + b_ * b_ * Q[2]
    u2 = a_ * c_ * Q[2]
    l1 = math.sqrt(b_ * b_ + (a_ - c_) * (a_ - c_))
    l2 = a_ + c_
    l3 = b_ * b_ - 4.0 * a_ * c_
    p1 = 2.0 * c_ * Q[0] - b_ * Q[1]
    p2 = 2.0 * a_ * Q[1] - b_ * Q[0]
    x0 = p1 / l3 / scale + cx
    y0 = p2 / l3 / scale + cy
    with np.errstate(invalid="ignore", divide="ignore"):
        A = math.sqrt(2.0) * np.sqrt((u1 - 4.0 * u2) / ((l1 - l2) * l3)) / scale
        B = math.sqrt(2.0) * np.sqrt(-1.0 * ((u1 - 4.0 * u2) / ((l1 + l2) * l3))) / scale
    if b_ == 0:
        theta = 0.0 if a_ < c_ else math.pi / 2.0
    else:
        theta = math.pi / 2.0 + 0.5 * math.atan2(b_, (a_ - c_))
    wd, ht = f32(2.0 * A), f32(2.0 * B)
    if wd > ht:
        wd, ht = ht, wd
        ang = f32(math.fmod(90.0 + theta * 180.0 / math.pi, 180.0))
    else:
        ang = f32(math.fmod(theta * 180.0 / math.pi, 180.0))
    return det, (float(f32(x0)), float(f32(y0)), float(wd), float(ht), float(ang))


def nodirect_fit_from_moments(m: Dict[str, float], scale: float, c32x: float, c32y: float):
    """cv::fitEllipseNoDirect through the normal equations, from moment sums of the float-centred points."""
    s = _scaled(m, scale)
    n = m["n"]
    # rows a = [-x^2, -y^2, -xy, x, y], rhs 10000
    AtA = np.array([
        [s["xxxx"], s["xxyy"], s["xxxy"], -s["xxx"], -s["xxy"]],
        [s["xxyy"], s["yyyy"], s["xyyy"], -s["xyy"], -s["yyy"]],
        [s["xxxy"], s["xyyy"], s["xxyy"], -s["xxy"], -s["xyy"]],
        [-s["xxx"], -s["xyy"], -s["xxy"], s["xx"], s["xy"]],
        [-s["xxy"], -s["yyy"], -s["xyy"], s["xy"], s["yy"]],
    ])
    Atb = 10000.0 * np.array([-s["xx"], -s["yy"], -s["xy"], s["x"], s["y"]])
    g = np.linalg.solve(AtA, Atb)
    r = np.linalg.solve(np.array([[2 * g[0], g[2]], [g[2], 2 * g[1]]]), np.array([g[3], g[4]]))
    rx, ry = r
    # refit rows [(x-rx)^2, (y-ry)^2, (x-rx)(y-ry)], rhs 1 : shifted moments by binomial expansion
    def sh(i, j):  # sum (x-rx)^i (y-ry)^j
        tot = 0.0
        for a in range(i + 1):
            for b in range(j + 1):
                key = "x" * a + "y" * b
                mv = n if key == "" else s[key]
                tot += math.comb(i, a) * math.comb(j, b) * (-rx) ** (i - a) * (-ry) ** (j - b) * mv
        return tot
    BtB = np.array([
        [sh(4, 0), sh(2, 2), sh(3, 1)],
        [sh(2, 2), sh(0, 4), sh(1, 3)],
        [sh(3, 1), sh(1, 3), sh(2, 2)],
    ])
    Btb = np.array([sh(2, 0), sh(0, 2), sh(1, 1)])
    g2 = np.linalg.solve(BtB, Btb)
    min_eps = 1e-8
    rp4 = -0.5 * math.atan2(g2[2], g2[1] - g2[0])
    if abs(g2[2]) > min_eps:
        t = g2[2] / math.sin(-2.0 * rp4)
    else:
        t = g2[1] - g2[0]
    rp2 = abs(g2[0] + g2[1] - t)
    if rp2 > min_eps:
        rp2 = math.sqrt(2.0 / rp2)
    rp3 = abs(g2[0] + g2[1] + t)
    if rp3 > min_eps:
        rp3 = math.sqrt(2.0 / rp3)
    bx = f32(f32(rx / scale) + f32(c32x))
    by = f32(f32(ry / scale) + f32(c32y))
    wd = f32(rp2 * 2 / scale)
    ht = f32(rp3 * 2 / scale)
    ang = f32(0.0)   # cv::fitEllipseNoDirect assigns box.angle only in the swap below: an unswapped box keeps RotatedRect's 0
    if wd > ht:
        wd, ht = ht, wd
        ang = f32(90 + rp4 * 180 / math.pi)
    if ang < -180:
        ang = f32(ang + f32(360))
    if ang > 360:
        ang = f32(ang - f32(360))
    return (float(bx), float(by), float(wd), float(ht), float(ang))


def fit_ellipse_direct(points: np.ndarray):
    """Order-free restatement of cv::fitEllipseDirect on an integer point multiset (N×2).
    Returns dict(box=(cx,cy,w,h,angle), det0, branch='direct'|'fallback')."""
    pts = np.asarray(points, np.float64)
    n = len(pts)
    cx, cy = pts[:, 0].sum() / n, pts[:, 1].sum() / n
    dx, dy = pts[:, 0] - cx, pts[:, 1] - cy
    s = float((np.abs(dx) + np.abs(dy)).sum())
    scale = 100.0 / (s if s > np.finfo(np.float32).eps else float(np.finfo(np.float32).eps))
    det, box = direct_fit_from_moments(_moments(dx, dy), scale, cx, cy)
    if box is not None:
        return dict(box=box, det0=det, branch="direct")
    # fallback: float centre, float subtraction (cv::fitEllipseNoDirect)
    p32 = np.asarray(points, np.float32)
    c32x = f32(f32(p32[:, 0].sum(dtype=np.float64)) / f32(n))  # exact while the running sum < 2^24
    c32y = f32(f32(p32[:, 1].sum(dtype=np.float64)) / f32(n))
    fx = (p32[:, 0] - c32x).astype(np.float32)
    fy = (p32[:, 1] - c32y).astype(np.float32)
    s2 = float((np.abs(fx) + np.abs(fy)).astype(np.float32).astype(np.float64).sum())
    scale2 = 100.0 / (s2 if s2 > np.finfo(np.float32).eps else float(np.finfo(np.float32).eps))
    box = nodirect_fit_from_moments(_moments(fx.astype(np.float64), fy.astype(np.float64)), scale2, float(c32x), float(c32y))
    return dict(box=box, det0=det, branch="fallback")


# --------------------------------------------------------------------------- small geometry helpers
def box_points(cx, cy, w, h, angle) -> np.ndarray:
    """cv::RotatedRect::points (A.9), float arithmetic."""
    ang = float(f32(angle)) * math.pi / 180.0
    b = f32(f32(math.cos(ang)) * f32(0.5))
    a = f32(f32(math.sin(ang)) * f32(0.5))
    cx, cy, w, h = f32(cx), f32(cy), f32(w), f32(h)
    p = np.empty((4, 2), np.float32)
    p[0, 0] = cx - a * h - b * w
    p[0, 1] = cy + b * h - a * w
    p[1, 0] = cx + a * h - b * w
    p[1, 1] = cy - b * h - a * w
    p[2, 0] = f32(2) * cx - p[0, 0]
    p[2, 1] = f32(2) * cy - p[0, 1]
    p[3, 0] = f32(2) * cx - p[1, 0]
    p[3, 1] = f32(2) * cy - p[1, 1]
    return p


def bounding_rect_f(pts: np.ndarray):
    """cv::boundingRect on float points (A.10)."""
    x0 = math.floor(float(pts[:, 0].min())); y0 = math.floor(float(pts[:, 1].min()))
    x1 = math.floor(float(pts[:, 0].max())); y1 = math.floor(float(pts[:, 1].max()))
    return (x0, y0, x1 - x0 + 1, y1 - y0 + 1)


# --------------------------------------------------------------------------- Bayer (A.7)
def bayer_bilinear_bgr(raw: np.ndarray, layout: int = 4) -> np.ndarray:
    """cv2.cvtColor(raw, COLOR_Bayer*2BGR) integer rule.  layout: Daheng code, 4 = BAYERBG (row 0 = B G B G)."""
    ch2x2 = {4: [[0, 1], [1, 2]], 2: [[1, 0], [2, 1]], 3: [[1, 2], [0, 1]], 1: [[2, 1], [1, 0]]}[layout]
    H, W = raw.shape
    r = raw.astype(np.int32)
    p = np.pad(r, 1, mode="edge")
    c = p[1:-1, 1:-1]
    lr = (p[1:-1, :-2] + p[1:-1, 2:] + 1) >> 1
    ud = (p[:-2, 1:-1] + p[2:, 1:-1] + 1) >> 1
    plus = (p[1:-1, :-2] + p[1:-1, 2:] + p[:-2, 1:-1] + p[2:, 1:-1] + 2) >> 2
    diag = (p[:-2, :-2] + p[:-2, 2:] + p[2:, :-2] + p[2:, 2:] + 2) >> 2
    out = np.zeros((H, W, 3), np.int32)
    yy, xx = np.mgrid[0:H, 0:W]
    site = np.array(ch2x2)[yy & 1, xx & 1]  # colour sampled at this site
    rowcol = np.array(ch2x2)[yy & 1, (xx & 1) ^ 1]  # colour of the horizontal neighbours
    for colour in (0, 1, 2):
        here = site == colour
        v = np.where(here, c, 0)
        if colour == 1:
            v = np.where(~here, plus, v)
        else:
            at_g = site == 1
            horiz = at_g & (rowcol == colour)
            vert = at_g & (rowcol != colour)
            opp = (~here) & (~at_g)
            v = np.where(horiz, lr, v)
            v = np.where(vert, ud, v)
            v = np.where(opp, diag, v)
        out[..., colour] = v
    out = out.astype(np.uint8)
    if H > 2:
        out[0] = out[1]
        out[-1] = out[-2]
    if W > 2:
        out[:, 0] = out[:, 1]
        out[:, -1] = out[:, -2]
    return out
