"""Seeded synthetic frame generator.

Produces the *inputs* every leg consumes (GPU path, oracle, bench): BGR frames
with red/blue light-bar armour plates plus a fixed set of coverage extras, and
the Bayer mosaics sampled from them.  Only plain numpy + cv2 rasterisers are
used so the CPU oracle and the GPU path see identical bytes.  Nothing here is
on the product data path; it only manufactures test/bench inputs.

Parameters of the detection path itself are the literals of the reference's
only caller, executable/main.cpp:172-176.
"""
from __future__ import annotations

import math

import numpy as np

try:  # cv2 is only needed to rasterise ellipses/lines for synthetic inputs
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

CAMP_RED, CAMP_BLUE, CAMP_GUIDELIGHT, CAMP_NEUTRAL = 0, 1, 2, -1  # include/core.h:20-23

# executable/main.cpp:172-176
MAIN_PARAMS = dict(
    target=CAMP_BLUE,
    lower_bound=80,
    tilt_max=70.0,
    ratio_range=(1.5, 80.0),
    area_range=(10.0, 99999.0),
    angle_difference_max=12.0,
    shear_max=22.0,
    lenght_ratio_max=0.4,
)


def _need_cv2():
    if cv2 is None:
        raise RuntimeError("cv2 is required to rasterise synthetic frames")


def make_frame(seed: int, width: int = 1280, height: int = 1024, plates: int = 8,
               blue: bool = True, extras: bool = True) -> np.ndarray:
    """One H×W×3 uint8 BGR frame (Appendix D recipe)."""
    _need_cv2()
    W, H, P = int(width), int(height), int(plates)
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 40, (H, W, 3), dtype=np.uint8)
    col = (255, 140, 40) if blue else (40, 140, 255)
    if P > 0:
        cols = int(math.ceil(math.sqrt(P * W / H)))
        rows = int(math.ceil(P / cols))
        cw, ch = W / cols, H / rows
        done = 0
        for r in range(rows):
            for c in range(cols):
                if done >= P:
                    break
                done += 1
                L = rng.uniform(0.18, 0.32) * min(cw, ch)
                Wd = max(5.0, L / 5)
                gap = L * rng.uniform(1.6, 2.2)
                tilt = rng.uniform(-12, 12)
                cx = (c + 0.5) * cw + rng.uniform(-0.05, 0.05) * cw
                cy = (r + 0.5) * ch + rng.uniform(-0.05, 0.05) * ch
                for s in (-1, +1):
                    ctr = (cx + s * gap / 2, cy + s * rng.uniform(-2, 2))
                    a = tilt + rng.uniform(-2, 2)
                    cv2.ellipse(img, (ctr, (Wd, L), a), col, -1)
                    d = (math.sin(math.radians(-a)) * (L / 2 - 3),
                         math.cos(math.radians(-a)) * (L / 2 - 3))
                    p0 = (int(ctr[0] - d[0]), int(ctr[1] - d[1]))
                    p1 = (int(ctr[0] + d[0]), int(ctr[1] + d[1]))
                    cv2.line(img, p0, p1, (255, 255, 255), 1)
    if extras:
        _draw_extras(img, rng, col)
    return img


def _draw_extras(img, rng, col):
    """Fixed per-frame coverage extras (specks, disc, flat bar, annulus+dot,
    border-touching blob, diagonal chain)."""
    H, W = img.shape[:2]
    # 6 specks of 1-2 x 1-3 px (rejected by size()<6 / area)
    for k in range(6):
        x = int(W * (0.30 + 0.07 * k)) % max(1, W - 4)
        y = int(H * 0.015) + 2
        w = 1 + (k & 1)
        h = 1 + (k % 3)
        img[y:y + h, x:x + w] = col
    # round disc (ratio < 1.5 => negative)
    cv2.circle(img, (int(0.07 * W), int(0.07 * H)), 9, col, -1)
    # near-horizontal bar (tilt > 70 => negative)
    cv2.ellipse(img, ((0.93 * W, 0.06 * H), (8, 60), 85), col, -1)
    # annulus with a dot inside (nesting: dot must not be reported)
    c = (int(0.06 * W), int(0.93 * H))
    cv2.circle(img, c, 16, col, 4)
    cv2.circle(img, c, 3, col, -1)
    # blob touching the left border
    cv2.ellipse(img, ((3, 0.5 * H), (8, 50), 0), col, -1)
    # 12-pixel diagonal chain (8- vs 4-connectivity)
    x0, y0 = int(0.9 * W), int(0.9 * H)
    for k in range(12):
        if 0 <= y0 + k < H and 0 <= x0 + k < W:
            img[y0 + k, x0 + k] = col


def make_stress_frame(seed: int, width: int = 4096, height: int = 3072, plates: int = 250) -> np.ndarray:
    """Config 4: 250 plates on a jittered grid -> ~500 light blobs."""
    return make_frame(seed, width, height, plates, blue=True, extras=True)


def plates_for_seed(seed: int) -> int:
    """Config 3: plates per frame uniform in [4, 20], derived from the seed only."""
    return int(np.random.default_rng(10_000_019 + seed).integers(4, 21))


def make_batch(seeds, width=1280, height=1024, plates=None, alternate_camp=False) -> np.ndarray:
    """Stack frames into one B×H×W×3 array.  plates=None -> plates_for_seed."""
    out = np.empty((len(seeds), height, width, 3), np.uint8)
    for i, s in enumerate(seeds):
        p = plates_for_seed(s) if plates is None else plates
        blue = True if not alternate_camp else (s % 2 == 0)
        out[i] = make_frame(s, width, height, p, blue=blue)
    return out


# Bayer layouts follow Daheng's DX_PIXEL_COLOR_FILTER (hardware/include/daheng/DxImageProc.h:54-61):
# the name gives the colours of the first two pixels of row 0.
BAYER_RG, BAYER_GB, BAYER_GR, BAYER_BG = 1, 2, 3, 4


def bayer_channel_index(layout: int):
    """Return a 2×2 array: channel index (0=B,1=G,2=R) sampled at (y&1, x&1)."""
    table = {
        BAYER_BG: [[0, 1], [1, 2]],
        BAYER_GB: [[1, 0], [2, 1]],
        BAYER_GR: [[1, 2], [0, 1]],
        BAYER_RG: [[2, 1], [1, 0]],
    }
    return np.array(table[layout], np.int64)


def bgr_to_bayer(img: np.ndarray, layout: int = BAYER_BG) -> np.ndarray:
    """Sample a BGR frame into an 8-bit mosaic (Appendix D 'Bayer variant')."""
    H, W = img.shape[:2]
    ch = bayer_channel_index(layout)
    yy, xx = np.mgrid[0:H, 0:W]
    idx = ch[yy & 1, xx & 1]
    return np.take_along_axis(img, idx[..., None], axis=2)[..., 0].copy()


def shape_mask(rng, W: int, H: int) -> np.ndarray:
    """Boolean mask of drawn outlines — rings in rings, C-shaped arcs holding other shapes in their concavity, box
    outlines, spirals, combs, dots, a little noise; some strokes erase and cut earlier shapes open.  Exercises holes,
    nesting and concavities of the external-contour discovery (the reference's contour call at src/imgproc.cpp:72) far more densely than light bars do."""
    import cv2
    img = np.zeros((H, W), np.uint8)
    for _ in range(int(rng.integers(3, 60))):
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        r = int(rng.integers(3, max(4, min(W, H) // 3)))
        th = int(rng.integers(1, 6))
        what = int(rng.integers(0, 6))
        col = 255 if rng.random() < 0.8 else 0
        if what == 0:
            cv2.circle(img, (cx, cy), r, col, th)
        elif what == 1:
            a0 = float(rng.uniform(0, 360))
            cv2.ellipse(img, (cx, cy), (r, max(2, int(r * rng.uniform(0.3, 1.0)))), float(rng.uniform(0, 180)), a0,
                        a0 + float(rng.uniform(180, 340)), col, th)
        elif what == 2:
            cv2.rectangle(img, (cx - r, cy - r // 2), (cx + r, cy + r // 2), col, th)
        elif what == 3:
            t = np.linspace(0, float(rng.uniform(2, 8)) * np.pi, 400)
            pts = np.stack([cx + r * t / t[-1] * np.cos(t), cy + r * t / t[-1] * np.sin(t)], 1).astype(np.int32)
            cv2.polylines(img, [pts.reshape(-1, 1, 2)], False, col, th)
        elif what == 4:
            step = int(rng.integers(3, 9))
            cv2.line(img, (cx - r, cy), (cx + r, cy), col, th)
            for x in range(cx - r, cx + r, step):
                cv2.line(img, (x, cy), (x, cy + int(rng.integers(-r, r + 1))), col, 1)
        else:
            cv2.circle(img, (cx, cy), int(rng.integers(1, 8)), col, -1)
    return (img > 0) | (rng.random((H, W)) < float(rng.uniform(0.0, 0.01)))
