"""Generates tests/golden/*.json from the REFERENCE ITSELF run here: oracle/_ref/librmcv_ref.so is /root/reference's own
src/core.cpp, src/objdetect.cpp, src/imgproc.cpp and src/mobility.cpp compiled unmodified (oracle/Makefile), with every cv::
call served by this image's OpenCV 4.13.0 (oracle/ref_bridge.py).  Every record says `"source": "_ref"`.

The reference ships no golden vectors; these fixtures are outputs of its compiled code on seeded synthetic
frames.  /root/reference does not exist on the GPU box, so they are what pins the oracle (tests/test_oracle.py, exact) and
the CUDA path (tests/test_gpu_golden.py) there.  Two things in a record are not the reference's code: the per-contour
`ellipse` (cv2.fitEllipseDirect called directly — the reference never exposes it) and the loop around the tracking
methods (a lambda inside main(), executable/main.cpp:60-85, restated in oracle/ref_bridge.py::tracking_step).
Run from the repo root where /root/reference exists:
    make -C oracle && python scripts/make_golden.py
"""
import json
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_bridge as RB  # noqa: E402
from oracle import rm_oracle as O  # noqa: E402  (constants and the cv2-only helpers)
from rmcv_b200 import synth  # noqa: E402

CASES = [
    dict(name="config1_seed1_8plates_blue", seed=1, width=1280, height=1024, plates=8, blue=True, target=1),
    dict(name="seed7_14plates_red", seed=7, width=1280, height=1024, plates=14, blue=False, target=0),
    dict(name="seed3_640x480_5plates_blue", seed=3, width=640, height=480, plates=5, blue=True, target=1),
]


def _f(a):
    """float32 values as exact Python floats."""
    return np.asarray(a, np.float32).astype(np.float64).tolist()


def record(case):
    ref = RB.get()
    img = synth.make_frame(case["seed"], case["width"], case["height"], case["plates"], blue=case["blue"])
    p = dict(synth.MAIN_PARAMS)
    p["target"] = case["target"]
    fr = RB.detect_frame(img, ref=ref, **p)
    rec = dict(source="_ref", case=case, camera=dict(matrix=O.MAIN_CAMMAT.tolist(), dist=O.MAIN_DISCOF.tolist(), exact_size=[27.0, 27.0]),
               frame_crc32=zlib.crc32(img.tobytes()), mask_crc32=zlib.crc32(fr.binary.tobytes()),
               mask_foreground=int((fr.binary > 0).sum()), contours=[], positive=[], armours=[])
    for c, status in zip(fr.contours, fr.status):
        e = O.fit_ellipse_direct(c) if status != 0 else None      # cv2 directly (see the module docstring)
        area = float(cv2.contourArea(c.reshape(-1, 1, 2)))
        rec["contours"].append(dict(first=[int(c[0][0]), int(c[0][1])], n=int(len(c)), area2=int(round(2 * area)),
                                    bbox=[int(c[:, 0].min()), int(c[:, 1].min()), int(np.ptp(c[:, 0]) + 1), int(np.ptp(c[:, 1]) + 1)],
                                    points_crc32=zlib.crc32(np.ascontiguousarray(c, np.int32).tobytes()), status=int(status),
                                    ellipse=None if e is None else _f([e.cx, e.cy, e.w, e.h, e.angle])))
    for b in fr.positive:
        angle, target, center, vertices, size = RB.blob_arrays(b)
        rec["positive"].append(dict(angle=float(angle), target=target, center=_f(center), size=_f(size), vertices=_f(vertices)))
    poses = []
    for a, (i, j) in zip(fr.armours, fr.pairs):
        icon, vertices, bbox = RB.armour_arrays(a)
        rvec, tvec = ref.solve_pnp(vertices, O.MAIN_CAMMAT, O.MAIN_DISCOF)          # next row f1: rm::solve_PnP
        icon20, _ = ref.affine_correction(img, icon)                               # next row f2: rm::affine_correction
        gates = O.pair_gates(_as_oracle_blob(fr.positive[i]), _as_oracle_blob(fr.positive[j]))   # diagnostics (oracle)
        poses.append(tvec)
        rec["armours"].append(dict(i=i, j=j, icon=_f(icon), vertices=_f(vertices), bounding_box=list(bbox), gates=list(gates),
                                   rvec=rvec.tolist(), tvec=tvec.tolist(), icon20_crc32=zlib.crc32(np.ascontiguousarray(icon20).tobytes())))
    # next row f3: the frame's armours drifting by (2, 1) px per frame through the tracking loop, 6 frames at 125 Hz
    tracking = []
    for n in range(6):
        obs = []
        for k, a in enumerate(fr.armours):
            bb = RB.armour_arrays(a)[2]
            obs.append(RB.RefTrackedArmour(ref, (bb[0] + 2 * n, bb[1] + n, bb[2], bb[3]), poses[k] + n, k % 7, 1000 + 8_000_000 * n))
        tracking = RB.tracking_step(tracking, obs)
    rec["tracking"] = []
    for t in tracking:
        state, cov, _ = t.state()
        rec["tracking"].append(dict(lost_count=t.lost_count, timestamp=t.timestamp, identity_max=list(t.identity_max()),
                                    state_post=state.tolist(), cov_post_diag=np.diag(cov).tolist()))
    # Bayer stand-in (config 2 front; the Daheng SDK is closed, so this part is cv2 only): mosaic -> bilinear -> pixel stage
    raw = synth.bgr_to_bayer(img, synth.BAYER_BG)
    _, bmask = ref.extract_color(O.bayer_to_bgr(raw, 4), case["target"], 80)
    rec["bayer_bg_mask_crc32"] = zlib.crc32(bmask.tobytes())
    return rec


def _as_oracle_blob(b):
    angle, target, center, vertices, size = RB.blob_arrays(b)
    return O.LightBlob(float(angle), target, (float(center[0]), float(center[1])), vertices, (float(size[0]), float(size[1])))


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for case in CASES:
        rec = record(case)
        with open(os.path.join(out, case["name"] + ".json"), "w") as fh:
            json.dump(rec, fh, indent=0, separators=(",", ":"))
        print(case["name"], len(rec["contours"]), "contours", len(rec["positive"]), "positive", len(rec["armours"]), "armours")
