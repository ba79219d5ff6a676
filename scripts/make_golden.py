"""Generates tests/golden/*.json from the cv2 oracle (oracle/rm_oracle.py) on seeded synthetic frames.

The reference ships no golden vectors; these fixtures freeze what the oracle (i.e. OpenCV 4.13.0's
arithmetic driven exactly as rmcv drives it) returns in this image, so that a different cv2 on another box, or an
accidental change to the oracle or the generator, is caught by tests/test_golden.py.  Run from the repo root:
    python scripts/make_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rm_oracle as O  # noqa: E402
from rmcv_b200 import synth  # noqa: E402

CASES = [
    dict(name="config1_seed1_8plates_blue", seed=1, width=1280, height=1024, plates=8, blue=True, target=1),
    dict(name="seed7_14plates_red", seed=7, width=1280, height=1024, plates=14, blue=False, target=0),
    dict(name="seed3_640x480_5plates_blue", seed=3, width=640, height=480, plates=5, blue=True, target=1),
]


def record(case):
    img = synth.make_frame(case["seed"], case["width"], case["height"], case["plates"], blue=case["blue"])
    p = dict(synth.MAIN_PARAMS)
    p["target"] = case["target"]
    fr = O.detect_frame(img, **p)
    rec = dict(case=case, camera=dict(matrix=O.MAIN_CAMMAT.tolist(), dist=O.MAIN_DISCOF.tolist(), exact_size=[27.0, 27.0]),
               frame_crc32=zlib.crc32(img.tobytes()), mask_crc32=zlib.crc32(fr.binary.tobytes()),
               mask_foreground=int((fr.binary > 0).sum()), contours=[], positive=[], armours=[])
    for c, v in zip(fr.contours, fr.verdicts):
        e = v.ellipse
        rec["contours"].append(dict(first=[int(c[0][0]), int(c[0][1])], n=v.n, area2=int(round(2 * v.area)),
                                    bbox=[int(c[:, 0].min()), int(c[:, 1].min()), int(np.ptp(c[:, 0]) + 1), int(np.ptp(c[:, 1]) + 1)],
                                    points_crc32=zlib.crc32(np.ascontiguousarray(c, np.int32).tobytes()), status=v.status,
                                    ellipse=None if e is None else [e.cx, e.cy, e.w, e.h, e.angle]))
    for b in fr.positive:
        rec["positive"].append(dict(angle=b.angle, center=list(b.center), size=list(b.size), vertices=b.vertices.tolist()))
    for a in fr.armours:
        rvec, tvec = O.solve_pnp(a.vertices)                                    # next row f1
        icon, _ = O.affine_correction(img, a.icon)                              # next row f2
        rec["armours"].append(dict(i=a.i, j=a.j, icon=a.icon.tolist(), vertices=a.vertices.tolist(), bounding_box=list(a.bounding_box),
                                   gates=list(a.gates), rvec=rvec.tolist(), tvec=tvec.tolist(),
                                   icon20_crc32=zlib.crc32(np.ascontiguousarray(icon).tobytes())))
    # next row f3: the frame's armours drifting by (2, 1) px per frame through the tracking loop, 6 frames at 125 Hz
    tracking = []
    for n in range(6):
        obs = [O.TrackedArmour((a.bounding_box[0] + 2 * n, a.bounding_box[1] + n, a.bounding_box[2], a.bounding_box[3]),
                               O.solve_pnp(a.vertices)[1] + n, k % 7, 1000 + 8_000_000 * n) for k, a in enumerate(fr.armours)]
        tracking = O.tracking_step(tracking, obs, 1e9)
    rec["tracking"] = [dict(lost_count=t.lost_count, timestamp=t.timestamp, history=sorted(t.identity_history.items()),
                            state_post=t.observer.statePost.ravel().tolist(), cov_post_diag=np.diag(t.observer.errorCovPost).tolist())
                       for t in tracking]
    # Bayer stand-in (config 2 front): mosaic -> cv2 bilinear -> pixel stage
    raw = synth.bgr_to_bayer(img, synth.BAYER_BG)
    bmask = O.extract_color_mask(O.bayer_to_bgr(raw, 4), case["target"], 80)
    rec["bayer_bg_mask_crc32"] = zlib.crc32(bmask.tobytes())
    return rec


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for case in CASES:
        rec = record(case)
        with open(os.path.join(out, case["name"] + ".json"), "w") as fh:
            json.dump(rec, fh, indent=0, separators=(",", ":"))
        print(case["name"], len(rec["contours"]), "contours", len(rec["positive"]), "positive", len(rec["armours"]), "armours")
