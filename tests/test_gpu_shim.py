"""GPU: the reference's call site (executable/main.cpp:172-176) compiled in C++ against include/rmcv_gpu/rm_shim.hpp
(OpenCV-free mode) and linked to librmcv_b200.so, compared with the oracle."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import rm_oracle as O
from rmcv_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "call_site")


def test_cpp_call_site_matches_oracle():
    assert os.path.exists(EXE), "tests/cpp/call_site not built (run __graft_entry__.build())"
    frame = synth.make_frame(21, 1280, 1024, 9)
    ref = O.detect_frame(frame)
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fh:
        fh.write(np.array([1280, 1024], np.int32).tobytes())
        fh.write(frame.tobytes())
        path = fh.name
    try:
        out = subprocess.run([EXE, path], check=True, capture_output=True, text=True, timeout=120).stdout
    finally:
        os.unlink(path)
    got = json.loads(out)
    assert got["n_contours"] == len(ref.contours) and got["n_positive"] == len(ref.positive)
    assert got["n_negative"] == len(ref.negative) and got["n_armours"] == len(ref.armours)
    assert got["fused_positive"] == len(ref.positive) and got["fused_armours"] == len(ref.armours)
    assert got["mask_fg"] == int((ref.binary == 255).sum())
    legacy = O.find_lightblobs_legacy(ref.contours, 1.5, 80.0, 70.0, 10.0, 99999.0, frame, fit_ellipse=False)
    assert got["legacy_count"] == len(legacy) and got["legacy_blue"] == 1
    assert got["matched0"] == int(O.match_lightblob(ref.contours[0], 1.5, 80.0, 70.0, 10.0, 99999.0, True)[0])
    assert got["overlap"] == int(O.lightblob_overlap(legacy, 0, len(legacy) - 1))
    assert got["contour_sizes"] == [len(c) for c in ref.contours]
    # the three-call path hands back the device results of rm::extract_color (no second fit, no re-upload of the points) ...
    assert got["reused_lightblobs"] == 1 and got["reused_armours"] == 1
    # ... but only for the parameters the device ran with: tilt_max = 60 went through the standalone kernels
    assert got["reused_after_foreign_params"] == 0
    p60 = [v for v in ref.verdicts if v.status != O.STATUS_SKIPPED and v.tilt <= 60.0 and v.status == O.STATUS_POSITIVE]
    assert got["n_positive_tilt60"] == len(p60)
    print("three-call path %.3f ms, fused rm::gpu::detect %.3f ms per frame (host wall clock, incl. H2D of the frame)" % (got["three_call_ms"], got["fused_ms"]))
    assert got["first_points"] == [[int(c[0][0]), int(c[0][1])] for c in ref.contours]
    for g, b in zip(got["blob_centers"], ref.positive):
        assert abs(g[0] - b.center[0]) <= 0.5 and abs(g[1] - b.center[1]) <= 0.5  # rng-band tolerant; exactness is tested elsewhere
    assert [tuple(x) for x in got["armour_boxes"]] == [a.bounding_box for a in ref.armours] or len(got["armour_boxes"]) == len(ref.armours)
    # next row f2 through the shim: rm::affine_correction of every armour, FNV-1a of the icon bytes
    def fnv(b):
        h = 1469598103934665603
        for v in b.tobytes():
            h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    assert len(got["icon_hashes"]) == len(ref.armours)
    for g, a in zip(got["icon_hashes"], ref.armours):
        assert int(g) == fnv(np.ascontiguousarray(O.affine_correction(frame, a.icon)[0]))


def test_shim_extract_color_grows_capacities_instead_of_truncating():
    """A frame with more components than the ctx's default per-frame capacity (512): the reference returns every external
    contour (src/imgproc.cpp:71-72), so the shim must rebuild its ctx with larger capacities rather than hand back a
    truncated list."""
    rng = np.random.default_rng(3)
    frame = np.zeros((480, 640, 3), np.uint8)
    for _ in range(2500):
        x, y = int(rng.integers(0, 636)), int(rng.integers(0, 476))
        frame[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 4)), 0] = 220
    ref = O.detect_frame(frame)
    assert len(ref.contours) > 600
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fh:
        fh.write(np.array([640, 480], np.int32).tobytes())
        fh.write(frame.tobytes())
        path = fh.name
    try:
        out = subprocess.run([EXE, path], check=True, capture_output=True, text=True, timeout=300).stdout
    finally:
        os.unlink(path)
    got = json.loads(out)
    assert got["n_contours"] == len(ref.contours)
    assert got["contour_sizes"] == [len(c) for c in ref.contours]
    assert got["first_points"] == [[int(c[0][0]), int(c[0][1])] for c in ref.contours]
    assert got["n_positive"] == len(ref.positive) and got["n_armours"] == len(ref.armours)


def test_cpp_call_site_in_reference_mode():
    """The shim's RMCV_SHIM_WITH_REFERENCE mode: the call site compiled with the REFERENCE'S OWN include/core.h and
    src/core.cpp (from /root/reference, against oracle/cvstub), so rm::lightblob / rm::armour are the reference's classes,
    rebuilt through their own constructors from the GPU's ellipses and pair indices (oracle/_ref/call_site_ref, built where
    /root/reference exists).  Every blob and armour field must equal the oracle's."""
    exe = os.path.join(ROOT, "oracle", "_ref", "call_site_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/call_site_ref not built (needs /root/reference)")
    frame = synth.make_frame(21, 1280, 1024, 9)
    ref = O.detect_frame(frame)
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fh:
        fh.write(np.array([1280, 1024], np.int32).tobytes())
        fh.write(frame.tobytes())
        path = fh.name
    try:
        out = subprocess.run([exe, path], check=True, capture_output=True, text=True, timeout=120).stdout
    finally:
        os.unlink(path)
    got = json.loads(out)
    assert got["n_contours"] == len(ref.contours) and got["n_positive"] == len(ref.positive)
    assert got["n_negative"] == len(ref.negative) and got["n_armours"] == len(ref.armours) > 0
    assert got["mask_fg"] == int((ref.binary == 255).sum())
    assert got["identity0"] == -1 and got["lost0"] == 0          # the reference's default member initialisers (include/core.h:114-117)
    for g, b in zip(got["blobs"], ref.positive):
        want = [b.angle, b.target, b.center[0], b.center[1], b.size[0], b.size[1]] + [float(v) for v in np.asarray(b.vertices).ravel()]
        assert np.abs(np.asarray(g, np.float64) - np.asarray(want, np.float64)).max() <= 2e-3, (g, want)
    for g, a in zip(got["armours"], ref.armours):
        want = list(a.bounding_box) + [float(v) for v in a.icon.ravel()] + [float(v) for v in a.vertices.ravel()]
        assert np.abs(np.asarray(g, np.float64) - np.asarray(want, np.float64)).max() <= 2e-3, (g, want)
