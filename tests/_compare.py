"""Comparison of the GPU path's results with the oracle's, with the tolerances of SURVEY.md §8(c).

Bit-exact: mask bytes, contour count, per-contour first pixel / point count / 2*area / bbox, ordered contour points,
label map.  Float: ellipse centre <= 1e-3 px, axes <= 1e-5 relative (+1e-4 px), angle <= 1e-3 deg (mod 180, not
compared for near-isotropic blobs); light-blob and armour vertices <= 2e-3 px; gate values <= 1e-3.
Carve-outs (counted, returned in the report):
  * `rng_band`: contours whose first direct-fit |det M| lies in [0.7e-10, 1e-10*(1+1e-6)]: cv::fitEllipseDirect
    itself is non-deterministic there (jittered retry from the global RNG) -> loose tolerance
    0.5 px / 0.5 % / 0.1 deg and no exact gate membership; when the oracle's draw landed on a jittered direct fit that
    is further away than that (thin ragged blobs), the GPU's answer must equal the reference's other possible outcome,
    cv::fitEllipseNoDirect on the same contour, to the full tolerance;
  * `near_gate`: contours / pairs whose gate quantity is within tolerance of its threshold;
  * `degenerate`: contours whose oracle ellipse is thinner than 2 px (or not finite): the contour points lie on two
    parallel lines, the conic through them is a line pair, cv::fitEllipse's least-squares systems are singular and its
    answer is rounding noise (SURVEY A.6 "degenerate input warning") -> geometry and verdict not compared.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from oracle import rm_oracle as O

TOL_CENTRE = 1e-3
TOL_AXIS_REL = 1e-5
TOL_AXIS_ABS = 1e-4
TOL_ANGLE = 1e-3
TOL_VERT = 2e-3
TOL_GATE = 1e-3
LOOSE = dict(centre=0.5, rel=5e-3, angle=0.1, vert=1.0)


@dataclass
class Report:
    frames: int = 0
    contours: int = 0
    fitted: int = 0
    direct: int = 0
    fallback: int = 0
    rng_band: int = 0
    near_gate: int = 0
    degenerate: int = 0
    long_fallback: int = 0    # fallback fits of contours whose coordinate sums reach 2^24 (RMCV_FIT_FALLBACK_LONG)
    flip_frames: int = 0      # frames with a verdict flip inside tolerance (still compared through the contours)
    blobs: int = 0
    armours: int = 0
    worst_centre: float = 0.0
    worst_axis_rel: float = 0.0
    worst_angle: float = 0.0
    worst_vertex: float = 0.0
    notes: list = field(default_factory=list)

    def merge(self, o: "Report"):
        for k in ("frames", "contours", "fitted", "direct", "fallback", "rng_band", "near_gate", "degenerate", "long_fallback", "flip_frames", "blobs", "armours"):
            setattr(self, k, getattr(self, k) + getattr(o, k))
        for k in ("worst_centre", "worst_axis_rel", "worst_angle", "worst_vertex"):
            setattr(self, k, max(getattr(self, k), getattr(o, k)))
        self.notes += o.notes


def angle_diff(a, b):
    return abs(((a - b) + 90.0) % 180.0 - 90.0)


def in_rng_band(det0: float) -> bool:
    return 0.7e-10 <= det0 <= 1.0e-10 * (1 + 1e-6)


def near_blob_gate(v: O.ContourVerdict, params) -> bool:
    """Is the oracle's decision for this contour within float tolerance of flipping?"""
    if v.status == O.STATUS_SKIPPED:
        return False
    r, t = v.ratio, v.tilt
    rel = 1e-5
    return (abs(r - params["ratio_range"][0]) <= rel * max(1, abs(r)) or abs(r - params["ratio_range"][1]) <= rel * max(1, abs(r))
            or abs(t - params["tilt_max"]) <= TOL_ANGLE)


def compare_frame(det, ref: O.FrameResult, params, where="") -> Report:
    """det: rmcv_b200.FrameDetections; ref: oracle FrameResult."""
    rep = Report(frames=1)
    assert det.flags == 0, f"{where}: overflow flags {det.flags}"
    assert len(det.contours) == len(ref.contours), f"{where}: contour count {len(det.contours)} != {len(ref.contours)}"
    rep.contours = len(ref.contours)
    loose_blob = {}   # oracle positive index -> loose?
    incomparable = set()   # contours whose geometry cannot be compared (degenerate fits, the RNG band's other outcome)
    flips = 0
    pos_idx = 0
    for k, (c, rc, v) in enumerate(zip(det.contours, ref.contours, ref.verdicts)):
        w = f"{where} contour {k}"
        assert c.first == (int(rc[0][0]), int(rc[0][1])), f"{w}: first pixel {c.first} != {tuple(rc[0])}"
        assert c.n_points == v.n, f"{w}: n_points {c.n_points} != {v.n}"
        assert c.area2 == int(round(2 * v.area)), f"{w}: area2 {c.area2} != {2 * v.area}"
        xs, ys = rc[:, 0], rc[:, 1]
        assert c.bbox == (int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)), f"{w}: bbox"
        if v.status == O.STATUS_SKIPPED:
            assert c.status == 0, f"{w}: status {c.status} for a skipped contour"
            continue
        assert c.status != 0, f"{w}: GPU skipped a contour the oracle fitted"
        rep.fitted += 1
        band = in_rng_band(c.det0)
        rep.rng_band += band
        if c.fit_branch == 1:
            rep.direct += 1
        else:
            rep.fallback += 1
        e = v.ellipse
        cx, cy, ew, eh, ea = c.ellipse
        if not (math.isfinite(e.w) and math.isfinite(e.h) and math.isfinite(e.cx) and math.isfinite(e.cy)) or e.w < 2.0:
            rep.degenerate += 1
            incomparable.add(k)
            if c.status != v.status:
                flips += 1
            if v.status == O.STATUS_POSITIVE:
                loose_blob[pos_idx] = True
                pos_idx += 1
            continue
        dc = max(abs(cx - e.cx), abs(cy - e.cy))
        ds = max(abs(ew - e.w) / max(e.w, 1e-9), abs(eh - e.h) / max(e.h, 1e-9))
        da = angle_diff(ea, e.angle) if e.h / max(e.w, 1e-9) >= 1 + 1e-4 else 0.0
        if c.fit_branch == 3 and not band:
            # RMCV_FIT_FALLBACK_LONG: cv::fitEllipseNoDirect sums the centre as a float Point2f point by point; past 2^24 that
            # rounds in contour order, the GPU returns the exactly rounded centre -> centre within 0.05 px, axes 1e-3 relative
            rep.long_fallback += 1
            assert dc <= 0.05 and ds <= 1e-3 and da <= 0.05, f"{w}: long-contour fallback off: {c.ellipse} vs {e}"
            incomparable.add(k)
        elif band:
            if not (dc <= LOOSE["centre"] and ds <= LOOSE["rel"] and da <= LOOSE["angle"]):
                # The reference's answer depends on its RNG here: a jittered retry of the direct fit when one succeeds, else
                # cv::fitEllipseNoDirect.  The two can be far apart for thin ragged blobs, so the GPU's (deterministic)
                # answer must then be the other outcome the reference can produce — the plain fallback — to full tolerance.
                import cv2
                (fx, fy), (fw, fh), fa = cv2.fitEllipse(np.ascontiguousarray(rc, np.int32).reshape(-1, 1, 2))
                ulp = float(np.spacing(np.float32(max(abs(fx), abs(fy), 1.0))))
                assert max(abs(cx - fx), abs(cy - fy)) <= max(TOL_CENTRE, 2 * ulp) and \
                    max(abs(ew - fw) / max(fw, 1e-9), abs(eh - fh) / max(fh, 1e-9)) <= TOL_AXIS_REL + TOL_AXIS_ABS / max(fw, 1e-9) and \
                    (angle_diff(ea, fa) <= TOL_ANGLE or fh / max(fw, 1e-9) < 1 + 1e-4), \
                    f"{w}: rng-band ellipse matches neither outcome of the reference: {c.ellipse} vs {e} / fallback {(fx, fy, fw, fh, fa)}"
                flips += 1      # the oracle's record of this contour comes from its other outcome
                incomparable.add(k)
        else:
            # tolerance floor: one fp32 ulp of the coordinate
            ulp = float(np.spacing(np.float32(max(abs(e.cx), abs(e.cy), 1.0))))
            assert dc <= max(TOL_CENTRE, 2 * ulp), f"{w}: centre off by {dc}: {c.ellipse} vs {e} det0={c.det0} branch={c.fit_branch}"
            assert ds <= TOL_AXIS_REL + TOL_AXIS_ABS / max(e.w, 1e-9), f"{w}: axes off by {ds}: {c.ellipse} vs {e} det0={c.det0}"
            assert da <= TOL_ANGLE, f"{w}: angle off by {da}: {c.ellipse} vs {e}"
            rep.worst_centre = max(rep.worst_centre, dc)
            rep.worst_axis_rel = max(rep.worst_axis_rel, ds)
            rep.worst_angle = max(rep.worst_angle, da)
        if c.status != v.status:
            if band or near_blob_gate(v, params):
                flips += 1
                rep.near_gate += 1
            else:
                raise AssertionError(f"{w}: status {c.status} != oracle {v.status} (ratio {v.ratio}, tilt {v.tilt})")
        if v.status == O.STATUS_POSITIVE:
            loose_blob[pos_idx] = band
            pos_idx += 1
    # ---- light blobs.  A verdict flip inside tolerance (counted above) changes the positive lists on one side only, so
    # blobs and pairs are matched through the CONTOUR they come from: everything that does not involve a flipped contour
    # is still compared in full.
    flipped = {k for k, (c, v) in enumerate(zip(det.contours, ref.verdicts)) if v.status != O.STATUS_SKIPPED and c.status != v.status}
    flipped |= incomparable
    if flips:
        rep.flip_frames = 1
        rep.notes.append(f"{where}: {flips} verdict flips inside tolerance ({len(flipped)} contours change lists)")
    det_pos = [k for k, c in enumerate(det.contours) if c.status == O.STATUS_POSITIVE]
    ref_pos = [k for k, v in enumerate(ref.verdicts) if v.status == O.STATUS_POSITIVE]
    assert len(det_pos) == len(det.positive), f"{where}: positive list length {len(det.positive)} != positive contours {len(det_pos)}"
    for di, k in enumerate(det_pos):
        assert det.contours[k].blob_index == di, f"{where}: contour {k} blob_index"
    if not flipped:
        assert len(det.positive) == len(ref.positive), f"{where}: positive count"
        assert det.n_negative == len(ref.negative), f"{where}: negative count"
    else:
        assert abs(len(det.positive) - len(ref.positive)) <= len(flipped) and abs(det.n_negative - len(ref.negative)) <= len(flipped), \
            f"{where}: list sizes differ by more than the flipped contours"
    det_of = {k: di for di, k in enumerate(det_pos)}
    ref_of = {k: ri for ri, k in enumerate(ref_pos)}
    loose_c = {ref_pos[ri]: bool(v) for ri, v in loose_blob.items() if ri < len(ref_pos)}   # by contour index
    common_c = [k for k in ref_pos if k in det_of and k not in incomparable]
    rep.blobs = len(common_c)
    for k in common_c:
        b, rb = det.positive[det_of[k]], ref.positive[ref_of[k]]
        w = f"{where} blob of contour {k}"
        lo = loose_c.get(k, False)
        tol = LOOSE["vert"] if lo else TOL_VERT
        assert b.target == rb.target
        assert angle_diff(b.angle, rb.angle) <= (LOOSE["angle"] if lo else TOL_ANGLE), f"{w}: angle"
        dv = float(np.max(np.abs(b.vertices - rb.vertices)))
        assert dv <= tol, f"{w}: vertices off by {dv}\n{b.vertices}\n{rb.vertices}"
        assert abs(b.size[0] - rb.size[0]) <= tol and abs(b.size[1] - rb.size[1]) <= tol, f"{w}: size"
        if not lo:
            rep.worst_vertex = max(rep.worst_vertex, dv)
    # ---- armours, keyed by the pair of contours
    det_pairs = [(a.i, a.j) for a in det.armours]
    assert det_pairs == sorted(det_pairs), f"{where}: armours not in lexicographic (i,j) order"
    rmap = {(ref_pos[a.i], ref_pos[a.j]): a for a in ref.armours}
    dmap = {(det_pos[a.i], det_pos[a.j]): a for a in det.armours}
    diff = set(rmap) ^ set(dmap)
    for (ci, cj) in diff:
        if ci in flipped or cj in flipped:
            continue            # one side does not have that blob at all
        g = O.pair_gates(ref.positive[ref_of[ci]], ref.positive[ref_of[cj]])
        near = (loose_c.get(ci) or loose_c.get(cj) or _near_pair_gate(g, params))
        assert near, f"{where}: armour pair of contours ({ci},{cj}) membership differs: gates {g}"
        rep.near_gate += 1
    if diff:
        rep.notes.append(f"{where}: {len(diff)} armour pairs differ inside tolerance / through flipped contours")
    common = [p for p in rmap if p in dmap and p[0] not in incomparable and p[1] not in incomparable]
    rep.armours = len(common)
    for p in common:
        a, ra = dmap[p], rmap[p]
        loose = loose_c.get(p[0]) or loose_c.get(p[1])
        tol = LOOSE["vert"] * 2 if loose else TOL_VERT
        assert float(np.max(np.abs(a.icon - ra.icon))) <= tol, f"{where}: armour {p} icon\n{a.icon}\n{ra.icon}"
        assert float(np.max(np.abs(a.vertices - ra.vertices))) <= tol, f"{where}: armour {p} vertices"
        if not loose:
            for q in range(6):
                assert abs(a.gates[q] - ra.gates[q]) <= TOL_GATE, f"{where}: armour {p} gate {q}: {a.gates} vs {ra.gates}"
            # bounding box = floor of icon extents: exact unless an icon coordinate sits on an integer
            if a.bounding_box != ra.bounding_box:
                ic = ra.icon
                near_int = np.min(np.abs(ic - np.round(ic))) <= TOL_VERT
                assert near_int, f"{where}: armour {p} bounding_box {a.bounding_box} != {ra.bounding_box}"
    return rep


def _near_pair_gate(g, params) -> bool:
    ad, si, sj, ratio, dy, dx, hsum = g
    t = TOL_GATE
    return (abs(ad - params["angle_difference_max"]) <= t or abs(si - params["shear_max"]) <= t or
            abs(sj - params["shear_max"]) <= t or abs(ratio - params["lenght_ratio_max"]) <= 1e-5 or
            abs(dy - hsum / 2) <= t or abs(dx - hsum * 2) <= t)


def oracle_params(p=None):
    from rmcv_b200 import synth
    d = dict(synth.MAIN_PARAMS)
    if p:
        d.update(p)
    return d
