"""f3 (next row): armour tracking on the GPU (rmcv_tracker_update: IoU association, 6-state Kalman filter, identity vote)
against the oracle's restatement of the reference's tracking loop (executable/main.cpp:57-88, src/core.cpp:51-161) through
cv2.KalmanFilter.  Integer state exact, fp64 filter state to 1e-9 relative."""
import os

import numpy as np
import pytest

import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O

pytestmark = pytest.mark.gpu
SEED_OFFSET = int(os.environ.get("RMCV_TEST_SEED", "0"))   # other random cases: RMCV_TEST_SEED=n pytest -m gpu ...
FREQ = 1e9   # cv::getTickFrequency() on Linux: nanoseconds


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=1280, max_height=1024, max_batch=2) as c:
        yield c


def armour_with_box(box):
    z = ((0.0, 0.0),) * 4
    return rb.Armour(z, z, tuple(float(v) for v in box), 0, 1, (0.0,) * 6)


def compare(gpu_tracks, ref_tracks, tracker, where):
    assert len(gpu_tracks) == len(ref_tracks), f"{where}: {len(gpu_tracks)} tracks, oracle {len(ref_tracks)}"
    for k, (g, r) in enumerate(zip(gpu_tracks, ref_tracks)):
        w = f"{where} track {k}"
        assert tuple(g.bbox) == tuple(float(v) for v in r.bounding_box), w
        assert g.timestamp == r.timestamp and g.lost_count == r.lost_count and g.identity == r.identity, w
        assert bool(g.initialized) == r.initialized, w
        hist = {g.hist_id[i]: g.hist_count[i] for i in range(g.n_hist)}
        assert hist == r.identity_history, w
        assert np.array_equal(np.array(g.position[:]), r.position), w
        for name, got, ref in (("statePost", g.state_post, r.observer.statePost), ("errorCovPost", g.cov_post, r.observer.errorCovPost),
                               ("measurement", g.meas, r.measurement)):
            got = np.array(got[:]); ref = np.asarray(ref, np.float64).reshape(-1)
            assert np.array_equal(np.isfinite(got), np.isfinite(ref)), f"{w} {name} finiteness"
            fin = np.isfinite(ref)
            scale = max(1.0, np.abs(ref[fin]).max()) if fin.any() else 1.0
            assert np.abs(got[fin] - ref[fin]).max() <= 1e-9 * scale if fin.any() else True, f"{w} {name}"
        if r.identity_history:
            gi, gp = tracker.identity_max(g)
            ri, rp = r.identity_max()
            assert gi == ri and abs(gp - rp) <= 1e-12, w


def run_sequence(ctx, frames, where):
    """frames: list of (timestamp, [(box, position, identity), ...])."""
    trk = rb.Tracker(ctx, capacity=64)
    ref = []
    try:
        for n, (ts, obs) in enumerate(frames):
            trk.update([armour_with_box(b) for b, _, _ in obs], [p for _, p, _ in obs], [i for _, _, i in obs], ts, FREQ)
            ref = O.tracking_step(ref, [O.TrackedArmour(b, p, i, ts) for b, p, i in obs], FREQ)
            compare(trk.read(), ref, trk, f"{where} frame {n}")
    finally:
        trk.close()
    return ref


def test_moving_targets_with_dropouts_births_and_deaths(ctx):
    """Three armours drifting at constant velocity with measurement noise; one disappears for good (its track is erased
    after 27 misses, and the erase skips the track behind it exactly like the reference's loop), one drops out for a few
    frames, a new one appears late; identities flicker."""
    rng = np.random.default_rng(1 + SEED_OFFSET)
    frames = []
    ts = 1_000_000
    for n in range(60):
        ts += int(8e6 + rng.integers(-2e5, 2e5))     # ~125 Hz
        obs = []
        for k, (x0, y0, vx) in enumerate(((100.0, 200.0, 1.5), (600.0, 300.0, -1.0), (900.0, 700.0, 0.5))):
            if k == 0 and n >= 12:
                continue                                  # gone for good
            if k == 1 and 20 <= n < 24:
                continue                                  # short dropout
            box = (x0 + vx * min(n, 10) + rng.normal(0, 0.3), y0 + rng.normal(0, 0.3), 80.0 + rng.normal(0, 0.5), 60.0)
            pos = (1000.0 + 10.0 * k + 3.0 * n + rng.normal(0, 0.5), 50.0 * k + rng.normal(0, 0.5), 2000.0 - 2.0 * n + rng.normal(0, 0.5))
            ident = int(rng.choice([k + 1, k + 1, k + 1, 5]))
            obs.append((box, pos, ident))
        if n >= 30:
            obs.append(((300.0, 900.0, 50.0, 40.0), (500.0 + n, 10.0, 1500.0), 4))
        order = rng.permutation(len(obs))
        frames.append((ts, [obs[i] for i in order]))
    ref = run_sequence(ctx, frames, "moving")
    assert len(ref) >= 3 and any(t.lost_count > 0 for t in ref)


def test_empty_frames_equal_timestamps_and_overlapping_boxes(ctx):
    """Empty frames change nothing; equal timestamps give dt = 0 (infinite velocity measurements, like the reference);
    two observations over the same track: the first maximum wins and the other opens a new track."""
    frames = [
        (10, [((10, 10, 50, 50), (1, 2, 3), 1), ((200, 200, 50, 50), (4, 5, 6), 2)]),
        (20, []),
        (30, [((12, 11, 50, 50), (1.5, 2.5, 3.5), 1), ((11, 10, 50, 50), (9, 9, 9), 3)]),
        (30, [((12, 11, 50, 50), (2.0, 3.0, 4.0), 1)]),
        (45, [((12, 11, 50, 50), (2.5, 3.5, 4.5), 1), ((0, 0, 0, 0), (0, 0, 0), -1), ((-30, -30, 50, 50), (7, 7, 7), 2)]),
    ]
    run_sequence(ctx, frames, "edge")


def test_tracks_of_detected_armours(ctx):
    """Boxes and positions from the detector itself: the armours of a synthetic frame, shifted frame to frame."""
    frame = synth.make_frame(77, 1280, 1024, 9)
    res = ctx.detect_batch_host(frame[None], rb.default_params())
    arm = ctx.frame_detections(res, 0).armours
    assert len(arm) >= 5
    poses = ctx.solve_pnp(arm, O.MAIN_CAMMAT, O.MAIN_DISCOF, (27.0, 27.0))
    frames = []
    for n in range(8):
        obs = [((a.bounding_box[0] + 2 * n, a.bounding_box[1] + n, a.bounding_box[2], a.bounding_box[3]), p[1] + n, k % 7)
               for k, (a, p) in enumerate(zip(arm, poses))]
        frames.append((1000 + 8_000_000 * n, obs))
    run_sequence(ctx, frames, "detected")


def test_capacity_and_arguments(ctx):
    trk = rb.Tracker(ctx, capacity=2)
    try:
        boxes = [((100 * k, 0, 50, 50), (k, k, k), k) for k in range(3)]
        with pytest.raises(rb.RmcvError):
            trk.update([armour_with_box(b) for b, _, _ in boxes], [p for _, p, _ in boxes], None, 5, FREQ)
        assert trk.read() == []
        trk.update([armour_with_box(b) for b, _, _ in boxes[:2]], [p for _, p, _ in boxes[:2]], None, 5, FREQ)
        assert len(trk.read()) == 2 and trk.read()[0].identity == -1
        with pytest.raises(rb.RmcvError):
            trk.update([armour_with_box(boxes[2][0])], [boxes[2][1]], None, 6, 0.0)
        trk.reset()
        assert trk.read() == []
    finally:
        trk.close()
