"""GPU parity at the sizes of BASELINE.json's configs 2-5 (config 1 is tests/test_gpu_detect.py::test_config1_single_frame):
  config 2  1440x1080 raw Bayer batch: debayer + colour difference + threshold + close (+ the rest of the path)
  config 3  1280x1024 full detection over a large batch: results are independent of batch position / chunking
  config 4  4096x3072 stress frame with ~500 light blobs
  config 5  batch-1 stream of frames (latency mode): identical results call after call, two calls in flight
Everything goes through the C ABI (ctypes) and is compared with the cv2 oracle."""
import zlib

import numpy as np
import pytest

import rmcv_b200 as rb
from oracle import rm_oracle as O
from rmcv_b200 import synth
from tests import _compare as CMP
from tests.test_gpu_detect import PRM, c_params, detect_and_compare

pytestmark = pytest.mark.gpu


def signature(ctx, res, f):
    """Everything the path returns for frame f, as a hashable tuple (floats compared bit for bit)."""
    d = ctx.frame_detections(res, f)
    return (tuple((c.first, c.n_points, c.area2, tuple(c.bbox), c.status, c.fit_branch, tuple(np.float32(c.ellipse).tobytes())) for c in d.contours),
            tuple(np.asarray(b.vertices, np.float32).tobytes() for b in d.positive),
            tuple((a.i, a.j, np.asarray(a.icon, np.float32).tobytes(), tuple(a.bounding_box)) for a in d.armours))


def test_config2_bayer_batch_full_path():
    """64 x 1440x1080 BGGR mosaics: mask bit-exact against cv2 demosaic + extract_color, detections against the oracle."""
    W, H, B, distinct = 1440, 1080, 64, 4
    # two frames with the coverage extras (after demosaicing the diagonal chain becomes a 2-px-wide band, a declared
    # degenerate fit: those frames are compared up to the contour statistics) and two without (compared in full)
    bgr = [synth.make_frame(100 + s, W, H, 10, extras=(s < 2)) for s in range(distinct)]
    raw = np.stack([synth.bgr_to_bayer(bgr[i % distinct], synth.BAYER_BG) for i in range(B)])
    with rb.Context(max_width=W, max_height=H, max_batch=B) as c:
        d_in = c.device_buffer(raw.nbytes); d_mask = c.device_buffer(B * H * W)
        d_in.upload(raw)
        c.bayer_detect_batch(d_in.ptr, W, H, B, rb.BAYER_BG, c_params(PRM), d_mask.ptr)
        res = c.fetch_results()
        masks = d_mask.download((B, H, W))
        refs = [O.detect_frame(O.bayer_to_bgr(raw[i], synth.BAYER_BG)) for i in range(distinct)]
        sigs = [signature(c, res, f) for f in range(B)]
        for f in range(B):
            ref = refs[f % distinct]
            assert np.array_equal(masks[f], ref.binary), f"frame {f}: mask differs from cv2 demosaic + extract_color"
            assert sigs[f] == sigs[f % distinct], f"frame {f}: result depends on the batch position"
        rep = CMP.Report()
        for f in range(distinct):
            rep.merge(CMP.compare_frame(c.frame_detections(res, f), refs[f], PRM, where=f"bayer frame {f}"))
        assert rep.blobs >= 30 and rep.armours >= 15, "the frames without extras must be compared down to the armours"
        d_in.free(); d_mask.free()


def test_config3_large_batch_is_position_independent():
    """A batch larger than one chunk (and than the SM count): every copy of a frame gives bit-identical results and
    masks, whatever its chunk and slot; the distinct frames are checked against the oracle."""
    W, H, distinct, B = 1280, 1024, 6, 300
    base = np.stack([synth.make_frame(200 + s, W, H, synth.plates_for_seed(200 + s)) for s in range(distinct)])
    order = np.random.default_rng(7).integers(0, distinct, B)
    frames = base[order]
    with rb.Context(max_width=W, max_height=H, max_batch=B, chunk_frames=128) as c:
        d_in = c.device_buffer(frames.nbytes); d_mask = c.device_buffer(B * H * W)
        d_in.upload(frames)
        c.detect_batch(d_in.ptr, W, H, B, c_params(PRM), d_mask.ptr)
        res = c.fetch_results()
        masks = d_mask.download((B, H, W))
        refs = [O.detect_frame(base[k]) for k in range(distinct)]
        crc = [zlib.crc32(r.binary.tobytes()) for r in refs]
        first = {}
        for f in range(B):
            k = int(order[f])
            assert zlib.crc32(masks[f].tobytes()) == crc[k], f"frame {f}: mask checksum differs from the oracle's"
            s = signature(c, res, f)
            if k not in first:
                first[k] = s
                CMP.compare_frame(c.frame_detections(res, f), refs[k], PRM, where=f"frame {f} (distinct {k})")
            assert s == first[k], f"frame {f}: result depends on the batch position"
        assert res.total_contours == sum(len(refs[int(k)].contours) for k in order)
        d_in.free(); d_mask.free()


def test_config4_stress_frame():
    """4096x3072 with 250 plates -> ~500 light blobs: labelling runs on global arrays (too many runs for shared memory),
    ~125k pairs go through the armour gates."""
    W, H = 4096, 3072
    frame = synth.make_stress_frame(3, W, H, 250)
    with rb.Context(max_width=W, max_height=H, max_batch=2, max_blobs_per_frame=1024, max_armours_per_frame=4096) as c:
        frames = np.stack([frame, frame[::-1].copy()])   # the flipped copy exercises different runs with the same load
        rep = detect_and_compare(c, frames, check_points=False, check_labels=True, what="stress")
        assert rep.contours >= 2 * 500 and rep.blobs >= 2 * 450 and rep.armours >= 2 * 250
        # ordered points of a sample of contours (all of them would be 1000 launches)
        ref = O.detect_frame(frames[1])
        res = c.detect_batch_host(frames, c_params(PRM))
        got = c.get_contours(1)
        assert len(got) == len(ref.contours)
        for k in range(0, len(got), 7):
            assert np.array_equal(got[k], ref.contours[k]), f"stress contour {k}"
        assert res.total_contours == rep.contours


def test_config5_stream_of_single_frames_and_two_calls_in_flight():
    """Latency mode: batch 1, frame after frame through one ctx; then the same stream with two calls in flight
    (call n+1 enqueued before call n is fetched): every fetch returns the right call's results."""
    W, H, N = 1280, 1024, 12
    frames = [synth.make_frame(300 + s, W, H, synth.plates_for_seed(300 + s)) for s in range(N)]
    refs = [O.detect_frame(f) for f in frames]
    with rb.Context(max_width=W, max_height=H, max_batch=1) as c:
        bufs = []
        for f in frames:
            b = c.device_buffer(f.nbytes); b.upload(f); bufs.append(b)
        d_mask = [c.device_buffer(H * W) for _ in range(2)]
        sync_sigs = []
        for k in range(N):
            c.detect_batch(bufs[k].ptr, W, H, 1, c_params(PRM), d_mask[0].ptr)
            res = c.fetch_results()
            CMP.compare_frame(c.frame_detections(res, 0), refs[k], PRM, where=f"stream frame {k}")
            assert np.array_equal(d_mask[0].download((H, W)), refs[k].binary)
            sync_sigs.append(signature(c, res, 0))
        # two in flight
        c.detect_batch(bufs[0].ptr, W, H, 1, c_params(PRM), d_mask[0].ptr)
        for k in range(1, N):
            c.detect_batch(bufs[k].ptr, W, H, 1, c_params(PRM), d_mask[k & 1].ptr)
            res = c.fetch_results()      # call k-1
            assert signature(c, res, 0) == sync_sigs[k - 1], f"pipelined fetch {k - 1} returned another call's results"
            assert np.array_equal(d_mask[(k - 1) & 1].download((H, W)), refs[k - 1].binary)
        res = c.fetch_results()
        assert signature(c, res, 0) == sync_sigs[N - 1]
        # a fetch without an unfetched call re-exposes the last results
        assert signature(c, c.fetch_results(), 0) == sync_sigs[N - 1]
        for b in bufs + d_mask:
            b.free()
