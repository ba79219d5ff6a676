"""GPU aid: SHA-256 over every result byte (frame infos, contour / blob / armour records, masks) of a fixed set of small
device-resident and host batches.  tests/test_gpu_fuzz.py runs it twice - with the latency-mode defaults (chained launches on
the slot stream, fits on the contour kernel's warps, 8-row emit bands) and with all three switched off (tuning is read once
per process) - and requires the same digest: the latency mode must not change a bit of the results."""
import hashlib, os, struct, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmcv_b200 as rb
from rmcv_b200 import synth

h = hashlib.sha256()
n_frames = 0


def absorb(c, res, B, masks):
    global n_frames
    for f in range(B):   # the records of each frame in frame order (the dense offsets depend on the order in which frames finish)
        fi = res.frames[f]
        h.update(struct.pack("5i", fi.n_contours, fi.n_positive, fi.n_negative, fi.n_armours, fi.flags))
        for k in range(fi.n_contours): h.update(bytes(res.contours[fi.contour_offset + k]))
        for k in range(fi.n_positive): h.update(bytes(res.blobs[fi.blob_offset + k]))
        for k in range(fi.n_armours): h.update(bytes(res.armours[fi.armour_offset + k]))
        n_frames += 1
    h.update(np.ascontiguousarray(masks).tobytes())


for (W, H, sizes) in ((1280, 1024, (1, 3, 5, 16)), (640, 480, (1, 2, 16)), (333, 77, (1, 7))):
    for B in sizes:
        frames = np.stack([synth.make_frame(4000 + 31 * B + s, W, H, synth.plates_for_seed(4000 + s), blue=(s % 2 == 0)) for s in range(B)])
        for target in (rb.CAMP_BLUE, rb.CAMP_RED):
            prm = rb.default_params(target=target)
            with rb.Context(max_width=W, max_height=H, max_batch=B) as c:
                # host path (masks downloaded by the library) ...
                masks = np.empty((B, H, W), np.uint8)
                res = c.detect_batch_host(frames, prm, masks)
                absorb(c, res, B, masks)
                # ... and the device-resident path, two calls in flight, three rounds (slot and result-set rotation)
                buf = c.device_buffer(frames.nbytes); buf.upload(frames)
                dm = c.device_buffer(B * H * W)
                for rnd in range(3):
                    c.detect_batch(buf.ptr, W, H, B, prm, dm.ptr)
                    c.detect_batch(buf.ptr, W, H, B, prm, dm.ptr)
                    r1 = c.fetch_results(); absorb(c, r1, B, dm.download((B, H, W)))
                    r2 = c.fetch_results(); absorb(c, r2, B, dm.download((B, H, W)))
print("latency-mode digest over %d frame results: %s" % (n_frames, h.hexdigest()))
