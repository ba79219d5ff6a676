"""GPU: one batch partitioned by frame across devices inside the library (rmcv_multi_*) is identical,
byte for byte, to one device running the whole batch, and both match the oracle; the raw-Bayer host entry point
(rmcv_bayer_detect_batch_host) matches the device-resident one.  With a single visible device the partition still runs as
two workers on that device, so slicing, per-slice calls and the frame-order merge are exercised on every box."""
import ctypes as C
import zlib

import numpy as np
import pytest

import rmcv_b200 as rb
from oracle import rm_oracle as O
from rmcv_b200 import shard, synth

pytestmark = pytest.mark.gpu
W, H = 1280, 1024


def n_devices() -> int:
    n = C.c_int(0)
    rb.load_library().rmcv_device_count(C.byref(n))
    return n.value


def frame_digests(res, masks):
    """Per frame: everything a caller can read, as bytes — counts, flags, contour / blob / armour records, mask CRC."""
    out = []
    for f in range(res.batch):
        fi = res.frames[f]
        rec = [bytes(memoryview(res.contours[fi.contour_offset + k])) for k in range(fi.n_contours)]
        rec += [bytes(memoryview(res.blobs[fi.blob_offset + k])) for k in range(fi.n_positive)]
        rec += [bytes(memoryview(res.armours[fi.armour_offset + k])) for k in range(fi.n_armours)]
        out.append((fi.n_contours, fi.n_positive, fi.n_negative, fi.n_armours, fi.flags, zlib.crc32(b"".join(rec)),
                    zlib.crc32(masks[f].tobytes())))
    return out


@pytest.fixture(scope="module")
def batch():
    seeds = list(range(7000, 7000 + 37))      # not a multiple of any device count on purpose
    return np.stack([synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in seeds])


@pytest.fixture(scope="module")
def single(batch):
    with rb.Context(max_width=W, max_height=H, max_batch=len(batch), chunk_frames=7) as c:
        masks = np.empty(batch.shape[:3], np.uint8)
        res = c.detect_batch_host(batch, rb.default_params(), masks)
        return frame_digests(res, masks), masks


def test_slices_match_the_python_rule():
    lib = rb.load_library()
    for B in (0, 1, 5, 37, 128, 1024, 1025):
        for G in (1, 2, 3, 4, 8):
            got = []
            for g in range(G):
                a, n = C.c_int(0), C.c_int(0)
                lib.rmcv_multi_slice(B, G, g, C.byref(a), C.byref(n))
                got.append((a.value, a.value + n.value))
            assert got == shard.all_slices(B, G), (B, G)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], "all"])
def test_partitioned_batch_equals_single_device_and_oracle(batch, single, devices):
    ndev = n_devices()
    if devices == "all":
        if ndev < 2:
            pytest.skip("one visible device: the multi-device run needs gpurun --gpus N")
        devices = list(range(ndev))
    want, want_masks = single
    with rb.MultiContext(devices, max_width=W, max_height=H, max_batch=len(batch), chunk_frames=5) as m:
        assert m.n_devices == len(devices)
        masks = np.zeros(batch.shape[:3], np.uint8)
        for rep in range(2):       # a second call reuses the workers and their contexts
            res = m.detect_batch_host(batch, rb.default_params(), masks)
            assert res.batch == len(batch)
            assert frame_digests(res, masks) == want
        # offsets of the merged arrays are dense and in frame order
        oc = ob = oa = 0
        for f in range(res.batch):
            fi = res.frames[f]
            assert (fi.contour_offset, fi.blob_offset, fi.armour_offset) == (oc, ob, oa)
            oc += fi.n_contours; ob += fi.n_positive; oa += fi.n_armours
        assert (res.total_contours, res.total_blobs, res.total_armours) == (oc, ob, oa)
        # a batch smaller than the device count leaves trailing devices idle
        small = m.detect_batch_host(batch[:1], rb.default_params())
        assert small.batch == 1 and small.frames[0].n_contours == want[0][0]
    for f in (0, 17, 36):
        ref = O.detect_frame(batch[f])
        assert np.array_equal(want_masks[f], ref.binary)
        assert want[f][:4] == (len(ref.contours), len(ref.positive), len(ref.negative), len(ref.armours))


def test_multi_rejects_bad_arguments(batch):
    with rb.MultiContext([0], max_width=W, max_height=H, max_batch=4) as m:
        with pytest.raises(rb.RmcvError):
            m.detect_batch_host(batch[:5], rb.default_params())      # above max_batch
    with pytest.raises(rb.RmcvError):
        rb.MultiContext([99], max_width=W, max_height=H, max_batch=4)


@pytest.mark.parametrize("layout", [rb.BAYER_BG, rb.BAYER_GB])
def test_bayer_host_entry_matches_device_entry_and_oracle(batch, layout):
    raw = np.stack([synth.bgr_to_bayer(batch[f], layout) for f in range(9)])
    prm = rb.default_params()
    with rb.Context(max_width=W, max_height=H, max_batch=len(raw), chunk_frames=4) as c:
        hm = np.empty(raw.shape, np.uint8)
        res = c.bayer_detect_batch_host(raw, layout, prm, hm)
        host = frame_digests(res, hm)
        d_in = c.device_buffer(raw.nbytes); d_out = c.device_buffer(raw.nbytes)
        d_in.upload(raw)
        c.bayer_detect_batch(d_in.ptr, W, H, len(raw), layout, prm, d_out.ptr)
        res2 = c.fetch_results()
        dev = frame_digests(res2, d_out.download(raw.shape))
        d_in.free(); d_out.free()
    assert host == dev
    for f in (0, 8):
        ref = O.detect_frame(O.bayer_to_bgr(raw[f], layout))
        assert np.array_equal(hm[f], ref.binary) and host[f][:4] == (len(ref.contours), len(ref.positive), len(ref.negative), len(ref.armours))
    with rb.MultiContext([0, 0], max_width=W, max_height=H, max_batch=len(raw), chunk_frames=4) as m:
        mm = np.empty(raw.shape, np.uint8)
        assert frame_digests(m.bayer_detect_batch_host(raw, layout, prm, mm), mm) == host
