"""GPU parity of the pixel stage (rm::extract_color up to the binary mask, src/imgproc.cpp:52-69) and of the Bayer
front, through the C ABI, against the cv2 oracle.  Bit-exact."""
import os

import numpy as np
import pytest

import rmcv_b200 as rb
from oracle import cv_restate as R
from oracle import rm_oracle as O
from rmcv_b200 import synth

pytestmark = pytest.mark.gpu
SEED_OFFSET = int(os.environ.get("RMCV_TEST_SEED", "0"))   # other random cases: RMCV_TEST_SEED=n pytest -m gpu ...


def run_mask(ctx, frames, target, lb, pitch=None):
    B, H, W, _ = frames.shape
    if pitch is None:
        src = frames
        pitch_b = W * 3
    else:
        src = np.zeros((B, H, pitch), np.uint8)
        src[:, :, :W * 3] = frames.reshape(B, H, W * 3)
        pitch_b = pitch
    d_in = ctx.device_buffer(src.nbytes)
    d_out = ctx.device_buffer(B * H * W)
    try:
        d_in.upload(src)
        ctx.extract_color_batch(d_in.ptr, W, H, B, target, lb, d_out.ptr, pitch=pitch_b, frame_stride=pitch_b * H)
        ctx.sync()
        mask = d_out.download((B, H, W))
        bits = []
        for f in range(B):  # only the last two chunks stay resident inside the ctx
            try:
                bits.append(ctx.get_bitmask(f, W, H))
            except rb.RmcvError as e:
                assert e.status == rb.abi.RMCV_ERR_STATE
                bits.append(None)
    finally:
        d_in.free(); d_out.free()
    return mask, bits


def unpack_bits(words, W):
    b = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
    return b[:, :W].astype(bool)


def check(ctx, frames, target, lb, pitch=None, what=""):
    mask, bits = run_mask(ctx, frames, target, lb, pitch)
    B, H, W, _ = frames.shape
    for f in range(B):
        ref = O.extract_color_mask(frames[f], target, lb)
        bad = np.argwhere(mask[f] != ref)
        assert bad.size == 0, f"{what} frame {f}: {len(bad)} mask bytes differ, first at (y,x)={bad[0]}, got {mask[f][tuple(bad[0])]}"
        assert set(np.unique(mask[f])) <= {0, 255}
        if bits[f] is not None:
            assert np.array_equal(unpack_bits(bits[f], W), ref > 0), f"{what} frame {f}: bit-packed mask differs"


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=1440, max_height=1080, max_batch=8) as c:
        yield c


def test_synthetic_1280x1024_blue_red(ctx):
    frames = np.stack([synth.make_frame(1, 1280, 1024, 8, blue=True), synth.make_frame(2, 1280, 1024, 12, blue=False)])
    check(ctx, frames, rb.CAMP_BLUE, 80, what="blue")
    check(ctx, frames, rb.CAMP_RED, 80, what="red")
    check(ctx, frames, rb.CAMP_GUIDELIGHT, 40, what="guide")
    check(ctx, frames, rb.CAMP_NEUTRAL, 80, what="neutral(-1) takes the red branch")


@pytest.mark.parametrize("lb", [-5, 0, 1, 79, 80, 81, 254, 255, 256, 1000])
def test_lower_bound_edges(ctx, lb):
    rng = np.random.default_rng(lb + 100)
    frames = rng.integers(0, 256, (1, 64, 96, 3), dtype=np.uint8)
    check(ctx, frames, rb.CAMP_BLUE, lb, what=f"lb={lb}")


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (5, 1), (2, 2), (3, 33), (17, 31), (37, 53), (64, 64), (100, 1279), (33, 16),
                                    (1080, 1440), (70, 1296), (9, 48), (130, 640)])
def test_random_noise_shapes(ctx, shape):
    H, W = shape
    rng = np.random.default_rng(H * 10007 + W)
    # dense noise: exercises every branch of the 3x3 close and the image borders
    frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    check(ctx, frames, rb.CAMP_BLUE, 60, what=f"{W}x{H}")
    # blobby noise near the threshold
    base = rng.integers(0, 2, (2, H, W, 1), dtype=np.uint8) * 120
    frames2 = np.concatenate([base + rng.integers(0, 20, (2, H, W, 1), dtype=np.uint8), base, np.zeros_like(base)], axis=3)
    check(ctx, frames2.astype(np.uint8), rb.CAMP_BLUE, 110, what=f"{W}x{H} blobby")


def test_pitched_rows_and_unaligned(ctx):
    rng = np.random.default_rng(7 + SEED_OFFSET)
    frames = rng.integers(0, 256, (2, 40, 320, 3), dtype=np.uint8)
    check(ctx, frames, rb.CAMP_BLUE, 70, pitch=320 * 3 + 64, what="pitch +64 (16-B aligned rows, bulk per-row copies)")
    check(ctx, frames, rb.CAMP_BLUE, 70, pitch=320 * 3 + 7, what="pitch +7 (unaligned rows, generic loader)")


def test_all_foreground_and_all_background(ctx):
    ones = np.zeros((1, 50, 200, 3), np.uint8); ones[..., 0] = 255
    check(ctx, ones, rb.CAMP_BLUE, 80, what="all fg")
    check(ctx, np.zeros((1, 50, 200, 3), np.uint8), rb.CAMP_BLUE, 80, what="all bg")


def test_batch_spanning_chunks():
    frames = np.stack([synth.make_frame(s, 320, 240, 3) for s in range(7)])
    with rb.Context(max_width=320, max_height=240, max_batch=8, chunk_frames=2) as c:
        mask, _ = run_mask(c, frames, rb.CAMP_BLUE, 80)
        for f in range(7):
            assert np.array_equal(mask[f], O.extract_color_mask(frames[f], rb.CAMP_BLUE, 80)), f"frame {f}"


@pytest.mark.parametrize("layout", [rb.BAYER_BG, rb.BAYER_GB, rb.BAYER_GR, rb.BAYER_RG])
@pytest.mark.parametrize("shape", [(1080, 1440), (37, 53), (64, 96), (3, 3), (4, 7)])
def test_bayer_front(ctx, layout, shape):
    H, W = shape
    if H >= 1000:
        img = synth.make_frame(5, W, H, 10)
    else:
        rng = np.random.default_rng(H * 31 + W + layout)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    raw = synth.bgr_to_bayer(img, layout)
    bgr = O.bayer_to_bgr(raw, layout)
    assert np.array_equal(bgr, R.bayer_bilinear_bgr(raw, layout)), "numpy restatement of the demosaic differs from cv2"
    for target, lb in ((rb.CAMP_BLUE, 80), (rb.CAMP_RED, 60), (rb.CAMP_GUIDELIGHT, 30)):
        ref = O.extract_color_mask(bgr, target, lb)
        d_in = ctx.device_buffer(raw.nbytes); d_out = ctx.device_buffer(H * W)
        try:
            d_in.upload(raw)
            ctx.bayer_extract_color_batch(d_in.ptr, W, H, 1, layout, target, lb, d_out.ptr)
            ctx.sync()
            mask = d_out.download((H, W))
        finally:
            d_in.free(); d_out.free()
        bad = np.argwhere(mask != ref)
        assert bad.size == 0, f"layout {layout} {W}x{H} target {target}: {len(bad)} differ, first {bad[0]}"


def test_invalid_arguments(ctx):
    d = ctx.device_buffer(1024)
    with pytest.raises(rb.RmcvError):
        ctx.extract_color_batch(d.ptr, 4000, 10, 1, 1, 80, d.ptr)  # above ctx maxima
    with pytest.raises(rb.RmcvError):
        ctx.extract_color_batch(d.ptr, 16, 16, 0, 1, 80, d.ptr)  # empty batch
    with pytest.raises(rb.RmcvError):
        ctx.extract_color_batch(None, 16, 16, 1, 1, 80, d.ptr)  # null frames
    with pytest.raises(rb.RmcvError):
        ctx.extract_color_batch(d.ptr, 16, 16, 1, 1, 80, d.ptr, pitch=10)  # pitch < row
    d.free()
