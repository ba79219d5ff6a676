"""Row a1, alternative pixel kernel (opt-in, RMCV_BGR_STRIP=1): the TMA-fed, register-resident BGR band-strip kernel
(rmcv_b200/csrc/bgr_bandstrip.cu) against the oracle's restatement
of rm::extract_color's mask (src/imgproc.cpp:52-69: split, saturating difference, inRange, 3x3 MORPH_CLOSE) and against
the shared-memory band kernel.  Bit-exact on the byte mask and on the bit mask the labelling stages read."""
import os

import numpy as np
import pytest

import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O

pytestmark = pytest.mark.gpu
VARIANT = "1"


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=2048, max_height=1200, max_batch=8) as c:
        yield c


def run(ctx, frames, target, lb, pitch=None, band=False, env=None):
    B, H, W, _ = frames.shape
    pitch = pitch or W * 3
    buf = np.zeros((B, H, pitch), np.uint8)
    buf[:, :, :W * 3] = frames.reshape(B, H, W * 3)
    d_in = ctx.device_buffer(buf.nbytes); d_out = ctx.device_buffer(B * H * W)
    saved = {k: os.environ.get(k) for k in ("RMCV_BGR_STRIP", "RMCV_BANDSTRIP_RC")}
    if not band:
        os.environ["RMCV_BGR_STRIP"] = VARIANT
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        d_in.upload(buf)
        ctx.extract_color_batch(d_in.ptr, W, H, B, target, lb, d_out.ptr, pitch=pitch)
        ctx.sync()
        mask = d_out.download((B, H, W))
        bits = [ctx.get_bitmask(f, W, H) for f in range(B)]
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
        d_in.free(); d_out.free()
    return mask, bits


def check(ctx, frames, target, lb, pitch=None, what="", env=None):
    mask, bits = run(ctx, frames, target, lb, pitch, env=env)
    B, H, W, _ = frames.shape
    for f in range(B):
        ref = O.extract_color_mask(frames[f], target, lb)
        bad = np.argwhere(mask[f] != ref)
        assert bad.size == 0, f"{what} {W}x{H} target {target} lb {lb} frame {f}: {len(bad)} differ, first {bad[0]}"
        packed = np.packbits(np.pad(ref > 0, ((0, 0), (0, (-W) % 32))), axis=1, bitorder="little").view(np.uint32)
        assert np.array_equal(bits[f], packed), f"{what} bit mask differs, frame {f}"


@pytest.mark.parametrize("shape", [(1, 32), (2, 48), (3, 32), (5, 480), (37, 496), (64, 1296), (130, 976), (200, 2048), (21, 1280)])
def test_random_frames(ctx, shape):
    """Uniform random bytes; widths of exactly one warp strip (480), one group more (496), W % 32 == 16 (1296, 976), two
    groups (32); heights of 1..5 rows and heights that cut bands; batches that do not fill a CTA's frame group; all three
    targets."""
    H, W = shape
    rng = np.random.default_rng(H * 131 + W * 7)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    for target, lb in ((rb.CAMP_BLUE, 80), (rb.CAMP_RED, 33), (rb.CAMP_GUIDELIGHT, 10), (rb.CAMP_BLUE, 1), (rb.CAMP_RED, 255)):
        check(ctx, frames, target, lb, what="random")


def test_rows_per_stage_and_band_edges(ctx):
    """Every ring-stage height (rows per TMA chunk) cuts the 32-row bands differently; heights around the band size."""
    rng = np.random.default_rng(5)
    for H in (31, 32, 33, 63, 65):
        frames = rng.integers(0, 256, (4, H, 160, 3), dtype=np.uint8)
        for rc in (1, 2, 3, 5):
            check(ctx, frames, rb.CAMP_BLUE, 60, what=f"H {H} rc {rc}", env={"RMCV_BANDSTRIP_RC": str(rc)})


def test_degenerate_bounds_and_extremes(ctx):
    rng = np.random.default_rng(6)
    frames = rng.integers(0, 256, (1, 40, 320, 3), dtype=np.uint8)
    for lb in (0, -5, 256, 300):
        check(ctx, frames, rb.CAMP_BLUE, lb, what="bounds")
    ones = np.zeros((1, 50, 208, 3), np.uint8); ones[..., 0] = 255
    check(ctx, ones, rb.CAMP_BLUE, 80, what="all foreground")
    check(ctx, np.zeros((1, 50, 208, 3), np.uint8), rb.CAMP_BLUE, 80, what="all background")
    dots = np.zeros((1, 33, 64, 3), np.uint8)
    dots[0, ::2, ::2, 0] = 255       # isolated pixels: the close must leave them alone
    dots[0, 0, :, 0] = 255; dots[0, -1, :, 0] = 255; dots[0, :, 0, 0] = 255; dots[0, :, -1, 0] = 255   # borders
    check(ctx, dots, rb.CAMP_BLUE, 80, what="dots + border lines")


def test_synthetic_frames_batch_and_pitch(ctx):
    frames = np.stack([synth.make_frame(s, 1280, 1024, 10, blue=(s % 2 == 0)) for s in (20, 21, 22)])
    check(ctx, frames[::2], rb.CAMP_BLUE, 80, what="synthetic blue")
    check(ctx, frames[1:2], rb.CAMP_RED, 80, what="synthetic red")
    check(ctx, frames[:1], rb.CAMP_BLUE, 80, pitch=1280 * 3 + 64, what="pitched")


def test_strip_kernel_equals_band_kernel(ctx):
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (3, 128, 640, 3), dtype=np.uint8)
    a, ba = run(ctx, frames, rb.CAMP_RED, 60)
    b, bb = run(ctx, frames, rb.CAMP_RED, 60, band=True)
    assert np.array_equal(a, b)
    for x, y in zip(ba, bb):
        assert np.array_equal(x, y)
