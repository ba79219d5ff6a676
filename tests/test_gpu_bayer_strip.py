"""Row a0 fast path: the register-resident Bayer strip kernel (rmcv_b200/csrc/bayer_strip.cu) against the oracle
(cv2 bilinear demosaic = declared stand-in for DxRaw8toRGB24, hardware/src/daheng.cpp:136-151, followed by the
restatement of rm::extract_color's mask, src/imgproc.cpp:52-69) and against the generic shared-memory kernel.
Bit-exact on the byte mask and on the bit mask the labelling stages read."""
import os

import numpy as np
import pytest

import rmcv_b200 as rb
from rmcv_b200 import synth
from oracle import rm_oracle as O

pytestmark = pytest.mark.gpu
SEED_OFFSET = int(os.environ.get("RMCV_TEST_SEED", "0"))   # other random cases: RMCV_TEST_SEED=n pytest -m gpu ...


@pytest.fixture(scope="module")
def ctx():
    with rb.Context(max_width=2048, max_height=1200, max_batch=8) as c:
        yield c


def run(ctx, raw, layout, target, lb, pitch=None, generic=False):
    B, H, W = raw.shape
    pitch = pitch or W
    buf = np.zeros((B, H, pitch), np.uint8)
    buf[:, :, :W] = raw
    d_in = ctx.device_buffer(buf.nbytes); d_out = ctx.device_buffer(B * H * W)
    old = os.environ.pop("RMCV_BAYER_GENERIC", None)
    if generic:
        os.environ["RMCV_BAYER_GENERIC"] = "1"
    try:
        d_in.upload(buf)
        ctx.bayer_extract_color_batch(d_in.ptr, W, H, B, layout, target, lb, d_out.ptr, pitch=pitch)
        ctx.sync()
        mask = d_out.download((B, H, W))
        bits = [ctx.get_bitmask(f, W, H) for f in range(B)]
    finally:
        os.environ.pop("RMCV_BAYER_GENERIC", None)
        if old is not None:
            os.environ["RMCV_BAYER_GENERIC"] = old
        d_in.free(); d_out.free()
    return mask, bits


def check(ctx, raw, layout, target, lb, pitch=None, what=""):
    mask, bits = run(ctx, raw, layout, target, lb, pitch)
    B, H, W = raw.shape
    for f in range(B):
        ref = O.extract_color_mask(O.bayer_to_bgr(raw[f], layout), target, lb)
        bad = np.argwhere(mask[f] != ref)
        assert bad.size == 0, f"{what} layout {layout} {W}x{H} target {target} lb {lb} frame {f}: {len(bad)} differ, first {bad[0]}"
        packed = np.packbits(np.pad(ref > 0, ((0, 0), (0, (-W) % 32))), axis=1, bitorder="little").view(np.uint32)
        assert np.array_equal(bits[f], packed), f"{what} bit mask differs, frame {f}"


@pytest.mark.parametrize("layout", [rb.BAYER_BG, rb.BAYER_GB, rb.BAYER_GR, rb.BAYER_RG])
@pytest.mark.parametrize("shape", [(4, 32), (6, 48), (66, 480), (64, 496), (130, 1296), (34, 976), (200, 2048)])
def test_random_mosaics(ctx, layout, shape):
    """Uniform random raw bytes (every interpolation pattern near its threshold somewhere), widths that are exactly one
    warp strip (480), one group more (496), W % 32 == 16 (1296, 976), two groups (32); heights that cut segments."""
    H, W = shape
    rng = np.random.default_rng(H * 131 + W * 7 + layout)
    raw = rng.integers(0, 256, (2, H, W), dtype=np.uint8)
    for target, lb in ((rb.CAMP_BLUE, 80), (rb.CAMP_RED, 33), (rb.CAMP_BLUE, 1), (rb.CAMP_RED, 255)):
        check(ctx, raw, layout, target, lb, what="random")


@pytest.mark.parametrize("target", [rb.CAMP_BLUE, rb.CAMP_RED])
def test_threshold_edges_and_saturation(ctx, target):
    """Values clustered around the threshold (differences of -2..+2 about lb after rounding) and the degenerate bounds
    (lower_bound <= 0: everything passes; > 255: nothing does)."""
    rng = np.random.default_rng(11 + SEED_OFFSET)
    H, W = 48, 256
    base = rng.integers(0, 170, (1, H, W), dtype=np.int32)
    raw = base.copy()
    raw[:, 0::2, 0::2] += 80 + rng.integers(-3, 4, (1, H // 2, W // 2))   # one diagonal lifted by about lb
    raw[:, 1::2, 1::2] += rng.integers(-3, 4, (1, H // 2, W // 2))
    raw = np.clip(raw, 0, 255).astype(np.uint8)
    for layout in (rb.BAYER_BG, rb.BAYER_RG):
        for lb in (80, 79, 81, 0, -5, 256, 300):
            check(ctx, raw, layout, target, lb, what="edges")


def test_synthetic_frames_batch_and_pitch(ctx):
    """Config 2 material (1440x1080 mosaics of the synthetic generator), a batch, and a pitched source."""
    W, H = 1440, 1080
    raw = np.stack([synth.bgr_to_bayer(synth.make_frame(s, W, H, 10), synth.BAYER_BG) for s in (5, 6, 7)])
    check(ctx, raw, rb.BAYER_BG, rb.CAMP_BLUE, 80, what="synthetic")
    check(ctx, raw[:1], rb.BAYER_BG, rb.CAMP_BLUE, 80, pitch=W + 48, what="pitched")
    rawr = np.stack([synth.bgr_to_bayer(synth.make_frame(9, W, H, 12, blue=False), synth.BAYER_GB)])
    check(ctx, rawr, rb.BAYER_GB, rb.CAMP_RED, 80, what="synthetic red")


def test_strip_kernel_equals_generic_kernel(ctx):
    """Same inputs through the generic shared-memory Bayer kernel (RMCV_BAYER_GENERIC=1) and the strip kernel."""
    rng = np.random.default_rng(3 + SEED_OFFSET)
    raw = rng.integers(0, 256, (3, 128, 640), dtype=np.uint8)
    for layout in (rb.BAYER_BG, rb.BAYER_GR):
        a, ba = run(ctx, raw, layout, rb.CAMP_BLUE, 60)
        b, bb = run(ctx, raw, layout, rb.CAMP_BLUE, 60, generic=True)
        assert np.array_equal(a, b)
        for x, y in zip(ba, bb):
            assert np.array_equal(x, y)
