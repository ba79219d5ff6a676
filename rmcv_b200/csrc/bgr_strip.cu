// K1 (alternative, opt-in: RMCV_BGR_STRIP=1) — register-resident pixel stage of rm::extract_color on interleaved BGR
// frames (src/imgproc.cpp:52-69): cv::split + saturating channel difference + inRange + 3x3 MORPH_CLOSE -> {0,255} byte mask
// and the same mask bit-packed for the labelling stages.  HBM-bound streaming kernel, no tensor cores.
//
//   * a WARP owns a strip of 30 groups of 16 pixels (+ one halo lane each side) of one frame and walks down a segment of
//     rows.  The warp's 1536 bytes of a row travel global -> shared with three fully coalesced cp.async (LDGSTS)
//     instructions (lane l copies the 16-byte chunks l, l+32, l+64) into a 4-stage ring, three rows ahead of the
//     arithmetic; after cp.async.wait_group + __syncwarp a lane reads its own 48 bytes back with three conflict-free
//     LDS.128;
//   * per 4 pixels 6 dp4a form c_a - c_b - lower_bound and funnel shifts collect the sign bits into a 16-bit threshold
//     word; the close runs on a 20-bit window in registers (strip.cuh) and the byte mask leaves through the bits -> bytes
//     table with one 16-byte streaming store per lane and row.
//
// About 7 thread instructions per pixel (the shared-memory band kernel in pixel.cu: 15.5).
#include "strip.cuh"

namespace rmcv {

namespace {

constexpr int kRows = 3;                       // rows in flight per warp
constexpr int kRing = kRows + 1;               // ring stages: the stage read in the previous row is the one being refilled
constexpr uint32_t kRowBytes = 32u * 48u;      // one row of a warp's ring
constexpr int kWarps = 8;
constexpr size_t kSmemBytes = 4096 + (size_t)kWarps * kRing * kRowBytes;   // table + rings (dynamic shared memory)

struct BgrStripParams {
    const uint8_t* src; size_t frame_stride;
    uint8_t* mask; size_t mask_frame_stride;   // mask may be null
    int pitch, mask_pitch;                     // row pitches in bytes (< 2^31)
    uint16_t* bits16;                          // bit mask viewed as 16-bit words, [batch][H][WB2]
    int W, H, NC, WB2;                         // NC = W / 16 groups per row, WB2 = 16-bit words per bit row
    int seg, nseg, nwx;                        // rows per segment, segments per frame, warps per strip row
    int total_warps;
    uint32_t coef[6];                          // dp4a coefficient words (signed bytes) for the 4 pixels of a 12-byte group
    int acc0;                                  // -lower_bound (or the constants that force all-0 / all-1)
};

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

// threshold word of 16 interleaved BGR pixels (bit x = pixel x passes)
__device__ __forceinline__ uint32_t thr16(const uint4 A, const uint4 B, const uint4 C, const BgrStripParams& p) {
    const uint32_t w[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
    const uint32_t c0 = p.coef[0], c1a = p.coef[1], c1b = p.coef[2], c2a = p.coef[3], c2b = p.coef[4], c3 = p.coef[5];
    const int acc0 = p.acc0;
    uint32_t nb = 0;  // sign bits of (diff - lb), pixel 15 first so that pixel 0 lands in bit 0
#pragma unroll
    for (int grp = 3; grp >= 0; --grp) {
        const uint32_t w0 = w[3 * grp], w1 = w[3 * grp + 1], w2 = w[3 * grp + 2];
        const int v3 = dp4a_us(w2, c3, acc0);
        const int v2 = dp4a_us(w1, c2a, dp4a_us(w2, c2b, acc0));
        const int v1 = dp4a_us(w0, c1a, dp4a_us(w1, c1b, acc0));
        const int v0 = dp4a_us(w0, c0, acc0);
        nb = __funnelshift_l((uint32_t)v3, nb, 1);
        nb = __funnelshift_l((uint32_t)v2, nb, 1);
        nb = __funnelshift_l((uint32_t)v1, nb, 1);
        nb = __funnelshift_l((uint32_t)v0, nb, 1);
    }
    return ~nb & 0xffffu;
}

struct Lane {
    strip::CloseLane k;
    const uint8_t* gp;       // chunk 0 of this lane in raw row `lr` (clamped into the image); chunks 1, 2 at +512, +1024
    int lr;
    uint32_t cmask;          // which of the lane's three chunks lie inside the row
    uint32_t wring;          // shared address of the warp's ring + lane * 16 (copy side)
    uint32_t rring;          // shared address of the warp's ring + lane * 48 (read side)
    uint32_t stage, fill;    // byte offsets of the stage read next / refilled next
    int r;                   // image row of the threshold word that enters next
    int left;                // rows this lane still has to store (0 for halo / idle lanes)
    int tail16;
};

// Starts the copy of the next raw row of the warp's strip into stage `off` (one cp.async group, possibly empty).
__device__ __forceinline__ void fetch_row(Lane& a, const BgrStripParams& p, uint32_t off) {
    const uint8_t* g = a.gp;
    a.gp += (unsigned)a.lr < (unsigned)(p.H - 1) ? p.pitch : 0;   // rows outside the image re-read the nearest row inside
    ++a.lr;
    if (a.cmask & 1u) strip::cp_async16(a.wring + off, g);
    if (a.cmask & 2u) strip::cp_async16(a.wring + off + 512u, g + 512);
    if (a.cmask & 4u) strip::cp_async16(a.wring + off + 1024u, g + 1024);
    strip::cp_commit();
}

template <bool STORE, bool MASK>
__device__ __forceinline__ void one_row(Lane& a, const BgrStripParams& p) {
    strip::cp_wait<kRows - 1>();
    __syncwarp();             // every lane's chunks of this row have landed, and every lane is done with the previous stage
    const uint32_t s = a.rring + a.stage;
    const uint4 A = strip::lds128(s), B = strip::lds128(s + 16u), C = strip::lds128(s + 32u);
    fetch_row(a, p, a.fill);
    a.fill = a.stage;
    a.stage += kRowBytes;
    if (a.stage == kRing * kRowBytes) a.stage = 0u;
    uint32_t t = thr16(A, B, C, p);
    if ((unsigned)a.r >= (unsigned)p.H) t = 0u;
    strip::push_row<STORE, MASK>(a.k, p.WB2, p.mask_pitch, t, (unsigned)(a.r - 1) >= (unsigned)p.H, a.left > 0,
                                 a.left > 0 && a.tail16);
    if (STORE) --a.left;
    ++a.r;
}

}  // namespace

template <bool MASK, int MINB>
__global__ void __launch_bounds__(kWarps * 32, MINB) bgr_strip_kernel(const BgrStripParams p) {
    extern __shared__ __align__(128) uint8_t smem[];         // [4 KB table area][kWarps rings]
    strip::lut_init(smem, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * kWarps + warp;
    if (wid >= p.total_warps) return;
    const int wx = wid % p.nwx, rest = wid / p.nwx;
    const int sg = rest % p.nseg, frame = rest / p.nseg;
    const int c = wx * 30 - 1 + lane;                        // pixel group of this lane (lanes 0 and 31: halo only)
    const bool valid = c >= 0 && c < p.NC;
    const int cc = min(max(c, 0), p.NC - 1);
    const int y0 = sg * p.seg;
    Lane a;
    a.k.h0 = a.k.h1 = 0u; a.k.e0 = a.k.e1 = 0u;
    uint32_t inside = 0xfffffu;
    if (c == 0) inside &= 0xffffcu;
    if (c == p.NC - 1) inside &= 0x3ffffu;
    a.k.inside = valid ? inside : 0u;
    a.k.lut = strip::lut_base(smem);
    a.k.mrow = MASK ? p.mask + (size_t)frame * p.mask_frame_stride + (size_t)y0 * p.mask_pitch + (size_t)cc * 16 : nullptr;
    a.k.brow = p.bits16 + ((size_t)frame * p.H + y0) * p.WB2 + cc;
    const bool writer = valid && lane >= 1 && lane <= 30;
    a.left = writer ? min(p.seg, p.H - y0) : 0;
    a.tail16 = (c == p.NC - 1 && p.WB2 > p.NC) ? 1 : 0;      // W % 32 == 16: the upper half of the last bit word is zero
    a.r = y0 - 2;
    a.lr = a.r;
    // byte offset in the row of this lane's chunk j: (wx * 30 - 1) * 48 + (j * 32 + lane) * 16
    const int x0 = (wx * 30 - 1) * 48 + lane * 16, rowbytes = p.W * 3;
    a.cmask = (x0 >= 0 && x0 < rowbytes ? 1u : 0u) | (x0 + 512 >= 0 && x0 + 512 < rowbytes ? 2u : 0u) |
              (x0 + 1024 >= 0 && x0 + 1024 < rowbytes ? 4u : 0u);
    a.gp = p.src + (size_t)frame * p.frame_stride + (size_t)min(max(a.lr, 0), p.H - 1) * p.pitch + (ptrdiff_t)x0;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem + 4096) + warp * (kRing * kRowBytes);
    a.wring = ring + lane * 16u;
    a.rring = ring + lane * 48u;
    a.stage = 0u;
    a.fill = kRows * kRowBytes;
#pragma unroll
    for (int st = 0; st < kRows; ++st) fetch_row(a, p, st * kRowBytes);
    // rows y0-2 .. y0+1 fill the pipeline; every later row releases one final row
#pragma unroll
    for (int i = 0; i < 4; ++i) one_row<false, MASK>(a, p);
#pragma unroll 2
    for (int i = 0; i < p.seg; ++i) one_row<true, MASK>(a, p);
    strip::cp_wait<0>();
}

// Fast path of the BGR pixel stage; cudaErrorNotSupported when the call does not qualify (the caller then runs the
// shared-memory band kernel of pixel.cu).
cudaError_t launch_bgr_strip(const PixelLaunch& L, int sm_count, cudaStream_t st, int64_t* launches) {
    if ((L.W & 15) || L.W < 32 || L.H < 1) return cudaErrorNotSupported;
    if ((L.pitch & 15) || (L.frame_stride & 15) || (((size_t)L.src) & 15)) return cudaErrorNotSupported;
    if (L.mask && ((L.mask_pitch & 15) || (L.mask_frame_stride & 15) || (((size_t)L.mask) & 15))) return cudaErrorNotSupported;
    if (L.pitch > 0x7fffffffu || L.mask_pitch > 0x7fffffffu) return cudaErrorNotSupported;
    BgrStripParams p;
    memset(&p, 0, sizeof(p));
    p.src = L.src; p.pitch = (int)L.pitch; p.frame_stride = L.frame_stride;
    p.mask = L.mask; p.mask_pitch = (int)L.mask_pitch; p.mask_frame_stride = L.mask_frame_stride;
    p.bits16 = reinterpret_cast<uint16_t*>(L.bits);
    p.W = L.W; p.H = L.H; p.NC = L.W / 16; p.WB2 = 2 * ((L.W + 31) / 32);
    p.nwx = (p.NC + 29) / 30;
    {   // plus / minus channel (src/imgproc.cpp:56-65) as dp4a coefficient words: byte k of a 12-byte group belongs to pixel
        // k/3, channel k%3; order: c0 (px0,w0) c1a (px1,w0) c1b (px1,w1) c2a (px2,w1) c2b (px2,w2) c3 (px3,w2)
        int ca, cb;
        if (L.target == RMCV_CAMP_GUIDELIGHT) { ca = 1; cb = 2; }
        else if (L.target == RMCV_CAMP_BLUE) { ca = 0; cb = 2; }
        else { ca = 2; cb = 0; }
        auto put = [&](int slot, int px, int word) {
            uint32_t v = 0;
            for (int b = 0; b < 4; ++b) {
                const int k = word * 4 + b;
                if (k / 3 != px) continue;
                const int ch = k % 3;
                const int coef = (ch == ca ? 1 : 0) - (ch == cb ? 1 : 0);
                v |= (uint32_t)(uint8_t)(int8_t)coef << (8 * b);
            }
            p.coef[slot] = v;
        };
        put(0, 0, 0); put(1, 1, 0); put(2, 1, 1); put(3, 2, 1); put(4, 2, 2); put(5, 3, 2);
        // v = diff + acc0 >= 0  <=>  sat_u8(diff) in [lb, 255].  The two-word pixels add acc0 once (inner dp4a).
        if (L.lower_bound <= 0) { for (int i = 0; i < 6; ++i) p.coef[i] = 0; p.acc0 = 0; }
        else if (L.lower_bound > 255) { for (int i = 0; i < 6; ++i) p.coef[i] = 0; p.acc0 = -1; }
        else p.acc0 = -L.lower_bound;
    }
    const char* eb = getenv("RMCV_STRIP_MINB");
    const int minb = eb ? atoi(eb) : 3;
    // segment height: tall segments amortise the four halo rows, but the warps of a launch should fill whole waves of the
    // resident warp slots: the height with the least waves x (rows + halo)
    int seg = 0;
    const char* es = getenv("RMCV_STRIP_SEG");
    if (es && atoi(es) > 0) seg = atoi(es);
    else {
        const long long slots = (long long)(minb >= 4 ? 4 : 3) * kWarps * sm_count;
        long long best = -1;
        for (int sg = 16; sg <= 128; ++sg) {
            const long long warps = (long long)L.batch * ((L.H + sg - 1) / sg) * p.nwx;
            const long long cost = ((warps + slots - 1) / slots) * (sg + 6);
            if (best < 0 || cost < best) { best = cost; seg = sg; }
        }
    }
    if (seg > L.H) seg = L.H;
    if (seg < 1) seg = 1;
    p.seg = seg; p.nseg = (L.H + seg - 1) / seg;
    const long long total = (long long)L.batch * p.nseg * p.nwx;
    if (total <= 0 || total > 0x7fffffffLL) return cudaErrorNotSupported;
    p.total_warps = (int)total;
    const unsigned grid = (unsigned)((p.total_warps + kWarps - 1) / kWarps);
    cudaError_t e = cudaSuccess;
#define RMCV_BGR_LAUNCH(M_, B_)                                                                                         \
    do {                                                                                                                \
        e = cudaFuncSetAttribute(bgr_strip_kernel<M_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes); \
        if (e == cudaSuccess) bgr_strip_kernel<M_, B_><<<grid, kWarps * 32, kSmemBytes, st>>>(p);                        \
    } while (0)
    if (L.mask) { if (minb >= 4) RMCV_BGR_LAUNCH(true, 4); else RMCV_BGR_LAUNCH(true, 3); }
    else { if (minb >= 4) RMCV_BGR_LAUNCH(false, 4); else RMCV_BGR_LAUNCH(false, 3); }
#undef RMCV_BGR_LAUNCH
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
