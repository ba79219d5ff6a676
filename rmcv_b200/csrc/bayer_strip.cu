// K1b — register-resident Bayer pixel stage (blue / red targets): the Bayer front (hardware/src/daheng.cpp:136-151,
// cv2 bilinear stand-in for DxRaw8toRGB24) fused with rm::extract_color's difference / threshold / 3x3 close
// (src/imgproc.cpp:52-69).  HBM-bound streaming kernel, no tensor cores, no shared-memory staging:
//
//   * pixel groups of 16 pixels are numbered through the whole call, slot = (frame, row segment, group), group fastest.
//     A WARP owns 30 consecutive slots (+ one halo lane each side) and every lane walks down its segment of rows;
//     neighbouring lanes are neighbouring groups of the same rows except across an image edge, where the close pads
//     anyway, so 30 of 32 lanes work whatever the image width.  A lane copies its 16 raw bytes of a row with cp.async
//     (LDGSTS; the warp's 512 contiguous bytes are one request) into its own slice of a 4-stage ring in shared memory,
//     four row pairs ahead of the arithmetic — no barrier, a lane only reads back what it asked for itself;
//   * only the two sampled planes matter (B and R; green is never read).  Samples are unpacked into 16-bit lanes, two
//     per register, the low lane for pixels 0..7 and the high lane for pixels 8..15 of the thread's group, so that
//     horizontal neighbours are whole registers and the final bits fall in order.  Per raw row: the samples U and
//     their horizontal pair sums H (+1).  Per output pixel the exact bilinear threshold test is one of
//         sampled P, quad M :  4 P - (H_up + H_down) + 3 - 4 lb >= 0                 (M = (sum + 2) >> 2)
//         quad P, sampled M :  (H_up + H_down) - 4 M - 4 lb     >= 0
//         pair P, pair M    :  a - (b & ~1) - 2 lb              >= 0                 (P = a >> 1, M = b >> 1)
//     evaluated on both lanes at once (IMAD / IADD3), biased by 2^(12 + x mod 4) so that the verdict of pixel x lands
//     on its own bit of the lane: three bit-selects gather four pixels, nine instructions sixteen;
//   * threshold words of the two neighbouring lanes arrive by warp shuffle; dilate and erode run on a 20-bit window in
//     registers (OpenCV's MORPH_CLOSE borders: dilate pads 0, erode pads 1), rows slide through registers;
//   * the byte mask leaves through a 256-entry bits -> 8 bytes table (the only shared memory, 2 KB) with one 16-byte
//     streaming store per lane and row; the bit mask for the labelling stages as one 16-bit store.
//
// Border rule of cv2's demosaic: row 0 shows row 1, row H-1 row H-2, column 0 column 1, column W-1 column
// W-2 — on threshold bits a replicate of the neighbouring interior bit.
#include "strip.cuh"

namespace rmcv {

namespace {

struct StripParams {
    const uint8_t* src; size_t frame_stride;
    uint8_t* mask; size_t mask_frame_stride;               // mask may be null
    int pitch, mask_pitch;                                 // row pitches in bytes (< 2^31)
    uint16_t* bits16;            // bit mask viewed as 16-bit words, [batch][H][WB2]
    int W, H, NC, WB2;           // NC = W / 16 pixel groups per row, WB2 = 16-bit words per bit row
    int seg, nseg;               // rows per segment (even), segments per frame
    int total_slots, total_warps;   // slot = (frame, segment, group), group fastest; a warp owns 30 consecutive slots
    uint32_t kSP[4], kSM[4], kN[4];   // per-lane constants of the three tests, by x mod 4 (both lanes)
    uint32_t force_or, force_and;     // lower_bound <= 0: all ones; > 255: all zeros
};

struct RowPrep {   // one raw row: samples U[j] = (x = 2j+s, x = 2j+8+s) and pair sums H[j] at the other x parity (+1)
    uint32_t U[4], H[4];
};

constexpr uint32_t kOne2 = 0x00010001u;

// Per-lane ring of raw rows in shared memory, filled by cp.async (LDGSTS): a lane only ever reads back the 16 bytes it
// asked for itself, so the groups need no barrier, only cp.async.wait_group.
constexpr int kStages = 4;                      // row pairs in flight per lane
constexpr uint32_t kStageBytes = 2u * 32u * 16u;   // two rows x 32 lanes x 16 B per warp and stage
using strip::bitsel;
using strip::cp_async16;
using strip::cp_commit;
using strip::cp_wait;
using strip::lds128;

// SX = x parity of the samples in this row
template <int SX>
__device__ __forceinline__ void prep_row(const uint4 w, RowPrep& r) {
    uint32_t e0, e1, e2, e3;
    if (SX == 0) { e0 = w.x & 0x00ff00ffu; e1 = w.y & 0x00ff00ffu; e2 = w.z & 0x00ff00ffu; e3 = w.w & 0x00ff00ffu; }
    else { e0 = __byte_perm(w.x, 0, 0x4341); e1 = __byte_perm(w.y, 0, 0x4341); e2 = __byte_perm(w.z, 0, 0x4341); e3 = __byte_perm(w.w, 0, 0x4341); }
    r.U[0] = __byte_perm(e0, e2, 0x5410);
    r.U[1] = __byte_perm(e0, e2, 0x7632);
    r.U[2] = __byte_perm(e1, e3, 0x5410);
    r.U[3] = __byte_perm(e1, e3, 0x7632);
    if (SX == 0) {   // H[j] at x = 2j+1: U(2j) + U(2j+2); the last one needs the right neighbour's first sample
        const uint32_t rn = __shfl_down_sync(0xffffffffu, r.U[0], 1);
        const uint32_t s3 = __funnelshift_r(r.U[0], rn, 16);    // (x = 8, x = 16)
        r.H[0] = r.U[0] + r.U[1] + kOne2;
        r.H[1] = r.U[1] + r.U[2] + kOne2;
        r.H[2] = r.U[2] + r.U[3] + kOne2;
        r.H[3] = r.U[3] + s3 + kOne2;
    } else {         // H[j] at x = 2j: U(2j-1) + U(2j+1); the first one needs the left neighbour's last sample
        const uint32_t ln = __shfl_up_sync(0xffffffffu, r.U[3], 1);
        const uint32_t s0 = __funnelshift_l(ln, r.U[3], 16);    // (x = -1, x = 7)
        r.H[0] = s0 + r.U[0] + kOne2;
        r.H[1] = r.U[0] + r.U[1] + kOne2;
        r.H[2] = r.U[1] + r.U[2] + kOne2;
        r.H[3] = r.U[2] + r.U[3] + kOne2;
    }
}

// Threshold word (16 pixels, bit x = pixel x of the group) of the row whose prep is C, rows above / below A and B.
// PROW: C holds samples of the plus channel.  SXC = x parity of the samples in row C.
template <bool PROW, int SXC>
__device__ __forceinline__ uint32_t thr_row(const RowPrep& A, const RowPrep& C, const RowPrep& B, const StripParams& p) {
    uint32_t sel[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xs = 2 * (j & 1) + SXC, xn = 2 * (j & 1) + (1 - SXC);   // x mod 4 of the sampled / the other site
        uint32_t ds, dn;
        if (PROW) {
            ds = C.U[j] * 4u + (p.kSP[xs] - A.H[j] - B.H[j]);
            dn = C.H[j] + p.kN[xn] - ((A.U[j] + B.U[j] + kOne2) & 0xfffefffeu);
        } else {
            ds = (A.H[j] + B.H[j] + p.kSM[xs]) - C.U[j] * 4u;
            dn = (A.U[j] + B.U[j] + kOne2) + p.kN[xn] - (C.H[j] & 0xfffefffeu);
        }
        sel[j] = bitsel(ds, dn, kOne2 << (12 + xs));
    }
    const uint32_t lo = bitsel(sel[0], sel[1], 0x30003000u);   // lanes: bits 12..15 = pixels 0..3 | 8..11
    const uint32_t hi = bitsel(sel[2], sel[3], 0x30003000u);   //        bits 12..15 = pixels 4..7 | 12..15
    const uint32_t z = bitsel(lo >> 4, hi, 0x0f000f00u);       // byte 1 = pixels 0..7, byte 3 = pixels 8..15
    return (__byte_perm(z, 0, 0x4431) | p.force_or) & p.force_and;
}

// Per-lane state of the walk down a row segment.
struct Walk {
    strip::CloseLane k;
    uint32_t fix_mask, fix_rot;  // border columns: bit 0 of group 0 shows bit 1 (rotate right 1), bit 15 of the last group bit 14
    const uint8_t* lp;           // raw row `lr` (clamped into the image) at the group of this lane
    int lr;
    uint32_t ring, stage;        // shared address of this lane's slot in stage 0; byte offset of the oldest stage
    int r;                       // image row (even) of the pair that enters next
    int left;                    // rows this lane still has to store (0 for halo / idle lanes)
    int tail16;                  // also zero the odd 16-bit word that ends a W % 32 == 16 bit row
};

// Address of the next raw row (rows outside the image read the nearest row inside; their results are never used).
__device__ __forceinline__ const uint8_t* next_row(Walk& k, const StripParams& p) {
    const uint8_t* v = k.lp;
    k.lp += (unsigned)k.lr < (unsigned)(p.H - 1) ? p.pitch : 0;
    ++k.lr;
    return v;
}
// Starts the copy of the next two raw rows into stage `off` of the lane's ring (one cp.async group).
__device__ __forceinline__ void fetch_pair(Walk& k, const StripParams& p, uint32_t off) {
    cp_async16(k.ring + off, next_row(k, p));
    cp_async16(k.ring + off + 512u, next_row(k, p));
    cp_commit();
}

// Row pair (r, r+1), r even.  On entry A = prep(r-1), C = prep(r) and the oldest stage of the ring holds the raw rows
// r+1, r+2.  On exit B = prep(r+1), N = prep(r+2) are the (A, C) of the next pair and the stage is being refilled with the
// rows r+1+2*kStages, r+2+2*kStages.
template <int PY, int PX, bool STORE, bool MASK>
__device__ __forceinline__ void row_pair(Walk& k, const StripParams& p, const RowPrep& A, const RowPrep& C,
                                         RowPrep& B, RowPrep& N) {
    constexpr bool kEvenIsP = PY == 0;
    constexpr int kSxEven = kEvenIsP ? PX : 1 - PX, kSxOdd = 1 - kSxEven;   // x parity of the samples in even / odd rows
    cp_wait<kStages - 1>();
    prep_row<kSxOdd>(lds128(k.ring + k.stage), B);
    prep_row<kSxEven>(lds128(k.ring + k.stage + 512u), N);
    fetch_pair(k, p, k.stage);
    k.stage = (k.stage + kStageBytes) & (kStages * kStageBytes - 1u);
    const int r = k.r;
    uint32_t t0 = thr_row<kEvenIsP, kSxEven>(A, C, B, p);
    uint32_t t1 = thr_row<!kEvenIsP, kSxOdd>(C, B, N, p);
    if (r == 0) t0 = t1;                                 // row 0 shows row 1
    if (r == p.H - 2) t1 = t0;                           // row H-1 shows row H-2
    t0 = bitsel(__funnelshift_r(t0, t0, k.fix_rot), t0, k.fix_mask);   // column 0 shows column 1, column W-1 column W-2
    t1 = bitsel(__funnelshift_r(t1, t1, k.fix_rot), t1, k.fix_mask);
    if ((unsigned)r >= (unsigned)p.H) { t0 = 0u; t1 = 0u; }   // a pair is either inside or outside the image (r, H even)
    strip::push_row<STORE, MASK>(k.k, p.WB2, p.mask_pitch, t0, (unsigned)(r - 1) >= (unsigned)p.H, k.left > 0,
                                 k.left > 0 && k.tail16);
    strip::push_row<STORE, MASK>(k.k, p.WB2, p.mask_pitch, t1, (unsigned)r >= (unsigned)p.H, k.left > 1,
                                 k.left > 1 && k.tail16);
    if (STORE) k.left -= 2;
    k.r = r + 2;
}

}  // namespace

// PY, PX: parity of the rows / columns that sample the plus channel (the minus channel sits on the opposite diagonal)
template <int PY, int PX, bool MASK, int MINB>
__global__ void __launch_bounds__(256, MINB) bayer_strip_kernel(const StripParams p) {
    __shared__ __align__(16) uint8_t s_lut_raw[4096];
    __shared__ __align__(16) uint8_t s_ring[8 * kStages * kStageBytes];   // 8 warps
    strip::lut_init(s_lut_raw, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= p.total_warps) return;
    const int slot = wid * 30 - 1 + lane;                    // lanes 0 and 31 only feed their neighbours
    const bool valid = slot >= 0 && slot < p.total_slots;
    const int sc = min(max(slot, 0), p.total_slots - 1);
    const int c = sc % p.NC, q = sc / p.NC;
    const int sg = q % p.nseg, frame = q / p.nseg;
    const int H = p.H, NC = p.NC;
    const int y0 = sg * p.seg;
    Walk k;
    k.k.h0 = k.k.h1 = 0u; k.k.e0 = k.k.e1 = 0u;
    uint32_t inside = 0xfffffu;
    if (c == 0) inside &= 0xffffcu;
    if (c == NC - 1) inside &= 0x3ffffu;
    k.k.inside = valid ? inside : 0u;
    k.fix_mask = c == 0 ? 1u : (c == NC - 1 ? 0x8000u : 0u);
    k.fix_rot = c == 0 ? 1u : 31u;
    const bool writer = valid && lane >= 1 && lane <= 30;
    k.left = writer ? min(p.seg, H - y0) : 0;
    k.tail16 = (c == NC - 1 && p.WB2 > NC) ? 1 : 0;          // W % 32 == 16: the upper half of the last bit word is zero
    k.k.lut = strip::lut_base(s_lut_raw);
    k.k.mrow = MASK ? p.mask + (size_t)frame * p.mask_frame_stride + (size_t)y0 * p.mask_pitch + (size_t)c * 16 : nullptr;
    k.k.brow = p.bits16 + ((size_t)frame * H + y0) * p.WB2 + c;

    constexpr int kSxEven = PY == 0 ? PX : 1 - PX, kSxOdd = 1 - kSxEven;
    RowPrep s0, s1, s2, s3;
    k.r = y0 - 2;
    k.lr = k.r - 1;
    k.lp = p.src + (size_t)frame * p.frame_stride + (size_t)min(max(k.lr, 0), H - 1) * p.pitch + (size_t)c * 16;
    k.ring = (uint32_t)__cvta_generic_to_shared(s_ring) + (threadIdx.x >> 5) * (kStages * kStageBytes) + lane * 16u;
    k.stage = 0u;
    {
        const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(next_row(k, p)));
        const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(next_row(k, p)));
#pragma unroll
        for (int st = 0; st < kStages; ++st) fetch_pair(k, p, st * kStageBytes);   // rows r+1 .. r+2*kStages
        prep_row<kSxOdd>(w0, s0);
        prep_row<kSxEven>(w1, s1);
    }
    // rows y0-2 .. y0+1 fill the pipeline; from the pair at y0+2 on, every pair releases two final rows
    row_pair<PY, PX, false, MASK>(k, p, s0, s1, s2, s3);
    row_pair<PY, PX, false, MASK>(k, p, s2, s3, s0, s1);
    for (int n = p.seg >> 1; n > 0; n -= 2) {               // seg / 2 storing pairs
        row_pair<PY, PX, true, MASK>(k, p, s0, s1, s2, s3);
        if (n == 1) break;
        row_pair<PY, PX, true, MASK>(k, p, s2, s3, s0, s1);
    }
    cp_wait<0>();
}

// Fast path of the Bayer pixel stage; returns cudaErrorNotSupported when the call does not qualify (the caller then
// runs the generic shared-memory kernel).
cudaError_t launch_bayer_strip(const PixelLaunch& L, int sm_count, cudaStream_t st, int64_t* launches) {
    if (L.target != RMCV_CAMP_BLUE && L.target != RMCV_CAMP_RED) return cudaErrorNotSupported;   // green needs the quincunx forms
    if ((L.W & 15) || (L.H & 1) || L.W < 32 || L.H < 4) return cudaErrorNotSupported;
    if ((L.pitch & 15) || (L.frame_stride & 15) || (((size_t)L.src) & 15)) return cudaErrorNotSupported;
    if (L.mask && ((L.mask_pitch & 15) || (L.mask_frame_stride & 15) || (((size_t)L.mask) & 15))) return cudaErrorNotSupported;
    int ch[2][2];   // colour sampled at (y&1, x&1): 0 = B, 1 = G, 2 = R
    switch (L.bayer_layout) {
        case RMCV_BAYER_BG: ch[0][0] = 0; ch[0][1] = 1; ch[1][0] = 1; ch[1][1] = 2; break;
        case RMCV_BAYER_GB: ch[0][0] = 1; ch[0][1] = 0; ch[1][0] = 2; ch[1][1] = 1; break;
        case RMCV_BAYER_GR: ch[0][0] = 1; ch[0][1] = 2; ch[1][0] = 0; ch[1][1] = 1; break;
        case RMCV_BAYER_RG: ch[0][0] = 2; ch[0][1] = 1; ch[1][0] = 1; ch[1][1] = 0; break;
        default: return cudaErrorInvalidValue;
    }
    const int plus = L.target == RMCV_CAMP_BLUE ? 0 : 2;   // src/imgproc.cpp:56-65: blue = B - R, otherwise R - B
    int py = 0, px = 0;
    for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) if (ch[y][x] == plus) { py = y; px = x; }

    StripParams p;
    memset(&p, 0, sizeof(p));
    if (L.pitch > 0x7fffffffu || L.mask_pitch > 0x7fffffffu) return cudaErrorNotSupported;
    p.src = L.src; p.pitch = (int)L.pitch; p.frame_stride = L.frame_stride;
    p.mask = L.mask; p.mask_pitch = (int)L.mask_pitch; p.mask_frame_stride = L.mask_frame_stride;
    p.bits16 = reinterpret_cast<uint16_t*>(L.bits);
    p.W = L.W; p.H = L.H; p.NC = L.W / 16; p.WB2 = 2 * ((L.W + 31) / 32);
    // segment height: tall segments amortise the six halo rows, but the warps of a launch should fill whole waves of the
    // resident warp slots (3 CTAs of 8 warps per SM): pick the even height with the least waves x (rows + halo)
    int seg = 0;
    const int eb0 = tuning().strip_minb > 0 ? tuning().strip_minb : 3;
    if (tuning().strip_seg > 0) seg = tuning().strip_seg & ~1;
    else {
        const long long slots = (eb0 >= 4 ? 32LL : 24LL) * sm_count;
        long long best = -1;
        for (int sg = 16; sg <= 128; sg += 2) {
            const long long warps = ((long long)L.batch * ((L.H + sg - 1) / sg) * p.NC + 29) / 30;
            const long long cost = ((warps + slots - 1) / slots) * (sg + 8);
            if (best < 0 || cost < best) { best = cost; seg = sg; }
        }
        // With enough work for two waves and more, short segments win although they re-read more halo rows: the CTAs of a
        // multi-wave launch overlap each other's ramp-up and drain, and the labelling kernels of a detect call get SM slots
        // sooner (cold sweep on 64 x 1440x1080: 24-row segments 4.21 TB/s, the one-wave choice of ~60 rows 3.99 TB/s).
        const long long warps24 = ((long long)L.batch * ((L.H + 23) / 24) * p.NC + 29) / 30;
        if (warps24 >= 2 * slots) seg = 24;
    }
    if (seg > L.H) seg = L.H;
    if (seg < 2) seg = 2;
    p.seg = seg; p.nseg = (L.H + seg - 1) / seg;
    const long long total_slots = (long long)L.batch * p.nseg * p.NC;
    if (total_slots <= 0 || total_slots > 0x3fffffffLL) return cudaErrorNotSupported;
    p.total_slots = (int)total_slots;
    p.total_warps = (int)((total_slots + 29) / 30);
    const long long total = p.total_warps;
    const int lb = L.lower_bound < 1 ? 1 : (L.lower_bound > 255 ? 255 : L.lower_bound);
    for (int x = 0; x < 4; ++x) {
        const uint32_t bias = 1u << (12 + x);
        p.kSP[x] = (bias + 3u - 4u * (uint32_t)lb) * kOne2;
        p.kSM[x] = (bias - 4u * (uint32_t)lb) * kOne2;
        p.kN[x] = (bias - 2u * (uint32_t)lb) * kOne2;
    }
    p.force_or = L.lower_bound <= 0 ? 0xffffu : 0u;
    p.force_and = L.lower_bound > 255 ? 0u : 0xffffu;
    const int wpb = 8;
    const unsigned grid = (unsigned)((total + wpb - 1) / wpb);
    const int which = (py * 2 + px) * 2 + (L.mask ? 1 : 0);
    const int minb = eb0;
#define RMCV_STRIP_CASE(n, PY_, PX_, M_)                                                                     \
    case n:                                                                                                  \
        if (minb >= 4) bayer_strip_kernel<PY_, PX_, M_, 4><<<grid, wpb * 32, 0, st>>>(p);                    \
        else bayer_strip_kernel<PY_, PX_, M_, 3><<<grid, wpb * 32, 0, st>>>(p);                              \
        break;
    switch (which) {
        RMCV_STRIP_CASE(0, 0, 0, false) RMCV_STRIP_CASE(1, 0, 0, true) RMCV_STRIP_CASE(2, 0, 1, false) RMCV_STRIP_CASE(3, 0, 1, true)
        RMCV_STRIP_CASE(4, 1, 0, false) RMCV_STRIP_CASE(5, 1, 0, true) RMCV_STRIP_CASE(6, 1, 1, false) RMCV_STRIP_CASE(7, 1, 1, true)
        default: return cudaErrorInvalidValue;
    }
#undef RMCV_STRIP_CASE
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
