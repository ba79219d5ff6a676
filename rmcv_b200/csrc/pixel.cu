// K1 — fused pixel stage of rm::extract_color (reference: src/imgproc.cpp:52-69):
//   cv::split + saturating channel difference + inRange + 3x3 MORPH_CLOSE  ->  {0,255} byte mask
//   (+ the same mask bit-packed, 1 bit/px, for the labelling stages).
// and its Bayer front (reference: hardware/src/daheng.cpp:136-151, DxRaw8toRGB24 stand-in).
//
// HBM-bound streaming kernel, no tensor cores (nothing here is a contraction):
//   * one CTA = one band of BH output rows of one frame, full image width (no x halo);
//   * raw rows travel global -> shared with 1-D TMA bulk copies (cp.async.bulk, UBLKCP in SASS) through a
//     S-stage ring guarded by mbarriers; a row band of a continuous frame is one contiguous copy;
//   * per 16-pixel group a thread reads 48 B (3 conflict-free LDS.128), forms B-R-lb (or any channel pair)
//     with 6 dp4a per 4 pixels and shifts the sign bits into a 16-bit threshold word;
//   * dilate/erode run on the bit rows in shared memory (3 OR / 3 AND of shifted words), border rules of
//     OpenCV's MORPH_CLOSE (dilate pads 0, erode pads 1);
//   * every input byte is read from HBM once (band halo rows are re-read through L2), every mask byte is
//     written once with 16-byte stores.
#include "common.cuh"
#include "emit_core.cuh"
#include "pairs.cuh"

namespace rmcv {

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16-B aligned).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
// 4 mask bits -> 4 bytes of 0x00/0xFF: spread the bits to the byte MSBs, then PRMT sign-replicate.
__device__ __forceinline__ uint32_t expand4(uint32_t nib) {
    uint32_t x = nib * 0x10204080u;
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

// ------------------------------------------------------------------------------------------ parameters
struct PixelParams {
    const uint8_t* src; size_t pitch, frame_stride;
    uint8_t* mask; size_t mask_pitch, mask_frame_stride;
    uint32_t* bits;
    int W, H, WB;
    int BH, bands;       // output rows per band, bands per frame
    int RC, S;           // rows per TMA chunk, ring stages
    int srow;            // shared-memory row stride in bytes
    int gpr;             // 16-pixel groups per row (BGR) / per row (Bayer)
    int halo;            // threshold-row halo of the close: 2
    uint32_t inv_wb, inv_gpr16;  // floor(2^32 / n) + 1 for n = WB / ceil(W/16): slot / n == umulhi(slot, inv); 0 when n == 1
    uint32_t coef[6];    // dp4a coefficient words (signed bytes) for the 4 pixels of a 12-byte group
    int acc0;            // -lower_bound (or the constants that force all-0 / all-1)
    int contiguous;      // pitch == row bytes: a chunk is one bulk copy
    int mask_vec;        // mask rows allow 16-byte stores
    uint32_t last_valid; // valid bits of the last word of a row
    // Bayer only
    int bayer;           // 0 = BGR
    int px, py;          // parity (x&1, y&1) of the site that samples channel `plus`... see kernel
    int plus_is_site;    // layout helpers, see bayer kernel
    int lb;
    uint8_t* band_flags; // one byte per (frame, band): 1 = the band's own rows hold foreground (null: not wanted)
    EmitParams em;       // fused pixel+emit kernel only: where the band's runs and boundary-pixel records go
};

struct Iter2D {  // walks idx = tid, tid+NT, ... over a [rows][cols] grid without divisions in the loop
    int r, c, dr, dc, cols;
    __device__ __forceinline__ Iter2D(int tid, int nt, int cols_) : cols(cols_) {
        r = tid / cols_; c = tid - r * cols_;
        dr = nt / cols_; dc = nt - dr * cols_;
    }
    __device__ __forceinline__ void next() {
        r += dr; c += dc;
        if (c >= cols) { c -= cols; ++r; }
    }
};

// Shared-memory carve-up (bytes): [S stages][S mbarriers (padded to 16 B)][t rows][d rows]
__host__ __device__ inline size_t pix_smem_bytes(int S, int RC, int srow, int BH, int WB, int halo_rows) {
    size_t TW = (size_t)WB + 2;
    size_t stage = (size_t)S * RC * srow;
    size_t bars = ((size_t)S * 8 + 15) & ~(size_t)15;
    size_t t = (size_t)(BH + 2 * halo_rows) * TW * 4;
    size_t d = (size_t)(BH + 2 * halo_rows - 2) * TW * 4;
    d += 16;
    return stage + bars + t + d;
}

// Compile-time geometry of a launch: kW = 0 reads everything from PixelParams; kW > 0 fixes the width, the band height,
// the rows per TMA chunk, the ring depth and the CTA size, so that the index arithmetic of a band (divisions by the word
// count, per-segment row counts, the 2-D walk) folds into constants.  Every CTA only sees 128 pixels per thread, which
// makes that set-up arithmetic a visible share of the kernel.
template <int kW_, int kBH_, int kRC_, int kS_, int kNT_>
struct PixGeom {
    static constexpr int kW = kW_, kBH = kBH_, kRC = kRC_, kS = kS_, kNT = kNT_;
};
using GeomRuntime = PixGeom<0, 0, 0, 0, 0>;
using Geom1280 = PixGeom<1280, 32, 4, 4, 320>;

// ------------------------------------------------------------------------------------------ morphology + stores
// t: (nout+4) x TW threshold words (row 0 <-> image row y0-2), zero outside the image.
// Writes the final mask of rows [y0, y0+nout) to global as bytes and as bit words (the labelling stages read the bits).
// Both passes are separable (3x1 on the bit row, then 1x3 down the rows) and a thread walks DOWN one word column over a
// segment of rows, so every word is loaded once and the row-combined value slides through three registers.
template <class G = GeomRuntime>
__device__ __forceinline__ void close_and_store(const PixelParams& p, uint32_t* t, uint32_t* d, int frame, int y0,
                                                int nout, int tid, int NT_) {
    constexpr bool kFixed = G::kW > 0;
    const int W = kFixed ? G::kW : p.W;
    const int NT = kFixed ? G::kNT : NT_;
    const int WB = kFixed ? (G::kW + 31) / 32 : p.WB, TW = WB + 2, H = p.H;
    const uint32_t valid = kFixed ? ((G::kW & 31) ? ((1u << (G::kW & 31)) - 1u) : 0xFFFFFFFFu) : p.last_valid;
    const int nseg = NT >= WB ? NT / WB : 1;      // row segments per word column
    auto col_of = [&](int slot, int* seg) -> int {  // slot -> (segment, word column); one umulhi instead of a division
        const int sg = kFixed ? slot / WB : (p.inv_wb ? (int)__umulhi((uint32_t)slot, p.inv_wb) : slot);
        *seg = sg;
        return slot - sg * WB;
    };
    // ---- dilate: rows y0-1 .. y0+nout (outside the image: all ones so that the erode ignores them)
    {
        const int nd = nout + 2, per = (nd + nseg - 1) / nseg;
        for (int slot = tid; slot < WB * nseg; slot += NT) {
            int seg;
            const int k = col_of(slot, &seg);
            const int i0 = seg * per, i1 = min(nd, i0 + per);
            if (i0 >= i1) continue;
            const bool last = k == WB - 1, prelast = k + 1 == WB - 1;
            auto hrow = [&](int ti) -> uint32_t {   // 3x1 OR of threshold row ti at word k (padded column k+1)
                const uint32_t* a = t + (size_t)ti * TW + k;
                uint32_t left = a[0], mid = a[1], right = a[2];
                if (last) mid &= valid;
                if (prelast) right &= valid;
                return mid | (mid << 1) | (left >> 31) | (mid >> 1) | (right << 31);
            };
            uint32_t h0 = hrow(i0), h1 = hrow(i0 + 1);
            uint32_t* dp = d + (size_t)i0 * TW + k + 1;
            for (int i = i0; i < i1; ++i, dp += TW) {
                const uint32_t h2 = hrow(i + 2);
                const int y = y0 - 1 + i;
                uint32_t dv = h0 | h1 | h2;
                if (last) dv |= ~valid;
                if (y < 0 || y >= H) dv = 0xFFFFFFFFu;
                dp[0] = dv;
                if (k == 0) dp[-1] = 0xFFFFFFFFu;
                if (last) dp[1] = 0xFFFFFFFFu;
                h0 = h1; h1 = h2;
            }
        }
    }
    __syncthreads();
    // ---- erode: rows y0 .. y0+nout-1; result into m (compact [nout][WB]) and to the global bit mask
    uint32_t* m = t;   // t and m alias: every thread reads d only and writes m; t was last read before the barrier
    uint32_t any_fg = 0u;
    {
        uint32_t* gbits = p.bits + ((size_t)frame * H + y0) * WB;
        const int per = (nout + nseg - 1) / nseg;
        for (int slot = tid; slot < WB * nseg; slot += NT) {
            int seg;
            const int k = col_of(slot, &seg);
            const int j0 = seg * per, j1 = min(nout, j0 + per);
            if (j0 >= j1) continue;
            const bool last = k == WB - 1;
            auto hrow = [&](int di) -> uint32_t {   // 3x1 AND of dilated row di at word k
                const uint32_t* a = d + (size_t)di * TW + k;
                const uint32_t left = a[0], mid = a[1], right = a[2];
                return mid & ((mid << 1) | (left >> 31)) & ((mid >> 1) | (right << 31));
            };
            uint32_t h0 = hrow(j0), h1 = hrow(j0 + 1);
            uint32_t* mp = m + (size_t)j0 * WB + k;
            uint32_t* gp = gbits + (size_t)j0 * WB + k;
            for (int j = j0; j < j1; ++j, mp += WB, gp += WB) {
                const uint32_t h2 = hrow(j + 2);
                uint32_t mv = h0 & h1 & h2;
                if (last) mv &= valid;
                mp[0] = mv;
                gp[0] = mv;
                any_fg |= mv;
                h0 = h1; h1 = h2;
            }
        }
    }
    {   // barrier + "does this band hold any foreground": lets the emit kernel skip empty bands without loading them
        const int any = __syncthreads_or(any_fg != 0u);
        if (tid == 0 && p.band_flags != nullptr) p.band_flags[(size_t)frame * p.bands + y0 / p.BH] = any ? 1 : 0;
    }
    // ---- byte mask: one 16-byte store per 16 pixels, a thread walks down one 16-pixel column
    if (p.mask != nullptr) {
        const uint16_t* m16 = reinterpret_cast<const uint16_t*>(m);
        const int gpr16 = (W + 15) >> 4;
        const int nsg = NT >= gpr16 ? NT / gpr16 : 1;
        uint8_t* gmask = p.mask + (size_t)frame * p.mask_frame_stride + (size_t)y0 * p.mask_pitch;
        for (int slot = tid; slot < gpr16 * nsg; slot += NT) {
            const int sg = kFixed ? slot / gpr16 : (p.inv_gpr16 ? (int)__umulhi((uint32_t)slot, p.inv_gpr16) : slot);
            const int g = slot - sg * gpr16;
            const bool vec = p.mask_vec && g * 16 + 16 <= W;
            uint8_t* dst = gmask + (size_t)sg * p.mask_pitch + (size_t)g * 16;
            const size_t dstep = (size_t)nsg * p.mask_pitch;
            const uint16_t* src = m16 + (size_t)sg * WB * 2 + g;
            for (int j = sg; j < nout; j += nsg, dst += dstep, src += (size_t)nsg * WB * 2) {
                const uint32_t b = *src;
                uint4 o;
                o.x = expand4(b & 15u);
                o.y = expand4((b >> 4) & 15u);
                o.z = expand4((b >> 8) & 15u);
                o.w = expand4(b >> 12);
                if (vec) {
                    __stcs(reinterpret_cast<uint4*>(dst), o);
                } else {
                    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
                    for (int q = 0; q < 16 && g * 16 + q < W; ++q) dst[q] = (uint8_t)(w[q >> 2] >> ((q & 3) * 8));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ fused close + emit
// The fixed-geometry band kernel with the emission of the labelling stage folded in (emit.cu's work without its launch and
// without re-reading the bit mask): the band computes its final mask rows y0-1 .. y0+nout (one more threshold row on each
// side than the plain kernel: halo 3), stores rows y0 .. y0+nout-1 as before, and cuts the runs and boundary-pixel records
// of its own rows straight from shared memory (emit_core.cuh).
// t: (nout+6) x TW threshold words (row 0 <-> image row y0-3).  work: the TMA ring, free once the last chunk is consumed.
template <class G>
__device__ __forceinline__ void close_store_emit(const PixelParams& p, uint32_t* t, uint32_t* d, uint8_t* work, int frame, int y0,
                                                 int nout, int tid) {
    constexpr int W = G::kW, NT = G::kNT, WB = (G::kW + 31) / 32, TW = WB + 2;
    constexpr uint32_t valid = (G::kW & 31) ? ((1u << (G::kW & 31)) - 1u) : 0xFFFFFFFFu;
    constexpr int nseg = NT >= WB ? NT / WB : 1;
    static_assert((WB & 3) == 0 && NT % 32 == 0, "fused emit: 16-byte groups of bit words, whole warps");
    const int H = p.H;
    // ---- dilate: rows y0-2 .. y0+nout+1 (outside the image: all ones so that the erode ignores them)
    {
        const int nd = nout + 4, per = (nd + nseg - 1) / nseg;
        for (int slot = tid; slot < WB * nseg; slot += NT) {
            const int seg = slot / WB, k = slot - seg * WB;
            const int i0 = seg * per, i1 = min(nd, i0 + per);
            if (i0 >= i1) continue;
            const bool last = k == WB - 1, prelast = k + 1 == WB - 1;
            auto hrow = [&](int ti) -> uint32_t {
                const uint32_t* a = t + (size_t)ti * TW + k;
                uint32_t left = a[0], mid = a[1], right = a[2];
                if (last) mid &= valid;
                if (prelast) right &= valid;
                return mid | (mid << 1) | (left >> 31) | (mid >> 1) | (right << 31);
            };
            uint32_t h0 = hrow(i0), h1 = hrow(i0 + 1);
            uint32_t* dp = d + (size_t)i0 * TW + k + 1;
            for (int i = i0; i < i1; ++i, dp += TW) {
                const uint32_t h2 = hrow(i + 2);
                const int y = y0 - 2 + i;
                uint32_t dv = h0 | h1 | h2;
                if (last) dv |= ~valid;
                if (y < 0 || y >= H) dv = 0xFFFFFFFFu;
                dp[0] = dv;
                if (k == 0) dp[-1] = 0xFFFFFFFFu;
                if (last) dp[1] = 0xFFFFFFFFu;
                h0 = h1; h1 = h2;
            }
        }
    }
    __syncthreads();
    // ---- erode: rows y0-1 .. y0+nout into mm (compact [nout+2][WB], aliases t); own rows also go to the global bit mask
    uint32_t* mm = t;
    {
        uint32_t* gbits = p.bits + ((size_t)frame * H + y0) * WB;
        const int ne = nout + 2, per = (ne + nseg - 1) / nseg;
        for (int slot = tid; slot < WB * nseg; slot += NT) {
            const int seg = slot / WB, k = slot - seg * WB;
            const int j0 = seg * per, j1 = min(ne, j0 + per);
            if (j0 >= j1) continue;
            const bool last = k == WB - 1;
            auto hrow = [&](int di) -> uint32_t {
                const uint32_t* a = d + (size_t)di * TW + k;
                const uint32_t left = a[0], mid = a[1], right = a[2];
                return mid & ((mid << 1) | (left >> 31)) & ((mid >> 1) | (right << 31));
            };
            uint32_t h0 = hrow(j0), h1 = hrow(j0 + 1);
            uint32_t* mp = mm + (size_t)j0 * WB + k;
            for (int j = j0; j < j1; ++j, mp += WB) {
                const uint32_t h2 = hrow(j + 2);
                uint32_t mv = h0 & h1 & h2;
                if (last) mv &= valid;
                const int y = y0 - 1 + j;
                if (y < 0 || y >= H) mv = 0u;                    // rows outside the image are background for the emission
                mp[0] = mv;
                if (j >= 1 && j <= nout) gbits[(size_t)(j - 1) * WB + k] = mv;
                h0 = h1; h1 = h2;
            }
        }
    }
    __syncthreads();
    uint32_t* m = mm + WB;   // row 0 of m <-> image row y0
    // ---- byte mask: one 16-byte store per 16 pixels, a thread walks down one 16-pixel column
    if (p.mask != nullptr) {
        const uint16_t* m16 = reinterpret_cast<const uint16_t*>(m);
        constexpr int gpr16 = (W + 15) >> 4;
        constexpr int nsg = NT >= gpr16 ? NT / gpr16 : 1;
        uint8_t* gmask = p.mask + (size_t)frame * p.mask_frame_stride + (size_t)y0 * p.mask_pitch;
        for (int slot = tid; slot < gpr16 * nsg; slot += NT) {
            const int sg = slot / gpr16, g = slot - sg * gpr16;
            const bool vec = p.mask_vec && g * 16 + 16 <= W;
            uint8_t* dst = gmask + (size_t)sg * p.mask_pitch + (size_t)g * 16;
            const size_t dstep = (size_t)nsg * p.mask_pitch;
            const uint16_t* src = m16 + (size_t)sg * WB * 2 + g;
            for (int j = sg; j < nout; j += nsg, dst += dstep, src += (size_t)nsg * WB * 2) {
                const uint32_t b = *src;
                uint4 o;
                o.x = expand4(b & 15u);
                o.y = expand4((b >> 4) & 15u);
                o.z = expand4((b >> 8) & 15u);
                o.w = expand4(b >> 12);
                if (vec) {
                    __stcs(reinterpret_cast<uint4*>(dst), o);
                } else {
                    for (int q = 0; q < 16 && g * 16 + q < W; ++q) {
                        const uint32_t w4 = q < 4 ? o.x : q < 8 ? o.y : q < 12 ? o.z : o.w;
                        dst[q] = (uint8_t)(w4 >> ((q & 3) * 8));
                    }
                }
            }
        }
    }
    // ---- emission of the band's own rows (the stage ring is free: every chunk has been consumed)
    __shared__ int s_wtot[NT / 32], s_base[2];
    constexpr int nwarps = NT / 32, cap = G::kBH * WB;
    long long* scratch = reinterpret_cast<long long*>(work);
    int* erun = reinterpret_cast<int*>(scratch + 36);
    int* erec = erun + cap + 1;
    uint16_t* list = reinterpret_cast<uint16_t*>(erec + cap + 1);
    const int lane = tid & 31, warp = tid >> 5;
    int pos = 0, n_ent = 0;
    {
        const uint4* src = reinterpret_cast<const uint4*>(mm);
        const int own_lo = WB >> 2, own_hi = ((nout + 1) * WB) >> 2;
        constexpr int per_w = ((((G::kBH * WB) >> 2) + nwarps - 1) / nwarps + 31) & ~31;
        static_assert(per_w <= 256, "eight rounds of nibbles per warp");
        const int q0 = min(own_hi, own_lo + warp * per_w), q1 = min(own_hi, q0 + per_w);
        uint32_t nibs = 0;
        int mine = 0, it = 0;
        for (int i = q0 + lane; i < q1; i += 32, ++it) {
            const uint4 v = src[i];
            const uint32_t nib = (v.x != 0u) | ((v.y != 0u) << 1) | ((v.z != 0u) << 2) | ((v.w != 0u) << 3);
            nibs |= nib << (4 * it);
            mine += __popc(nib);
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0) s_wtot[warp] = mine;
        __syncthreads();
        for (int w = 0; w < nwarps; ++w) { if (w < warp) pos += s_wtot[w]; n_ent += s_wtot[w]; }
        it = 0;
        for (int i0 = q0; i0 < q1; i0 += 32, ++it) {
            uint32_t nib = (nibs >> (4 * it)) & 15u;
            const int c = __popc(nib);
            int incl = c;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            int at = pos + incl - c;
            const int wbase = ((i0 + lane) << 2) - WB;   // word index relative to the band's first own row
            while (nib) {
                const int b = __ffs(nib) - 1;
                nib &= nib - 1;
                list[at++] = (uint16_t)(wbase + b);
            }
            pos += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    emit_tail<NT>(p.em, frame, y0, nout, n_ent, m, list, erun, erec, scratch, s_base, tid);
}

// ------------------------------------------------------------------------------------------ BGR kernel
RMCV_GSTAMP_ARRAY(g_ns_pixel)
template <bool kBulk, class G = GeomRuntime, bool kEmit = false>
__global__ void __launch_bounds__(512) pixel_bgr_kernel(const PixelParams p) {
    static_assert(!kEmit || G::kW > 0, "the fused emission needs the fixed geometry");
    extern __shared__ __align__(128) uint8_t smem[];
    RMCV_GSTAMP_BEGIN(g_ns_pixel, 0);
    chain_begin();   // a chained emit kernel (small chunks) may become resident now; it waits for this grid to complete
    constexpr bool kFixed = G::kW > 0;
    const int tid = threadIdx.x, NT = kFixed ? G::kNT : (int)blockDim.x;
    const int frame = blockIdx.x / p.bands, band = blockIdx.x - frame * p.bands;
    const int W = kFixed ? G::kW : p.W, H = p.H, WB = kFixed ? (G::kW + 31) / 32 : p.WB, TW = WB + 2;
    const int BH = kFixed ? G::kBH : p.BH, RC = kFixed ? G::kRC : p.RC, S = kFixed ? G::kS : p.S;
    const int srow = kFixed ? ((G::kW + 15) / 16) * 48 : p.srow;
    const int y0 = band * BH;
    const int nout = min(BH, H - y0);
    const int hl = kEmit ? 3 : p.halo;                    // the fused emission needs the final rows y0-1 and y0+nout too
    const int ty0 = y0 - hl;                              // image row of t row 0
    const int cy0 = max(0, ty0), cy1 = min(H, y0 + nout + hl);
    const int nchunks = (cy1 - cy0 + RC - 1) / RC;
    const size_t stage_bytes = (size_t)RC * srow;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint32_t* t = reinterpret_cast<uint32_t*>(smem + (size_t)S * stage_bytes + (((size_t)S * 8 + 15) & ~(size_t)15));
    uint32_t* d = t + (size_t)(BH + 2 * hl) * TW;
    uint16_t* t16 = reinterpret_cast<uint16_t*>(t);

    for (int i = tid; i < (BH + 2 * hl) * TW; i += NT) t[i] = 0u;
    if (!kBulk) {  // zero the row padding of the stage buffers once (generic widths)
        for (int i = tid; i < (int)(S * stage_bytes / 4); i += NT) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    }
    if (kBulk && tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint8_t* fsrc = p.src + (size_t)frame * p.frame_stride;
    const uint32_t rowbytes = (uint32_t)W * 3u;
    auto issue = [&](int c, int s) {  // one thread; chunk c into stage s
        const int r0 = cy0 + c * RC, nr = min(RC, cy1 - r0);
        uint8_t* dst = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&bars[s], (uint32_t)nr * rowbytes);
        if (p.contiguous) {
            bulk_g2s(dst, fsrc + (size_t)r0 * p.pitch, (uint32_t)nr * rowbytes, &bars[s]);
        } else {
            for (int r = 0; r < nr; ++r) bulk_g2s(dst + (size_t)r * srow, fsrc + (size_t)(r0 + r) * p.pitch, rowbytes, &bars[s]);
        }
    };
    if (kBulk && tid == 0) {
        for (int c = 0; c < S && c < nchunks; ++c) issue(c, c);
    }

    const int gpr = kFixed ? (G::kW + 15) / 16 : p.gpr;
    const uint32_t c0 = p.coef[0], c1a = p.coef[1], c1b = p.coef[2], c2a = p.coef[3], c2b = p.coef[4], c3 = p.coef[5];
    const int acc0 = p.acc0;
    const Iter2D it0(tid, NT, gpr);

    int s = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = cy0 + c * RC, nr = min(RC, cy1 - r0);
        const uint8_t* stage = smem + (size_t)s * stage_bytes;
        if (kBulk) {
            mbar_wait(&bars[s], phase);
        } else {
            uint8_t* wstage = smem + (size_t)s * stage_bytes;
            for (int r = 0; r < nr; ++r) {
                const uint8_t* g = fsrc + (size_t)(r0 + r) * p.pitch;
                for (uint32_t i = tid; i < rowbytes; i += NT) wstage[(size_t)r * srow + i] = __ldg(g + i);
            }
            __syncthreads();
        }
        for (Iter2D it = it0; it.r < nr; it.next()) {
            const uint4* q = reinterpret_cast<const uint4*>(stage + (size_t)it.r * srow + (size_t)it.c * 48);
            const uint4 A = q[0], B = q[1], C = q[2];
            const uint32_t w[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
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
            const int ty = r0 + it.r - ty0;
            t16[(size_t)ty * TW * 2 + 2 + it.c] = (uint16_t)(~nb);
        }
        __syncthreads();  // stage s fully consumed (and t rows of this chunk visible)
        if (kBulk && tid == 0 && c + S < nchunks) issue(c + S, s);
        if (++s == S) { s = 0; phase ^= 1u; }
    }
    if constexpr (kEmit) close_store_emit<G>(p, t, d, smem, frame, y0, nout, tid);
    else close_and_store<G>(p, t, d, frame, y0, nout, tid, NT);
    RMCV_GSTAMP_END(g_ns_pixel, 0);
}

// ------------------------------------------------------------------------------------------ Bayer kernel
// Raw 8-bit mosaic, 1 B/px.  For the colour difference only two planes are needed; they are rebuilt with the
// OpenCV bilinear rule: sample at its own site, (a+b+1)>>1 from two neighbours, (a+b+c+d+2)>>2
// from four; then row 0 := row 1, row H-1 := row H-2, column 0 := column 1, column W-1 := column W-2, which on
// threshold bits is a replicate of the neighbouring interior bit.
struct BayerSite {  // which of the 4 interpolation patterns yields channel c at parity (py,px)
    // 0 = own sample, 1 = horizontal pair, 2 = vertical pair, 3 = four diagonals, 4 = four plus-neighbours
    uint8_t plus[2][2], minus[2][2];
};

// The threshold test of one pixel as integer linear forms over its 3x3 raw neighbourhood.  With P = (Sp + rp) >> kp and
// M = (Sm + rm) >> km (kp, km in {0,1,2}: own sample / pair / quad) the test sat(P - M) >= lb, lb >= 1, is
//   km == 0:  Sp - (Sm << kp) + rp - (lb << kp) >= 0
//   kp == 0:  (Sp << km) - Sm - c0 >= 0,             c0 = (lb << km) - (1 << km) + rm + 1
//   else:     (((Sp + rp) >> kp) << km) - Sm - c0 >= 0
// i.e.  V + ((Q + rq) >> kq << ks) >= 0  with V, Q linear in the nine bytes: three dp4a each (one per raw row), the
// coefficient bytes placed where the pixel's neighbours sit in a word holding the bytes x-1 .. x+2 (even x) or
// x-2 .. x+1 (odd x).
struct BayerProg {
    uint32_t vu, vc, vd;   // V coefficients (packed s8x4) for the rows y-1, y, y+1
    uint32_t qu, qc, qd;   // Q coefficients; unused when kq == 0
    int32_t v0;            // constant of V
    int32_t rq, kq, ks;
};
// Shape of a pixel program, so that zero coefficient words cost nothing: 0 = V over all three rows, no Q (a sampled
// site against the four diagonals); 1 = Q on the centre row, V on the rows above/below, pair rounding (kq = ks = 1: a
// green site between two `plus` samples of its row); 2 = Q on the rows above/below, V on the centre row (a green site
// between two `plus` samples of its column); 3 = generic.
struct BayerProgs { BayerProg g[2][2]; int type[2][2]; };  // [row parity][x parity]

template <int T>
__device__ __forceinline__ int bayer_pixel(const BayerProg& g, uint32_t wu, uint32_t wc, uint32_t wd) {
    if (T == 0) return dp4a_us(wu, g.vu, dp4a_us(wc, g.vc, dp4a_us(wd, g.vd, g.v0)));
    if (T == 1) return dp4a_us(wu, g.vu, dp4a_us(wd, g.vd, g.v0)) + (dp4a_us(wc, g.qc, g.rq) & ~1);
    if (T == 2) return dp4a_us(wc, g.vc, g.v0) + (dp4a_us(wu, g.qu, dp4a_us(wd, g.qd, g.rq)) & ~1);
    int v = dp4a_us(wu, g.vu, dp4a_us(wc, g.vc, dp4a_us(wd, g.vd, g.v0)));
    if (g.kq) v += (dp4a_us(wu, g.qu, dp4a_us(wc, g.qc, dp4a_us(wd, g.qd, g.rq))) >> g.kq) << g.ks;
    return v;
}

// Threshold bits of one raw row: lanes over 16-pixel items.  Odd-x pixels sit at byte 1 of an aligned word (their
// neighbours at bytes 0 and 2), even-x pixels at byte 2; the pixels at bytes 3 and 0 use the word shifted by 16 bits,
// where they sit at bytes 1 and 2 again.  TE / TO = program shapes of the even-x / odd-x pixels.
template <int TE, int TO>
__device__ __forceinline__ void bayer_row(const uint32_t* rm, const uint32_t* r0, const uint32_t* rp, int nw, int items, int lane,
                                          const BayerProg& pe, const BayerProg& po, uint16_t* trow, uint32_t force_or,
                                          uint32_t force_and) {
    for (int it = lane; it < items; it += 32) {
        const int wi = it * 4;
        const uint4 U = *reinterpret_cast<const uint4*>(rm + wi), C = *reinterpret_cast<const uint4*>(r0 + wi),
                    D = *reinterpret_cast<const uint4*>(rp + wi);
        const bool hasp = wi > 0, hasn = wi + 4 < nw;
        const uint32_t u[6] = {hasp ? rm[wi - 1] : 0u, U.x, U.y, U.z, U.w, hasn ? rm[wi + 4] : 0u};
        const uint32_t c[6] = {hasp ? r0[wi - 1] : 0u, C.x, C.y, C.z, C.w, hasn ? r0[wi + 4] : 0u};
        const uint32_t d[6] = {hasp ? rp[wi - 1] : 0u, D.x, D.y, D.z, D.w, hasn ? rp[wi + 4] : 0u};
        uint32_t nb = 0;  // sign bits (1 = below the threshold), pixel 15 first so that pixel 0 lands in bit 0
#pragma unroll
        for (int m = 3; m >= 0; --m) {
            // su/sc/sd: bytes (w[m].2, w[m].3, w[m+1].0, w[m+1].1) -> pixel 4m+3 (odd x) and pixel 4m+4 (even x, next group)
            const uint32_t su = __funnelshift_r(u[m + 1], u[m + 2], 16), sc = __funnelshift_r(c[m + 1], c[m + 2], 16),
                           sd = __funnelshift_r(d[m + 1], d[m + 2], 16);
            nb = __funnelshift_l((uint32_t)bayer_pixel<TO>(po, su, sc, sd), nb, 1);                    // pixel 4m+3
            nb = __funnelshift_l((uint32_t)bayer_pixel<TE>(pe, u[m + 1], c[m + 1], d[m + 1]), nb, 1);  // pixel 4m+2
            nb = __funnelshift_l((uint32_t)bayer_pixel<TO>(po, u[m + 1], c[m + 1], d[m + 1]), nb, 1);  // pixel 4m+1
            const uint32_t tu = __funnelshift_r(u[m], u[m + 1], 16), tc = __funnelshift_r(c[m], c[m + 1], 16),
                           td = __funnelshift_r(d[m], d[m + 1], 16);
            nb = __funnelshift_l((uint32_t)bayer_pixel<TE>(pe, tu, tc, td), nb, 1);                    // pixel 4m
        }
        trow[it] = (uint16_t)((~nb | force_or) & force_and);
    }
}

__global__ void __launch_bounds__(512) pixel_bayer_kernel(const PixelParams p, const BayerProgs progs) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int frame = blockIdx.x / p.bands, band = blockIdx.x - frame * p.bands;
    const int W = p.W, H = p.H, WB = p.WB, TW = WB + 2, BH = p.BH;
    const int y0 = band * BH;
    const int nout = min(BH, H - y0);
    const int hl = p.halo;
    const int ty0 = y0 - hl;
    // threshold rows needed: [y0-hl, y0+nout+hl) clipped; raw rows: one more on each side, clamped
    const int cy0 = max(0, ty0), cy1 = min(H, y0 + nout + hl);
    const int ry0 = max(0, min(cy0, H - 2) - 1), ry1 = min(H, max(cy1 - 1, 1) + 2);  // raw rows yc-1 .. yc+1, yc clamped to [1, H-2]
    const int nraw = ry1 - ry0;
    const size_t raw_bytes = ((size_t)(BH + 2 * hl + 2) * p.srow + 15) & ~(size_t)15;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + raw_bytes);  // after the raw rows
    uint32_t* t = reinterpret_cast<uint32_t*>(smem + raw_bytes + 128);   // room for 16 mbarriers (<= 128 raw rows)
    uint32_t* d = t + (size_t)(BH + 2 * hl) * TW;

    for (int i = tid; i < (BH + 2 * hl) * TW; i += NT) t[i] = 0u;
    const uint8_t* fsrc = p.src + (size_t)frame * p.frame_stride;
    const bool bulk = p.contiguous >= 0 && (p.srow == W) && ((W & 15) == 0) && ((p.pitch & 15) == 0) &&
                      ((((size_t)fsrc) & 15) == 0);
    constexpr int kLoadRows = 8;   // raw rows per mbarrier: threshold rows start as soon as their three raw rows are in
    const int nloads = (nraw + kLoadRows - 1) / kLoadRows;
    if (bulk) {
        if (tid == 0) {
            for (int c = 0; c < nloads; ++c) mbar_init(&bars[c], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            for (int c = 0; c < nloads; ++c) {
                const int r = c * kLoadRows, nr = min(kLoadRows, nraw - r);
                mbar_expect_tx(&bars[c], (uint32_t)nr * (uint32_t)W);
                if (p.pitch == (size_t)W) {
                    bulk_g2s(smem + (size_t)r * p.srow, fsrc + (size_t)(ry0 + r) * p.pitch, (uint32_t)nr * (uint32_t)W, &bars[c]);
                } else {
                    for (int q = 0; q < nr; ++q)
                        bulk_g2s(smem + (size_t)(r + q) * p.srow, fsrc + (size_t)(ry0 + r + q) * p.pitch, (uint32_t)W, &bars[c]);
                }
            }
        }
    } else {
        for (int r = 0; r < nraw; ++r) {
            const uint8_t* g = fsrc + (size_t)(ry0 + r) * p.pitch;
            for (int i = tid; i < W; i += NT) smem[(size_t)r * p.srow + i] = __ldg(g + i);
        }
        __syncthreads();
    }
    // threshold bits: one warp per row (the pixel programs depend on the row parity only), lanes over 16-pixel items;
    // per pixel three dp4a on aligned / 16-bit-shifted words.  Border rows replicate the neighbouring interior row
    // (recomputed), border columns are fixed up afterwards.
    const int lb = p.lb;
    const int nw = p.srow >> 2;  // 32-bit words per shared-memory row
    const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int items = (W + 15) >> 4;
    const uint32_t force_or = lb <= 0 ? 0xffffu : 0u, force_and = lb > 255 ? 0u : 0xffffu;
    uint16_t* t16 = reinterpret_cast<uint16_t*>(t);
    for (int rr = warp; rr < cy1 - cy0; rr += nwarps) {
        const int y = cy0 + rr;
        const int yc = min(max(y, 1), H - 2);          // row whose interior values this row shows
        const uint32_t* rm = reinterpret_cast<const uint32_t*>(smem + (size_t)(yc - 1 - ry0) * p.srow);
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(smem + (size_t)(yc - ry0) * p.srow);
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(smem + (size_t)(yc + 1 - ry0) * p.srow);
        const BayerProg pe = progs.g[yc & 1][0], po = progs.g[yc & 1][1];
        const int te = progs.type[yc & 1][0], to = progs.type[yc & 1][1];
        uint16_t* trow = t16 + ((size_t)(y - ty0) * TW + 1) * 2;
        if (bulk) {   // the loads that hold raw rows yc-1 .. yc+1
            const int ca = (yc - 1 - ry0) / kLoadRows, cb = (yc + 1 - ry0) / kLoadRows;
            mbar_wait(&bars[ca], 0);
            if (cb != ca) mbar_wait(&bars[cb], 0);
        }
        // warp-uniform dispatch on the row's pair of program shapes
        if (te == 0 && to == 1) bayer_row<0, 1>(rm, r0, rp, nw, items, lane, pe, po, trow, force_or, force_and);
        else if (te == 1 && to == 0) bayer_row<1, 0>(rm, r0, rp, nw, items, lane, pe, po, trow, force_or, force_and);
        else if (te == 2 && to == 0) bayer_row<2, 0>(rm, r0, rp, nw, items, lane, pe, po, trow, force_or, force_and);
        else if (te == 0 && to == 2) bayer_row<0, 2>(rm, r0, rp, nw, items, lane, pe, po, trow, force_or, force_and);
        else bayer_row<3, 3>(rm, r0, rp, nw, items, lane, pe, po, trow, force_or, force_and);
    }
    __syncthreads();
    // border columns: x = 0 shows x = 1, x = W-1 shows x = W-2 (after the row replicate, which the clamped row did)
    for (int rr = tid; rr < cy1 - cy0; rr += NT) {
        uint32_t* trow = t + (size_t)(cy0 + rr - ty0) * TW + 1;
        uint32_t w0 = trow[0];
        w0 = (w0 & ~1u) | ((w0 >> 1) & 1u);
        trow[0] = w0;
        const int xl = W - 1, xs = W - 2;
        const uint32_t bit = (trow[xs >> 5] >> (xs & 31)) & 1u;
        trow[xl >> 5] = (trow[xl >> 5] & ~(1u << (xl & 31))) | (bit << (xl & 31));
    }
    __syncthreads();
    close_and_store(p, t, d, frame, y0, nout, tid, NT);
}

// ------------------------------------------------------------------------------------------ host launcher
cudaError_t launch_pixel_stage(const PixelLaunch& L, int sm_count, cudaStream_t st, int64_t* launches) {
    PixelParams p;
    memset(&p, 0, sizeof(p));
    p.src = L.src; p.pitch = L.pitch; p.frame_stride = L.frame_stride;
    p.mask = L.mask; p.mask_pitch = L.mask_pitch; p.mask_frame_stride = L.mask_frame_stride;
    p.bits = L.bits;
    p.W = L.W; p.H = L.H; p.WB = (L.W + 31) / 32;
    p.last_valid = (L.W & 31) ? ((1u << (L.W & 31)) - 1u) : 0xFFFFFFFFu;
    p.mask_vec = (L.mask != nullptr) && ((L.mask_pitch & 15) == 0) && ((L.mask_frame_stride & 15) == 0) &&
                 ((((size_t)L.mask) & 15) == 0);
    p.lb = L.lower_bound;
    p.halo = 2;
    const int hl = p.halo;
    {
        const unsigned g16 = (unsigned)((L.W + 15) / 16);
        p.inv_wb = p.WB > 1 ? (uint32_t)((1ull << 32) / (unsigned)p.WB) + 1u : 0u;
        p.inv_gpr16 = g16 > 1 ? (uint32_t)((1ull << 32) / g16) + 1u : 0u;
    }

    // band height: tall bands amortise the 4 halo rows; small batches need more, shorter bands to fill 148 SMs
    const Tuning& tune = tuning();
    int BH = tune.pix_bh > 0 ? tune.pix_bh : 0;
    if (BH <= 0) {
        BH = 32;
        while (BH > 8 && (long long)L.batch * ((L.H + BH - 1) / BH) < 4LL * sm_count) BH >>= 1;
    }
    if (BH > L.H) BH = L.H;
    p.BH = BH;
    p.bands = (L.H + BH - 1) / BH;
    const long long grid = (long long)L.batch * p.bands;
    if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorInvalidValue;

    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);

    if (L.bayer_layout == 0) {
        // ---- BGR.  The band-strip kernel (bgr_bandstrip.cu: TMA-fed bands consumed by register-resident lanes) needs half the
        // instructions per pixel, but as it draws more bandwidth per unit of time in the pipeline it lengthens the memory
        // latency the labelling kernels beside it see, and the whole path is faster with the band kernel below
        // (DESIGN.md 4.3).  RMCV_BGR_STRIP=1 selects the band-strip kernel.
        if (tune.bgr_strip != 0) {
            const cudaError_t se = launch_bgr_bandstrip(L, sm_count, st, launches);
            if (se != cudaErrorNotSupported) return se;
        }
        int a, b;  // plus / minus channel (src/imgproc.cpp:56-65)
        if (L.target == RMCV_CAMP_GUIDELIGHT) { a = 1; b = 2; }
        else if (L.target == RMCV_CAMP_BLUE) { a = 0; b = 2; }
        else { a = 2; b = 0; }
        // byte k of a 12-byte group belongs to pixel k/3, channel k%3; coefficient words are per (pixel, word)
        // order: c0 (px0,w0) c1a (px1,w0) c1b (px1,w1) c2a (px2,w1) c2b (px2,w2) c3 (px3,w2)
        auto put = [&](int slot, int px, int word) {
            uint32_t v = 0;
            for (int lane = 0; lane < 4; ++lane) {
                int k = word * 4 + lane;
                if (k / 3 != px) continue;
                int ch = k % 3;
                int coef = (ch == a ? 1 : 0) - (ch == b ? 1 : 0);
                v |= (uint32_t)(uint8_t)(int8_t)coef << (8 * lane);
            }
            p.coef[slot] = v;
        };
        put(0, 0, 0); put(1, 1, 0); put(2, 1, 1); put(3, 2, 1); put(4, 2, 2); put(5, 3, 2);
        // v = diff + acc0 >= 0  <=>  sat_u8(diff) in [lb, 255].  The two-word pixels add acc0 once (inner dp4a).
        if (L.lower_bound <= 0) { for (int i = 0; i < 6; ++i) p.coef[i] = 0; p.acc0 = 0; }
        else if (L.lower_bound > 255) { for (int i = 0; i < 6; ++i) p.coef[i] = 0; p.acc0 = -1; }
        else p.acc0 = -L.lower_bound;

        p.gpr = (L.W + 15) / 16;
        p.srow = p.gpr * 48;
        p.band_flags = L.band_flags;
        if (L.flags_bh) *L.flags_bh = L.band_flags ? p.BH : 0;
        const bool bulk = ((L.W & 15) == 0) && ((L.pitch & 15) == 0) && ((L.frame_stride & 15) == 0) &&
                          ((((size_t)L.src) & 15) == 0) && tune.pix_nobulk == 0;
        p.contiguous = (L.pitch == (size_t)L.W * 3) ? 1 : 0;
        int RC = tune.pix_rc > 0 ? tune.pix_rc : 0;
        if (RC <= 0) RC = max(1, min(16, 16384 / p.srow));  // ~16 KB per TMA chunk (sweep: gpurun_out/sweep.log)
        if (RC > BH + 2 * hl) RC = BH + 2 * hl;
        int S = tune.pix_s > 0 ? tune.pix_s : 4;
        int NT = tune.pix_nt > 0 ? tune.pix_nt : 0;
        if (NT <= 0) {  // one 16-pixel group per thread per chunk, rounded up to whole warps
            NT = ((RC * p.gpr + 31) / 32) * 32;
            NT = max(128, min(512, NT));
        }
        p.RC = RC; p.S = S;
        size_t smem = pix_smem_bytes(S, RC, p.srow, BH, p.WB, hl);
        while (smem > (size_t)max_smem && p.S > 1) { --p.S; smem = pix_smem_bytes(p.S, RC, p.srow, BH, p.WB, hl); }
        while (smem > (size_t)max_smem && p.RC > 1) { --p.RC; smem = pix_smem_bytes(p.S, p.RC, p.srow, BH, p.WB, hl); }
        if (smem > (size_t)max_smem) return cudaErrorInvalidConfiguration;
        cudaError_t e;
        const bool fixed1280 = bulk && L.W == Geom1280::kW && p.BH == Geom1280::kBH && p.RC == Geom1280::kRC && p.S == Geom1280::kS &&
                               NT == Geom1280::kNT && tune.pix_generic == 0;
        if (fixed1280 && L.emit != nullptr && tune.fused_emit > 0 && L.H >= 3) {
            // RMCV_FUSED_EMIT=1: the same band kernel with the labelling stage's emission folded in (halo 3): one launch less per
            // chunk and the bit mask is not read back.  Bit-identical results; measured: the fused kernel takes
            // 0.92 ms per 1024 frames alone against 0.81 + 0.16 ms for the two kernels, but the pipelined step only moves from
            // 1.245 to 1.234 ms, so the plain kernel (at the HBM roofline on its own) stays the default.
            const EmitLaunch& E = *L.emit;
            p.em.bits = L.bits; p.em.W = L.W; p.em.H = L.H; p.em.WB = p.WB; p.em.BH = p.BH; p.em.bands = p.bands;
            p.em.inv_wb = p.inv_wb;
            p.em.rows = E.rows; p.em.run_x = E.run_x; p.em.run_y = E.run_y; p.em.counters = E.counters; p.em.R = E.R;
            p.em.recs = E.recs; p.em.PC = E.PC;
            const size_t smem_e = pix_smem_bytes(S, RC, p.srow, BH, p.WB, 3);
            if (smem_e <= (size_t)max_smem) {
                e = cudaFuncSetAttribute(pixel_bgr_kernel<true, Geom1280, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e);
                if (e != cudaSuccess) return e;
                pixel_bgr_kernel<true, Geom1280, true><<<(unsigned)grid, NT, smem_e, st>>>(p);
                if (launches) ++*launches;
                if (L.emit_done) *L.emit_done = 1;
                return cudaGetLastError();
            }
        }
        if (fixed1280) {   // the default configuration with its geometry folded into the code
            e = cudaFuncSetAttribute(pixel_bgr_kernel<true, Geom1280>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            pixel_bgr_kernel<true, Geom1280><<<(unsigned)grid, NT, smem, st>>>(p);
        } else if (bulk) {
            e = cudaFuncSetAttribute(pixel_bgr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            pixel_bgr_kernel<true><<<(unsigned)grid, NT, smem, st>>>(p);
        } else {
            e = cudaFuncSetAttribute(pixel_bgr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            pixel_bgr_kernel<false><<<(unsigned)grid, NT, smem, st>>>(p);
        }
        if (launches) ++*launches;
        return cudaGetLastError();
    }

    // ---- Bayer: blue / red targets on 16-pixel-aligned frames take the register-resident strip kernel (bayer_strip.cu)
    if (tune.bayer_generic == 0) {
        const cudaError_t se = launch_bayer_strip(L, sm_count, st, launches);
        if (se != cudaErrorNotSupported) return se;
    }
    BayerSite site;
    {
        // colour sampled at (y&1, x&1) for the Daheng layouts (0=B,1=G,2=R)
        int ch[2][2];
        switch (L.bayer_layout) {
            case RMCV_BAYER_BG: ch[0][0] = 0; ch[0][1] = 1; ch[1][0] = 1; ch[1][1] = 2; break;
            case RMCV_BAYER_GB: ch[0][0] = 1; ch[0][1] = 0; ch[1][0] = 2; ch[1][1] = 1; break;
            case RMCV_BAYER_GR: ch[0][0] = 1; ch[0][1] = 2; ch[1][0] = 0; ch[1][1] = 1; break;
            case RMCV_BAYER_RG: ch[0][0] = 2; ch[0][1] = 1; ch[1][0] = 1; ch[1][1] = 0; break;
            default: return cudaErrorInvalidValue;
        }
        int a, b;
        if (L.target == RMCV_CAMP_GUIDELIGHT) { a = 1; b = 2; }
        else if (L.target == RMCV_CAMP_BLUE) { a = 0; b = 2; }
        else { a = 2; b = 0; }
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                auto mode = [&](int c) -> uint8_t {
                    int s = ch[py][px];
                    if (s == c) return 0;
                    if (c == 1) return 4;                     // green at a red/blue site: plus neighbours
                    if (s == 1) return ch[py][px ^ 1] == c ? 1 : 2;  // at a green site: row or column pair
                    return 3;                                 // opposite colour: diagonals
                };
                site.plus[py][px] = mode(a);
                site.minus[py][px] = mode(b);
            }
    }
    BayerProgs progs;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            // weights of the interpolation patterns over (row -1/0/+1, column -1/0/+1), and their (shift, rounding)
            auto weights = [](int mode, int w[3][3], int* k, int* r) {
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w[i][j] = 0;
                switch (mode) {
                    case 0: w[1][1] = 1; *k = 0; *r = 0; break;
                    case 1: w[1][0] = w[1][2] = 1; *k = 1; *r = 1; break;
                    case 2: w[0][1] = w[2][1] = 1; *k = 1; *r = 1; break;
                    case 3: w[0][0] = w[0][2] = w[2][0] = w[2][2] = 1; *k = 2; *r = 2; break;
                    default: w[1][0] = w[1][2] = w[0][1] = w[2][1] = 1; *k = 2; *r = 2; break;
                }
            };
            int wp[3][3], wm[3][3], kp, rp, km, rm;
            weights(site.plus[py][px], wp, &kp, &rp);
            weights(site.minus[py][px], wm, &km, &rm);
            const int lbv = L.lower_bound;
            int V[3][3], Q[3][3];
            BayerProg g;
            memset(&g, 0, sizeof(g));
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { V[i][j] = 0; Q[i][j] = 0; }
            if (km == 0) {
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = wp[i][j] - (wm[i][j] << kp);
                g.v0 = rp - (lbv << kp);
            } else if (kp == 0) {
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (wp[i][j] << km) - wm[i][j];
                g.v0 = -((lbv << km) - (1 << km) + rm + 1);
            } else {
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { V[i][j] = -wm[i][j]; Q[i][j] = wp[i][j]; }
                g.v0 = -((lbv << km) - (1 << km) + rm + 1);
                g.rq = rp; g.kq = kp; g.ks = km;
            }
            // byte lanes: an odd-x pixel sits at byte 1 of its word (neighbours at bytes 0 and 2), an even-x pixel at byte 2
            auto pack = [&](const int row[3]) -> uint32_t {
                uint32_t v = 0;
                for (int j = 0; j < 3; ++j) v |= (uint32_t)(uint8_t)(int8_t)row[j] << (8 * (j + (px == 0 ? 1 : 0)));
                return v;
            };
            g.vu = pack(V[0]); g.vc = pack(V[1]); g.vd = pack(V[2]);
            g.qu = pack(Q[0]); g.qc = pack(Q[1]); g.qd = pack(Q[2]);
            progs.g[py][px] = g;
            int type = 3;
            if (g.kq == 0) type = 0;
            else if (g.kq == 1 && g.ks == 1 && !g.vc && !g.qu && !g.qd) type = 1;
            else if (g.kq == 1 && g.ks == 1 && !g.vu && !g.vd && !g.qc) type = 2;
            progs.type[py][px] = type;
        }
    p.bayer = 1;
    p.gpr = p.WB;
    p.srow = (L.W + 15) & ~15;
    if ((L.W & 15) == 0) p.srow = L.W;
    p.contiguous = 0;
    int NT = tune.pix_nt > 0 ? tune.pix_nt : 256;
    if (L.W < 3 || L.H < 3) return cudaErrorInvalidValue;
    size_t raw_bytes = ((size_t)(BH + 2 * hl + 2) * p.srow + 15) & ~(size_t)15;
    size_t smem = raw_bytes + 128 + (size_t)(BH + 2 * hl) * (p.WB + 2) * 4 + (size_t)(BH + 2 * hl - 2) * (p.WB + 2) * 4 + 256;
    while (smem > (size_t)max_smem && p.BH > 1) {
        p.BH = max(1, p.BH / 2); BH = p.BH;
        p.bands = (L.H + BH - 1) / BH;
        raw_bytes = ((size_t)(BH + 2 * hl + 2) * p.srow + 15) & ~(size_t)15;
        smem = raw_bytes + 128 + (size_t)(BH + 2 * hl) * (p.WB + 2) * 4 + (size_t)(BH + 2 * hl - 2) * (p.WB + 2) * 4 + 256;
    }
    const long long grid2 = (long long)L.batch * p.bands;
    cudaError_t e = cudaFuncSetAttribute(pixel_bayer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pixel_bayer_kernel<<<(unsigned)grid2, NT, smem, st>>>(p, progs);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
RMCV_GSTAMP_GETTER(rmcv_debug_ns_pixel, rmcv::g_ns_pixel)
