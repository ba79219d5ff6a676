// Building blocks of the register-resident pixel kernels (bayer_strip.cu, bgr_bandstrip.cu):
// a lane owns 16 pixels of a row, walks down a segment of rows, receives the threshold words of its two neighbours by warp
// shuffle and runs the 3x3 close of rm::extract_color (src/imgproc.cpp:67-69, OpenCV MORPH_CLOSE borders: dilate pads 0,
// erode pads 1) on a 20-bit window in registers.
#pragma once
#include "common.cuh"

namespace rmcv {
namespace strip {

// ---- cp.async (LDGSTS) ring helpers: a lane only reads back bytes it asked for itself, so groups need no barrier
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, UBLKCP in SASS): bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) { return (a & m) | (b & ~m); }

// ---- bits -> bytes table: 256 entries of 8 bytes (0x00 / 0xFF per mask bit) on a 2 KB boundary of the shared window, so
// that index and base combine with OR.  `raw` must hold 4 KB.  Call with all threads of the CTA, then __syncthreads().
__device__ __forceinline__ uint32_t lut_base(const void* raw) {
    return ((uint32_t)__cvta_generic_to_shared(raw) + 2047u) & ~2047u;
}
__device__ __forceinline__ void lut_init(const void* raw, uint32_t tid0) {
    for (uint32_t tid = tid0; tid < 256u; tid += blockDim.x) {
        auto expand4 = [](uint32_t nib) {   // 4 bits -> 4 bytes: bits to the byte MSBs, PRMT sign-replicate
            uint32_t r;
            asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(nib * 0x10204080u));
            return r;
        };
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(lut_base(raw) + tid * 8u), "r"(expand4(tid & 15u)), "r"(expand4(tid >> 4))
                     : "memory");
    }
}

// ---- the close, one lane
struct CloseLane {
    uint32_t h0, h1, e0, e1;     // sliding rows: two horizontally dilated, two horizontally eroded rows
    uint32_t inside;             // window bit i <-> x = 16c - 2 + i: the bits inside the image (0 for an idle lane)
    uint8_t* mrow;               // byte-mask address of the next row to leave
    uint16_t* brow;              // bit-mask address of the next row to leave
    uint32_t lut;                // shared address of the bits -> bytes table
};

// Threshold row r (16 bits, 0 outside the image) enters, the final row r-2 leaves.  above_outside: row r-1 lies outside the
// image (the erode pads ones there).  do_store / do_tail: this lane writes row r-2 / also zeroes the odd 16-bit word that
// ends a W % 32 == 16 bit row.
template <bool STORE, bool MASK>
__device__ __forceinline__ void push_row(CloseLane& k, int wb2, int mask_pitch, uint32_t t, bool above_outside, bool do_store,
                                         bool do_tail) {
    const uint32_t tl = __shfl_up_sync(0xffffffffu, t, 1), tr = __shfl_down_sync(0xffffffffu, t, 1);
    const uint32_t w = ((tl >> 14) | (t << 2) | (tr << 18)) & k.inside;
    const uint32_t h = w | (w << 1) | (w >> 1);
    uint32_t d = k.h0 | k.h1 | h | ~k.inside;               // dilated row r-1; columns outside the image read as ones
    k.h0 = k.h1; k.h1 = h;
    if (above_outside) d = 0xffffffffu;
    const uint32_t e = d & (d << 1) & (d >> 1);
    const uint32_t m = k.e0 & k.e1 & e;                     // final row r-2 in window bits 2..17
    k.e0 = k.e1; k.e1 = e;
    if (STORE) {
        if (do_store) {
            k.brow[0] = (uint16_t)(m >> 2);
            if (do_tail) k.brow[1] = 0;
            if (MASK) {   // table entries are 8 bytes: pixels 0..7 at (m >> 2 & 255) * 8, pixels 8..15 at (m >> 10 & 255) * 8
                const uint2 a = lds64(((m << 1) & 0x7f8u) | k.lut), b = lds64(((m >> 7) & 0x7f8u) | k.lut);
                __stcs(reinterpret_cast<uint4*>(k.mrow), make_uint4(a.x, a.y, b.x, b.y));
            }
        }
        k.brow += wb2;
        if (MASK) k.mrow += mask_pitch;
    }
}

}  // namespace strip
}  // namespace rmcv
