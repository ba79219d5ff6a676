// K1 (alternative, opt-in: RMCV_BGR_STRIP=1) — pixel stage of rm::extract_color on BGR frames (src/imgproc.cpp:52-69)
// that combines the two designs of this library: rows arrive like in the band kernel of pixel.cu — a CTA owns a band of 32
// output rows, a producer warp feeds a ring of shared-memory stages with large 1-D TMA bulk copies (full / empty mbarriers)
// — and are consumed like in the strip kernels: a lane owns a 16-pixel group, forms its threshold word with dp4a, gets its
// neighbours' words by warp shuffle and runs the 3x3 close in registers (strip.cuh), so no threshold / dilated rows ever
// touch shared memory and there is no block-wide barrier in the loop.  A CTA covers the same band of FPC consecutive frames
// (240 groups at 1280 pixels: 8 consumer warps of 30 groups + halo lanes), which keeps 30 of 32 lanes busy for any width.
#include "strip.cuh"

namespace rmcv {

namespace {

constexpr int kStages = 3;      // ring stages
constexpr int kBand = 32;       // output rows per CTA
constexpr int kMaxWarps = 16;   // consumer warps per CTA

struct BandStripParams {
    const uint8_t* src; size_t frame_stride;
    uint8_t* mask; size_t mask_frame_stride;   // mask may be null
    int pitch, mask_pitch;
    uint16_t* bits16;
    int W, H, NC, WB2;
    int batch, bands, fpc;        // frames, bands per frame, frames per CTA
    int nwarps;                   // consumer warps per CTA
    int rc;                       // rows per stage
    uint32_t frame_stage_bytes;   // rc * W * 3
    uint32_t stage_bytes;         // fpc * frame_stage_bytes
    uint32_t coef[6]; int acc0;
};

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t thr16(const uint4 A, const uint4 B, const uint4 C, const BandStripParams& p) {
    const uint32_t w[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
    const uint32_t c0 = p.coef[0], c1a = p.coef[1], c1b = p.coef[2], c2a = p.coef[3], c2b = p.coef[4], c3 = p.coef[5];
    const int acc0 = p.acc0;
    uint32_t nb = 0;
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

}  // namespace

template <bool MASK>
__global__ void __launch_bounds__((kMaxWarps + 1) * 32) bgr_bandstrip_kernel(const BandStripParams p) {
    extern __shared__ __align__(128) uint8_t smem[];         // [4 KB table][2 * kStages mbarriers][stages]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int band = blockIdx.x % p.bands, fgroup = blockIdx.x / p.bands;
    const int f0 = fgroup * p.fpc, nf = min(p.fpc, p.batch - f0);
    const int y0 = band * kBand, nout = min(kBand, p.H - y0);
    const int ra = max(y0 - 2, 0), rb = min(y0 + nout + 2, p.H);      // image rows the band needs
    const int nchunks = (rb - ra + p.rc - 1) / p.rc;
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem + 4096);   // full[s] at +16 s, empty[s] at +16 s + 8
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem + 4096 + 128);
    strip::lut_init(smem, tid);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { strip::mbar_init(bars + 16 * s, 1); strip::mbar_init(bars + 16 * s + 8, p.nwarps); }
        strip::mbar_fence_init();
    }
    __syncthreads();

    if (warp == p.nwarps) {
        // ---- producer: one lane feeds the ring with one bulk copy per frame and chunk
        if (lane == 0) {
            const uint32_t rowbytes = (uint32_t)p.W * 3u;
            for (int k = 0; k < nchunks; ++k) {
                const int s = k % kStages;
                if (k >= kStages) strip::mbar_wait(bars + 16 * s + 8, ((k / kStages) - 1) & 1);   // consumers released the stage
                const int r0 = ra + k * p.rc, nr = min(p.rc, rb - r0);
                strip::mbar_expect_tx(bars + 16 * s, (uint32_t)nf * (uint32_t)nr * rowbytes);
                for (int f = 0; f < nf; ++f) {
                    const uint8_t* g = p.src + (size_t)(f0 + f) * p.frame_stride + (size_t)r0 * p.pitch;
                    const uint32_t dst = stage0 + s * p.stage_bytes + f * p.frame_stage_bytes;
                    if ((uint32_t)p.pitch == rowbytes) strip::bulk_g2s(dst, g, (uint32_t)nr * rowbytes, bars + 16 * s);
                    else for (int r = 0; r < nr; ++r) strip::bulk_g2s(dst + r * rowbytes, g + (size_t)r * p.pitch, rowbytes, bars + 16 * s);
                }
            }
        }
        return;
    }

    // ---- consumers: lane = one 16-pixel group of one of the CTA's frames
    const int slot = warp * 30 - 1 + lane;                   // lanes 0 and 31 only feed their neighbours
    const int nslots = nf * p.NC;
    const bool valid = slot >= 0 && slot < nslots;
    const int sc = min(max(slot, 0), nslots - 1);
    const int fr = sc / p.NC, c = sc - fr * p.NC;
    strip::CloseLane k;
    k.h0 = k.h1 = 0u; k.e0 = k.e1 = 0u;
    uint32_t inside = 0xfffffu;
    if (c == 0) inside &= 0xffffcu;
    if (c == p.NC - 1) inside &= 0x3ffffu;
    k.inside = valid ? inside : 0u;
    k.lut = strip::lut_base(smem);
    const int frame = f0 + fr;
    k.mrow = MASK ? p.mask + (size_t)frame * p.mask_frame_stride + (size_t)y0 * p.mask_pitch + (size_t)c * 16 : nullptr;
    k.brow = p.bits16 + ((size_t)frame * p.H + y0) * p.WB2 + c;
    const bool writer = valid && lane >= 1 && lane <= 30;
    const bool tail = writer && c == p.NC - 1 && p.WB2 > p.NC;
    const uint32_t lane_off = (uint32_t)fr * p.frame_stage_bytes + (uint32_t)c * 48u;
    const uint32_t rowbytes = (uint32_t)p.W * 3u;

    // rows y0-2 .. y0+nout+1 enter; rows outside the image enter as zeros
    int r = y0 - 2;
    for (; r < ra; ++r) strip::push_row<false, MASK>(k, p.WB2, p.mask_pitch, 0u, (unsigned)(r - 1) >= (unsigned)p.H, false, false);
    for (int kc = 0; kc < nchunks; ++kc) {
        const int s = kc % kStages;
        strip::mbar_wait(bars + 16 * s, (kc / kStages) & 1);
        const int r0 = ra + kc * p.rc, nr = min(p.rc, rb - r0);
        uint32_t a = stage0 + s * p.stage_bytes + lane_off;
        for (int i = 0; i < nr; ++i, a += rowbytes, ++r) {
            const uint4 A = strip::lds128(a), B = strip::lds128(a + 16u), C = strip::lds128(a + 32u);
            const uint32_t t = thr16(A, B, C, p);
            if (r >= y0 + 2) strip::push_row<true, MASK>(k, p.WB2, p.mask_pitch, t, false, writer, tail);
            else strip::push_row<false, MASK>(k, p.WB2, p.mask_pitch, t, (unsigned)(r - 1) >= (unsigned)p.H, false, false);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 16 * s + 8) : "memory");
    }
    for (; r < y0 + nout + 2; ++r) {
        const bool above = (unsigned)(r - 1) >= (unsigned)p.H;
        if (r >= y0 + 2) strip::push_row<true, MASK>(k, p.WB2, p.mask_pitch, 0u, above, writer, tail);
        else strip::push_row<false, MASK>(k, p.WB2, p.mask_pitch, 0u, above, false, false);
    }
}

// cudaErrorNotSupported when the call does not qualify (the caller then runs the band kernel of pixel.cu)
cudaError_t launch_bgr_bandstrip(const PixelLaunch& L, int sm_count, cudaStream_t st, int64_t* launches) {
    (void)sm_count;
    if ((L.W & 15) || L.W < 32 || L.H < 1) return cudaErrorNotSupported;
    if ((L.pitch & 15) || (L.frame_stride & 15) || (((size_t)L.src) & 15)) return cudaErrorNotSupported;
    if (L.mask && ((L.mask_pitch & 15) || (L.mask_frame_stride & 15) || (((size_t)L.mask) & 15))) return cudaErrorNotSupported;
    if (L.pitch > 0x7fffffffu || L.mask_pitch > 0x7fffffffu) return cudaErrorNotSupported;
    BandStripParams p;
    memset(&p, 0, sizeof(p));
    p.src = L.src; p.pitch = (int)L.pitch; p.frame_stride = L.frame_stride;
    p.mask = L.mask; p.mask_pitch = (int)L.mask_pitch; p.mask_frame_stride = L.mask_frame_stride;
    p.bits16 = reinterpret_cast<uint16_t*>(L.bits);
    p.W = L.W; p.H = L.H; p.NC = L.W / 16; p.WB2 = 2 * ((L.W + 31) / 32);
    p.batch = L.batch; p.bands = (L.H + kBand - 1) / kBand;
    p.fpc = 240 / p.NC; if (p.fpc < 1) p.fpc = 1; if (p.fpc > L.batch) p.fpc = L.batch;
    p.nwarps = (p.fpc * p.NC + 29) / 30;
    if (p.nwarps > kMaxWarps) return cudaErrorNotSupported;
    {
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
        if (L.lower_bound <= 0) { for (int i = 0; i < 6; ++i) p.coef[i] = 0; p.acc0 = 0; }
        else if (L.lower_bound > 255) { for (int i = 0; i < 6; ++i) p.coef[i] = 0; p.acc0 = -1; }
        else p.acc0 = -L.lower_bound;
    }
    p.rc = tuning().bandstrip_rc > 0 ? tuning().bandstrip_rc : 2;
    if (p.rc < 1) p.rc = 1;
    p.frame_stage_bytes = (uint32_t)p.rc * (uint32_t)L.W * 3u;
    p.stage_bytes = (uint32_t)p.fpc * p.frame_stage_bytes;
    const size_t smem = 4096 + 128 + (size_t)kStages * p.stage_bytes;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_smem) return cudaErrorNotSupported;
    const long long grid = (long long)p.bands * ((L.batch + p.fpc - 1) / p.fpc);
    if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorNotSupported;
    cudaError_t e;
    const int threads = (p.nwarps + 1) * 32;
    if (L.mask) {
        e = cudaFuncSetAttribute(bgr_bandstrip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        bgr_bandstrip_kernel<true><<<(unsigned)grid, threads, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(bgr_bandstrip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        bgr_bandstrip_kernel<false><<<(unsigned)grid, threads, smem, st>>>(p);
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
