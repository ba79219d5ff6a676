// K_E — emission for the labelling stage (runs + boundary-pixel records) from the bit-packed final mask.
// Replaces the scan half of cv::findContours (reference: src/imgproc.cpp:71-72): the foreground is cut into maximal
// horizontal runs, and every foreground pixel with a background 4-neighbour — the only pixels a contour can visit — is
// recorded with its 3x3 neighbourhood.
//
// One CTA = one band of BH rows of one frame, read from the bit mask the pixel kernel wrote (1/8 B per pixel, still in
// L2), plus one row above and below.  In raster order inside the band:
//   runs     every maximal horizontal run; rows[y] = (first, end) keeps every row addressable;
//   records  {x | y<<16, 8-neighbourhood | run index << 8} of every boundary pixel.
// Foreground is sparse, so the band's non-zero words are first compacted (ballots) into a raster-ordered list and
// everything else works on that list: run starts and boundary pixels are counted per entry, ranked by ONE block scan of
// the packed counts, and the band claims its ranges of the frame's arrays with ONE 64-bit atomicAdd (bands land in
// arrival order).  Records are fetched balanced: every thread takes "its" candidates by binary search over the
// per-entry prefix, so that a blob cap does not serialise on one thread.  Small CTAs with a few KB of shared memory: many are resident per SM, which hides
// the scan/atomic latency that used to sit in the tail of the HBM-bound pixel kernel.
#include "emit_core.cuh"

namespace rmcv {

// kVec: WB % 4 == 0 -> the band is loaded with 128-bit loads and the non-zero test rides on the load.
RMCV_GSTAMP_ARRAY(g_ns_emit)
template <bool kVec>
__global__ void __launch_bounds__(128) emit_kernel(const EmitParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    RMCV_GSTAMP_BEGIN(g_ns_emit, 0);
    __shared__ int s_wtot[4], s_base[2];
    constexpr int NT = 128, nwarps = 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y, band = blockIdx.x;
    const int H = p.H, WB = p.WB;
    const int y0 = band * p.BH;
    const int nout = min(p.BH, H - y0);
    chain_begin();
    chain_wait();
    if (p.band_flags != nullptr && p.band_flags[(size_t)frame * p.bands + band] == 0) {
        // the pixel kernel saw no foreground in this band's rows: no runs, no records, nothing to load
        int2* rows = p.rows + (size_t)frame * H;
        for (int j = threadIdx.x; j < nout; j += 128) rows[y0 + j] = make_int2(0, 0);
        RMCV_GSTAMP_END(g_ns_emit, 0);
        return;
    }
    const int nwords = nout * WB, cap = p.BH * WB;
    auto row_of = [&](int idx) -> int { return p.inv_wb ? (int)__umulhi((uint32_t)idx, p.inv_wb) : idx; };  // idx / WB
    long long* scratch = reinterpret_cast<long long*>(smem);        // 34 long longs of scan scratch
    int* erun = reinterpret_cast<int*>(scratch + 36);                // [cap + 1] runs before entry e (exclusive prefix)
    int* erec = erun + cap + 1;                                      // [cap + 1] records before entry e
    uint32_t* mm = reinterpret_cast<uint32_t*>(erec + cap + 1 + ((2 * (cap + 1)) & 3 ? 4 - ((2 * (cap + 1)) & 3) : 0));
                                                                     // [(BH+2)][WB] rows y0-1 .. y0+nout, 16-byte aligned
    uint32_t* m = mm + WB;                                           // row 0 of m <-> image row y0
    uint16_t* list = reinterpret_cast<uint16_t*>(mm + (size_t)(p.BH + 2) * WB);  // [cap] non-zero words, raster order
    const uint32_t* gb = p.bits + (size_t)frame * H * WB;
    // rows y0-1 .. y0+nout are contiguous in the bit mask; rows outside the image read as background
    const int lo_i = y0 > 0 ? 0 : WB, hi_i = y0 + nout < H ? (nout + 2) * WB : (nout + 1) * WB;
    const long long off = ((long long)y0 - 1) * WB;
    int pos = 0, n_ent = 0;
    if (kVec) {
        // ---- load + the non-zero words of the band in raster order (foreground is sparse: everything below works on
        // this list).  A warp owns a contiguous range of 16-byte groups; the non-zero nibbles stay in a register
        // between the counting round and the writing round.
        const uint4* src = reinterpret_cast<const uint4*>(gb + off);
        uint4* dst = reinterpret_cast<uint4*>(mm);
        const int nq = ((nout + 2) * WB) >> 2, lo_q = lo_i >> 2, hi_q = hi_i >> 2;
        const int own_lo = WB >> 2, own_hi = ((nout + 1) * WB) >> 2;
        const int per_w = (((nq + nwarps - 1) >> 2) + 31) & ~31;      // at most 256: eight rounds of nibbles
        const int q0 = min(nq, warp * per_w), q1 = min(nq, q0 + per_w);
        uint32_t nibs = 0;
        int mine = 0, it = 0;
        for (int i = q0 + lane; i < q1; i += 32, ++it) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (i >= lo_q && i < hi_q) v = __ldg(src + i);
            dst[i] = v;
            if (i >= own_lo && i < own_hi) {
                const uint32_t nib = (v.x != 0u) | ((v.y != 0u) << 1) | ((v.z != 0u) << 2) | ((v.w != 0u) << 3);
                nibs |= nib << (4 * it);
                mine += __popc(nib);
            }
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
    } else {
        for (int i = tid; i < (nout + 2) * WB; i += NT) mm[i] = (i >= lo_i && i < hi_i) ? __ldg(gb + (off + i)) : 0u;
        __syncthreads();
        const int chunk = (((nwords + nwarps - 1) >> 2) + 31) & ~31;  // words per warp, whole ballots
        const int c0 = min(nwords, warp * chunk), c1 = min(nwords, c0 + chunk);
        int mine = 0;
        for (int i = c0; i < c1; i += 32) {
            const int idx = i + lane;
            mine += __popc(__ballot_sync(0xffffffffu, idx < c1 && m[idx] != 0u));
        }
        if (lane == 0) s_wtot[warp] = mine;
        __syncthreads();
        for (int w = 0; w < nwarps; ++w) { if (w < warp) pos += s_wtot[w]; n_ent += s_wtot[w]; }
        for (int i = c0; i < c1; i += 32) {
            const int idx = i + lane;
            const bool nz = idx < c1 && m[idx] != 0u;
            const unsigned bal = __ballot_sync(0xffffffffu, nz);
            if (nz) list[pos + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)idx;
            pos += __popc(bal);
        }
    }
    emit_tail<NT>(p, frame, y0, nout, n_ent, m, list, erun, erec, scratch, s_base, tid);
    RMCV_GSTAMP_END(g_ns_emit, 0);
}

cudaError_t launch_emit(const EmitLaunch& L, cudaStream_t st, int64_t* launches, bool chained) {
    EmitParams p;
    p.bits = L.bits; p.W = L.W; p.H = L.H; p.WB = (L.W + 31) / 32;
    p.rows = L.rows; p.run_x = L.run_x; p.run_y = L.run_y; p.counters = L.counters; p.R = L.R;
    p.recs = L.recs; p.PC = L.PC;
    // band height: (BH + 2) * WB <= 4096 words keeps the per-warp nibble register within eight rounds and the word
    // indices within 16 bits; taller bands amortise the scan / claim / row-table steps
    const int bh_max = 4096 / p.WB - 2;
    // (chunks of a few frames: 8-row bands - four times the CTAs, each with a quarter of the scan / claim / write chain)
    int BH = L.batch <= small_batch_limit() ? 8 : 32;
    if (tuning().emit_bh > 0) BH = tuning().emit_bh;
    if (BH > bh_max) BH = bh_max;
    if (BH < 1) BH = 1;
    if (BH > L.H) BH = L.H;
    p.BH = BH; p.bands = (L.H + BH - 1) / BH;
    p.inv_wb = p.WB > 1 ? (uint32_t)((1ull << 32) / (unsigned)p.WB) + 1u : 0u;   // 0: WB == 1
    p.band_flags = (L.band_flags != nullptr && L.flag_bh == p.BH && L.flag_bands == p.bands) ? L.band_flags : nullptr;
    if (L.batch <= 0 || L.batch > 65535) return cudaErrorInvalidValue;
    const size_t cap = (size_t)BH * p.WB;
    const size_t smem = 36 * 8 + 2 * (cap + 1) * 4 + 16 + (size_t)(BH + 2) * p.WB * 4 + cap * 2 + 16;
    const bool vec = (p.WB & 3) == 0;
    size_t smem_l = smem;
    if (tuning().chain_pad > 0 && (size_t)tuning().chain_pad > smem_l) smem_l = (size_t)tuning().chain_pad;
    if (smem_l > 48 * 1024) {
        cudaError_t e = vec ? cudaFuncSetAttribute(emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l)
                            : cudaFuncSetAttribute(emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(p.bands, L.batch);
    const cudaError_t e = vec ? launch_chained<EmitParams>(emit_kernel<true>, grid, dim3(128), smem_l, st, chained, p)
                              : launch_chained<EmitParams>(emit_kernel<false>, grid, dim3(128), smem_l, st, chained, p);
    if (launches) ++*launches;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace rmcv
RMCV_GSTAMP_GETTER(rmcv_debug_ns_emit, rmcv::g_ns_emit)
