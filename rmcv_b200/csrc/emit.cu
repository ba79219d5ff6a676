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
#include "common.cuh"
#include "pairs.cuh"

namespace rmcv {

struct EmitParams {
    const uint32_t* bits;   // [frames][H][WB]
    int W, H, WB, BH, bands;
    uint32_t inv_wb;        // floor(2^32 / WB) + 1: idx / WB == umulhi(idx, inv_wb) for idx < 2^20; 0 when WB == 1
    int2* rows; uint32_t* run_x; uint16_t* run_y; FrameCounters* counters; int R;
    uint2* recs; int PC;
};

// kVec: WB % 4 == 0 -> the band is loaded with 128-bit loads and the non-zero test rides on the load.
template <bool kVec>
__global__ void __launch_bounds__(128) emit_kernel(const EmitParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_wtot[4], s_base[2];
    constexpr int NT = 128, nwarps = 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y, band = blockIdx.x;
    const int H = p.H, WB = p.WB;
    const int y0 = band * p.BH;
    const int nout = min(p.BH, H - y0);
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
    if (n_ent == 0) {   // an empty band (block-uniform): its rows hold no runs, nothing to rank or claim
        int2* rows = p.rows + (size_t)frame * H;
        for (int j = tid; j < nout; j += NT) rows[y0 + j] = make_int2(0, 0);
        return;
    }
    __syncthreads();
    auto boundary_word = [&](int idx, int k) -> uint32_t {
        const uint32_t* c = m + idx;
        const uint32_t w = c[0];
        const uint32_t prev = k > 0 ? (c[-1] >> 31) : 0u;
        const uint32_t next = k + 1 < WB ? (c[1] & 1u) : 0u;
        return w & ~(c[-WB] & c[WB] & ((w << 1) | prev) & ((w >> 1) | (next << 31)));
    };
    // ---- per entry: run starts and boundary pixels, ranked by a block scan of the packed counts
    long long carry = 0;
    for (int e0 = 0; e0 < n_ent; e0 += NT) {
        const int e = e0 + tid;
        long long v = 0;
        if (e < n_ent) {
            const int idx = list[e], k = idx - row_of(idx) * WB;
            const uint32_t w = m[idx];
            const uint32_t prev = k > 0 ? (m[idx - 1] >> 31) : 0u;
            v = (long long)__popc(w & ~((w << 1) | prev)) | ((long long)__popc(boundary_word(idx, k)) << 32);
        }
        long long total;
        const long long ex = block_excl_scan64(v, &total, scratch) + carry;
        if (e < n_ent) { erun[e] = (int)(ex & 0xffffffffll); erec[e] = (int)(ex >> 32); }
        carry += total;
        __syncthreads();   // scratch is reused by the next round
    }
    const int run_total = (int)(carry & 0xffffffffll), rec_total = (int)(carry >> 32);
    if (tid == 0) {
        erun[n_ent] = run_total; erec[n_ent] = rec_total;
        static_assert(offsetof(FrameCounters, n_recs) == offsetof(FrameCounters, n_runs) + 4, "n_runs/n_recs must pack into 64 bits");
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(&p.counters[frame].n_runs),
                                                 (unsigned long long)carry);
        s_base[0] = (int)(old & 0xffffffffull);
        s_base[1] = (int)(old >> 32);
    }
    __syncthreads();
    const int run_base = s_base[0], rec_base = s_base[1];
    // ---- rows[y] = (first, end) of the row's runs: entries are sorted by word index, so a row is a range of entries
    {
        int2* rows = p.rows + (size_t)frame * H;
        for (int j = tid; j < nout; j += NT) {
            int lo = 0, hi = n_ent;                  // first entry with idx >= j*WB
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)list[mid] < j * WB) lo = mid + 1; else hi = mid; }
            const int first = lo;
            hi = n_ent;                              // first entry with idx >= (j+1)*WB
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)list[mid] < (j + 1) * WB) lo = mid + 1; else hi = mid; }
            rows[y0 + j] = make_int2(run_base + erun[first], run_base + erun[lo]);
        }
    }
    // ---- runs: xs at the start bits, xe at the end bits (a run entering from the previous word is still open)
    {
        const int R = p.R;
        uint16_t* run_x16 = reinterpret_cast<uint16_t*>(p.run_x + (size_t)frame * R);
        uint16_t* run_y = p.run_y + (size_t)frame * R;
        for (int e = tid; e < n_ent; e += NT) {
            const int idx = list[e], j = row_of(idx), k = idx - j * WB;
            const uint32_t w = m[idx];
            const uint32_t prev = k > 0 ? (m[idx - 1] >> 31) : 0u;
            const uint32_t next = k + 1 < WB ? (m[idx + 1] & 1u) : 0u;
            uint32_t starts = w & ~((w << 1) | prev);
            uint32_t ends = w & ~((w >> 1) | (next << 31));
            int rs = run_base + erun[e];
            int re = rs - (int)(prev & w & 1u);
            const int y = y0 + j;
            while (starts) {
                const int b = __ffs(starts) - 1;
                starts &= starts - 1;
                if (rs < R) { run_x16[2 * rs] = (uint16_t)(k * 32 + b); run_y[rs] = (uint16_t)y; }
                ++rs;
            }
            while (ends) {
                const int b = __ffs(ends) - 1;
                ends &= ends - 1;
                if (re < R) run_x16[2 * re + 1] = (uint16_t)(k * 32 + b);
                ++re;
            }
        }
    }
    // ---- records, balanced: thread q takes candidate q (binary search over the per-entry prefix)
    if (p.recs != nullptr) {
        uint2* recs = p.recs + (size_t)frame * p.PC;
        for (int q = tid; q < rec_total; q += NT) {
            int lo = 0, hi = n_ent;               // last entry with erec <= q
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (erec[mid] <= q) lo = mid; else hi = mid;
            }
            const int idx = list[lo], j = row_of(idx), k = idx - j * WB;
            uint32_t b = boundary_word(idx, k);
            for (int skip = q - erec[lo]; skip > 0; --skip) b &= b - 1;
            const int i = __ffs(b) - 1;
            const uint32_t* c = m + idx;
            auto win = [&](const uint32_t* r) -> uint64_t {
                const uint32_t prev = k > 0 ? (r[-1] >> 31) : 0u;
                const uint32_t next = k + 1 < WB ? (r[1] & 1u) : 0u;
                return (uint64_t)prev | ((uint64_t)r[0] << 1) | ((uint64_t)next << 33);
            };
            const uint32_t u3 = (uint32_t)(win(c - WB) >> i) & 7u, c3 = (uint32_t)(win(c) >> i) & 7u, d3 = (uint32_t)(win(c + WB) >> i) & 7u;
            const uint32_t nb = u3 | ((c3 & 1u) << 3) | ((c3 >> 2) << 4) | (d3 << 5);
            // the run that owns the pixel: the last one started at or before bit i, else the one entering from the previous word
            const uint32_t w = c[0];
            const uint32_t pv = k > 0 ? (c[-1] >> 31) : 0u;
            const uint32_t starts = w & ~((w << 1) | pv);
            const int run = run_base + erun[lo] + __popc(starts & (0xffffffffu >> (31 - i))) - 1;
            if (rec_base + q < p.PC)
                recs[rec_base + q] = make_uint2((uint32_t)(k * 32 + i) | ((uint32_t)(y0 + j) << 16), nb | ((uint32_t)run << 8));
        }
    }
}

cudaError_t launch_emit(const EmitLaunch& L, cudaStream_t st, int64_t* launches) {
    EmitParams p;
    p.bits = L.bits; p.W = L.W; p.H = L.H; p.WB = (L.W + 31) / 32;
    p.rows = L.rows; p.run_x = L.run_x; p.run_y = L.run_y; p.counters = L.counters; p.R = L.R;
    p.recs = L.recs; p.PC = L.PC;
    // band height: (BH + 2) * WB <= 4096 words keeps the per-warp nibble register within eight rounds and the word
    // indices within 16 bits; taller bands amortise the scan / claim / row-table steps
    const int bh_max = 4096 / p.WB - 2;
    int BH = 32;
    if (tuning().emit_bh > 0) BH = tuning().emit_bh;
    if (BH > bh_max) BH = bh_max;
    if (BH < 1) BH = 1;
    if (BH > L.H) BH = L.H;
    p.BH = BH; p.bands = (L.H + BH - 1) / BH;
    p.inv_wb = p.WB > 1 ? (uint32_t)((1ull << 32) / (unsigned)p.WB) + 1u : 0u;   // 0: WB == 1
    if (L.batch <= 0 || L.batch > 65535) return cudaErrorInvalidValue;
    const size_t cap = (size_t)BH * p.WB;
    const size_t smem = 36 * 8 + 2 * (cap + 1) * 4 + 16 + (size_t)(BH + 2) * p.WB * 4 + cap * 2 + 16;
    const bool vec = (p.WB & 3) == 0;
    if (smem > 48 * 1024) {
        cudaError_t e = vec ? cudaFuncSetAttribute(emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(p.bands, L.batch);
    if (vec) emit_kernel<true><<<grid, 128, smem, st>>>(p);
    else emit_kernel<false><<<grid, 128, smem, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
