// Shared by the emit kernel (emit.cu) and the fused pixel+emit kernel (pixel.cu): what happens once a band's final mask rows
// (with one row above and below) sit in shared memory and its non-zero words have been listed in raster order.
#pragma once
#include "common.cuh"
#include "pairs.cuh"

namespace rmcv {

struct EmitParams {
    const uint32_t* bits;   // [frames][H][WB]
    int W, H, WB, BH, bands;
    uint32_t inv_wb;        // floor(2^32 / WB) + 1: idx / WB == umulhi(idx, inv_wb) for idx < 2^20; 0 when WB == 1
    int2* rows; uint32_t* run_x; uint16_t* run_y; FrameCounters* counters; int R;
    uint2* recs; int PC;
    const uint8_t* band_flags;   // one byte per (frame, band) or null: 0 = the band's own rows are all background
};


// m: row 0 <-> image row y0 (rows -1 and nout are readable); list[n_ent]: indices (row * WB + word) of the band's non-zero
// words in raster order; erun / erec: [n_ent + 1] ints; scratch: 34 long longs; s_base: 2 ints.  All shared memory.
// Block-wide (NT threads, all of them call it); n_ent is block-uniform.
template <int NT>
__device__ __forceinline__ void emit_tail(const EmitParams& p, int frame, int y0, int nout, int n_ent, const uint32_t* m,
                                          const uint16_t* list, int* erun, int* erec, long long* scratch, int* s_base, int tid) {
    const int H = p.H, WB = p.WB;
    auto row_of = [&](int idx) -> int { return p.inv_wb ? (int)__umulhi((uint32_t)idx, p.inv_wb) : idx; };  // idx / WB
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

}  // namespace rmcv
