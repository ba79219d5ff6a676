// K2/K3 — connected-component labelling of the bit-packed mask; replaces the component discovery half of
// cv::findContours(RETR_EXTERNAL) (reference: src/imgproc.cpp:71-72; semantics SURVEY A.2-A.5).
//
// Run based: the mask is ~2 % foreground, so everything after the pixel stage works on horizontal runs.
//   runs_kernel     one CTA per frame: per-row run counts (popc of run-start bits), block scan -> row offsets,
//                   then emits (xs,xe,y) per run and initialises the two union-find forests.
//   union_kernel    one thread per run: 8-connected foreground unions with the previous row; 4-connected unions
//                   of the background gap to the left of the run with the gaps of the rows above and below
//                   (node 0 = "outer" background = connected to the image border).
//   flatten_kernel  path-compress both forests; per-root bbox / first-pixel via atomics; enumerate components;
//                   paint hole gaps (background not connected to the border) into the `hole` bit plane so that
//                   the contour pass can tell hole borders and nested components from external borders.
// Lock-free union-find with atomicMin (Playne & Hawick style): roots only ever decrease.
#include "common.cuh"

namespace rmcv {

__device__ __forceinline__ int uf_find(const int32_t* parent, int x) {
    while (true) {
        const int p = *reinterpret_cast<const volatile int32_t*>(parent + x);
        if (p == x) return x;
        x = p;
    }
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under b
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;  // somebody else re-parented a first; retry from there
    }
}

// ------------------------------------------------------------------------------------------ runs
__global__ void __launch_bounds__(1024) runs_kernel(Geometry g, SlotBuffers sb) {
    extern __shared__ int32_t sh_cnt[];  // [H] run count per row, then exclusive offsets
    __shared__ int32_t sh_warp[32];
    __shared__ int32_t sh_total;
    const int frame = blockIdx.x;
    const int W = g.W, H = g.H, WB = g.WB, R = g.R;
    const uint32_t* bits = sb.bits + (size_t)frame * H * WB;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;

    // pass 1: run starts per row
    for (int y = warp; y < H; y += nwarps) {
        const uint32_t* row = bits + (size_t)y * WB;
        int cnt = 0;
        for (int k = lane; k < WB; k += 32) {
            const uint32_t w = row[k];
            const uint32_t prev = k > 0 ? (row[k - 1] >> 31) : 0u;
            cnt += __popc(w & ~((w << 1) | prev));
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) sh_cnt[y] = cnt;
    }
    __syncthreads();
    // block exclusive scan over H rows: each thread owns a contiguous slice
    const int per = (H + NT - 1) / NT;
    const int b0 = min(H, tid * per), b1 = min(H, b0 + per);
    int local = 0;
    for (int y = b0; y < b1; ++y) local += sh_cnt[y];
    int incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) sh_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = lane < nwarps ? sh_warp[lane] : 0;
        int inc2 = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += u;
        }
        sh_warp[lane] = inc2 - v;  // exclusive warp offsets
        if (lane == 31) sh_total = inc2;
    }
    __syncthreads();
    int run = sh_warp[warp] + incl - local;
    int32_t* row_off = sb.row_off + (size_t)frame * (H + 1);
    for (int y = b0; y < b1; ++y) {
        const int c = sh_cnt[y];
        sh_cnt[y] = run;
        row_off[y] = run;
        run += c;
    }
    __syncthreads();
    const int total = sh_total;
    if (tid == 0) {
        row_off[H] = total;
        FrameCounters& fc = sb.counters[frame];
        fc.n_runs = min(total, R);
        fc.n_comps = 0;
        fc.n_holes = 0;
        fc.flags = total > R ? RMCV_FRAME_OVERFLOW_RUNS : 0;
        fc.n_contours = fc.n_positive = fc.n_negative = fc.n_armours = 0;
        sb.gparent[(size_t)frame * (R + 1)] = 0;
    }
    // pass 2: emit runs.  Start bits and end bits are ranked independently; the k-th start pairs with the k-th end.
    uint16_t* run_x16 = reinterpret_cast<uint16_t*>(sb.run_x + (size_t)frame * R);
    int32_t* run_y = sb.run_y + (size_t)frame * R;
    int32_t* parent = sb.parent + (size_t)frame * R;
    int32_t* gparent = sb.gparent + (size_t)frame * (R + 1);
    RunStat* rstat = sb.rstat + (size_t)frame * R;
    for (int y = warp; y < H; y += nwarps) {
        const uint32_t* row = bits + (size_t)y * WB;
        const int off = sh_cnt[y];
        int carry_s = 0, carry_e = 0;
        for (int base = 0; base < WB; base += 32) {
            const int k = base + lane;
            uint32_t w = 0, prev = 0, next = 0;
            if (k < WB) {
                w = row[k];
                prev = k > 0 ? (row[k - 1] >> 31) : 0u;
                next = k + 1 < WB ? (row[k + 1] & 1u) : 0u;
            }
            uint32_t starts = w & ~((w << 1) | prev);
            uint32_t ends = w & ~((w >> 1) | (next << 31));
            const int ns = __popc(starts), ne = __popc(ends);
            int is = ns, ie = ne;
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, is, o);
                const int b = __shfl_up_sync(0xffffffffu, ie, o);
                if (lane >= o) { is += a; ie += b; }
            }
            int rs = off + carry_s + is - ns;
            int re = off + carry_e + ie - ne;
            while (starts) {
                const int bpos = __ffs(starts) - 1;
                starts &= starts - 1;
                if (rs < R) {
                    const int xs = k * 32 + bpos;
                    run_x16[2 * rs] = (uint16_t)xs;
                    run_y[rs] = y;
                    parent[rs] = rs;
                    // gap to the left of this run: outer when it touches the border (first run of the row, or
                    // a run in the first/last image row); otherwise its own node rs+1
                    const bool first_in_row = (rs == off);
                    gparent[rs + 1] = (first_in_row || y == 0 || y == H - 1) ? 0 : rs + 1;
                    rstat[rs].x0 = xs; rstat[rs].y0 = y; rstat[rs].y1 = y; rstat[rs].firstkey = y * W + xs;
                }
                ++rs;
            }
            while (ends) {
                const int bpos = __ffs(ends) - 1;
                ends &= ends - 1;
                if (re < R) {
                    const int xe = k * 32 + bpos;
                    run_x16[2 * re + 1] = (uint16_t)xe;
                    rstat[re].x1 = xe;
                }
                ++re;
            }
            carry_s += __shfl_sync(0xffffffffu, is, 31);
            carry_e += __shfl_sync(0xffffffffu, ie, 31);
        }
    }
}

// ------------------------------------------------------------------------------------------ unions
struct FrameView {
    const uint32_t* run_x; const int32_t* run_y; const int32_t* row_off;
    int32_t* parent; int32_t* gparent;
    int n_runs, W, H;
};

__device__ __forceinline__ int row_begin(const FrameView& f, int y, int n_runs) { return min(f.row_off[y], n_runs); }

// first run index in [lo,hi) with xe >= x
__device__ __forceinline__ int lower_bound_xe(const uint32_t* run_x, int lo, int hi, int x) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(run_x[mid] >> 16) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// number of runs in [lo,hi) with xs <= x  (returned as an index: first run with xs > x)
__device__ __forceinline__ int upper_bound_xs(const uint32_t* run_x, int lo, int hi, int x) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(run_x[mid] & 0xffffu) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Unions the interior gap node `gnode` = [a,b] of row y with the 4-connected gaps of row yy.
__device__ __forceinline__ void gap_union_row(const FrameView& f, int32_t* gparent, int gnode, int a, int b, int yy) {
    const int lo = row_begin(f, yy, f.n_runs), hi = row_begin(f, yy + 1, f.n_runs);
    // gap kk of row yy lies between run lo+kk-1 and run lo+kk.  First gap whose end >= a:
    int idx = upper_bound_xs(f.run_x, lo, hi, a);  // runs with xs <= a precede; gap index = idx - lo
    // walk gaps while gap.start <= b
    for (int r = idx;; ++r) {
        // gap between run r-1 and run r (r == lo: left border gap; r == hi: right border gap)
        const int ga = (r == lo) ? 0 : (int)(f.run_x[r - 1] >> 16) + 1;
        if (ga > b) break;
        const int gb = (r == hi) ? f.W - 1 : (int)(f.run_x[r] & 0xffffu) - 1;
        if (gb >= a && ga <= gb) {
            const bool border = (r == lo) || (r == hi) || yy == 0 || yy == f.H - 1;
            uf_union(gparent, gnode, border ? 0 : r + 1);
        }
        if (r == hi) break;
    }
}

__global__ void __launch_bounds__(256) union_kernel(Geometry g, SlotBuffers sb) {
    const int frame = blockIdx.y;
    const int R = g.R, H = g.H;
    FrameView f;
    f.run_x = sb.run_x + (size_t)frame * R;
    f.run_y = sb.run_y + (size_t)frame * R;
    f.row_off = sb.row_off + (size_t)frame * (H + 1);
    f.parent = sb.parent + (size_t)frame * R;
    f.gparent = sb.gparent + (size_t)frame * (R + 1);
    f.n_runs = sb.counters[frame].n_runs;
    f.W = g.W; f.H = H;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < f.n_runs; r += gridDim.x * blockDim.x) {
        const uint32_t rx = f.run_x[r];
        const int xs = (int)(rx & 0xffffu), xe = (int)(rx >> 16), y = f.run_y[r];
        // ---- foreground, 8-connectivity: runs of row y-1 overlapping [xs-1, xe+1]
        if (y > 0) {
            const int lo = row_begin(f, y - 1, f.n_runs), hi = row_begin(f, y, f.n_runs);
            for (int p = lower_bound_xe(f.run_x, lo, hi, xs - 1); p < hi; ++p) {
                if ((int)(f.run_x[p] & 0xffffu) > xe + 1) break;
                uf_union(f.parent, r, p);
            }
        }
        // ---- background gap to the left of the run, 4-connectivity, rows y-1 and y+1
        const int row_lo = row_begin(f, y, f.n_runs);
        if (r > row_lo && y > 0 && y < H - 1) {
            const int a = (int)(f.run_x[r - 1] >> 16) + 1, b = xs - 1;
            gap_union_row(f, f.gparent, r + 1, a, b, y - 1);
            gap_union_row(f, f.gparent, r + 1, a, b, y + 1);
        }
    }
}

// ------------------------------------------------------------------------------------------ flatten
__global__ void __launch_bounds__(256) flatten_kernel(Geometry g, SlotBuffers sb) {
    const int frame = blockIdx.y;
    const int R = g.R, H = g.H, WB = g.WB, C = g.C;
    const uint32_t* run_x = sb.run_x + (size_t)frame * R;
    const int32_t* run_y = sb.run_y + (size_t)frame * R;
    const int32_t* row_off = sb.row_off + (size_t)frame * (H + 1);
    int32_t* parent = sb.parent + (size_t)frame * R;
    int32_t* gparent = sb.gparent + (size_t)frame * (R + 1);
    RunStat* rstat = sb.rstat + (size_t)frame * R;
    uint32_t* hole = sb.hole + (size_t)frame * H * WB;
    FrameCounters& fc = sb.counters[frame];
    const int n_runs = fc.n_runs;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
        const int root = uf_find(parent, r);
        if (root != r) {
            parent[r] = root;
            const uint32_t rx = run_x[r];
            const int xs = (int)(rx & 0xffffu), xe = (int)(rx >> 16), y = run_y[r];
            RunStat* s = rstat + root;
            atomicMin(&s->x0, xs); atomicMax(&s->x1, xe);
            atomicMin(&s->y0, y); atomicMax(&s->y1, y);
            atomicMin(&s->firstkey, y * g.W + xs);
        } else {
            const int cid = atomicAdd(&fc.n_comps, 1);
            if (cid < C) sb.comp_root[(size_t)frame * C + cid] = r;
            else atomicOr(&fc.flags, RMCV_FRAME_OVERFLOW_BLOBS);
        }
        // gap to the left
        const int groot = uf_find(gparent, r + 1);
        gparent[r + 1] = groot;
        if (groot != 0) {  // hole: paint [xe_prev+1, xs-1] of row y
            const int y = run_y[r];
            const int lo = min(row_off[y], n_runs);
            if (r > lo) {
                const int a = (int)(run_x[r - 1] >> 16) + 1, b = (int)(run_x[r] & 0xffffu) - 1;
                for (int k = a >> 5; k <= (b >> 5); ++k) {
                    const int l = max(a, k * 32) - k * 32, h = min(b, k * 32 + 31) - k * 32;
                    const uint32_t m = (h == 31 ? 0xffffffffu : ((1u << (h + 1)) - 1u)) & ~((1u << l) - 1u);
                    atomicOr(hole + (size_t)y * WB + k, m);
                }
                atomicAdd(&fc.n_holes, 1);
            }
        }
    }
}

// Restores the all-zero invariant of the hole plane after the blob pass.
__global__ void __launch_bounds__(256) unpaint_kernel(Geometry g, SlotBuffers sb) {
    const int frame = blockIdx.y;
    const int R = g.R, H = g.H, WB = g.WB;
    FrameCounters& fc = sb.counters[frame];
    if (fc.n_holes == 0) return;
    const uint32_t* run_x = sb.run_x + (size_t)frame * R;
    const int32_t* run_y = sb.run_y + (size_t)frame * R;
    const int32_t* row_off = sb.row_off + (size_t)frame * (H + 1);
    const int32_t* gparent = sb.gparent + (size_t)frame * (R + 1);
    uint32_t* hole = sb.hole + (size_t)frame * H * WB;
    const int n_runs = fc.n_runs;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
        if (gparent[r + 1] == 0) continue;
        const int y = run_y[r];
        const int lo = min(row_off[y], n_runs);
        if (r <= lo) continue;
        const int a = (int)(run_x[r - 1] >> 16) + 1, b = (int)(run_x[r] & 0xffffu) - 1;
        for (int k = a >> 5; k <= (b >> 5); ++k) hole[(size_t)y * WB + k] = 0u;
    }
}

static int blocks_per_frame(const Geometry& g) {
    long long px = (long long)g.W * g.H;
    int b = (int)(px / (256LL * 1024LL));
    return b < 2 ? 2 : (b > 64 ? 64 : b);
}

cudaError_t launch_runs(const LabelLaunch& L, cudaStream_t st, int64_t* launches) {
    const size_t smem = (size_t)L.g.H * sizeof(int32_t);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        cudaError_t e = cudaFuncSetAttribute(runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    runs_kernel<<<L.frames, 1024, smem, st>>>(L.g, *L.sb);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_label(const LabelLaunch& L, cudaStream_t st, int64_t* launches) {
    dim3 grid(blocks_per_frame(L.g), L.frames);
    union_kernel<<<grid, 256, 0, st>>>(L.g, *L.sb);
    flatten_kernel<<<grid, 256, 0, st>>>(L.g, *L.sb);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_unpaint(const LabelLaunch& L, cudaStream_t st, int64_t* launches) {
    dim3 grid(blocks_per_frame(L.g), L.frames);
    unpaint_kernel<<<grid, 256, 0, st>>>(L.g, *L.sb);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
