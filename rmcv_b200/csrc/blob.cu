// K4 — per-component contour statistics, ellipse fit and light-blob gates.  Replaces, per external contour,
//   contour.size() / cv::contourArea / cv::fitEllipseDirect / the ratio+tilt gates / rm::lightblob ctor
//   (reference: src/objdetect.cpp:62-84, src/core.cpp:9-19) without ever materialising the contour.
//
// One warp per 8-connected component.  Lanes take the rows of the component's bounding box; for every run of
// the component the boundary pixels are visited bit-parallel (candidates = run & ~(N&S&W&E)), their 3x3
// neighbourhood indexes a 256-entry arc table and every emitted arc is one contour point with
// multiplicity plus one directed edge for the shoelace sum.  Arcs whose background side is a hole (A.4) are
// skipped, so nested components end with n == 0 and hole borders never contribute.
//   pass A: n, sum x, sum y (exact ints), shoelace cross sum (int64)
//   pass B: L1 spread s and the 14 centred moment sums in fp64 -> direct fit (Halir-Flusser)
//   pass C: only when |det M| <= 1e-10: float-centred sums -> fitEllipseNoDirect (OpenCV's fallback)
// Warp-shuffle reductions, no atomics, deterministic.
#include "blob_math.cuh"
#include "common.cuh"

namespace rmcv {

// Arc table indexed by  NW | N<<1 | NE<<2 | W<<3 | E<<4 | SW<<5 | S<<6 | SE<<7  (bit set = foreground).
// Entry: bits 0-2 arc count m; arc i at bits 3+5i: low 2 bits = 4-neighbour to test for "hole" (0=E,1=N,2=W,3=S),
// high 3 bits = direction of the edge target q (0..7 = E,NE,N,NW,W,SW,S,SE); bit 31 = isolated pixel (no edge).
__constant__ uint32_t c_arc_lut[256];
__constant__ int8_t c_dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int8_t c_dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};

void upload_luts() {
    uint32_t lut[256];
    for (int idx = 0; idx < 256; ++idx) {
        // un-permute to direction order E,NE,N,NW,W,SW,S,SE
        int fg[8];
        fg[3] = (idx >> 0) & 1; fg[2] = (idx >> 1) & 1; fg[1] = (idx >> 2) & 1;
        fg[4] = (idx >> 3) & 1; fg[0] = (idx >> 4) & 1;
        fg[5] = (idx >> 5) & 1; fg[6] = (idx >> 6) & 1; fg[7] = (idx >> 7) & 1;
        int any = 0;
        for (int k = 0; k < 8; ++k) any |= fg[k];
        if (!any) { lut[idx] = (1u << 31) | 1u; continue; }
        int start = 0;
        while (!fg[start]) ++start;
        uint32_t v = 0; int m = 0;
        int k = (start + 1) & 7, steps = 0;
        while (steps < 8) {
            if (!fg[k]) {
                int four = -1;
                while (!fg[k]) {
                    if ((k & 1) == 0 && four < 0) four = k >> 1;
                    k = (k + 1) & 7; ++steps;
                }
                if (four >= 0) { v |= (uint32_t)(four | (k << 2)) << (3 + 5 * m); ++m; }
            } else {
                k = (k + 1) & 7; ++steps;
            }
        }
        lut[idx] = v | (uint32_t)m;
    }
    cudaMemcpyToSymbol(c_arc_lut, lut, sizeof(lut));
}

__device__ __forceinline__ uint64_t window(const uint32_t* row, int k, int WB) {
    const uint32_t w = row[k];
    const uint32_t prev = k > 0 ? (row[k - 1] >> 31) : 0u;
    const uint32_t next = k + 1 < WB ? (row[k + 1] & 1u) : 0u;
    return (uint64_t)prev | ((uint64_t)w << 1) | ((uint64_t)next << 33);
}

// Calls emit(x, y, dx, dy) for every contour point contributed by the run [xs,xe] of row y.
template <class F>
__device__ __forceinline__ void run_contour_points(const uint32_t* bits, const uint32_t* hole, int H, int WB, int y,
                                                   int xs, int xe, F&& emit) {
    const uint32_t* rc = bits + (size_t)y * WB;
    const uint32_t* ru = bits + (size_t)(y - 1) * WB;
    const uint32_t* rd = bits + (size_t)(y + 1) * WB;
    for (int k = xs >> 5; k <= (xe >> 5); ++k) {
        const int l = max(xs, k * 32) - k * 32, h = min(xe, k * 32 + 31) - k * 32;
        const uint32_t runmask = (h == 31 ? 0xffffffffu : ((1u << (h + 1)) - 1u)) & ~((1u << l) - 1u);
        const uint64_t cw = window(rc, k, WB);
        const uint64_t uw = y > 0 ? window(ru, k, WB) : 0ull;
        const uint64_t dw = y < H - 1 ? window(rd, k, WB) : 0ull;
        uint32_t cand = runmask & ~((uint32_t)(uw >> 1) & (uint32_t)(dw >> 1) & (uint32_t)cw & (uint32_t)(cw >> 2));
        uint64_t hc = 0, hu = 0, hd = 0;
        if (hole != nullptr && cand) {
            hc = window(hole + (size_t)y * WB, k, WB);
            hu = y > 0 ? window(hole + (size_t)(y - 1) * WB, k, WB) : 0ull;
            hd = y < H - 1 ? window(hole + (size_t)(y + 1) * WB, k, WB) : 0ull;
        }
        while (cand) {
            const int i = __ffs(cand) - 1;
            cand &= cand - 1;
            const uint32_t u3 = (uint32_t)(uw >> i) & 7u, c3 = (uint32_t)(cw >> i) & 7u, d3 = (uint32_t)(dw >> i) & 7u;
            const uint32_t idx = u3 | ((c3 & 1u) << 3) | ((c3 >> 2) << 4) | (d3 << 5);
            const uint32_t ent = c_arc_lut[idx];
            const int m = ent & 7;
            const bool iso = (ent >> 31) != 0;
            // hole flags of the 4-neighbours: E, N, W, S
            const uint32_t h4 = ((uint32_t)(hc >> (i + 2)) & 1u) | (((uint32_t)(hu >> (i + 1)) & 1u) << 1) |
                                (((uint32_t)(hc >> i) & 1u) << 2) | (((uint32_t)(hd >> (i + 1)) & 1u) << 3);
            const int x = k * 32 + i;
            for (int a = 0; a < m; ++a) {
                const uint32_t arc = (ent >> (3 + 5 * a)) & 31u;
                if ((h4 >> (arc & 3u)) & 1u) continue;
                const int q = arc >> 2;
                emit(x, y, iso ? 0 : (int)c_dx[q], iso ? 0 : (int)c_dy[q]);
            }
        }
    }
}

// Visits every run of component `root` inside its bbox rows, lanes striding over rows.
template <class F>
__device__ __forceinline__ void component_points(const uint32_t* bits, const uint32_t* hole, const uint32_t* run_x,
                                                 const int32_t* parent, const int32_t* row_off, int n_runs, int H,
                                                 int WB, int root, int bx0, int by0, int bx1, int by1, int lane,
                                                 F&& emit) {
    for (int y = by0 + lane; y <= by1; y += 32) {
        const int lo = min(row_off[y], n_runs), hi = min(row_off[y + 1], n_runs);
        for (int r = lo; r < hi; ++r) {
            const uint32_t rx = run_x[r];
            const int xs = (int)(rx & 0xffffu), xe = (int)(rx >> 16);
            if (xe < bx0) continue;
            if (xs > bx1) break;
            if (parent[r] != root) continue;
            run_contour_points(bits, hole, H, WB, y, xs, xe, emit);
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void warp_sum(Moments& m) {
    m.n = warp_sum(m.n);
    m.x = warp_sum(m.x); m.y = warp_sum(m.y);
    m.xx = warp_sum(m.xx); m.xy = warp_sum(m.xy); m.yy = warp_sum(m.yy);
    m.xxx = warp_sum(m.xxx); m.xxy = warp_sum(m.xxy); m.xyy = warp_sum(m.xyy); m.yyy = warp_sum(m.yyy);
    m.xxxx = warp_sum(m.xxxx); m.xxxy = warp_sum(m.xxxy); m.xxyy = warp_sum(m.xxyy);
    m.xyyy = warp_sum(m.xyyy); m.yyyy = warp_sum(m.yyyy);
}

// Decision + fit shared by the detect path (points regenerated from the mask) and rm::filter_lightblobs on
// caller-supplied contours.  `pass(fn)` must call fn(x, y) for every contour point (with multiplicity),
// partitioned over the lanes of the warp.
template <class PassFn>
__device__ __forceinline__ void fit_and_gate(int n, long long sum_x, long long sum_y, long long cross, const rmcv_params& prm,
                                             PassFn&& pass, int* status, int* branch, float* det0_out,
                                             rmcv_rotated_rect* ell, rmcv_lightblob* blob) {
    *status = RMCV_CONTOUR_SKIPPED;
    *branch = RMCV_FIT_NONE;
    *det0_out = 0.f;
    ell->cx = ell->cy = ell->w = ell->h = ell->angle = 0.f;
    const long long area2 = cross < 0 ? -cross : cross;
    const double area = (double)area2 * 0.5;
    if (n < 6 || !(area >= prm.area_min && area <= prm.area_max)) return;  // src/objdetect.cpp:64
    // ---- direct branch (centre in double)
    const double cx = (double)sum_x / (double)n, cy = (double)sum_y / (double)n;
    Moments m;
    moments_zero(m);
    double s = 0.0;
    pass([&](int x, int y) {
        const double dx = (double)x - cx, dy = (double)y - cy;
        s += fabs(dx) + fabs(dy);
        moments_add(m, dx, dy);
    });
    s = warp_sum(s);
    warp_sum(m);
    double scale = 100.0 / (s > RMCV_FLT_EPSILON ? s : RMCV_FLT_EPSILON);
    double det = 0.0;
    bool ok = direct_fit(m, scale, cx, cy, ell, &det);
    *det0_out = (float)det;
    if (ok) {
        *branch = RMCV_FIT_DIRECT;
    } else {
        // ---- fallback branch: cv::fitEllipseNoDirect keeps the centre and the centred points in float
        const float c32x = __fdiv_rn((float)sum_x, (float)n), c32y = __fdiv_rn((float)sum_y, (float)n);
        moments_zero(m);
        double s2 = 0.0;
        pass([&](int x, int y) {
            const float fx = __fsub_rn((float)x, c32x), fy = __fsub_rn((float)y, c32y);
            s2 += (double)__fadd_rn(fabsf(fx), fabsf(fy));
            moments_add(m, (double)fx, (double)fy);
        });
        s2 = warp_sum(s2);
        warp_sum(m);
        scale = 100.0 / (s2 > RMCV_FLT_EPSILON ? s2 : RMCV_FLT_EPSILON);
        nodirect_fit(m, scale, c32x, c32y, ell);
        *branch = RMCV_FIT_FALLBACK;
    }
    *status = blob_gates(*ell, prm);
    if (*status == RMCV_CONTOUR_POSITIVE) make_lightblob(*ell, prm.target, blob);
}

__global__ void __launch_bounds__(256) blob_kernel(Geometry g, SlotBuffers sb, rmcv_params prm) {
    const int frame = blockIdx.y;
    const int R = g.R, H = g.H, WB = g.WB, C = g.C, W = g.W;
    const FrameCounters& fc = sb.counters[frame];
    const int n_comps = min(fc.n_comps, C), n_runs = fc.n_runs;
    const uint32_t* bits = sb.bits + (size_t)frame * H * WB;
    const uint32_t* hole = fc.n_holes > 0 ? sb.hole + (size_t)frame * H * WB : nullptr;
    const uint32_t* run_x = sb.run_x + (size_t)frame * R;
    const int32_t* parent = sb.parent + (size_t)frame * R;
    const int32_t* row_off = sb.row_off + (size_t)frame * (H + 1);
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5), nw = gridDim.x * warps_per_block;
    for (int cid = gw; cid < n_comps; cid += nw) {
        const int root = sb.comp_root[(size_t)frame * C + cid];
        const RunStat st = sb.rstat[(size_t)frame * R + root];
        // ---- pass A: exact integer statistics
        int n = 0;
        long long sx = 0, sy = 0, cross = 0;
        component_points(bits, hole, run_x, parent, row_off, n_runs, H, WB, root, st.x0, st.y0, st.x1, st.y1, lane,
                         [&](int x, int y, int dx, int dy) {
                             ++n; sx += x; sy += y;
                             cross += (long long)x * dy - (long long)y * dx;
                         });
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        sx = warp_sum(sx); sy = warp_sum(sy); cross = warp_sum(cross);
        CompRec rec;
        rec.firstkey = n > 0 ? st.firstkey : -1;
        rec.n_points = n;
        rec.area2 = cross < 0 ? -cross : cross;
        rec.bbox[0] = st.x0; rec.bbox[1] = st.y0; rec.bbox[2] = st.x1; rec.bbox[3] = st.y1;
        rec.status = -1;
        rec.fit_branch = RMCV_FIT_NONE;
        rec.det0 = 0.f;
        memset(&rec.blob, 0, sizeof(rec.blob));
        memset(&rec.ellipse, 0, sizeof(rec.ellipse));
        if (n > 0) {  // external component (warp-uniform)
            auto pass = [&](auto&& fn) {
                component_points(bits, hole, run_x, parent, row_off, n_runs, H, WB, root, st.x0, st.y0, st.x1, st.y1,
                                 lane, [&](int x, int y, int, int) { fn(x, y); });
            };
            fit_and_gate(n, sx, sy, cross, prm, pass, &rec.status, &rec.fit_branch, &rec.det0, &rec.ellipse, &rec.blob);
        }
        if (lane == 0) sb.comps[(size_t)frame * C + cid] = rec;
        (void)W;
    }
}

// ------------------------------------------------------------------------------------------ standalone a2
// rm::filter_lightblobs on caller-supplied ordered contours: one warp per contour.
__global__ void __launch_bounds__(256) filter_lightblobs_kernel(const int32_t* xy, const int32_t* off, int n_contours,
                                                                rmcv_params prm, rmcv_contour_info* infos,
                                                                rmcv_lightblob* blobs) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= n_contours) return;
    const int p0 = off[gw], p1 = off[gw + 1], n = p1 - p0;
    const int32_t* pts = xy + 2 * (size_t)p0;
    long long sx = 0, sy = 0, cross = 0;
    int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
    for (int i = lane; i < n; i += 32) {
        const int x = pts[2 * i], y = pts[2 * i + 1];
        const int j = i == 0 ? n - 1 : i - 1;  // cv::contourArea: sum over (prev, cur)
        const int px = pts[2 * j], py = pts[2 * j + 1];
        sx += x; sy += y;
        cross += (long long)px * y - (long long)py * x;
        x0 = min(x0, x); y0 = min(y0, y); x1 = max(x1, x); y1 = max(y1, y);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); cross = warp_sum(cross);
    for (int o = 16; o > 0; o >>= 1) {
        x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    rmcv_contour_info info;
    memset(&info, 0, sizeof(info));
    rmcv_lightblob blob;
    memset(&blob, 0, sizeof(blob));
    auto pass = [&](auto&& fn) {
        for (int i = lane; i < n; i += 32) fn(pts[2 * i], pts[2 * i + 1]);
    };
    fit_and_gate(n, sx, sy, cross, prm, pass, &info.status, &info.fit_branch, &info.det0, &info.ellipse, &blob);
    if (lane == 0) {
        info.first_x = n > 0 ? pts[0] : 0; info.first_y = n > 0 ? pts[1] : 0;
        info.n_points = n;
        info.area2 = cross < 0 ? -cross : cross;
        if (n > 0) { info.bbox[0] = x0; info.bbox[1] = y0; info.bbox[2] = x1 - x0 + 1; info.bbox[3] = y1 - y0 + 1; }
        info.blob_index = -1;
        infos[gw] = info;
        blobs[gw] = blob;
    }
}

__global__ void make_lightblobs_kernel(const rmcv_rotated_rect* boxes, int n, int target, rmcv_lightblob* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) make_lightblob(boxes[i], target, out + i);
}

// ------------------------------------------------------------------------------------------ launchers
cudaError_t launch_blobs(const LabelLaunch& L, const rmcv_params& prm, cudaStream_t st, int64_t* launches) {
    dim3 grid(4, L.frames);
    if (L.g.C > 2048) grid.x = 16;
    blob_kernel<<<grid, 256, 0, st>>>(L.g, *L.sb, prm);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_filter_lightblobs(const int32_t* d_xy, const int32_t* d_off, int n, const rmcv_params& prm,
                                     rmcv_contour_info* d_infos, rmcv_lightblob* d_blobs, cudaStream_t st,
                                     int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    const int warps_per_block = 8;
    filter_lightblobs_kernel<<<(n + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(d_xy, d_off, n, prm, d_infos,
                                                                                          d_blobs);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_make_lightblobs(const rmcv_rotated_rect* d_boxes, int n, int target, rmcv_lightblob* d_out,
                                   cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    make_lightblobs_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_boxes, n, target, d_out);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
