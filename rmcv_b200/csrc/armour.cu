// Standalone rm::filter_armours (a4/a5) on caller-supplied light blobs, and the on-demand helpers behind
// rmcv_get_contour(s) / rmcv_get_label_map.  The batched path lives in frame.cu.
#include "blob_math.cuh"
#include "common.cuh"
#include "pairs.cuh"

namespace rmcv {

// ------------------------------------------------------------------------------------------ standalone a4
__global__ void __launch_bounds__(256) filter_armours_kernel(const rmcv_lightblob* blobs, int P, rmcv_params prm,
                                                             rmcv_armour* out, int cap, int32_t* count) {
    __shared__ int sh_scan[33];
    const int tid = threadIdx.x, NT = blockDim.x;
    const long long npairs = (long long)P * (P - 1) / 2;
    int base = 0;
    for (long long k0 = 0; k0 < npairs; k0 += NT) {
        const long long k = k0 + tid;
        bool pass = false;
        int i = 0, j = 0;
        float gates[6];
        if (k < npairs) {
            pair_from_index(k, P, &i, &j);
            pass = pair_passes(blobs[i], blobs[j], prm);       // cheap gates first; the gate values only for survivors
            if (pass) pair_gates(blobs[i], blobs[j], prm, gates);
        }
        int total;
        const int pos = base + block_excl_scan(pass ? 1 : 0, &total, sh_scan);
        if (pass && pos < cap) {
            rmcv_armour a;
            make_armour(blobs[i], blobs[j], &a);
            a.i = i; a.j = j;
            for (int q = 0; q < 6; ++q) a.gates[q] = gates[q];
            out[pos] = a;
        }
        base += total;
    }
    if (tid == 0) *count = base;
}

cudaError_t launch_filter_armours(const rmcv_lightblob* d_blobs, int n, const rmcv_params& prm, rmcv_armour* d_out,
                                  int cap, int32_t* d_count, cudaStream_t st, int64_t* launches) {
    filter_armours_kernel<<<1, 256, 0, st>>>(d_blobs, n, prm, d_out, cap, d_count);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ ordered contour
// Suzuki border following from the raster-first pixel of a component: direction codes 0..7 =
// E,NE,N,NW,W,SW,S,SE; first neighbour found rotating clockwise from W; then, from the direction pointing back to
// the previous pixel, rotate counter-clockwise to the next foreground pixel.  One thread; on-demand API only.
__device__ __forceinline__ bool bit_at(const uint32_t* bits, int W, int H, int WB, int x, int y) {
    if (x < 0 || y < 0 || x >= W || y >= H) return false;
    return (bits[(size_t)y * WB + (x >> 5)] >> (x & 31)) & 1u;
}

__device__ __forceinline__ int trace_border(const uint32_t* bits, int W, int H, int WB, int x0, int y0, int32_t* xy, int cap) {
    const int dxs[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dys[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int n = 0;
    auto put = [&](int x, int y) { if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; } ++n; };
    // find the first neighbour: start at W (4), rotate clockwise (s-1)
    int s = 4;
    int i1x = -1, i1y = -1;
    bool found = false;
    do {
        s = (s - 1) & 7;
        const int nx = x0 + dxs[s], ny = y0 + dys[s];
        if (bit_at(bits, W, H, WB, nx, ny)) { found = true; i1x = nx; i1y = ny; break; }
    } while (s != 4);
    if (!found) { put(x0, y0); return n; }
    int cx = x0, cy = y0;    // current pixel i3
    int sdir = s;            // direction from the current pixel to the previous one (initially towards i1)
    while (true) {
        int k = sdir, nx = cx, ny = cy;
        for (int t = 0; t < 8; ++t) {
            k = (k + 1) & 7;
            nx = cx + dxs[k]; ny = cy + dys[k];
            if (bit_at(bits, W, H, WB, nx, ny)) break;
        }
        put(cx, cy);
        if (nx == x0 && ny == y0 && cx == i1x && cy == i1y) break;
        cx = nx; cy = ny;
        sdir = (k + 4) & 7;
        if (n > (1 << 26)) break;  // safety
    }
    return n;
}

__global__ void trace_contour_kernel(Geometry g, const uint32_t* bits, int x0, int y0, int32_t* xy, int cap, int32_t* n_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    *n_out = trace_border(bits, g.W, g.H, g.WB, x0, y0, xy, cap);
}

// one thread per contour; starts = (x,y) pairs, offsets = exclusive prefix of the known point counts
__global__ void trace_all_kernel(Geometry g, const uint32_t* bits, const int32_t* starts, const int32_t* offsets, int n_contours,
                                 int32_t* xy) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_contours) return;
    const int o = offsets[k], cap = offsets[k + 1] - o;
    trace_border(bits, g.W, g.H, g.WB, starts[2 * k], starts[2 * k + 1], xy + 2 * (size_t)o, cap);
}

cudaError_t launch_trace_all(const Geometry& g, const uint32_t* bits, const int32_t* d_starts, const int32_t* d_offsets,
                             int n_contours, int32_t* d_xy, cudaStream_t st, int64_t* launches) {
    if (n_contours <= 0) return cudaSuccess;
    trace_all_kernel<<<(n_contours + 31) / 32, 32, 0, st>>>(g, bits, d_starts, d_offsets, n_contours, d_xy);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_trace_contour(const Geometry& g, const uint32_t* bits, int x0, int y0, int32_t* d_xy, int cap,
                                 int32_t* d_n, cudaStream_t st, int64_t* launches) {
    trace_contour_kernel<<<1, 32, 0, st>>>(g, bits, x0, y0, d_xy, cap, d_n);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ label map
// labels[y][x] = output index of the external contour owning the pixel, -1 elsewhere.
__global__ void label_map_kernel(Geometry g, SlotBuffers sb, int frame, int32_t* labels) {
    const int R = g.R, C = g.C, W = g.W;
    const FrameCounters& fc = sb.counters[frame];
    const int n_runs = fc.n_runs, n_comps = min(fc.n_comps, C);
    const uint32_t* run_x = sb.run_x + (size_t)frame * R;
    const uint16_t* run_y = sb.run_y + (size_t)frame * R;
    const int32_t* parent = sb.parent + (size_t)frame * R;
    const CompRec* comps = sb.comps + (size_t)frame * C;
    const int32_t* comp_root = sb.comp_root + (size_t)frame * C;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
        const int root = parent[r];
        // component id of the root, then its rank among the externals
        int cid = -1;
        for (int c = 0; c < n_comps; ++c) if (comp_root[c] == root) { cid = c; break; }
        int lab = -1;
        if (cid >= 0 && comps[cid].firstkey >= 0) {
            const int key = comps[cid].firstkey;
            lab = 0;
            for (int c = 0; c < n_comps; ++c) lab += comps[c].firstkey > key;
        }
        const uint32_t rx = run_x[r];
        const int xs = (int)(rx & 0xffffu), xe = (int)(rx >> 16), y = run_y[r];
        for (int x = xs; x <= xe; ++x) labels[(size_t)y * W + x] = lab;
    }
}

cudaError_t launch_label_map(const Geometry& g, SlotBuffers* sb, int frame, int32_t* d_labels, cudaStream_t st,
                             int64_t* launches) {
    cudaError_t e = cudaMemsetAsync(d_labels, 0xff, (size_t)g.W * g.H * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    label_map_kernel<<<64, 256, 0, st>>>(g, *sb, frame, d_labels);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
