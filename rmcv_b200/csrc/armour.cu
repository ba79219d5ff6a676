// K5/K6 — output ordering, rm::filter_armours and result write-out.
//   order_pairs_kernel   one CTA per frame: ranks the external components into cv::findContours order
//                        (reverse raster order of the first pixel), the positives likewise, then
//                        evaluates the O(P^2) pair gates of rm::filter_armours (src/objdetect.cpp:122-163) and
//                        builds rm::armour geometry (src/core.cpp:21-49) with an order-preserving block compaction
//                        (lexicographic (i,j), exactly the reference's push_back order).
//   scan_kernel          exclusive scan of the per-frame counts of the chunk -> dense offsets.
//   compact_kernel       copies the per-frame slots into the dense, host-mapped (pinned) result arrays with 16-byte
//                        stores: results reach the host without a cudaMemcpy whose size would need a sync to know.
// Plus the on-demand helpers behind rmcv_get_contour / rmcv_get_label_map and the standalone a4 entry point.
#include "blob_math.cuh"
#include "common.cuh"

namespace rmcv {

__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sh /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    __syncthreads();  // protect sh from the previous use
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarps ? sh[lane] : 0;
        int i2 = w;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) i2 += u;
        }
        sh[lane] = i2 - w;
        if (lane == 31) sh[32] = i2;
    }
    __syncthreads();
    *total = sh[32];
    return sh[warp] + incl - v;
}

// pair index k (row-major over i<j) -> (i, j)
__device__ __forceinline__ void pair_from_index(long long k, int P, int* pi, int* pj) {
    // rows: i has (P-1-i) pairs; offset(i) = i*(2P-i-1)/2
    const double b = 2.0 * P - 1.0;
    int i = (int)floor((b - sqrt(b * b - 8.0 * (double)k)) * 0.5);
    if (i < 0) i = 0;
    if (i > P - 2) i = P - 2;
    auto offs = [P](long long ii) { return ii * (2LL * P - ii - 1) / 2; };
    while (i > 0 && offs(i) > k) --i;
    while (i < P - 2 && offs(i + 1) <= k) ++i;
    *pi = i;
    *pj = (int)(k - offs(i)) + i + 1;
}

__global__ void __launch_bounds__(256) order_pairs_kernel(Geometry g, SlotBuffers sb, rmcv_params prm) {
    extern __shared__ int32_t sh_dyn[];  // keys[C] | rank_pos[C] (index of positives in order)
    __shared__ int sh_scan[33];
    __shared__ int sh_np, sh_nc, sh_nn;
    const int frame = blockIdx.x;
    const int C = g.C, A = g.A, W = g.W;
    FrameCounters& fc = sb.counters[frame];
    const int n_comps = min(fc.n_comps, C);
    const CompRec* comps = sb.comps + (size_t)frame * C;
    rmcv_contour_info* oc = sb.s_contours + (size_t)frame * C;
    rmcv_lightblob* ob = sb.s_blobs + (size_t)frame * C;
    rmcv_armour* oa = sb.s_armours + (size_t)frame * A;
    int32_t* keys = sh_dyn;
    int32_t* stat = sh_dyn + C;
    const int tid = threadIdx.x, NT = blockDim.x;
    if (tid == 0) { sh_np = 0; sh_nc = 0; sh_nn = 0; }
    for (int i = tid; i < n_comps; i += NT) {
        keys[i] = comps[i].firstkey;
        stat[i] = comps[i].status;
    }
    __syncthreads();
    // rank = number of external components with a larger first-pixel key (reverse raster order)
    for (int i = tid; i < n_comps; i += NT) {
        const int key = keys[i];
        if (key < 0) continue;
        const int st = stat[i];
        int rank = 0, prank = 0;
        for (int j = 0; j < n_comps; ++j) {
            const int kj = keys[j];
            if (kj > key) { ++rank; prank += (stat[j] == RMCV_CONTOUR_POSITIVE); }
        }
        const CompRec& c = comps[i];
        rmcv_contour_info info;
        info.first_x = key % W; info.first_y = key / W;
        info.n_points = c.n_points;
        info.status = st;
        info.area2 = c.area2;
        info.bbox[0] = c.bbox[0]; info.bbox[1] = c.bbox[1];
        info.bbox[2] = c.bbox[2] - c.bbox[0] + 1; info.bbox[3] = c.bbox[3] - c.bbox[1] + 1;
        info.ellipse = c.ellipse;
        info.fit_branch = c.fit_branch;
        info.det0 = c.det0;
        info.blob_index = st == RMCV_CONTOUR_POSITIVE ? prank : -1;
        oc[rank] = info;
        atomicAdd(&sh_nc, 1);
        if (st == RMCV_CONTOUR_POSITIVE) { ob[prank] = c.blob; atomicAdd(&sh_np, 1); }
        else if (st == RMCV_CONTOUR_NEGATIVE) atomicAdd(&sh_nn, 1);
    }
    __syncthreads();
    const int P = sh_np;
    // ---- pairs, in lexicographic (i,j) order
    const long long npairs = (long long)P * (P - 1) / 2;
    int base = 0;
    bool overflow = false;
    for (long long k0 = 0; k0 < npairs; k0 += NT) {
        const long long k = k0 + tid;
        bool pass = false;
        int i = 0, j = 0;
        float gates[6];
        if (k < npairs) {
            pair_from_index(k, P, &i, &j);
            pass = pair_gates(ob[i], ob[j], prm, gates);
        }
        int total;
        const int pos = base + block_excl_scan(pass ? 1 : 0, &total, sh_scan);
        if (pass) {
            if (pos < A) {
                rmcv_armour a;
                make_armour(ob[i], ob[j], &a);
                a.i = i; a.j = j;
                for (int q = 0; q < 6; ++q) a.gates[q] = gates[q];
                oa[pos] = a;
            } else {
                overflow = true;
            }
        }
        base += total;
    }
    if (overflow) atomicOr(&fc.flags, RMCV_FRAME_OVERFLOW_ARMOURS);
    if (tid == 0) {
        fc.n_contours = sh_nc;
        fc.n_positive = P;
        fc.n_negative = sh_nn;
        fc.n_armours = min(base, A);
    }
}

// One CTA: exclusive scan over the frames of the chunk; writes rmcv_frame_info (host-mapped).
__global__ void __launch_bounds__(1024) scan_kernel(int frames, SlotBuffers sb, int frame_base, int C_out, int A_out,
                                                    rmcv_frame_info* o_frames) {
    __shared__ int sh_scan[33];
    int bc = 0, bb = 0, ba = 0;
    for (int f0 = 0; f0 < frames; f0 += blockDim.x) {
        const int f = f0 + threadIdx.x;
        int nc = 0, nb = 0, na = 0, nn = 0, flags = 0;
        if (f < frames) {
            const FrameCounters& fc = sb.counters[f];
            nc = fc.n_contours; nb = fc.n_positive; na = fc.n_armours; nn = fc.n_negative; flags = fc.flags;
        }
        int tc, tb, ta;
        const int oc = bc + block_excl_scan(nc, &tc, sh_scan);
        const int ob = bb + block_excl_scan(nb, &tb, sh_scan);
        const int oa = ba + block_excl_scan(na, &ta, sh_scan);
        if (f < frames) {
            rmcv_frame_info fi;
            fi.n_contours = nc; fi.n_positive = nb; fi.n_negative = nn; fi.n_armours = na;
            // dense inside the chunk; the chunk's region starts at frame_base * per-frame capacity
            fi.contour_offset = frame_base * C_out + oc;
            fi.blob_offset = frame_base * C_out + ob;
            fi.armour_offset = frame_base * A_out + oa;
            fi.flags = flags;
            o_frames[frame_base + f] = fi;
        }
        bc += tc; bb += tb; ba += ta;
    }
}

__device__ __forceinline__ void copy_words(void* dst, const void* src, size_t bytes, int tid, int nt) {
    // both 8-byte aligned (struct sizes are multiples of 8); use 8-byte words
    const uint64_t* s = reinterpret_cast<const uint64_t*>(src);
    uint64_t* d = reinterpret_cast<uint64_t*>(dst);
    const size_t n = bytes / 8;
    for (size_t i = tid; i < n; i += nt) d[i] = s[i];
}

__global__ void __launch_bounds__(128) compact_kernel(Geometry g, SlotBuffers sb, int frame_base,
                                                      const rmcv_frame_info* o_frames, rmcv_contour_info* o_contours,
                                                      rmcv_lightblob* o_blobs, rmcv_armour* o_armours) {
    const int f = blockIdx.x;
    const rmcv_frame_info fi = o_frames[frame_base + f];
    copy_words(o_contours + fi.contour_offset, sb.s_contours + (size_t)f * g.C, (size_t)fi.n_contours * sizeof(rmcv_contour_info),
               threadIdx.x, blockDim.x);
    copy_words(o_blobs + fi.blob_offset, sb.s_blobs + (size_t)f * g.C, (size_t)fi.n_positive * sizeof(rmcv_lightblob),
               threadIdx.x, blockDim.x);
    copy_words(o_armours + fi.armour_offset, sb.s_armours + (size_t)f * g.A, (size_t)fi.n_armours * sizeof(rmcv_armour),
               threadIdx.x, blockDim.x);
}

cudaError_t launch_armours(const OutputLaunch& L, const rmcv_params& prm, cudaStream_t st, int64_t* launches) {
    const size_t smem = (size_t)L.g.C * 2 * sizeof(int32_t);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        cudaError_t e = cudaFuncSetAttribute(order_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    order_pairs_kernel<<<L.frames, 256, smem, st>>>(L.g, *L.sb, prm);
    scan_kernel<<<1, 1024, 0, st>>>(L.frames, *L.sb, L.frame_base, L.C_out, L.A_out, L.o_frames);
    compact_kernel<<<L.frames, 128, 0, st>>>(L.g, *L.sb, L.frame_base, L.o_frames, L.o_contours, L.o_blobs, L.o_armours);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

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
            pass = pair_gates(blobs[i], blobs[j], prm, gates);
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

__global__ void trace_contour_kernel(Geometry g, const uint32_t* bits, int x0, int y0, int32_t* xy, int cap, int32_t* n_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int dxs[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dys[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    const int W = g.W, H = g.H, WB = g.WB;
    int n = 0;
    auto put = [&](int x, int y) { if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; } ++n; };
    // find the first neighbour: start at W (4), rotate clockwise (s-1)
    int s = 4, s_end = 4;
    int i1x = -1, i1y = -1;
    bool found = false;
    do {
        s = (s - 1) & 7;
        const int nx = x0 + dxs[s], ny = y0 + dys[s];
        if (bit_at(bits, W, H, WB, nx, ny)) { found = true; i1x = nx; i1y = ny; break; }
    } while (s != s_end);
    if (!found) { put(x0, y0); *n_out = n; return; }
    int cx = x0, cy = y0;           // current pixel i3
    int px = i1x, py = i1y;         // "previous" pixel i2 (initially i1)
    // direction from current to previous
    while (true) {
        int sdir = 0;
        for (int k = 0; k < 8; ++k) if (cx + dxs[k] == px && cy + dys[k] == py) sdir = k;
        int k = sdir, nx = cx, ny = cy;
        for (int t = 0; t < 8; ++t) {
            k = (k + 1) & 7;
            nx = cx + dxs[k]; ny = cy + dys[k];
            if (bit_at(bits, W, H, WB, nx, ny)) break;
        }
        put(cx, cy);
        if (nx == x0 && ny == y0 && cx == i1x && cy == i1y) break;
        px = cx; py = cy;
        cx = nx; cy = ny;
        if (n > (1 << 24)) break;  // safety
    }
    *n_out = n;
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
    const int32_t* run_y = sb.run_y + (size_t)frame * R;
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
