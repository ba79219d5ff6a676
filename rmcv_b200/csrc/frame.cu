// Labelling and measurement stages of the detection path (everything after the pixel kernel), one launch each per chunk:
//   K_E  emit.cu   runs + boundary-pixel records from the bit mask;
//   K_L  labelling, one CTA per frame in shared memory: every run links to the first run it touches in the row above
//        (8-connected foreground), the links are collapsed by pointer jumping (log2(depth) rounds, no atomics) and only
//        the rare extra contacts (a run touching several runs above) go through a lock-free union-find — replaces the
//        component discovery of cv::findContours(RETR_EXTERNAL) (reference: src/imgproc.cpp:71-72).
//        Holes: #holes = #components - #runs + #run contacts (Euler relation on the run graph); only frames that have a
//        hole label the background gaps too (4-connected, same link + jump + union scheme; node 0 = background connected
//        to the image border), so that arcs facing a hole are skipped and nested components end with n == 0;
//   K_C  contour sums, one CTA per frame with little shared memory: the records are bucketed by component and one warp
//        per component accumulates, as EXACT integers, contour.size(), the shoelace sum of cv::contourArea, bbox and
//        the 14 moment sums of the contour point multiset from the local 3x3 arc rule — the contour is
//        never materialised;
//   K_F  one thread per component of the whole chunk: cv::fitEllipseDirect incl. its fallback, the ratio/tilt gates and
//        the rm::lightblob ctor (reference: src/objdetect.cpp:62-84, src/core.cpp:9-19);
//   K_O  one small CTA per frame: ordering into cv::findContours order, the O(P^2) pair gates of rm::filter_armours and
//        rm::armour geometry (reference: src/objdetect.cpp:114-166, src/core.cpp:21-49) with an order-preserving block
//        compaction, dense write-out straight into pinned, device-mapped host memory.
// Each kernel has a homogeneous resource profile, so each fills the SMs on its own.
#include <cooperative_groups.h>

#include "blob_math.cuh"
#include "common.cuh"
#include "pairs.cuh"
#include "warp_fit.cuh"

namespace rmcv {

namespace cg = cooperative_groups;

// Phase stamps for latency anatomy (scripts/phase_stamps.py builds a -DRMCV_STAMPS copy of the library; the product has none).
#ifdef RMCV_STAMPS
__device__ long long g_stamps[4][32];
#define STAMP(k, i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_stamps[k][i] = clock64(); } while (0)
#else
#define STAMP(k, i) do { } while (0)
#endif
RMCV_GSTAMP_ARRAY(g_ns_frame)

// A frame is labelled by ONE CTA (CS == 1) or, when a chunk holds fewer frames than the GPU has SMs (4096x3072 stress
// frames, batch-1 latency), by a thread-block CLUSTER of CS CTAs: the per-frame phases are bound by memory latency, so
// CS times the threads mean CS times the loads in flight.  The cluster works on the frame's global arrays; what a single
// CTA keeps in shared-memory scalars lives in the shared memory of cluster rank 0 (distributed shared memory), block-wide
// barriers become cluster barriers (barrier.cluster, release/acquire at cluster scope).
template <int CS>
struct Team {
    int rank, k_or;
    __device__ __forceinline__ Team() : rank(0), k_or(0) {
        if (CS > 1) rank = (int)cg::this_cluster().block_rank();
    }
    __device__ __forceinline__ void sync() const {
        if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
    }
    template <class T>
    __device__ __forceinline__ T* on0(T* p) const {      // the same shared-memory object in cluster rank 0
        if (CS > 1) return cg::this_cluster().map_shared_rank(p, 0);
        return p;
    }
    // cluster-wide OR of a per-thread flag; flags3 = three ints of shared memory (rank 0's are used), zero at kernel start
    __device__ __forceinline__ int any(int v, int* flags3) {
        if (CS == 1) return __syncthreads_or(v);
        const int mine = __syncthreads_or(v);
        int* f0 = on0(flags3);
        const int slot = k_or % 3;
        if (threadIdx.x == 0) {
            if (rank == 0) f0[(slot + 1) % 3] = 0;        // the slot of the next call; last read two barriers ago
            if (mine) atomicOr(f0 + slot, 1);
        }
        sync();
        ++k_or;
        return *reinterpret_cast<volatile int*>(f0 + slot);
    }
};

// Arc table indexed by  NW | N<<1 | NE<<2 | W<<3 | E<<4 | SW<<5 | S<<6 | SE<<7  (bit set = foreground).
// Entry: bits 0-2 arc count m; arc i at bits 3+5i: low 2 bits = 4-neighbour to test for "hole" (0=E,1=N,2=W,3=S),
// high 3 bits = direction of the edge target q (0..7 = E,NE,N,NW,W,SW,S,SE); bit 31 = isolated pixel (no edge).
__constant__ uint32_t c_arc_lut[256];
// direction q -> (dx, dy), two bits per direction holding d+1 (a register literal: divergent lanes would serialise on a
// constant-memory table)
__device__ __forceinline__ int dir_dx(int q) { return (int)((0x901Au >> (2 * q)) & 3u) - 1; }
__device__ __forceinline__ int dir_dy(int q) { return (int)((0xA901u >> (2 * q)) & 3u) - 1; }

void upload_luts() {
    uint32_t lut[256];
    for (int idx = 0; idx < 256; ++idx) {
        int fg[8];  // un-permute to direction order E,NE,N,NW,W,SW,S,SE
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

// ------------------------------------------------------------------------------------------ union-find
__device__ __forceinline__ int uf_find(int32_t* parent, int x) {
    // path halving with atomicMin: links only ever move towards smaller indices, so concurrent unions stay valid
    while (true) {
        const int p = *reinterpret_cast<volatile int32_t*>(parent + x);
        if (p == x) return x;
        const int gp = *reinterpret_cast<volatile int32_t*>(parent + p);
        if (gp != p) atomicMin(parent + x, gp);
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
        a = old;
    }
}

struct Runs {  // run storage of one frame (shared or global memory, same code)
    const int2* rows;      // per row: [first, end)
    const uint32_t* run_x; // xs | xe<<16
    const uint16_t* run_y;
    int32_t* link;         // foreground forest, then the flat label (root run) of every run
    int32_t* glink;        // background-gap forest: node r+1 = gap to the left of run r; node 0 = outer background
    int16_t* cid;          // first: "joined with the previous run of my row" flag; then the component id of the run
    uint8_t* jp;           // gap node flags (alias the link region / the global bucket array before the bucketing):
    uint8_t* jo;           //   jp = joined with the previous gap of its row, jo = joined with the outer background
    int n_runs, W, H;
};

// first run index in [lo,hi) with xe >= x
__device__ __forceinline__ int lower_bound_xe(const uint32_t* run_x, int lo, int hi, int x) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(run_x[mid] >> 16) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first run index in [lo,hi) with xs > x
__device__ __forceinline__ int upper_bound_xs(const uint32_t* run_x, int lo, int hi, int x) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(run_x[mid] & 0xffffu) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Collapses a forest of links (every node points at an ancestor or at itself) until every node points at its root.
// Reads may see values written concurrently, which are ancestors too, and a root never changes, so a single CTA needs no
// barrier between rounds: every thread jumps its own nodes until their parent is a root (log2(depth) steps while the other
// threads make the same progress).  A cluster keeps the lock-step rounds (its members see each other through L2).
template <class TeamT>
__device__ __forceinline__ void pointer_jump(int32_t* link, int first, int end, int tid, int NT, TeamT& team, int* flags3, bool lockstep) {
    volatile int32_t* v = link;
    if (!lockstep) {
        for (int r0 = first + tid; r0 < end; r0 += 4 * NT) {   // four independent chains per iteration (latency-bound)
            int l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int r = r0 + u * NT; l[u] = r < end ? v[r] : -1; }
            bool busy = true;
            while (busy) {
                busy = false;
                int ll[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) ll[u] = l[u] >= 0 ? v[l[u]] : -1;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ll[u] != l[u]) { l[u] = ll[u]; v[r0 + u * NT] = ll[u]; busy = true; }
            }
        }
        team.sync();
        return;
    }
    while (true) {
        int changed = 0;
        for (int r0 = first + tid; r0 < end; r0 += 4 * NT) {
            int l[4], ll[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int r = r0 + u * NT; l[u] = r < end ? v[r] : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) ll[u] = l[u] >= 0 ? v[l[u]] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ll[u] != l[u]) { v[r0 + u * NT] = ll[u]; changed = 1; }
        }
        if (!team.any(changed, flags3)) break;
    }
}

// ------------------------------------------------------------------------------------------ K_L: labelling
// One CTA per frame; runs, links and labels live in shared memory (frames with more runs than fit use global arrays).
struct LabelParams {
    Geometry g;
    SlotBuffers sb;
    int Rs;      // run capacity of the shared-memory arrays
};

// Shared memory of the label kernel.  The "link" region is reused over the kernel's life: foreground links -> gap join
// flags + per-component hole flags -> record counts / start offsets.
__host__ __device__ inline size_t label_link_bytes(int Rs, int C) {
    size_t a = (size_t)Rs * sizeof(int32_t);
    const size_t b = (((size_t)2 * (Rs + 2) + 3) & ~(size_t)3) + (size_t)C * sizeof(int32_t) + 8;  // jp | jo | hole flags
    const size_t c = ((size_t)2 * C + 2) * sizeof(int32_t);                                        // cnt | start
    if (b > a) a = b;
    if (c > a) a = c;
    return (a + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t label_smem_bytes(int H, int Rs, int C) {
    size_t b = 0;
    b += ((size_t)H * sizeof(ushort2) + 15) & ~(size_t)15;  // rows (16-bit run indices: shared-memory mode only)
    b += (size_t)Rs * sizeof(uint32_t);            // run_x
    b += label_link_bytes(Rs, C);                  // link / gap flags / record counts
    b += ((size_t)Rs + 2) * sizeof(int32_t);       // glink
    b += (size_t)Rs * sizeof(uint16_t);            // run_y
    b += (size_t)Rs * sizeof(int16_t);             // cid
    b = (b + 3) & ~(size_t)3;
    b += (size_t)C * sizeof(int32_t);              // Euler term of every component (single-CTA frames)
    return b + 64;
}

// NTMAX / MINB: 256 threads, four CTAs per SM for ordinary frames (the runs fit in shared memory); 1024 threads, one CTA per
// SM for frames whose run arrays stay in global memory (4096x3072 stress frames), where the phases are bound by L2 latency
// and more threads per frame mean more loads in flight.
template <int CS>
__device__ __forceinline__ void label_body(const LabelParams& p, uint8_t* smem) {
    __shared__ int sh_scan[33];
    __shared__ int s_ncomp_, s_nadj_, s_flags_, s_hb_[4], s_or_[3];
    Team<CS> team;
    STAMP(0, 0);
    RMCV_GSTAMP_BEGIN(g_ns_frame, 0);
    // shared scalars of the frame: rank 0's copies (CS == 1: this CTA's)
    int& s_ncomp = *team.on0(&s_ncomp_); int& s_nadj = *team.on0(&s_nadj_); int& s_flags = *team.on0(&s_flags_);
    int* s_hb = team.on0(s_hb_);
    const Geometry& g = p.g;
    const int W = g.W, H = g.H, R = g.R, C = g.C, Rs = p.Rs;
    const int frame = blockIdx.x / CS;
    const int ltid = threadIdx.x, LNT = blockDim.x;               // within this CTA
    const int tid = team.rank * LNT + ltid, NT = LNT * CS, lane = ltid & 31;   // within the frame's team
    const SlotBuffers& sb = p.sb;
    FrameCounters& fc = sb.counters[frame];
    chain_begin();

    uint8_t* q = smem;
    ushort2* s_rows16 = reinterpret_cast<ushort2*>(q); q += ((size_t)H * sizeof(ushort2) + 15) & ~(size_t)15;
    uint32_t* s_run_x = reinterpret_cast<uint32_t*>(q); q += (size_t)Rs * sizeof(uint32_t);
    int32_t* s_link = reinterpret_cast<int32_t*>(q);
    int32_t* s_hole = reinterpret_cast<int32_t*>(q + (((size_t)2 * (Rs + 2) + 3) & ~(size_t)3));  // per-component hole flags (gap phase)
    int32_t* s_cnt = team.on0(s_link);             // record counts / start offsets (after the gap phase): rank 0's
    int32_t* s_start = s_cnt + C;
    q += label_link_bytes(Rs, C);
    int32_t* s_glink = reinterpret_cast<int32_t*>(q); q += ((size_t)Rs + 2) * sizeof(int32_t);
    uint16_t* s_run_y = reinterpret_cast<uint16_t*>(q); q += (size_t)Rs * sizeof(uint16_t);
    int16_t* s_cid = reinterpret_cast<int16_t*>(q); q += (size_t)Rs * sizeof(int16_t);
    int32_t* s_euler = reinterpret_cast<int32_t*>(smem + (((size_t)(q - smem) + 3) & ~(size_t)3));

    chain_wait();
    const int raw_runs = fc.n_runs;
    const int n_runs = min(raw_runs, R);
    const bool in_smem = CS == 1 && n_runs <= Rs;   // a team works on the global arrays
    const uint32_t* g_run_x = sb.run_x + (size_t)frame * R;
    const uint16_t* g_run_y = sb.run_y + (size_t)frame * R;
    int32_t* g_parent = sb.parent + (size_t)frame * R;
    int32_t* g_glink = sb.gparent + (size_t)frame * (R + 2);
    int16_t* g_cid = sb.run_cid + (size_t)frame * R;
    int2* g_rows = sb.rows + (size_t)frame * H;
    int32_t* g_comp_root = sb.comp_root + (size_t)frame * C;
    int32_t* g_comp_cnt = sb.comp_cnt + (size_t)frame * C;     // out: bit 0 = has holes of its own, bit 1 = lies in a hole
    // Euler term 1 - runs + contacts of every component (> 0: it has holes): shared-memory atomics for a single CTA, the
    // global array for a cluster
    int32_t* ecnt = CS == 1 ? s_euler : g_comp_cnt;

    Runs f;
    f.rows = nullptr;
    // rows[y] = (first, end) run of row y: 16-bit copies in shared memory when the runs live there, else the global array
    auto row = [&](int y) -> int2 {
        if (in_smem) { const ushort2 v = s_rows16[y]; return make_int2((int)v.x, (int)v.y); }
        return g_rows[y];
    };
    f.run_x = in_smem ? s_run_x : g_run_x;
    f.run_y = in_smem ? s_run_y : g_run_y;
    f.link = in_smem ? s_link : g_parent;
    f.glink = in_smem ? s_glink : g_glink;
    f.cid = in_smem ? s_cid : g_cid;
    f.jp = in_smem ? reinterpret_cast<uint8_t*>(s_link) : reinterpret_cast<uint8_t*>(sb.sorted + (size_t)frame * g.SC);
    f.jo = f.jp + (size_t)n_runs + 2;
    f.n_runs = n_runs; f.W = W; f.H = H;

    if (tid == 0) {
        s_ncomp = 0; s_nadj = 0;
        s_flags = raw_runs > R ? RMCV_FRAME_OVERFLOW_RUNS : 0;
        s_or_[0] = s_or_[1] = s_or_[2] = 0;
    }
    for (int y = tid; y < H; y += NT) {
        int2 rr = g_rows[y];
        if (rr.x > n_runs || rr.y > n_runs) {  // only after a run overflow: keep every later stage inside the arrays
            rr.x = min(rr.x, n_runs); rr.y = min(rr.y, n_runs);
            g_rows[y] = rr;
        }
        if (in_smem) s_rows16[y] = make_ushort2((unsigned short)rr.x, (unsigned short)rr.y);
    }
    for (int r = tid; r < n_runs; r += NT) {
        if (in_smem) { s_run_x[r] = g_run_x[r]; s_run_y[r] = g_run_y[r]; }
        f.cid[r] = 0;
    }
    team.sync();
    STAMP(0, 1);
    // ---- foreground links (8-connectivity): the first run of row y-1 overlapping [xs-1, xe+1] becomes the parent;
    // every further touched run is flagged "joined with the run before it" (they are consecutive in their row)
    {
        int adj = 0;
        for (int r = tid; r < n_runs; r += NT) {
            const uint32_t rx = f.run_x[r];
            const int xs = (int)(rx & 0xffffu), xe = (int)(rx >> 16), y = f.run_y[r];
            int lk = r;
            if (y > 0) {
                const int2 pr = row(y - 1);
                const int p0 = lower_bound_xe(f.run_x, pr.x, pr.y, xs - 1);
                int pp = p0;
                while (pp < pr.y && (int)(f.run_x[pp] & 0xffffu) <= xe + 1) {
                    if (pp > p0) f.cid[pp] = 1;
                    ++pp;
                }
                if (pp > p0) { lk = p0; adj += pp - p0; }
                f.glink[r + 1] = pp - p0;  // contacts of this run, summed per component below (glink is free until the gap phase)
            } else {
                f.glink[r + 1] = 0;
            }
            f.link[r] = lk;
        }
        if (adj) atomicAdd(&s_nadj, adj);
    }
    team.sync();
    STAMP(0, 2);
    pointer_jump(f.link, 0, n_runs, tid, NT, team, s_or_, CS > 1);
    STAMP(0, 3);
    for (int r = tid; r < n_runs; r += NT)
        if (f.cid[r]) uf_union(f.link, r, r - 1);
    team.sync();
    STAMP(0, 4);
    // ---- flatten, enumerate components
    for (int r = tid; r < n_runs; r += NT) {
        const int root = uf_find(f.link, r);
        if (root == r) {
            const int c = atomicAdd(&s_ncomp, 1);
            if (c < C) {
                g_comp_root[c] = r; ecnt[c] = 1;   // the Euler term 1 - runs + contacts, accumulated below
                f.cid[r] = (int16_t)c;
            } else {
                f.cid[r] = -1;
                atomicOr(&s_flags, RMCV_FRAME_OVERFLOW_BLOBS);
            }
        } else {
            f.link[r] = root;
        }
    }
    team.sync();
    STAMP(0, 5);
    const int n_comps = min(s_ncomp, C);
    const int n_holes = s_ncomp - n_runs + s_nadj;  // Euler relation on the run graph
    const bool has_holes = n_holes > 0;
    for (int r0 = 0; r0 < n_runs; r0 += NT) {  // component id of every run; per component: contacts - runs (warp-aggregated)
        const int r = r0 + tid;
        int c = -1, adj = 0;
        if (r < n_runs) {
            const int root = f.link[r];
            c = f.cid[root];
            adj = f.glink[r + 1] - 1;
            if (root != r) f.cid[r] = (int16_t)c;
            if (in_smem) { g_parent[r] = root; g_cid[r] = (int16_t)c; }   // read by the contour kernel / rmcv_get_label_map
        }
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        adj = __reduce_add_sync(peers, adj);
        if (c >= 0 && lane == __ffs(peers) - 1) atomicAdd(&ecnt[c], adj);
    }
    team.sync();
    STAMP(0, 6);
    // ---- background gaps (4-connectivity), only when the frame has a hole
    if (has_holes) {
        // A hole lies strictly inside the bounding box of the component that encloses it, so a gap that is not strictly
        // inside the union of the bounding boxes of the components with holes is outer background without any search.
        if (tid == 0) { s_hb[0] = INT32_MAX; s_hb[1] = INT32_MAX; s_hb[2] = -1; s_hb[3] = -1; f.glink[0] = 0; }
        for (int i = tid; i < (2 * (n_runs + 2) + 3) / 4; i += NT) reinterpret_cast<uint32_t*>(f.jp)[i] = 0u;
        for (int c = ltid; c < n_comps; c += LNT) s_hole[c] = ecnt[c];   // every CTA of the team keeps its own copy
        team.sync();
        STAMP(0, 12);
        {
            int bx0 = INT32_MAX, by0 = INT32_MAX, bx1 = -1, by1 = -1;
            for (int r = tid; r < n_runs; r += NT) {
                const int c = f.cid[r];
                if (c >= 0 && s_hole[c] > 0) {
                    const uint32_t rx = f.run_x[r];
                    const int y = (int)f.run_y[r];
                    bx0 = min(bx0, (int)(rx & 0xffffu)); bx1 = max(bx1, (int)(rx >> 16));
                    by0 = min(by0, y); by1 = max(by1, y);
                }
            }
            bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
            bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
            if (lane == 0 && bx1 >= 0) {
                atomicMin(&s_hb[0], bx0); atomicMax(&s_hb[2], bx1);
                atomicMin(&s_hb[1], by0); atomicMax(&s_hb[3], by1);
            }
        }
        team.sync();
        STAMP(0, 13);
        const int hx0 = s_hb[0], hy0 = s_hb[1], hx1 = s_hb[2], hy1 = s_hb[3];
        for (int r = tid; r < n_runs; r += NT) {
            const int y = f.run_y[r];
            const int2 rr = row(y);
            if (r == rr.x) { f.glink[r + 1] = 0; continue; }   // left border gap (the gaps above it test for it themselves)
            const int a = (int)(f.run_x[r - 1] >> 16) + 1, b = (int)(f.run_x[r] & 0xffffu) - 1;
            if (y <= hy0 || y >= hy1 || a <= hx0 || b >= hx1) {
                // not strictly inside the box: outer background without any search.  The gaps of the row above that it
                // touches and that do lie inside the box are connected to it, and nothing else tells them so.
                f.glink[r + 1] = 0;
                if (y - 1 > hy0 && y - 1 < hy1 && b > hx0 && a < hx1) {
                    const int2 pr = row(y - 1);
                    const int lo = pr.x, hi = pr.y;
                    for (int k = upper_bound_xs(f.run_x, lo, hi, a);; ++k) {
                        const int ga = (k == lo) ? 0 : (int)(f.run_x[k - 1] >> 16) + 1;
                        if (ga > b) break;
                        const int gb = (k == hi) ? W - 1 : (int)(f.run_x[k] & 0xffffu) - 1;
                        if (gb >= a && ga <= gb && k != lo && k != hi) f.jo[k + 1] = 1;
                        if (k == hi) break;
                    }
                }
                continue;
            }
            bool outer = false;
            int first = -1;
            {   // row above: every overlapped gap is connected to this one (and so to each other)
                const int2 pr = row(y - 1);
                const int lo = pr.x, hi = pr.y;
                bool prev_overlapped = false;
                for (int k = upper_bound_xs(f.run_x, lo, hi, a);; ++k) {
                    // gap k lies between run k-1 and run k (k == lo: left border gap; k == hi: right border gap)
                    const int ga = (k == lo) ? 0 : (int)(f.run_x[k - 1] >> 16) + 1;
                    if (ga > b) break;
                    const int gb = (k == hi) ? W - 1 : (int)(f.run_x[k] & 0xffffu) - 1;
                    if (gb >= a && ga <= gb) {
                        if (k == lo || k == hi || y - 1 == 0) {
                            outer = true;
                        } else {
                            if (first < 0) first = k + 1;
                            if (prev_overlapped) f.jp[k + 1] = 1;  // node k (the gap before it) may alias the outer background
                        }
                        prev_overlapped = true;
                    }
                    if (k == hi) break;
                }
            }
            {   // row below: its interior gaps look up themselves; only its border gaps have no node
                const int2 nr = row(y + 1);
                if (nr.x == nr.y) {
                    outer = true;
                } else if (y + 1 == H - 1) {  // every gap of the last row is outer: connected unless one run covers [a,b]
                    const int k = lower_bound_xe(f.run_x, nr.x, nr.y, a);
                    const bool covered = k < nr.y && (int)(f.run_x[k] & 0xffffu) <= a && (int)(f.run_x[k] >> 16) >= b;
                    if (!covered) outer = true;
                } else if ((int)(f.run_x[nr.x] & 0xffffu) > a || (int)(f.run_x[nr.y - 1] >> 16) < b) {
                    outer = true;
                }
            }
            if (outer && first >= 0) f.jo[first] = 1;
            f.glink[r + 1] = outer ? 0 : (first >= 0 ? first : r + 1);
        }
        team.sync();
        STAMP(0, 14);
        pointer_jump(f.glink, 1, n_runs + 1, tid, NT, team, s_or_, CS > 1);
        STAMP(0, 15);
        for (int gnode = 1 + tid; gnode <= n_runs; gnode += NT) {
            if (f.jp[gnode]) uf_union(f.glink, gnode, gnode - 1);
            if (f.jo[gnode]) uf_union(f.glink, gnode, 0);
        }
        team.sync();
        STAMP(0, 16);
        for (int gnode = tid; gnode <= n_runs; gnode += NT) {
            const int root = gnode == 0 ? 0 : uf_find(f.glink, gnode);
            f.glink[gnode] = root;
            if (in_smem) g_glink[gnode] = root;  // read by the contour kernel's hole tests
        }
    }
    team.sync();
    STAMP(0, 7);
    // A component without holes faces ONE background region: outer, or a hole of another component (nested: RETR_EXTERNAL
    // drops it).  The pixel above the first pixel of its root run is background of that region (a root run touches no run
    // above), so one gap look-up decides; the contour kernel reads the verdict.
    for (int c = tid; c < n_comps; c += NT) {
        const int own = ecnt[c] > 0;
        int nested = 0;
        if (has_holes && !own) {
            const int root = g_comp_root[c];
            const int x = (int)(f.run_x[root] & 0xffffu), yy = (int)f.run_y[root] - 1;
            if (x > 0 && yy > 0 && x < W - 1 && yy < H - 1) {
                const int2 rr = row(yy);
                const int k = upper_bound_xs(f.run_x, rr.x, rr.y, x);   // the gap between run k-1 and run k
                if (k != rr.x && k != rr.y) nested = f.glink[k + 1] != 0;
            }
        }
        g_comp_cnt[c] = own | (nested << 1);
        s_cnt[c] = 0;
    }
    // ---- boundary-pixel records -> components (through the run each record is tagged with): the records are counted
    // per component, start = exclusive scan of the counts, and a warp-aggregated scatter writes them
    // bucketed by component (so that the contour kernel streams each component's records linearly).
    const int raw_recs = fc.n_recs;
    const int n_recs = min(raw_recs, g.PC);
    const uint2* recs = sb.recs + (size_t)frame * g.PC;
    uint2* recs2 = sb.recs2 + (size_t)frame * g.PC;
    int32_t* g_start = sb.comp_start + (size_t)frame * (C + 1);
    if (tid == 0 && raw_recs > g.PC) s_flags |= RMCV_FRAME_OVERFLOW_POINTS;
    team.sync();
    STAMP(0, 8);
    auto comp_of = [&](const uint2 rec) -> int {  // the emit kernel tagged the record with its run
        const int r = (int)(rec.y >> 8);
        return r < n_runs ? (int)f.cid[r] : -1;
    };
    for (int i0 = 0; i0 < n_recs; i0 += 4 * NT) {   // four loads in flight per thread
        uint2 rec[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * NT + tid; rec[u] = i < n_recs ? recs[i] : make_uint2(0u, 0xffffff00u); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = comp_of(rec[u]);
            const unsigned peers = __match_any_sync(0xffffffffu, c);
            if (c >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[c], __popc(peers));
        }
    }
    team.sync();
    STAMP(0, 9);
    if (team.rank == 0) {   // block-wide scan by the first CTA of the team
        int carry = 0;
        for (int c0 = 0; c0 < n_comps; c0 += LNT) {
            const int c = c0 + ltid;
            const int v = c < n_comps ? s_cnt[c] : 0;
            int total;
            const int ex = block_excl_scan(v, &total, sh_scan);
            if (c < n_comps) { s_start[c] = carry + ex; g_start[c] = carry + ex; s_cnt[c] = 0; }
            carry += total;
        }
        if (ltid == 0) g_start[n_comps] = carry;
    }
    team.sync();
    STAMP(0, 10);
    for (int i0 = 0; i0 < n_recs; i0 += 4 * NT) {
        uint2 rec[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * NT + tid; rec[u] = i < n_recs ? recs[i] : make_uint2(0u, 0xffffff00u); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = comp_of(rec[u]);
            const unsigned peers = __match_any_sync(0xffffffffu, c);
            int base = 0;
            const int leader = __ffs(peers) - 1;
            if (c >= 0 && lane == leader) base = atomicAdd(&s_cnt[c], __popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (c >= 0) recs2[s_start[c] + base + __popc(peers & ((1u << lane) - 1u))] = rec[u];
        }
    }
    if (tid == 0) { fc.n_comps = n_comps; fc.n_holes = n_holes; fc.flags = s_flags; }
    if (CS > 1) team.sync();   // rank 0's shared memory is read by the whole team until here
#ifdef RMCV_STAMPS
    __syncthreads();
    STAMP(0, 11);
    RMCV_GSTAMP_END(g_ns_frame, 0);
#endif
}

template <int NTMAX, int MINB, int CS>
__global__ void __launch_bounds__(NTMAX, MINB) label_kernel(const LabelParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    label_body<CS>(p, smem);
}

// ------------------------------------------------------------------------------------------ K_C: contour sums
// One warp per component over the whole chunk, lanes over its boundary-pixel records (bucketed by the label kernel):
// accumulates, as exact integers, everything cv::contourArea / cv::fitEllipseDirect need.  No per-frame state in shared
// memory, so tens of warps are resident per SM and the global-memory latency of the look-ups is hidden.
struct ContourParams {
    Geometry g;
    SlotBuffers sb;
    rmcv_params prm;
};

struct GRuns {  // read-only view of one frame's runs and gap labels in global memory
    const int2* rows; const uint32_t* run_x; const int32_t* glink; int W, H;
};
__device__ __forceinline__ bool g_is_hole(const GRuns& f, int x, int yy) {
    if (x <= 0 || yy <= 0 || x >= f.W - 1 || yy >= f.H - 1) return false;
    const int2 rr = __ldg(f.rows + yy);
    int lo = rr.x, hi = rr.y;  // first run with xs > x: the gap lies between run r-1 and run r
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(__ldg(f.run_x + mid) & 0xffffu) <= x) lo = mid + 1; else hi = mid;
    }
    if (lo == rr.x || lo == rr.y) return false;  // touches the left / right border
    return __ldg(f.glink + lo + 1) != 0;
}

// One component's sums -> its record: cv::fitEllipseDirect incl. its fallback, the ratio/tilt gates and the rm::lightblob ctor
// (reference: src/objdetect.cpp:62-84, src/core.cpp:9-19).  One thread - or, kWarp, all 32 lanes of a warp with the same `a`
// (the warp-cooperative direct fit of blob_math.cuh; every lane ends with the same record, `leader` raises the frame flag).
template <bool kWarp>
__device__ __forceinline__ void fit_component(const CompAcc& a, const rmcv_params& prm, int32_t* frame_flags, CompRec& rec, bool leader = true) {
    const int n = (int)a.n;
    rec.firstkey = n > 0 ? a.firstkey : -1;
    rec.n_points = n;
    rec.area2 = a.cross < 0 ? -a.cross : a.cross;
    rec.bbox[0] = a.bbox[0]; rec.bbox[1] = a.bbox[1]; rec.bbox[2] = a.bbox[2]; rec.bbox[3] = a.bbox[3];
    rec.status = -1;
    rec.fit_branch = RMCV_FIT_NONE;
    rec.det0 = 0.f;
    memset(&rec.blob, 0, sizeof(rec.blob));
    memset(&rec.ellipse, 0, sizeof(rec.ellipse));
    if (n > 0) {  // external component
        ContourSums cs;
        cs.n = a.n; cs.sx = a.sx; cs.sy = a.sy; cs.cross = a.cross;
        cs.xx = a.xx; cs.xy = a.xy; cs.yy = a.yy; cs.xxx = a.xxx; cs.xxy = a.xxy; cs.xyy = a.xyy; cs.yyy = a.yyy;
        cs.xxxx = a.xxxx; cs.xxxy = a.xxxy; cs.xxyy = a.xxyy; cs.xyyy = a.xyyy; cs.yyyy = a.yyyy;
        cs.s_int = a.s_int; cs.ox = a.ox; cs.oy = a.oy;
        fit_contour_t<kWarp>(cs, prm, &rec.status, &rec.fit_branch, &rec.det0, &rec.ellipse, &rec.blob);
        // the 4th-order sums are exact 64-bit integers about the component's root pixel: n * extent^4 must stay below 2^62
        // (never reached with the reference's area_max = 99999; a caller who raises it gets the frame flagged, not a wrong fit)
        if (a.fitted) {
            const double ext = (double)max(a.bbox[2] - a.bbox[0], a.bbox[3] - a.bbox[1]) + 1.0;
            if (leader && (double)a.n * ext * ext * ext * ext > 4.6e18) atomicOr(frame_flags, RMCV_FRAME_OVERFLOW_MOMENTS);
        }
    }
}

// c_first / c_stride: this warp's first component and the number of warps working on the frame
// kFit: lane 0 goes straight on to the component's fit (chunks of a few frames: one kernel boundary less, and the fits of
// the small components start while the long ones are still summing; the other lanes idle, so not for throughput)
template <int kFit>   // 0: sums only; 1: lane 0 fits; 2: the warp fits (warp-cooperative direct fit)
__device__ __forceinline__ void contour_body(const Geometry& g, const SlotBuffers& sb, const rmcv_params& prm, int frame, int c_first, int c_stride) {
    __shared__ uint32_t s_lut[256];
    const int W = g.W, H = g.H, R = g.R, C = g.C;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
    const FrameCounters& fc = sb.counters[frame];
    chain_begin();
    STAMP(1, 0);
    RMCV_GSTAMP_BEGIN(g_ns_frame, 1);
    for (int i = tid; i < 256; i += NT) s_lut[i] = c_arc_lut[i];   // (before the wait: overlaps the label kernel's tail)
    __syncthreads();
    chain_wait();
    const int n_comps = fc.n_comps;
    STAMP(1, 1);

    GRuns f;
    f.rows = sb.rows + (size_t)frame * H;
    f.run_x = sb.run_x + (size_t)frame * R;
    f.glink = sb.gparent + (size_t)frame * (R + 2);
    f.W = W; f.H = H;
    const uint16_t* run_y = sb.run_y + (size_t)frame * R;
    const int32_t* comp_root = sb.comp_root + (size_t)frame * C;
    const int32_t* comp_holes = sb.comp_cnt + (size_t)frame * C;
    const int32_t* g_start = sb.comp_start + (size_t)frame * (C + 1);
    const uint2* recs = sb.recs2 + (size_t)frame * g.PC;
    // ---- per component (one warp each, claimed dynamically), lanes over its boundary pixels.  A pixel contributes one
    // contour point per arc of the 3x3 rule.
    CompAcc* accs = sb.acc + (size_t)frame * C;
    for (int c = c_first; c < n_comps; c += c_stride) {
        const int base = g_start[c], cnt = g_start[c + 1] - base;
        const int root = comp_root[c];
        const int ox = (int)(f.run_x[root] & 0xffffu), oy = (int)run_y[root];
        // A component without holes faces ONE background region: outer (all arcs count) or a hole of another component
        // (nested: RETR_EXTERNAL drops it).  The pixel above the first pixel of the root run is background of that region
        // (a root run touches no run above).  Only components with holes of their own test every arc.
        const int hole_flags = comp_holes[c];      // label kernel: bit 0 = holes of its own, bit 1 = lies in a hole
        const bool own_holes = (hole_flags & 1) != 0;
        const bool nested = (hole_flags & 2) != 0;
        const int ncnt = nested ? 0 : cnt;
        STAMP(1, 2);
        // one record -> (multiplicity k, sum of the edge directions of its counted arcs)
        auto arcs_of = [&](uint32_t nb, int x, int y, int* dxs, int* dys) -> int {
            const uint32_t ent = s_lut[nb & 0xffu];
            const int m = ent & 7;
            const bool iso = (ent >> 31) != 0;
            int k = 0, sdx = 0, sdy = 0;
            for (int a = 0; a < m; ++a) {
                const uint32_t arc = (ent >> (3 + 5 * a)) & 31u;
                if (own_holes) {
                    const int t4 = arc & 3u;  // 0=E,1=N,2=W,3=S
                    if (g_is_hole(f, x + (t4 == 0) - (t4 == 2), y + (t4 == 3) - (t4 == 1))) continue;
                }
                ++k;
                if (!iso) { sdx += dir_dx(arc >> 2); sdy += dir_dy(arc >> 2); }
            }
            *dxs = sdx; *dys = sdy;
            return k;
        };
        int n = 0, x0 = INT32_MAX, y0 = INT32_MAX, x1 = -1, y1 = -1, fk = INT32_MAX;
        long long sx = 0, sy = 0, cross = 0;
        long long m20 = 0, m11 = 0, m02 = 0, m30 = 0, m21 = 0, m12 = 0, m03 = 0, m40 = 0, m31 = 0, m22 = 0, m13 = 0, m04 = 0;
        for (int i = lane; i < ncnt; i += 32) {
            const uint2 rec = recs[base + i];
            const int x = (int)(rec.x & 0xffffu), y = (int)(rec.x >> 16);
            x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y);
            fk = min(fk, y * W + x);
            int sdx, sdy;
            const int k = arcs_of(rec.y, x, y, &sdx, &sdy);
            n += k; sx += k * x; sy += k * y;
            cross += (long long)x * sdy - (long long)y * sdx;
            // relative coordinates stay below 2^15 (frames are at most 32767 px per side): 2nd-order products fit in
            // int32 and every higher product is one 32x32->64 multiply-add
            const int ex = x - ox, ey = y - oy;
            const int exx = ex * ex, exy = ex * ey, eyy = ey * ey;
            m20 += (long long)k * exx; m11 += (long long)k * exy; m02 += (long long)k * eyy;
            const int kex = k * ex, key = k * ey;
            m30 += (long long)exx * kex; m21 += (long long)exx * key; m12 += (long long)eyy * kex; m03 += (long long)eyy * key;
            const long long q40 = (long long)exx * exx, q31 = (long long)exx * exy, q22 = (long long)exx * eyy,
                            q13 = (long long)exy * eyy, q04 = (long long)eyy * eyy;
            m40 += k * q40; m31 += k * q31; m22 += k * q22; m13 += k * q13; m04 += k * q04;
        }
        n = __reduce_add_sync(0xffffffffu, n);
        x0 = __reduce_min_sync(0xffffffffu, x0); y0 = __reduce_min_sync(0xffffffffu, y0);
        x1 = __reduce_max_sync(0xffffffffu, x1); y1 = __reduce_max_sync(0xffffffffu, y1);
        fk = __reduce_min_sync(0xffffffffu, fk);
        sx = warp_sum(sx); sy = warp_sum(sy); cross = warp_sum(cross);
        STAMP(1, 3);
        long long s_int = 0;
        const bool fitted = contour_is_fitted(n, cross, prm);  // warp-uniform
        if (fitted) {
            m20 = warp_sum(m20); m11 = warp_sum(m11); m02 = warp_sum(m02);
            m30 = warp_sum(m30); m21 = warp_sum(m21); m12 = warp_sum(m12); m03 = warp_sum(m03);
            m40 = warp_sum(m40); m31 = warp_sum(m31); m22 = warp_sum(m22); m13 = warp_sum(m13); m04 = warp_sum(m04);
            // second pass: n * (L1 spread about the mean), exact
            for (int i = lane; i < ncnt; i += 32) {
                const uint2 rec = recs[base + i];
                const int x = (int)(rec.x & 0xffffu), y = (int)(rec.x >> 16);
                int sdx, sdy;
                const int k = arcs_of(rec.y, x, y, &sdx, &sdy);
                s_int += k * (llabs((long long)n * x - sx) + llabs((long long)n * y - sy));
            }
            s_int = warp_sum(s_int);
        }
        STAMP(1, 4);
        if (kFit || lane == 0) {   // (every lane holds the warp totals; with the fit on board all of them go on)
            CompAcc a;
            a.n = n; a.sx = sx; a.sy = sy; a.cross = cross;
            a.xx = m20; a.xy = m11; a.yy = m02; a.xxx = m30; a.xxy = m21; a.xyy = m12; a.yyy = m03;
            a.xxxx = m40; a.xxxy = m31; a.xxyy = m22; a.xyyy = m13; a.yyyy = m04;
            a.s_int = s_int;
            a.ox = ox; a.oy = oy;
            a.bbox[0] = x0; a.bbox[1] = y0; a.bbox[2] = x1; a.bbox[3] = y1;
            a.firstkey = fk; a.fitted = fitted ? 1 : 0;
            if (kFit) {
                CompRec rec;
                if (kFit == 2) fit_component<true>(a, prm, &sb.counters[frame].flags, rec, lane == 0);
                else if (lane == 0) fit_component<false>(a, prm, &sb.counters[frame].flags, rec);
                if (lane == 0) sb.comps[(size_t)frame * C + c] = rec;
            } else {
                accs[c] = a;
            }
        }
        STAMP(1, 5);
    }
    RMCV_GSTAMP_END(g_ns_frame, 1);
}

template <int kFit>
__global__ void __launch_bounds__(128) contour_kernel(const ContourParams p) {
    const int wpc = blockDim.x >> 5;
    contour_body<kFit>(p.g, p.sb, p.prm, blockIdx.x, blockIdx.y * wpc + (threadIdx.x >> 5), gridDim.y * wpc);
}

// ------------------------------------------------------------------------------------------ K_F: fits
// One thread per component over the whole chunk: cv::fitEllipseDirect incl. its fallback, the ratio/tilt gates and the
// rm::lightblob ctor (reference: src/objdetect.cpp:62-84, src/core.cpp:9-19) from the integer sums.
struct FitParams {
    Geometry g;
    SlotBuffers sb;
    rmcv_params prm;
};

__device__ __forceinline__ void fit_body(const Geometry& g, const SlotBuffers& sb, const rmcv_params& prm, int frame, int c_first, int c_stride) {
    const int C = g.C;
    STAMP(2, 0);
    RMCV_GSTAMP_BEGIN(g_ns_frame, 2);
    chain_begin();
    chain_wait();
    const int n_comps = sb.counters[frame].n_comps;
    // a frame has ~36 components: two CTAs of 64 threads per frame cover it in one round (a grid over the capacity C would
    // launch C/64 CTAs per frame that find nothing to do)
    for (int c = c_first; c < n_comps; c += c_stride) {
        const CompAcc& a = sb.acc[(size_t)frame * C + c];
        CompRec rec;
        STAMP(2, 1);
        fit_component<false>(a, prm, &sb.counters[frame].flags, rec);
        STAMP(2, 2);
        sb.comps[(size_t)frame * C + c] = rec;
        STAMP(2, 3);
    }
    RMCV_GSTAMP_END(g_ns_frame, 2);
}

__global__ void __launch_bounds__(64) fit_kernel(const FitParams p) {
    fit_body(p.g, p.sb, p.prm, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------ K_O: order, pairs, output
// One CTA per frame: ordering into cv::findContours order, the O(P^2) pair gates of rm::filter_armours and rm::armour
// geometry (reference: src/objdetect.cpp:114-166, src/core.cpp:21-49) with an order-preserving block compaction, and
// the dense write-out into pinned, device-mapped host memory (space claimed with one atomicAdd per array).
struct OrderParams {
    Geometry g;
    SlotBuffers sb;
    rmcv_params prm;
    int frame_base;
    rmcv_frame_info* o_frames;
    rmcv_contour_info* o_contours;
    rmcv_lightblob* o_blobs;
    rmcv_armour* o_armours;
    int frames;  // frames in this chunk (index of the allocator entry in sb.counters)
    int defer_copy;  // the records are posted by writeout_kernel (large frames: many CTAs per frame share the PCIe writes)
    int stage_blobs; // the positives are staged in shared memory for the pair loops (wide variant, when they fit)
};

// NTMAX: 128 threads for ordinary frames; 512 for frames with large capacities (stress frames: hundreds of blobs, ~125k
// pairs, ~130 KB of records to post over PCIe per frame from a handful of CTAs).  CS > 1: a cluster of CS CTAs per frame
// (Team, above) shares the O(n^2) ordering and the O(P^2) pair loops when the chunk leaves most SMs idle.
template <int CS>
__device__ __forceinline__ void order_body(const OrderParams& p, uint8_t* smem) {
    __shared__ int sh_scan[33];
    __shared__ int s_np_, s_nc_, s_nn_, s_flags_, s_total_, s_off_[3];
    Team<CS> team;
    STAMP(3, 0);
    RMCV_GSTAMP_BEGIN(g_ns_frame, 3);
    int& s_np = *team.on0(&s_np_); int& s_nc = *team.on0(&s_nc_); int& s_nn = *team.on0(&s_nn_);
    int& s_flags = *team.on0(&s_flags_); int& s_total = *team.on0(&s_total_);
    int* s_off = team.on0(s_off_);
    const Geometry& g = p.g;
    const int W = g.W, C = g.C, A = g.A;
    const int frame = blockIdx.x / CS;
    const int ltid = threadIdx.x, LNT = blockDim.x;               // within this CTA
    const int tid = team.rank * LNT + ltid, NT = LNT * CS;        // within the frame's team
    const SlotBuffers& sb = p.sb;
    FrameCounters& fc = sb.counters[frame];
    chain_begin();
    chain_wait();
    const int n_comps = fc.n_comps;
    int32_t* s_keys = reinterpret_cast<int32_t*>(smem);          // every CTA of a team keeps its own copy
    int32_t* s_status = s_keys + C;
    const CompRec* comps = sb.comps + (size_t)frame * C;
    if (tid == 0) { s_np = 0; s_nc = 0; s_nn = 0; s_total = 0; s_flags = fc.flags; }
    for (int c = ltid; c < n_comps; c += LNT) { s_keys[c] = comps[c].firstkey; s_status[c] = comps[c].status; }
    team.sync();
    STAMP(3, 1);
    // ---- order: rank = number of external components with a larger first-pixel key (reverse raster order)
    rmcv_contour_info* oc = sb.s_contours + (size_t)frame * C;
    rmcv_lightblob* ob = sb.s_blobs + (size_t)frame * C;
    rmcv_armour* oa = sb.s_armours + (size_t)frame * A;
    {
        int nc = 0, np = 0, nn = 0;
        for (int i = tid; i < n_comps; i += NT) {
            const int key = s_keys[i];
            if (key < 0) continue;
            const int stt = s_status[i];
            int rank = 0, prank = 0;
            for (int j = 0; j < n_comps; ++j) {
                if (s_keys[j] > key) { ++rank; prank += (s_status[j] == RMCV_CONTOUR_POSITIVE); }
            }
            const CompRec& c = comps[i];
            rmcv_contour_info info;
            info.first_x = key % W; info.first_y = key / W;
            info.n_points = c.n_points;
            info.status = stt;
            info.area2 = c.area2;
            info.bbox[0] = c.bbox[0]; info.bbox[1] = c.bbox[1];
            info.bbox[2] = c.bbox[2] - c.bbox[0] + 1; info.bbox[3] = c.bbox[3] - c.bbox[1] + 1;
            info.ellipse = c.ellipse;
            info.fit_branch = c.fit_branch;
            info.det0 = c.det0;
            info.blob_index = stt == RMCV_CONTOUR_POSITIVE ? prank : -1;
            oc[rank] = info;
            ++nc;
            if (stt == RMCV_CONTOUR_POSITIVE) { ob[prank] = c.blob; ++np; }
            else if (stt == RMCV_CONTOUR_NEGATIVE) ++nn;
        }
        if (nc) atomicAdd(&s_nc, nc);
        if (np) atomicAdd(&s_np, np);
        if (nn) atomicAdd(&s_nn, nn);
    }
    team.sync();
    STAMP(3, 2);
    // ---- pairs in lexicographic (i,j) order (src/objdetect.cpp:122-163)
    const int P = s_np;
    const rmcv_lightblob* sblob = ob;            // ordinary frames: ~25 positives, read in place (L1/L2)
    if (p.stage_blobs) {                         // large capacities: staged in shared memory for the O(P^2) loops
        rmcv_lightblob* st = reinterpret_cast<rmcv_lightblob*>(s_status + C);
        copy_words(st, ob, (size_t)P * sizeof(rmcv_lightblob), ltid, LNT);
        sblob = st;
        __syncthreads();
    }
    const long long npairs = (long long)P * (P - 1) / 2;
    if (npairs <= 8 * LNT) {
        // few pairs (an ordinary frame): one pair per thread and round, order-preserving block compaction (first CTA)
        if (team.rank == 0) {
            int base = 0;
            for (long long k0 = 0; k0 < npairs; k0 += LNT) {
                const long long k = k0 + ltid;
                bool pass = false;
                int i = 0, j = 0;
                float gates[6];
                if (k < npairs) {
                    pair_from_index(k, P, &i, &j);
                    if (LNT >= 512) {   // a handful of frames, one pair per thread: the gate values at once (same verdict), not
                                        // the cheap verdict first and the values, with a second atan2, for the survivors
                        pass = pair_gates(sblob[i], sblob[j], p.prm, gates);
                    } else {
                        pass = pair_passes(sblob[i], sblob[j], p.prm);
                        if (pass) pair_gates(sblob[i], sblob[j], p.prm, gates);
                    }
                }
                int total;
                const int pos = base + block_excl_scan(pass ? 1 : 0, &total, sh_scan);
                if (pass) {
                    if (pos < A) {
                        rmcv_armour a;
                        make_armour(sblob[i], sblob[j], &a);
                        a.i = i; a.j = j;
                        for (int t = 0; t < 6; ++t) a.gates[t] = gates[t];
                        oa[pos] = a;
                    } else {
                        atomicOr(&s_flags, RMCV_FRAME_OVERFLOW_ARMOURS);
                    }
                }
                base += total;
            }
            if (ltid == 0) s_total = base;
        }
        team.sync();
    } else {
        // many pairs (stress frames, ~125k): no barrier inside the O(P^2) loop.  A warp takes rows i and P-2-i of the pair
        // triangle together (P-1 pairs per unit) with its lanes over j; pass 1 counts the pairs of a row that pass the gates,
        // a scan of the counts gives every row its first output slot, pass 2 records the passing pairs there in j order.
        int* s_rowpos = team.on0(s_keys);        // the ordering keys are no longer needed; rank 0's array
        const int nrows = P - 1, nunits = (nrows + 1) / 2;
        const int lane = ltid & 31, warp = tid >> 5, nwarps = NT >> 5;
        for (int u = warp; u < nunits; u += nwarps) {
            for (int half = 0; half < 2; ++half) {
                const int i = half == 0 ? u : nrows - 1 - u;
                if (half == 1 && i == u) break;
                const rmcv_lightblob bi = sblob[i];
                int cnt = 0;
                for (int j0 = i + 1; j0 < P; j0 += 32) {
                    const int j = j0 + lane;
                    const bool pass = j < P && pair_passes(bi, sblob[j], p.prm);
                    cnt += __popc(__ballot_sync(0xffffffffu, pass));
                }
                if (lane == 0) s_rowpos[i] = cnt;
            }
        }
        team.sync();
        if (team.rank == 0) {
            int base = 0;
            for (int r0 = 0; r0 < nrows; r0 += LNT) {
                const int r = r0 + ltid;
                const int v = r < nrows ? s_rowpos[r] : 0;
                int total;
                const int ex = block_excl_scan(v, &total, sh_scan);
                if (r < nrows) s_rowpos[r] = base + ex;
                base += total;
            }
            if (ltid == 0) s_total = base;
        }
        team.sync();
        for (int u = warp; u < nunits; u += nwarps) {
            for (int half = 0; half < 2; ++half) {
                const int i = half == 0 ? u : nrows - 1 - u;
                if (half == 1 && i == u) break;
                const rmcv_lightblob bi = sblob[i];
                int pos = s_rowpos[i];
                for (int j0 = i + 1; j0 < P; j0 += 32) {
                    const int j = j0 + lane;
                    const bool pass = j < P && pair_passes(bi, sblob[j], p.prm);
                    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const int my = pos + __popc(bal & ((1u << lane) - 1u));
                        if (my < A) { oa[my].i = i; oa[my].j = j; }   // the pair; the armour is built below by all threads
                        else atomicOr(&s_flags, RMCV_FRAME_OVERFLOW_ARMOURS);
                    }
                    pos += __popc(bal);
                }
            }
        }
        team.sync();
        // the armours themselves (double-precision trigonometry): one per thread and round, no divergence
        const int n_pass = min(s_total, A);
        for (int k = tid; k < n_pass; k += NT) {
            const int i = oa[k].i, j = oa[k].j;
            float gates[6];
            pair_gates(sblob[i], sblob[j], p.prm, gates);
            rmcv_armour a;
            make_armour(sblob[i], sblob[j], &a);
            a.i = i; a.j = j;
            for (int t = 0; t < 6; ++t) a.gates[t] = gates[t];
            oa[k] = a;
        }
    }
    const int n_arm = min(s_total, A);
    STAMP(3, 3);
    // ---- claim dense space in the chunk's region of the pinned result arrays, write out
    if (tid == 0) {
        FrameCounters& al = sb.counters[p.frames];
        s_off[0] = atomicAdd(&al.n_runs, s_nc);
        s_off[1] = atomicAdd(&al.n_comps, P);
        s_off[2] = atomicAdd(&al.n_holes, n_arm);
        fc.flags = s_flags;
        fc.n_contours = s_nc; fc.n_positive = P; fc.n_negative = s_nn; fc.n_armours = n_arm;
    }
    team.sync();
    STAMP(3, 4);
    const int nc_all = s_nc;
    const size_t base_c = (size_t)p.frame_base * C + s_off[0];
    const size_t base_b = (size_t)p.frame_base * C + s_off[1];
    const size_t base_a = (size_t)p.frame_base * A + s_off[2];
    if (tid == 0) {
        rmcv_frame_info fi;
        fi.n_contours = nc_all; fi.n_positive = P; fi.n_negative = s_nn; fi.n_armours = n_arm;
        fi.contour_offset = (int32_t)base_c; fi.blob_offset = (int32_t)base_b; fi.armour_offset = (int32_t)base_a;
        fi.flags = s_flags;
        p.o_frames[p.frame_base + frame] = fi;
        sb.arm_offset[4 * frame] = (int32_t)base_a;
        sb.arm_offset[4 * frame + 1] = (int32_t)base_c;
        sb.arm_offset[4 * frame + 2] = (int32_t)base_b;
    }
    if (!p.defer_copy) {
        copy_words(p.o_contours + base_c, oc, (size_t)nc_all * sizeof(rmcv_contour_info), tid, NT);
        copy_words(p.o_blobs + base_b, ob, (size_t)P * sizeof(rmcv_lightblob), tid, NT);
        copy_words(p.o_armours + base_a, oa, (size_t)n_arm * sizeof(rmcv_armour), tid, NT);
    }
    if (CS > 1) team.sync();   // rank 0's shared memory is read by the whole team until here
#ifdef RMCV_STAMPS
    __syncthreads();
    STAMP(3, 5);
    RMCV_GSTAMP_END(g_ns_frame, 3);
#endif
}

template <int NTMAX, int CS>
__global__ void __launch_bounds__(NTMAX) order_kernel(const OrderParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    order_body<CS>(p, smem);
}

// Dense write-out of a large frame's records into the pinned result arrays, kWriteSplit CTAs per frame: with a handful of
// frames per chunk the single order CTA of a frame would be the only one posting its ~130 KB over PCIe.
constexpr int kWriteSplit = 16;
__global__ void __launch_bounds__(256) writeout_kernel(const OrderParams p) {
    const int frame = blockIdx.y, part = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const SlotBuffers& sb = p.sb;
    const FrameCounters& fc = sb.counters[frame];
    const int C = p.g.C, A = p.g.A;
    chain_begin();
    chain_wait();
    auto slice = [&](uint8_t* dst, const uint8_t* src, size_t records, size_t rec_bytes) {
        const size_t per = (records + kWriteSplit - 1) / kWriteSplit;
        const size_t r0 = (size_t)part * per, r1 = r0 + per < records ? r0 + per : records;
        if (r0 < r1) copy_words(dst + r0 * rec_bytes, src + r0 * rec_bytes, (r1 - r0) * rec_bytes, tid, NT);
    };
    slice(reinterpret_cast<uint8_t*>(p.o_armours + sb.arm_offset[4 * frame]), reinterpret_cast<const uint8_t*>(sb.s_armours + (size_t)frame * A),
          (size_t)fc.n_armours, sizeof(rmcv_armour));
    slice(reinterpret_cast<uint8_t*>(p.o_contours + sb.arm_offset[4 * frame + 1]), reinterpret_cast<const uint8_t*>(sb.s_contours + (size_t)frame * C),
          (size_t)fc.n_contours, sizeof(rmcv_contour_info));
    slice(reinterpret_cast<uint8_t*>(p.o_blobs + sb.arm_offset[4 * frame + 2]), reinterpret_cast<const uint8_t*>(sb.s_blobs + (size_t)frame * C),
          (size_t)fc.n_positive, sizeof(rmcv_lightblob));
}

// Team launches: a thread-block cluster of CS CTAs per frame.  All clusters of a launch should be co-resident (a second
// wave doubles the kernel's time), and a cluster lives inside one GPC, so the largest CS in {8, 4, 2} whose clusters all
// fit at once is chosen (cudaOccupancyMaxActiveClusters); 0 = no team.
template <class Params>
static int pick_team(void (*k8)(Params), void (*k4)(Params), void (*k2)(Params), int frames, int threads, size_t smem) {
    void (*ks[3])(Params) = {k8, k4, k2};
    const int cs[3] = {8, 4, 2};
    for (int i = 0; i < 3; ++i) {
        if (cudaFuncSetAttribute(ks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)frames * cs[i]); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs[i]; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, ks[i], &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
        if (n >= frames) return cs[i];
    }
    return 0;
}
template <class Params>
static cudaError_t launch_team(void (*k)(Params), int cs, int frames, int threads, size_t smem, cudaStream_t st, const Params& p,
                               bool chained = false) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)frames * cs); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // a link of a small chunk's chain (common.cuh)
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = chained ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, k, p);
}

// Enqueues the labelling stages of one chunk.  stage_done(i) is called after each kernel (profiling events).
cudaError_t launch_frames(const FrameLaunch& L, const rmcv_params& prm, int max_smem_optin, cudaStream_t st, int64_t* launches,
                          void (*stage_done)(void*, int, cudaStream_t), void* stage_arg) {
    cudaError_t e;
    auto done = [&](int stage) { if (stage_done) stage_done(stage_arg, stage, st); };
    const Tuning& tune = tuning();
    const int small_batch = small_batch_limit();   // at most this many frames: per-frame kernels take their wide variants
    const bool chained = L.chained != 0;
    if (!L.emit_done) {   // K_E (unless the pixel kernel emitted the runs / records itself)
        EmitLaunch el;
        el.bits = L.sb->bits; el.W = L.g.W; el.H = L.g.H; el.batch = L.frames;
        el.rows = L.sb->rows; el.run_x = L.sb->run_x; el.run_y = L.sb->run_y; el.counters = L.sb->counters; el.R = L.g.R;
        el.recs = L.sb->recs; el.PC = L.g.PC;
        if (L.flags_bh > 0) { el.band_flags = L.sb->band_flags; el.flag_bh = L.flags_bh; el.flag_bands = (L.g.H + L.flags_bh - 1) / L.flags_bh; }
        e = launch_emit(el, st, launches, chained);
        if (e != cudaSuccess) return e;
        done(RMCV_STAGE_EMIT);
    } else {
        done(RMCV_STAGE_EMIT);
    }
    {   // K_L
        LabelParams p;
        p.g = L.g; p.sb = *L.sb;
        int Rs = tune.frame_rs >= 0 ? tune.frame_rs : 3072;  // four CTAs per SM at 1280x1024 (gpurun_out/exp_rs*.json)
        if (Rs > L.g.R) Rs = L.g.R;
        size_t smem = label_smem_bytes(L.g.H, Rs, L.g.C);
        while (smem > (size_t)max_smem_optin && Rs > 0) { Rs = Rs > 1024 ? Rs - 1024 : 0; smem = label_smem_bytes(L.g.H, Rs, L.g.C); }
        if (smem > (size_t)max_smem_optin) return cudaErrorInvalidConfiguration;
        p.Rs = Rs;
        if (tune.chain_pad > 0 && (size_t)tune.chain_pad > smem && tune.chain_pad <= max_smem_optin) smem = (size_t)tune.chain_pad;
        if (tune.label_minsmem > 0) {   // experiment: cap the CTAs per SM by padding shared memory
            const size_t m = (size_t)tune.label_minsmem;
            if (m > smem && m <= (size_t)max_smem_optin) smem = m;
        }
        // 1024 threads: frames whose runs stay in global memory (run indices beyond 16 bits), and small batches, where a
        // frame's latency matters and the SMs are idle anyway
        const bool big = (L.g.R > 65535 || L.frames <= small_batch) && !tune.label_small;
        // a team of 8 CTAs per frame: frames whose runs live in global memory anyway (above 2 Mpx), when the chunk leaves
        // most SMs idle.  RMCV_WIDE_LABEL=0 switches it off, =1 forces it for every global-memory frame.
        int cs = 0;
        const size_t smem_team = label_smem_bytes(L.g.H, 0, L.g.C);
        if (L.g.R > 65535 && tune.wide_label != 0 && L.frames < 148)
            cs = pick_team<LabelParams>(label_kernel<1024, 1, 8>, label_kernel<1024, 1, 4>, label_kernel<1024, 1, 2>, L.frames, 1024, smem_team);
        if (cs) {
            p.Rs = 0;
            e = launch_team<LabelParams>(cs == 8 ? label_kernel<1024, 1, 8> : cs == 4 ? label_kernel<1024, 1, 4> : label_kernel<1024, 1, 2>,
                                         cs, L.frames, 1024, smem_team, st, p, chained);
            if (e != cudaSuccess) return e;
        } else if (big) {
            e = cudaFuncSetAttribute(label_kernel<1024, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = launch_chained<LabelParams>(label_kernel<1024, 1, 1>, dim3(L.frames), dim3(1024), smem, st, chained, p);
            if (e != cudaSuccess) return e;
        } else {
            e = cudaFuncSetAttribute(label_kernel<256, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = launch_chained<LabelParams>(label_kernel<256, 4, 1>, dim3(L.frames), dim3(256), smem, st, chained, p);
            if (e != cudaSuccess) return e;
        }
        if (launches) ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        done(RMCV_STAGE_LABEL);
    }
    // chunks of a few ordinary frames (latency mode): the fits ride on the contour kernel's warps
    const int tiny = tune.fit_in_contour >= 0 ? tune.fit_in_contour : small_batch;   // frames; 0 = never
    const bool fit_in_contour = L.frames <= tiny && L.g.C <= 512;
    {   // K_C
        ContourParams p;
        p.g = L.g; p.sb = *L.sb; p.prm = prm;
        // CTAs of four warps per frame: eight for small chunks (a warp per component and round: latency), two for large ones
        // (1024 frames: 2048 fat CTAs instead of 8192 thin ones, 1.188 -> 1.175 ms per step; gpu_exp_t.sh)
        // (a handful of frames: sixteen, so that every component of an ordinary frame has a warp in the first round)
        int gy = tune.contour_gy > 0 ? tune.contour_gy : (L.frames >= 512 ? 1 : L.frames >= 128 ? 2 : (L.frames <= 4 || fit_in_contour) ? 16 : 8);
        if (gy * 4 > L.g.C) gy = (L.g.C + 3) / 4;
        if (gy < 1) gy = 1;
        dim3 grid(L.frames, gy);
        const size_t pad = tune.chain_pad > 0 ? (size_t)tune.chain_pad : 0;
        if (fit_in_contour) {
            if ((e = launch_chained<ContourParams>(tune.warp_fit == 0 ? contour_kernel<1> : contour_kernel<2>, grid, dim3(128), 0, st, chained, p)) != cudaSuccess) return e;
        } else {
            if (pad > 48 * 1024) cudaFuncSetAttribute(contour_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
            if ((e = launch_chained<ContourParams>(contour_kernel<0>, grid, dim3(128), pad, st, chained, p)) != cudaSuccess) return e;
        }
        if (launches) ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        done(RMCV_STAGE_CONTOUR);
    }
    if (fit_in_contour) {
        done(RMCV_STAGE_FIT);
    } else {   // K_F
        FitParams p;
        p.g = L.g; p.sb = *L.sb; p.prm = prm;
        // ordinary frames: two CTAs per frame; large capacities (stress frames, ~510 components): one per 64 components
        const int gx_full = (L.g.C + 63) / 64;
        dim3 grid(L.g.C > 512 ? gx_full : (gx_full < 2 ? gx_full : 2), L.frames);
        const size_t pad = tune.chain_pad > 0 ? (size_t)tune.chain_pad : 0;
        if (pad > 48 * 1024) cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        if ((e = launch_chained<FitParams>(fit_kernel, grid, dim3(64), pad, st, chained, p)) != cudaSuccess) return e;
        if (launches) ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        done(RMCV_STAGE_FIT);
    }
    {   // K_O on its own stream: its write-out over PCIe stalls on store back-pressure while occupying almost no SM
        // resources, so the next chunk's kernels need not queue behind it
        cudaStream_t so = L.st_out ? L.st_out : st;
        if (so != st) {
            if ((e = cudaEventRecord(L.sb->ev_fit, st)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(so, L.sb->ev_fit, 0)) != cudaSuccess) return e;
        }
        OrderParams p;
        p.g = L.g; p.sb = *L.sb; p.prm = prm; p.frame_base = L.frame_base;
        p.o_frames = L.o_frames; p.o_contours = L.o_contours; p.o_blobs = L.o_blobs; p.o_armours = L.o_armours;
        p.frames = L.frames;
        const bool big = L.g.R > 65535 || L.g.C > 512 || L.frames <= small_batch;
        size_t smem = (size_t)2 * L.g.C * 4 + 16;   // keys, status
        p.stage_blobs = big && smem + (size_t)L.g.C * sizeof(rmcv_lightblob) <= (size_t)max_smem_optin ? 1 : 0;
        if (p.stage_blobs) smem += (size_t)L.g.C * sizeof(rmcv_lightblob);
        if (smem > (size_t)max_smem_optin) return cudaErrorInvalidConfiguration;
        if (tune.chain_pad > 0 && (size_t)tune.chain_pad > smem) smem = (size_t)tune.chain_pad;
        p.defer_copy = (L.g.R > 65535 || L.g.C > 512) ? 1 : 0;
        // a team of 8 CTAs per frame for frames with large capacities when the chunk leaves most SMs idle (see label)
        int cs = 0;
        if (p.defer_copy && tune.wide_label != 0 && L.frames < 148)
            cs = pick_team<OrderParams>(order_kernel<512, 8>, order_kernel<512, 4>, order_kernel<512, 2>, L.frames, 512, smem);
        if (!cs && smem > 48 * 1024) {
            e = big ? cudaFuncSetAttribute(order_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                    : cudaFuncSetAttribute(order_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        if (cs) {
            e = launch_team<OrderParams>(cs == 8 ? order_kernel<512, 8> : cs == 4 ? order_kernel<512, 4> : order_kernel<512, 2>, cs, L.frames, 512,
                                         smem, so, p, chained && so == st);
            if (e != cudaSuccess) return e;
        } else if (big) {
            if ((e = launch_chained<OrderParams>(order_kernel<512, 1>, dim3(L.frames), dim3(512), smem, so, chained && so == st, p)) != cudaSuccess) return e;
        } else {
            if ((e = launch_chained<OrderParams>(order_kernel<128, 1>, dim3(L.frames), dim3(128), smem, so, chained && so == st, p)) != cudaSuccess) return e;
        }
        if (p.defer_copy == 1) {
            if (launches) ++*launches;
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = launch_chained<OrderParams>(writeout_kernel, dim3(kWriteSplit, L.frames), dim3(256), 0, so, chained && so == st, p)) != cudaSuccess) return e;
        }
        if (launches) ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (L.o_poses && L.camera) {   // f1 fused: rm::solve_PnP for every armour of the chunk
            if ((e = launch_chunk_poses(*L.sb, L.frames, L.g.A, L.o_poses, *L.camera, so, launches)) != cudaSuccess) return e;
        }
        if (stage_done) stage_done(stage_arg, RMCV_STAGE_ORDER, so);
    }
    return cudaSuccess;
}

#ifdef RMCV_STAMPS
}  // namespace rmcv
extern "C" int rmcv_debug_stamps(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, rmcv::g_stamps, sizeof(rmcv::g_stamps));
}
RMCV_GSTAMP_GETTER(rmcv_debug_ns_frame, rmcv::g_ns_frame)
extern "C" int rmcv_debug_fit_marks(long long* out) { return (int)cudaMemcpyFromSymbol(out, rmcv_fit_marks, sizeof(rmcv_fit_marks)); }
namespace rmcv {
#endif

// ------------------------------------------------------------------------------------------ standalone a2 / a3
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
