// f1: rm::solve_PnP per armour (reference: src/mobility.cpp:166-190, called at executable/main.cpp:183-192)
// = cv::solvePnP(SOLVEPNP_IPPE_SQUARE) on armour.vertices, followed by the caller's camera -> world transform of tvec.
// One thread per armour, fp64; the arithmetic lives in pnp_math.cuh.
#include "common.cuh"
#include "pnp_math.cuh"

namespace rmcv {

struct PnpParams {
    const rmcv_armour* armours; int n;
    double K[9], dist[5], M[16];
    int has_M;
    float w, h, roi_x, roi_y;
    rmcv_pose* out;
};

__global__ void __launch_bounds__(64) pnp_kernel(const PnpParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float pts[4][2];
    for (int k = 0; k < 4; ++k) { pts[k][0] = p.armours[i].vertices[k][0]; pts[k][1] = p.armours[i].vertices[k][1]; }
    PnpResult r;
    rmcv_pose o;
    memset(&o, 0, sizeof(o));
    o.ok = solve_pnp_square(pts, p.K, p.dist, p.w, p.h, p.roi_x, p.roi_y, &r) ? 1 : 0;
    if (o.ok) {
        for (int k = 0; k < 3; ++k) { o.rvec[k] = r.rvec[k]; o.tvec[k] = r.tvec[k]; }
        o.reproj_err = r.err;
        for (int k = 0; k < 3; ++k)   // executable/main.cpp:186-192: world = M * [tvec; 1]
            o.position[k] = p.has_M ? p.M[4 * k] * r.tvec[0] + p.M[4 * k + 1] * r.tvec[1] + p.M[4 * k + 2] * r.tvec[2] + p.M[4 * k + 3]
                                    : r.tvec[k];
    }
    p.out[i] = o;
}

// Fused variant: thread (frame, k) solves armour k of the frame straight from / into the dense result arrays.
struct ChunkPoseParams {
    const FrameCounters* counters; const int32_t* arm_offset;
    const rmcv_armour* armours;   // [frames][A] device staging of the write-out kernel
    rmcv_pose* poses;             // dense pinned result array
    int A;
    CameraSetup cam;
};

__global__ void __launch_bounds__(64) chunk_pose_kernel(const ChunkPoseParams p) {
    const int frame = blockIdx.x;
    const int n = p.counters[frame].n_armours;
    const size_t base = (size_t)p.arm_offset[4 * frame];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const rmcv_armour& a = p.armours[(size_t)frame * p.A + k];
        float pts[4][2];
        for (int i = 0; i < 4; ++i) { pts[i][0] = a.vertices[i][0]; pts[i][1] = a.vertices[i][1]; }
        PnpResult r;
        rmcv_pose o;
        memset(&o, 0, sizeof(o));
        o.ok = solve_pnp_square(pts, p.cam.K, p.cam.dist, p.cam.w, p.cam.h, 0.f, 0.f, &r) ? 1 : 0;
        if (o.ok) {
            for (int i = 0; i < 3; ++i) { o.rvec[i] = r.rvec[i]; o.tvec[i] = r.tvec[i]; }
            o.reproj_err = r.err;
            for (int i = 0; i < 3; ++i)
                o.position[i] = p.cam.has_M ? p.cam.M[4 * i] * r.tvec[0] + p.cam.M[4 * i + 1] * r.tvec[1] + p.cam.M[4 * i + 2] * r.tvec[2] + p.cam.M[4 * i + 3]
                                            : r.tvec[i];
        }
        p.poses[base + k] = o;
    }
}

cudaError_t launch_chunk_poses(const SlotBuffers& sb, int frames, int A, rmcv_pose* o_poses, const CameraSetup& cam,
                               cudaStream_t st, int64_t* launches) {
    if (frames <= 0) return cudaSuccess;
    ChunkPoseParams p;
    p.counters = sb.counters; p.arm_offset = sb.arm_offset; p.armours = sb.s_armours; p.poses = o_poses; p.A = A; p.cam = cam;
    chunk_pose_kernel<<<frames, 64, 0, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_pnp(const rmcv_armour* d_armours, int n, const double K[9], const double dist[5], float w, float h,
                       float roi_x, float roi_y, const double* cam2world, rmcv_pose* d_out, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    PnpParams p;
    p.armours = d_armours; p.n = n;
    for (int i = 0; i < 9; ++i) p.K[i] = K[i];
    for (int i = 0; i < 5; ++i) p.dist[i] = dist ? dist[i] : 0.0;
    p.has_M = cam2world != nullptr;
    for (int i = 0; i < 16; ++i) p.M[i] = cam2world ? cam2world[i] : 0.0;
    p.w = w; p.h = h; p.roi_x = roi_x; p.roi_y = roi_y;
    p.out = d_out;
    pnp_kernel<<<(n + 63) / 64, 64, 0, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
