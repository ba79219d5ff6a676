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
