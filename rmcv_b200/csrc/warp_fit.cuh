// Warp-cooperative helpers shared by the contour kernel, the standalone rm::filter_lightblobs kernel and the legacy
// kernels: warp sums and the point-based ellipse fit (cv::fitEllipseDirect incl. its fallback, src/objdetect.cpp:68).
#pragma once
#include "blob_math.cuh"

namespace rmcv {

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
// caller-supplied contours.  `pass(fn)` must call fn(x, y) for every contour point (with multiplicity), partitioned
// over the lanes of the warp.  All lanes return identical results.
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
    const bool ok = direct_fit(m, scale, cx, cy, ell, &det);
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

}  // namespace rmcv
