// TEST INFRASTRUCTURE ONLY — never linked into librmcv_b200.so.
// Compiles rmcv_b200/csrc/blob_math.cuh (the per-blob / per-pair device arithmetic) for the host so that the
// numerics can be checked against the cv2 oracle in a GPU-less container (tests/test_hostmath.py).
#include <cstring>
#include "../../rmcv_b200/csrc/blob_math.cuh"
#include "../../rmcv_b200/csrc/pnp_math.cuh"
#include "../../rmcv_b200/csrc/calipers.cuh"
#include <vector>

using namespace rmcv;

extern "C" {

// Mirrors what the blob kernel does with a contour point multiset: mean, L1 spread, centred moments, direct fit,
// fallback fit with float centring.  Returns the RMCV_FIT_* branch.
int hm_fit_points(const int32_t* xy, int n, rmcv_rotated_rect* box, double* det0) {
    double sx = 0, sy = 0;
    for (int i = 0; i < n; ++i) { sx += xy[2 * i]; sy += xy[2 * i + 1]; }
    const double cx = sx / n, cy = sy / n;
    Moments m; moments_zero(m);
    double s = 0;
    for (int i = 0; i < n; ++i) {
        const double dx = xy[2 * i] - cx, dy = xy[2 * i + 1] - cy;
        s += fabs(dx) + fabs(dy);
        moments_add(m, dx, dy);
    }
    double scale = 100.0 / (s > RMCV_FLT_EPSILON ? s : RMCV_FLT_EPSILON);
    if (direct_fit(m, scale, cx, cy, box, det0)) return RMCV_FIT_DIRECT;
    const float c32x = (float)sx / (float)n, c32y = (float)sy / (float)n;  // exact while sums < 2^24
    moments_zero(m);
    double s2 = 0;
    for (int i = 0; i < n; ++i) {
        const float fx = fsub((float)xy[2 * i], c32x), fy = fsub((float)xy[2 * i + 1], c32y);
        s2 += (double)fadd(fabsf(fx), fabsf(fy));
        moments_add(m, (double)fx, (double)fy);
    }
    scale = 100.0 / (s2 > RMCV_FLT_EPSILON ? s2 : RMCV_FLT_EPSILON);
    nodirect_fit(m, scale, c32x, c32y, box);
    return RMCV_FIT_FALLBACK;
}

// The integer-sum route the kernels take: exact sums about an origin, then fit_contour.
int hm_fit_points_int(const int32_t* xy, int n, int ox, int oy, const rmcv_params* prm, rmcv_rotated_rect* box, double* det0, int* status) {
    ContourSums c;
    sums_zero(c, ox, oy);
    for (int i = 0; i < n; ++i) {
        sums_add_point(c, xy[2 * i], xy[2 * i + 1]);
        const int j = i == 0 ? n - 1 : i - 1;
        c.cross += (long long)xy[2 * j] * xy[2 * i + 1] - (long long)xy[2 * j + 1] * xy[2 * i];
    }
    for (int i = 0; i < n; ++i) {
        const long long ax = c.n * xy[2 * i] - c.sx, ay = c.n * xy[2 * i + 1] - c.sy;
        c.s_int += (ax < 0 ? -ax : ax) + (ay < 0 ? -ay : ay);
    }
    int branch; float d0; rmcv_lightblob blob;
    fit_contour(c, *prm, status, &branch, &d0, box, &blob);
    *det0 = (double)d0;
    return branch;
}

void hm_make_lightblob(const rmcv_rotated_rect* box, int target, rmcv_lightblob* out) { make_lightblob(*box, target, out); }
int hm_blob_gates(const rmcv_rotated_rect* e, const rmcv_params* p) { return blob_gates(*e, *p); }
int hm_pair_gates(const rmcv_lightblob* a, const rmcv_lightblob* b, const rmcv_params* p, float* gates) {
    return pair_gates(*a, *b, *p, gates) ? 1 : 0;
}
void hm_make_armour(const rmcv_lightblob* a, const rmcv_lightblob* b, rmcv_armour* out) {
    memset(out, 0, sizeof(*out));
    make_armour(*a, *b, out);
}
int hm_solve_pnp(const float* pts, const double* K, const double* dist, float w, float h, float rx, float ry, double* rvec, double* tvec) {
    PnpResult r;
    const bool ok = solve_pnp_square(reinterpret_cast<const float (*)[2]>(pts), K, dist, w, h, rx, ry, &r);
    for (int i = 0; i < 3; ++i) { rvec[i] = r.rvec[i]; tvec[i] = r.tvec[i]; }
    return ok ? 1 : 0;
}

// cv::minAreaRect the way legacy.cu evaluates it: gift-wrapped hull (smallest contour index among coincident points),
// OpenCV's hull order, literal float32 rotating calipers (calipers.cuh).
void hm_min_area_rect(const int32_t* xy, int n, rmcv_rotated_rect* box) {
    memset(box, 0, sizeof(*box));
    if (n <= 0) return;
    std::vector<int32_t> hull((size_t)3 * n + 3), tmp((size_t)3 * n + 3);
    int s = 0;
    for (int i = 1; i < n; ++i)
        if (xy[2 * i + 1] < xy[2 * s + 1] || (xy[2 * i + 1] == xy[2 * s + 1] && xy[2 * i] < xy[2 * s])) s = i;
    const int sx = xy[2 * s], sy = xy[2 * s + 1];
    int h = 0, cx = sx, cy = sy, ci = s;
    while (h < n) {
        hull[3 * h] = cx; hull[3 * h + 1] = cy; hull[3 * h + 2] = ci;
        ++h;
        int bx = cx, by = cy, bi = -1;
        for (int i = 0; i < n; ++i) {
            const int px = xy[2 * i], py = xy[2 * i + 1];
            if (px == cx && py == cy) continue;
            if (bi < 0 || (px == bx && py == by ? false : hull_better_wrap(cx, cy, bx, by, px, py))) { bx = px; by = py; bi = i; }
        }
        if (bi < 0 || (bx == sx && by == sy)) break;
        cx = bx; cy = by; ci = bi;
    }
    if (h >= 3) hull_to_cv_order(hull.data(), h, tmp.data());
    else if (h == 2 && (hull[3] > hull[0] || (hull[3] == hull[0] && hull[4] > hull[1]))) {   // OpenCV: (max x, max y) first
        for (int q = 0; q < 3; ++q) { const int32_t t = hull[q]; hull[q] = hull[3 + q]; hull[3 + q] = t; }
    }
    min_area_rect_from_hull(hull.data(), h, box);
}
}
