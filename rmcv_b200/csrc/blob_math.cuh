// Per-blob and per-pair arithmetic of the hot path, written once as host+device inline functions:
//   cv::fitEllipseDirect (+ its fitEllipseNoDirect fallback)        src/objdetect.cpp:68   
//   rm::lightblob ctor + reorder_vertices + RotatedRect::points      src/core.cpp:9-19, 265-283
//   rm::filter_armours gates                                         src/objdetect.cpp:131-159
//   rm::armour ctor + PointDistance/ExtendCord/CalcPerspective       src/core.cpp:21-49, 285-404
// The CUDA kernels call these on the device.  The same header compiles under g++ so that
// tests/hostmath (test infrastructure only, never linked into the product library) can check the
// numerics against the cv2 oracle in a GPU-less container.
//
// fp32 results must round exactly like the reference's C++ (no FMA contraction), hence the explicit
// __f*_rn intrinsics on the device; transcendental/sqrt/pow calls follow the reference's overload
// resolution (double intermediates).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/rmcv_b200.h"

#if defined(__CUDACC__)
#define RMCV_HD __host__ __device__ __forceinline__
#else
#define RMCV_HD inline
#endif

// Section marks of the ellipse fit for scripts/phase_stamps.py (a -DRMCV_STAMPS build only; the product has none).
#if defined(RMCV_STAMPS) && defined(__CUDACC__)
static __device__ long long rmcv_fit_marks[16];
#if defined(__CUDA_ARCH__)
#define RMCV_FIT_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) rmcv_fit_marks[i] = clock64(); } while (0)
#else
#define RMCV_FIT_MARK(i) do { } while (0)
#endif
#else
#define RMCV_FIT_MARK(i) do { } while (0)
#endif

namespace rmcv {

#if defined(__CUDA_ARCH__)
RMCV_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
RMCV_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
RMCV_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
RMCV_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
RMCV_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
RMCV_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
#else
RMCV_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
RMCV_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
RMCV_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
RMCV_HD float fdiv(float a, float b) { volatile float r = a / b; return r; }
RMCV_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
RMCV_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
#endif

#define RMCV_PI 3.1415926535897932384626433832795
#define RMCV_FLT_EPSILON 1.1920928955078125e-07

// Centred, unscaled moment sums of the contour point multiset.
struct Moments {
    double n;
    double x, y;
    double xx, xy, yy;
    double xxx, xxy, xyy, yyy;
    double xxxx, xxxy, xxyy, xyyy, yyyy;
};

RMCV_HD void moments_zero(Moments& m) {
    m.n = m.x = m.y = m.xx = m.xy = m.yy = m.xxx = m.xxy = m.xyy = m.yyy = 0.0;
    m.xxxx = m.xxxy = m.xxyy = m.xyyy = m.yyyy = 0.0;
}

RMCV_HD void moments_add(Moments& m, double dx, double dy) {
    const double xx = dx * dx, xy = dx * dy, yy = dy * dy;
    m.n += 1.0;
    m.x += dx; m.y += dy;
    m.xx += xx; m.xy += xy; m.yy += yy;
    m.xxx += xx * dx; m.xxy += xx * dy; m.xyy += dx * yy; m.yyy += yy * dy;
    m.xxxx += xx * xx; m.xxxy += xx * xy; m.xxyy += xx * yy; m.xyyy += xy * yy; m.yyyy += yy * yy;
}

RMCV_HD double det3(const double a[3][3]) {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}

RMCV_HD void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Real eigenvalues of a 3x3 matrix (roots of the characteristic cubic), Newton-polished.  Returns count.
RMCV_HD int eig3_values(const double M[3][3], double lam[3]) {
    const double tr = M[0][0] + M[1][1] + M[2][2];
    const double c1 = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) + (M[0][0] * M[2][2] - M[0][2] * M[2][0]) +
                      (M[1][1] * M[2][2] - M[1][2] * M[2][1]);
    const double dt = det3(M);
    // l^3 - tr l^2 + c1 l - dt = 0 ; substitute l = t + tr/3
    const double a = -tr, b = c1, c = -dt;
    const double p = b - a * a / 3.0;
    const double q = 2.0 * a * a * a / 27.0 - a * b / 3.0 + c;
    const double disc = q * q / 4.0 + p * p * p / 27.0;
    int nroots;
    if (disc <= 0.0 && p < 0.0) {
        const double r = sqrt(-p / 3.0);
        double cosarg = -q / (2.0 * r * r * r);
        cosarg = cosarg > 1.0 ? 1.0 : (cosarg < -1.0 ? -1.0 : cosarg);
        const double phi = acos(cosarg);
        for (int k = 0; k < 3; ++k) lam[k] = 2.0 * r * cos((phi - 2.0 * RMCV_PI * k) / 3.0) - a / 3.0;
        nroots = 3;
    } else {
        const double sq = sqrt(disc > 0.0 ? disc : 0.0);
        const double u = cbrt(-q / 2.0 + sq), v = cbrt(-q / 2.0 - sq);
        lam[0] = u + v - a / 3.0;
        nroots = 1;
    }
    for (int k = 0; k < nroots; ++k) {  // Newton polish on the cubic
        double l = lam[k];
        for (int it = 0; it < 3; ++it) {
            const double f = ((l + a) * l + b) * l + c;
            const double fp = (3.0 * l + 2.0 * a) * l + b;
            if (fp == 0.0) break;
            const double step = f / fp;
            if (!(fabs(step) < fabs(l) * 1e-3 + 1e-300)) break;  // ill-conditioned (clustered roots): keep closed form
            l -= step;
        }
        lam[k] = l;
    }
    return nroots;
}

// Unit eigenvector of M for eigenvalue l (null vector of M - l I by the best-conditioned cross product).
RMCV_HD void eig3_vector(const double M[3][3], double l, double v[3]) {
    double r0[3] = {M[0][0] - l, M[0][1], M[0][2]};
    double r1[3] = {M[1][0], M[1][1] - l, M[1][2]};
    double r2[3] = {M[2][0], M[2][1], M[2][2] - l};
    double c01[3], c02[3], c12[3];
    cross3(r0, r1, c01); cross3(r0, r2, c02); cross3(r1, r2, c12);
    const double n01 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
    const double n02 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
    const double n12 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
    const double* best = c01; double nb = n01;
    if (n02 > nb) { best = c02; nb = n02; }
    if (n12 > nb) { best = c12; nb = n12; }
    const double inv = nb > 0.0 ? 1.0 / sqrt(nb) : 0.0;
    v[0] = best[0] * inv; v[1] = best[1] * inv; v[2] = best[2] * inv;
}

// cv::fitEllipseDirect's first attempt (Halir-Flusser) from the moment sums.  *det_out = |det M|.
// Returns false when |det M| <= 1e-10 (the reference then retries with jitter / falls back; we fall back).
RMCV_HD bool direct_fit(const Moments& m, double scale, double cx, double cy, rmcv_rotated_rect* box, double* det_out) {
    const double n = m.n, s1 = scale, s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2;
    const double inv_n = 1.0 / n;
    const double X = m.x * s1 * inv_n, Y = m.y * s1 * inv_n;
    const double XX = m.xx * s2 * inv_n, XY = m.xy * s2 * inv_n, YY = m.yy * s2 * inv_n;
    const double XXX = m.xxx * s3 * inv_n, XXY = m.xxy * s3 * inv_n, XYY = m.xyy * s3 * inv_n, YYY = m.yyy * s3 * inv_n;
    const double XXXX = m.xxxx * s4 * inv_n, XXXY = m.xxxy * s4 * inv_n, XXYY = m.xxyy * s4 * inv_n,
                 XYYY = m.xyyy * s4 * inv_n, YYYY = m.yyyy * s4 * inv_n;
    const double S1[3][3] = {{XXXX, XXXY, XXYY}, {XXXY, XXYY, XYYY}, {XXYY, XYYY, YYYY}};
    const double S2[3][3] = {{XXX, XXY, XX}, {XXY, XYY, XY}, {XYY, YYY, YY}};
    const double S3[3][3] = {{XX, XY, X}, {XY, YY, Y}, {X, Y, 1.0}};
    const double Ts = det3(S3);
    double adj[3][3];  // adjugate of the symmetric S3
    adj[0][0] = S3[1][1] * S3[2][2] - S3[1][2] * S3[2][1];
    adj[0][1] = S3[0][2] * S3[2][1] - S3[0][1] * S3[2][2];
    adj[0][2] = S3[0][1] * S3[1][2] - S3[0][2] * S3[1][1];
    adj[1][0] = adj[0][1];
    adj[1][1] = S3[0][0] * S3[2][2] - S3[0][2] * S3[2][0];
    adj[1][2] = S3[0][2] * S3[1][0] - S3[0][0] * S3[1][2];
    adj[2][0] = adj[0][2];
    adj[2][1] = adj[1][2];
    adj[2][2] = S3[0][0] * S3[1][1] - S3[0][1] * S3[1][0];
    double TM[3][3];  // -adj(S3) * S2^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) TM[i][j] = -(adj[i][0] * S2[j][0] + adj[i][1] * S2[j][1] + adj[i][2] * S2[j][2]);
    double Mp[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Mp[i][j] = S1[i][j] + (S2[i][0] * TM[0][j] + S2[i][1] * TM[1][j] + S2[i][2] * TM[2][j]) / Ts;
    double M[3][3];
    for (int j = 0; j < 3; ++j) {
        M[0][j] = Mp[2][j] / 2.0;
        M[1][j] = -Mp[1][j];
        M[2][j] = Mp[0][j] / 2.0;
    }
    const double det = fabs(det3(M));
    *det_out = det;
    RMCV_FIT_MARK(2);
    if (!(det > 1.0e-10)) return false;

    double lam[3], pv[3] = {0, 0, 0};
    const int nl = eig3_values(M, lam);
    RMCV_FIT_MARK(3);
    double best_cond = 0.0;
    for (int k = 0; k < nl; ++k) {
        double v[3];
        eig3_vector(M, lam[k], v);
        const double cond = 4.0 * v[0] * v[2] - v[1] * v[1];
        if (k == 0 || cond > best_cond) { best_cond = cond; pv[0] = v[0]; pv[1] = v[1]; pv[2] = v[2]; }
    }
    RMCV_FIT_MARK(4);
    double norm = sqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    const int sg = (pv[0] < 0.0 ? -1 : 1) * (pv[1] < 0.0 ? -1 : 1) * (pv[2] < 0.0 ? -1 : 1);
    if (sg <= 0) norm = -norm;
    pv[0] /= norm; pv[1] /= norm; pv[2] /= norm;
    const double Q0 = (TM[0][0] * pv[0] + TM[0][1] * pv[1] + TM[0][2] * pv[2]) / Ts;
    const double Q1 = (TM[1][0] * pv[0] + TM[1][1] * pv[1] + TM[1][2] * pv[2]) / Ts;
    const double Q2 = (TM[2][0] * pv[0] + TM[2][1] * pv[1] + TM[2][2] * pv[2]) / Ts;
    const double a_ = pv[0], b_ = pv[1], c_ = pv[2];
    const double u1 = c_ * Q0 * Q0 - b_ * Q0 * Q1 + a_ * Q1 * Q1 + b_ * b_ * Q2;
    const double u2 = a_ * c_ * Q2;
    const double l1 = sqrt(b_ * b_ + (a_ - c_) * (a_ - c_));
    const double l2 = a_ + c_;
    const double l3 = b_ * b_ - 4.0 * a_ * c_;
    const double p1 = 2.0 * c_ * Q0 - b_ * Q1;
    const double p2 = 2.0 * a_ * Q1 - b_ * Q0;
    const double x0 = p1 / l3 / scale + cx;
    const double y0 = p2 / l3 / scale + cy;
    const double A = sqrt(2.0) * sqrt((u1 - 4.0 * u2) / ((l1 - l2) * l3)) / scale;
    const double B = sqrt(2.0) * sqrt(-1.0 * ((u1 - 4.0 * u2) / ((l1 + l2) * l3))) / scale;
    double theta;
    if (b_ == 0.0) theta = (a_ < c_) ? 0.0 : RMCV_PI / 2.0;
    else theta = RMCV_PI / 2.0 + 0.5 * atan2(b_, (a_ - c_));
    float wd = (float)(2.0 * A), ht = (float)(2.0 * B), ang;
    if (wd > ht) {
        const float tmp = wd; wd = ht; ht = tmp;
        ang = (float)fmod(90.0 + theta * 180.0 / RMCV_PI, 180.0);
    } else {
        ang = (float)fmod(theta * 180.0 / RMCV_PI, 180.0);
    }
    box->cx = (float)x0; box->cy = (float)y0; box->w = wd; box->h = ht; box->angle = ang;
    RMCV_FIT_MARK(5);
    return true;
}

#if defined(__CUDACC__)
// Warp-cooperative twin of direct_fit for the latency path (contour kernel with the fit on board, frame.cu): ALL 32 lanes call
// it with identical arguments and all return the identical result.  The independent pieces that are long fp64 dependency
// chains run on different lanes - the nine divisions of M', the three eigenvalues (cos, Newton polish) with their
// eigenvectors, the centre and axis pairs - and are exchanged with shuffles.  Every value is computed by the very expression
// direct_fit uses, so the two agree bit for bit (scripts/latency_mode_digest_gpu.py under pytest -m gpu compares them).
__device__ __forceinline__ bool direct_fit_warp(const Moments& m, double scale, double cx, double cy, rmcv_rotated_rect* box,
                                                double* det_out) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double n = m.n, s1 = scale, s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2;
    const double inv_n = 1.0 / n;
    const double X = m.x * s1 * inv_n, Y = m.y * s1 * inv_n;
    const double XX = m.xx * s2 * inv_n, XY = m.xy * s2 * inv_n, YY = m.yy * s2 * inv_n;
    const double XXX = m.xxx * s3 * inv_n, XXY = m.xxy * s3 * inv_n, XYY = m.xyy * s3 * inv_n, YYY = m.yyy * s3 * inv_n;
    const double XXXX = m.xxxx * s4 * inv_n, XXXY = m.xxxy * s4 * inv_n, XXYY = m.xxyy * s4 * inv_n,
                 XYYY = m.xyyy * s4 * inv_n, YYYY = m.yyyy * s4 * inv_n;
    const double S1[3][3] = {{XXXX, XXXY, XXYY}, {XXXY, XXYY, XYYY}, {XXYY, XYYY, YYYY}};
    const double S2[3][3] = {{XXX, XXY, XX}, {XXY, XYY, XY}, {XYY, YYY, YY}};
    const double S3[3][3] = {{XX, XY, X}, {XY, YY, Y}, {X, Y, 1.0}};
    const double Ts = det3(S3);
    double adj[3][3];
    adj[0][0] = S3[1][1] * S3[2][2] - S3[1][2] * S3[2][1];
    adj[0][1] = S3[0][2] * S3[2][1] - S3[0][1] * S3[2][2];
    adj[0][2] = S3[0][1] * S3[1][2] - S3[0][2] * S3[1][1];
    adj[1][0] = adj[0][1];
    adj[1][1] = S3[0][0] * S3[2][2] - S3[0][2] * S3[2][0];
    adj[1][2] = S3[0][2] * S3[1][0] - S3[0][0] * S3[1][2];
    adj[2][0] = adj[0][2];
    adj[2][1] = adj[1][2];
    adj[2][2] = S3[0][0] * S3[1][1] - S3[0][1] * S3[1][0];
    double TM[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) TM[i][j] = -(adj[i][0] * S2[j][0] + adj[i][1] * S2[j][1] + adj[i][2] * S2[j][2]);
    // M'[i][j] on lane 3 i + j
    auto pick = [](double a0, double a1, double a2, int k) -> double { return k == 0 ? a0 : (k == 1 ? a1 : a2); };
    double Mp[3][3];
    {
        const int e = lane < 9 ? lane : 0, ei = e / 3, ej = e - 3 * ei;
        const double a0 = pick(S2[0][0], S2[1][0], S2[2][0], ei), a1 = pick(S2[0][1], S2[1][1], S2[2][1], ei),
                     a2 = pick(S2[0][2], S2[1][2], S2[2][2], ei);
        const double b0 = pick(TM[0][0], TM[0][1], TM[0][2], ej), b1 = pick(TM[1][0], TM[1][1], TM[1][2], ej),
                     b2 = pick(TM[2][0], TM[2][1], TM[2][2], ej);
        const double s1e = pick(pick(S1[0][0], S1[0][1], S1[0][2], ej), pick(S1[1][0], S1[1][1], S1[1][2], ej),
                                pick(S1[2][0], S1[2][1], S1[2][2], ej), ei);
        const double mine = s1e + (a0 * b0 + a1 * b1 + a2 * b2) / Ts;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Mp[i][j] = __shfl_sync(full, mine, 3 * i + j);
    }
    double M[3][3];
    for (int j = 0; j < 3; ++j) {
        M[0][j] = Mp[2][j] / 2.0;
        M[1][j] = -Mp[1][j];
        M[2][j] = Mp[0][j] / 2.0;
    }
    const double det = fabs(det3(M));
    *det_out = det;
    if (!(det > 1.0e-10)) return false;

    // eigenvalue k, its Newton polish and its eigenvector on lane k (eig3_values / eig3_vector, one root per lane)
    int nl;
    double lam_k;
    {
        const double tr = M[0][0] + M[1][1] + M[2][2];
        const double c1 = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) + (M[0][0] * M[2][2] - M[0][2] * M[2][0]) +
                          (M[1][1] * M[2][2] - M[1][2] * M[2][1]);
        const double dt = det3(M);
        const double a = -tr, b = c1, c = -dt;
        const double p = b - a * a / 3.0;
        const double q = 2.0 * a * a * a / 27.0 - a * b / 3.0 + c;
        const double disc = q * q / 4.0 + p * p * p / 27.0;
        const int k = lane < 3 ? lane : 0;
        if (disc <= 0.0 && p < 0.0) {
            const double r = sqrt(-p / 3.0);
            double cosarg = -q / (2.0 * r * r * r);
            cosarg = cosarg > 1.0 ? 1.0 : (cosarg < -1.0 ? -1.0 : cosarg);
            const double phi = acos(cosarg);
            lam_k = 2.0 * r * cos((phi - 2.0 * RMCV_PI * k) / 3.0) - a / 3.0;
            nl = 3;
        } else {
            const double sq = sqrt(disc > 0.0 ? disc : 0.0);
            const double u = cbrt(-q / 2.0 + sq), v = cbrt(-q / 2.0 - sq);
            lam_k = u + v - a / 3.0;
            nl = 1;
        }
        double l = lam_k;
        for (int it = 0; it < 3; ++it) {
            const double f = ((l + a) * l + b) * l + c;
            const double fp = (3.0 * l + 2.0 * a) * l + b;
            if (fp == 0.0) break;
            const double step = f / fp;
            if (!(fabs(step) < fabs(l) * 1e-3 + 1e-300)) break;
            l -= step;
        }
        lam_k = l;
    }
    double pv[3] = {0, 0, 0};
    {
        double v[3];
        eig3_vector(M, lam_k, v);
        const double cond_k = 4.0 * v[0] * v[2] - v[1] * v[1];
        double best_cond = 0.0;
        for (int k = 0; k < nl; ++k) {
            const double w0 = __shfl_sync(full, v[0], k), w1 = __shfl_sync(full, v[1], k), w2 = __shfl_sync(full, v[2], k);
            const double cond = __shfl_sync(full, cond_k, k);
            if (k == 0 || cond > best_cond) { best_cond = cond; pv[0] = w0; pv[1] = w1; pv[2] = w2; }
        }
    }
    double norm = sqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    const int sg = (pv[0] < 0.0 ? -1 : 1) * (pv[1] < 0.0 ? -1 : 1) * (pv[2] < 0.0 ? -1 : 1);
    if (sg <= 0) norm = -norm;
    pv[0] /= norm; pv[1] /= norm; pv[2] /= norm;
    const double Q0 = (TM[0][0] * pv[0] + TM[0][1] * pv[1] + TM[0][2] * pv[2]) / Ts;
    const double Q1 = (TM[1][0] * pv[0] + TM[1][1] * pv[1] + TM[1][2] * pv[2]) / Ts;
    const double Q2 = (TM[2][0] * pv[0] + TM[2][1] * pv[1] + TM[2][2] * pv[2]) / Ts;
    const double a_ = pv[0], b_ = pv[1], c_ = pv[2];
    const double u1 = c_ * Q0 * Q0 - b_ * Q0 * Q1 + a_ * Q1 * Q1 + b_ * b_ * Q2;
    const double u2 = a_ * c_ * Q2;
    const double l1 = sqrt(b_ * b_ + (a_ - c_) * (a_ - c_));
    const double l2 = a_ + c_;
    const double l3 = b_ * b_ - 4.0 * a_ * c_;
    const double p1 = 2.0 * c_ * Q0 - b_ * Q1;
    const double p2 = 2.0 * a_ * Q1 - b_ * Q0;
    // centre: x0 on even lanes, y0 on odd lanes (the same chain of two divisions on different operands)
    const bool odd = (lane & 1) != 0;
    const double ctr = (odd ? p2 : p1) / l3 / scale + (odd ? cy : cx);
    const double x0 = __shfl_sync(full, ctr, 0), y0 = __shfl_sync(full, ctr, 1);
    // semi-axes: A on even lanes, B on odd lanes.  B = sqrt(2) * sqrt(-1.0 * (num / ((l1 + l2) * l3))) / scale
    const double num = u1 - 4.0 * u2;
    const double quo = num / ((odd ? (l1 + l2) : (l1 - l2)) * l3);
    const double ax = sqrt(2.0) * sqrt(odd ? -1.0 * quo : quo) / scale;
    const double A = __shfl_sync(full, ax, 0), B = __shfl_sync(full, ax, 1);
    double theta;
    if (b_ == 0.0) theta = (a_ < c_) ? 0.0 : RMCV_PI / 2.0;
    else theta = RMCV_PI / 2.0 + 0.5 * atan2(b_, (a_ - c_));
    float wd = (float)(2.0 * A), ht = (float)(2.0 * B), ang;
    if (wd > ht) {
        const float tmp = wd; wd = ht; ht = tmp;
        ang = (float)fmod(90.0 + theta * 180.0 / RMCV_PI, 180.0);
    } else {
        ang = (float)fmod(theta * 180.0 / RMCV_PI, 180.0);
    }
    box->cx = (float)x0; box->cy = (float)y0; box->w = wd; box->h = ht; box->angle = ang;
    return true;
}
#endif  // __CUDACC__

// Solve A x = b (n <= 5) in place by Gaussian elimination with partial pivoting.
template <int N>
RMCV_HD void solve_n(double A[N][N], double b[N], double x[N]) {
    for (int c = 0; c < N; ++c) {
        int piv = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < N; ++r)
            if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
        if (piv != c) {
            for (int k = 0; k < N; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            const double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const double inv = 1.0 / A[c][c];
        for (int r = c + 1; r < N; ++r) {
            const double f = A[r][c] * inv;
            for (int k = c; k < N; ++k) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    for (int r = N - 1; r >= 0; --r) {
        double s = b[r];
        for (int k = r + 1; k < N; ++k) s -= A[r][k] * x[k];
        x[r] = s / A[r][r];
    }
}

RMCV_HD double ipow(double v, int e) {
    double r = 1.0;
    for (int i = 0; i < e; ++i) r *= v;
    return r;
}

// cv::fitEllipseNoDirect from the moment sums of the float-centred points (normal equations).
RMCV_HD void nodirect_fit(const Moments& m, double scale, float c32x, float c32y, rmcv_rotated_rect* box) {
    const double n = m.n, s1 = scale, s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2;
    // scaled sums S[i][j] = sum x^i y^j
    double S[5][5];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) S[i][j] = 0.0;
    S[0][0] = n;
    S[1][0] = m.x * s1; S[0][1] = m.y * s1;
    S[2][0] = m.xx * s2; S[1][1] = m.xy * s2; S[0][2] = m.yy * s2;
    S[3][0] = m.xxx * s3; S[2][1] = m.xxy * s3; S[1][2] = m.xyy * s3; S[0][3] = m.yyy * s3;
    S[4][0] = m.xxxx * s4; S[3][1] = m.xxxy * s4; S[2][2] = m.xxyy * s4; S[1][3] = m.xyyy * s4; S[0][4] = m.yyyy * s4;
    // rows [-x^2, -y^2, -xy, x, y], rhs 10000
    double AtA[5][5] = {
        {S[4][0], S[2][2], S[3][1], -S[3][0], -S[2][1]},
        {S[2][2], S[0][4], S[1][3], -S[1][2], -S[0][3]},
        {S[3][1], S[1][3], S[2][2], -S[2][1], -S[1][2]},
        {-S[3][0], -S[1][2], -S[2][1], S[2][0], S[1][1]},
        {-S[2][1], -S[0][3], -S[1][2], S[1][1], S[0][2]},
    };
    double Atb[5] = {-10000.0 * S[2][0], -10000.0 * S[0][2], -10000.0 * S[1][1], 10000.0 * S[1][0], 10000.0 * S[0][1]};
    double g[5];
    solve_n<5>(AtA, Atb, g);
    double A2[2][2] = {{2.0 * g[0], g[2]}, {g[2], 2.0 * g[1]}};
    double b2[2] = {g[3], g[4]};
    double r[2];
    solve_n<2>(A2, b2, r);
    const double rx = r[0], ry = r[1];
    // shifted sums  sum (x-rx)^i (y-ry)^j  by binomial expansion
    const double binom[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};
    auto sh = [&](int i, int j) {
        double tot = 0.0;
        for (int a = 0; a <= i; ++a)
            for (int b = 0; b <= j; ++b) tot += binom[i][a] * binom[j][b] * ipow(-rx, i - a) * ipow(-ry, j - b) * S[a][b];
        return tot;
    };
    const double s40 = sh(4, 0), s22 = sh(2, 2), s31 = sh(3, 1), s04 = sh(0, 4), s13 = sh(1, 3);
    double BtB[3][3] = {{s40, s22, s31}, {s22, s04, s13}, {s31, s13, s22}};
    double Btb[3] = {sh(2, 0), sh(0, 2), sh(1, 1)};
    double g2[3];
    solve_n<3>(BtB, Btb, g2);
    const double min_eps = 1e-8;
    const double rp4 = -0.5 * atan2(g2[2], g2[1] - g2[0]);
    double t;
    if (fabs(g2[2]) > min_eps) t = g2[2] / sin(-2.0 * rp4);
    else t = g2[1] - g2[0];
    double rp2 = fabs(g2[0] + g2[1] - t);
    if (rp2 > min_eps) rp2 = sqrt(2.0 / rp2);
    double rp3 = fabs(g2[0] + g2[1] + t);
    if (rp3 > min_eps) rp3 = sqrt(2.0 / rp3);
    box->cx = fadd((float)(rx / scale), c32x);
    box->cy = fadd((float)(ry / scale), c32y);
    float wd = (float)(rp2 * 2.0 / scale), ht = (float)(rp3 * 2.0 / scale);
    // cv::fitEllipseNoDirect assigns box.angle only when it swaps the axes, so the angle of an unswapped box stays at
    // cv::RotatedRect's default 0.  With |g2[2]| > min_eps t = hypot(g2[2], g2[1] - g2[0]) > 0 and the swap always happens;
    // the unswapped case is the axis-aligned ellipse taller than wide (t = g2[1] - g2[0] < 0), e.g. an exactly
    // mirror-symmetric upright light bar.
    float ang = 0.f;
    if (wd > ht) {
        const float tmp = wd; wd = ht; ht = tmp;
        ang = (float)(90.0 + rp4 * 180.0 / RMCV_PI);
    }
    if (ang < -180.f) ang = fadd(ang, 360.f);
    if (ang > 360.f) ang = fsub(ang, 360.f);
    box->w = wd; box->h = ht; box->angle = ang;
}

// ---------------------------------------------------------------------------------------------- lightblob
// cv::RotatedRect::points
RMCV_HD void rotated_rect_points(const rmcv_rotated_rect& r, float pt[4][2]) {
    const double ang = (double)r.angle * RMCV_PI / 180.0;
    const float b = fmul((float)cos(ang), 0.5f);
    const float a = fmul((float)sin(ang), 0.5f);
    pt[0][0] = fsub(fsub(r.cx, fmul(a, r.h)), fmul(b, r.w));
    pt[0][1] = fsub(fadd(r.cy, fmul(b, r.h)), fmul(a, r.w));
    pt[1][0] = fsub(fadd(r.cx, fmul(a, r.h)), fmul(b, r.w));
    pt[1][1] = fsub(fsub(r.cy, fmul(b, r.h)), fmul(a, r.w));
    pt[2][0] = fsub(fmul(2.f, r.cx), pt[0][0]);
    pt[2][1] = fsub(fmul(2.f, r.cy), pt[0][1]);
    pt[3][0] = fsub(fmul(2.f, r.cx), pt[1][0]);
    pt[3][1] = fsub(fmul(2.f, r.cy), pt[1][1]);
}

// rm::lightblob::lightblob + rm::utils::reorder_vertices (src/core.cpp:9-19, 265-283)
RMCV_HD void make_lightblob(const rmcv_rotated_rect& box, int target, rmcv_lightblob* out) {
    out->angle = box.angle > 90.f ? fsub(box.angle, 90.f) : fadd(box.angle, 90.f);
    out->target = target;
    out->center[0] = box.cx;
    out->center[1] = box.cy;
    float t[4][2];
    rotated_rect_points(box, t);
    for (int i = 1; i < 4; ++i) {  // std::sort on 4 elements == insertion sort (stable) by y
        const float vx = t[i][0], vy = t[i][1];
        int j = i - 1;
        while (j >= 0 && vy < t[j][1]) { t[j + 1][0] = t[j][0]; t[j + 1][1] = t[j][1]; --j; }
        t[j + 1][0] = vx; t[j + 1][1] = vy;
    }
    const bool swap_up = t[0][0] < t[1][0], swap_down = t[2][0] < t[3][0];
    const int i0 = swap_down ? 2 : 3, i1 = swap_up ? 0 : 1, i2 = swap_up ? 1 : 0, i3 = swap_down ? 3 : 2;
    out->vertices[0][0] = t[i0][0]; out->vertices[0][1] = t[i0][1];
    out->vertices[1][0] = t[i1][0]; out->vertices[1][1] = t[i1][1];
    out->vertices[2][0] = t[i2][0]; out->vertices[2][1] = t[i2][1];
    out->vertices[3][0] = t[i3][0]; out->vertices[3][1] = t[i3][1];
    out->size[0] = fminf(box.h, box.w);
    out->size[1] = fmaxf(box.h, box.w);
}

// Loop body of rm::filter_lightblobs after the fit (src/objdetect.cpp:71-80).  Returns RMCV_CONTOUR_*.
RMCV_HD int blob_gates(const rmcv_rotated_rect& e, const rmcv_params& prm) {
    bool negative = false;
    const float ratio = fdiv(fmaxf(e.w, e.h), fminf(e.w, e.h));
    if (!(ratio >= prm.ratio_min && ratio <= prm.ratio_max)) negative = true;
    const float angle = e.angle > 90.f ? fsub(e.angle, 90.f) : fadd(e.angle, 90.f);
    if (fabsf(fsub(angle, 90.f)) > prm.tilt_max) negative = true;
    return negative ? RMCV_CONTOUR_NEGATIVE : RMCV_CONTOUR_POSITIVE;
}

// ---------------------------------------------------------------------------------------------- armour
// rm::utils::PointDistance(Point2f) (src/core.cpp:285-288)
RMCV_HD float point_distance(const float p1[2], const float p2[2]) {
    const double dx = (double)fsub(p1[0], p2[0]), dy = (double)fsub(p1[1], p2[1]);
    return (float)sqrt(dadd(dmul(dx, dx), dmul(dy, dy)));
}

// rm::utils::ExtendCord (src/core.cpp:295-380)
RMCV_HD void extend_cord(const float pt1[2], const float pt2[2], float d, float dst1[2], float dst2[2]) {
    if (pt1[0] == pt2[0]) {
        dst1[0] = pt1[0]; dst2[0] = pt1[0];
        if (pt1[1] > pt2[1]) { dst1[1] = fadd(pt1[1], d); dst2[1] = fsub(pt2[1], d); }
        else { dst1[1] = fsub(pt1[1], d); dst2[1] = fadd(pt2[1], d); }
    } else if (pt1[1] == pt2[1]) {
        dst1[1] = pt1[1]; dst2[1] = pt1[1];
        if (pt1[0] > pt2[0]) { dst1[0] = fadd(pt1[0], d); dst2[0] = fsub(pt2[0], d); }
        else { dst1[0] = fsub(pt1[0], d); dst2[0] = fadd(pt2[0], d); }
    } else {
        const float k = fdiv(fsub(pt1[1], pt2[1]), fsub(pt1[0], pt2[0]));
        const float theta = (float)atan2((double)fabsf(fsub(pt1[1], pt2[1])), (double)fabsf(fsub(pt1[0], pt2[0])));
        const float zoomY = (float)dmul(sin((double)theta), (double)d);
        const float zoomX = (float)dmul(cos((double)theta), (double)d);
        if (k > 0.f) {
            if (pt1[0] > pt2[0]) {
                dst1[0] = fadd(pt1[0], zoomX); dst1[1] = fadd(pt1[1], zoomY);
                dst2[0] = fsub(pt2[0], zoomX); dst2[1] = fsub(pt2[1], zoomY);
            } else {
                dst1[0] = fsub(pt1[0], zoomX); dst1[1] = fsub(pt1[1], zoomY);
                dst2[0] = fadd(pt2[0], zoomX); dst2[1] = fadd(pt2[1], zoomY);
            }
        } else {
            if (pt1[0] < pt2[0]) {
                dst1[0] = fsub(pt1[0], zoomX); dst1[1] = fadd(pt1[1], zoomY);
                dst2[0] = fadd(pt2[0], zoomX); dst2[1] = fsub(pt2[1], zoomY);
            } else {
                dst1[0] = fadd(pt1[0], zoomX); dst1[1] = fsub(pt1[1], zoomY);
                dst2[0] = fsub(pt2[0], zoomX); dst2[1] = fadd(pt2[1], zoomY);
            }
        }
    }
}

RMCV_HD float c_round_f(float v) { return (float)round((double)v); }

// rm::armour::armour geometry (src/core.cpp:21-49)
RMCV_HD void make_armour(const rmcv_lightblob& b0, const rmcv_lightblob& b1, rmcv_armour* out) {
    const rmcv_lightblob* L = &b0;
    const rmcv_lightblob* R = &b1;
    if (b1.center[0] < b0.center[0]) { L = &b1; R = &b0; }
    float v[4][2];
    v[0][0] = L->vertices[3][0]; v[0][1] = L->vertices[3][1];
    v[1][0] = L->vertices[2][0]; v[1][1] = L->vertices[2][1];
    v[2][0] = R->vertices[1][0]; v[2][1] = R->vertices[1][1];
    v[3][0] = R->vertices[0][0]; v[3][1] = R->vertices[0][1];
    const float dl = point_distance(v[0], v[1]);
    const float dr = point_distance(v[2], v[3]);
    const float offl = c_round_f(fdiv(fsub(fdiv(dl, 0.5f), dl), 2.f));
    const float offr = c_round_f(fdiv(fsub(fdiv(dr, 0.5f), dr), 2.f));
    extend_cord(v[0], v[1], offl, out->icon[0], out->icon[1]);
    extend_cord(v[3], v[2], offr, out->icon[3], out->icon[2]);
    // cv::boundingRect(vector<Point2f>)
    float minx = out->icon[0][0], maxx = minx, miny = out->icon[0][1], maxy = miny;
    for (int i = 1; i < 4; ++i) {
        minx = fminf(minx, out->icon[i][0]); maxx = fmaxf(maxx, out->icon[i][0]);
        miny = fminf(miny, out->icon[i][1]); maxy = fmaxf(maxy, out->icon[i][1]);
    }
    const int ix0 = (int)floorf(minx), iy0 = (int)floorf(miny), ix1 = (int)floorf(maxx), iy1 = (int)floorf(maxy);
    out->bounding_box[0] = (float)ix0; out->bounding_box[1] = (float)iy0;
    out->bounding_box[2] = (float)(ix1 - ix0 + 1); out->bounding_box[3] = (float)(iy1 - iy0 + 1);
    // rm::utils::CalcPerspective(vertices, vertices, 1.0f) (src/core.cpp:382-399)
    const float maxh = (float)fmax((double)dl, (double)dr);
    const float sw = fmul(maxh, 1.0f), shh = maxh;
    float c01[2] = {fadd(fdiv(v[0][0], 2.f), fdiv(v[1][0], 2.f)), fadd(fdiv(v[0][1], 2.f), fdiv(v[1][1], 2.f))};
    float c23[2] = {fadd(fdiv(v[2][0], 2.f), fdiv(v[3][0], 2.f)), fadd(fdiv(v[2][1], 2.f), fdiv(v[3][1], 2.f))};
    const float cxx = fadd(fdiv(c01[0], 2.f), fdiv(c23[0], 2.f)), cyy = fadd(fdiv(c01[1], 2.f), fdiv(c23[1], 2.f));
    out->vertices[0][0] = fsub(cxx, fdiv(sw, 2.f)); out->vertices[0][1] = fsub(cyy, fdiv(shh, 2.f));
    out->vertices[1][0] = fsub(cxx, fdiv(sw, 2.f)); out->vertices[1][1] = fadd(cyy, fdiv(shh, 2.f));
    out->vertices[2][0] = fadd(cxx, fdiv(sw, 2.f)); out->vertices[2][1] = fadd(cyy, fdiv(shh, 2.f));
    out->vertices[3][0] = fadd(cxx, fdiv(sw, 2.f)); out->vertices[3][1] = fsub(cyy, fdiv(shh, 2.f));
}

// Gates of rm::filter_armours for one pair (src/objdetect.cpp:124-159).  gates = {|dangle|, shear_i, shear_j,
// min/max height, |dcy|, |dcx|}.  Returns true when the pair becomes an armour.
RMCV_HD bool pair_gates(const rmcv_lightblob& bi, const rmcv_lightblob& bj, const rmcv_params& prm, float gates[6]) {
    const float ad = fabsf(fsub(bi.angle, bj.angle));
    const float y = fabsf(fsub(bi.center[1], bj.center[1]));
    const float x = fabsf(fsub(bi.center[0], bj.center[0]));
    const float pif = 3.14159274101257324f;  // static_cast<float>(CV_PI)
    const float rect_angle = (float)(atan2((double)y, (double)x) * 180.0 / (double)pif);
    const float si = fabsf(bi.angle > 90.f ? fsub(fabsf(fsub(bi.angle, rect_angle)), 90.f)
                                           : fsub(fabsf(fsub(fsub(180.f, bi.angle), rect_angle)), 90.f));
    const float sj = fabsf(bj.angle > 90.f ? fsub(fabsf(fsub(bj.angle, rect_angle)), 90.f)
                                           : fsub(fabsf(fsub(fsub(180.f, bj.angle), rect_angle)), 90.f));
    const float hi = bi.size[1], hj = bj.size[1];
    const float ratio = fdiv(fminf(hi, hj), fmaxf(hi, hj));
    const float hsum = fadd(hi, hj);
    gates[0] = ad; gates[1] = si; gates[2] = sj; gates[3] = ratio; gates[4] = y; gates[5] = x;
    if (bi.target != prm.target || bj.target != prm.target) return false;
    if (ad > prm.angle_difference_max) return false;
    if (si > prm.shear_max || sj > prm.shear_max) return false;
    if (ratio < prm.lenght_ratio_max) return false;
    if (y > fdiv(hsum, 2.f)) return false;
    if (x > fmul(hsum, 2.f)) return false;
    return true;
}

// The verdict of pair_gates alone, cheap comparisons first: the double-precision atan2 behind the shear gate is only
// evaluated for pairs that pass every other gate (the verdict is a conjunction of the same comparisons, so the order in
// which they are tried does not change it).  The O(P^2) loops call this and build the gate values for the few survivors.
RMCV_HD bool pair_passes(const rmcv_lightblob& bi, const rmcv_lightblob& bj, const rmcv_params& prm) {
    if (bi.target != prm.target || bj.target != prm.target) return false;
    if (fabsf(fsub(bi.angle, bj.angle)) > prm.angle_difference_max) return false;
    const float hi = bi.size[1], hj = bj.size[1];
    const float hsum = fadd(hi, hj);
    const float y = fabsf(fsub(bi.center[1], bj.center[1]));
    if (y > fmul(hsum, 0.5f)) return false;                      // hsum / 2.f: halving is exact
    const float x = fabsf(fsub(bi.center[0], bj.center[0]));
    if (x > fmul(hsum, 2.f)) return false;
    if (fdiv(fminf(hi, hj), fmaxf(hi, hj)) < prm.lenght_ratio_max) return false;
    const float pif = 3.14159274101257324f;
    const float rect_angle = (float)(atan2((double)y, (double)x) * 180.0 / (double)pif);
    const float si = fabsf(bi.angle > 90.f ? fsub(fabsf(fsub(bi.angle, rect_angle)), 90.f)
                                           : fsub(fabsf(fsub(fsub(180.f, bi.angle), rect_angle)), 90.f));
    const float sj = fabsf(bj.angle > 90.f ? fsub(fabsf(fsub(bj.angle, rect_angle)), 90.f)
                                           : fsub(fabsf(fsub(fsub(180.f, bj.angle), rect_angle)), 90.f));
    return !(si > prm.shear_max || sj > prm.shear_max);
}

// ---------------------------------------------------------------------------------------------- integer contour sums
// Everything cv::contourArea / cv::fitEllipseDirect need from a contour is a sum over its point multiset.  The kernels
// accumulate them as EXACT integers (order-free, deterministic, atomics-friendly); the fit converts them to the centred
// double sums the reference accumulates point by point.
struct ContourSums {
    long long n;                  // contour.size()
    long long sx, sy;             // sum x, sum y (absolute pixel coordinates)
    long long cross;              // shoelace sum: sum over contour edges p->q of x_p*y_q - x_q*y_p  (= +-2*contourArea)
    long long xx, xy, yy;         // sums of dx^i dy^j with dx = x - ox, dy = y - oy
    long long xxx, xxy, xyy, yyy;
    long long xxxx, xxxy, xxyy, xyyy, yyyy;
    long long s_int;              // sum |n*x - sx| + |n*y - sy|  (= n * sum(|x-cx|+|y-cy|), exact)
    int ox, oy;                   // origin of the relative coordinates
};

RMCV_HD void sums_zero(ContourSums& c, int ox, int oy) {
    c.n = c.sx = c.sy = c.cross = 0;
    c.xx = c.xy = c.yy = c.xxx = c.xxy = c.xyy = c.yyy = 0;
    c.xxxx = c.xxxy = c.xxyy = c.xyyy = c.yyyy = 0;
    c.s_int = 0;
    c.ox = ox; c.oy = oy;
}

RMCV_HD void sums_add_point(ContourSums& c, int x, int y) {
    const long long dx = x - c.ox, dy = y - c.oy;
    const long long xx = dx * dx, xy = dx * dy, yy = dy * dy;
    c.n += 1; c.sx += x; c.sy += y;
    c.xx += xx; c.xy += xy; c.yy += yy;
    c.xxx += xx * dx; c.xxy += xx * dy; c.xyy += dx * yy; c.yyy += yy * dy;
    c.xxxx += xx * xx; c.xxxy += xx * xy; c.xxyy += xx * yy; c.xyyy += xy * yy; c.yyyy += yy * yy;
}

// Moment sums about (O + (a, b)) from the sums R about O: binomial shift in double.
RMCV_HD void shift_moments(const Moments& R, double a, double b, Moments* m) {
    const double n = R.n;
    const double R10 = R.x, R01 = R.y, R20 = R.xx, R11 = R.xy, R02 = R.yy;
    const double R30 = R.xxx, R21 = R.xxy, R12 = R.xyy, R03 = R.yyy;
    const double R40 = R.xxxx, R31 = R.xxxy, R22 = R.xxyy, R13 = R.xyyy, R04 = R.yyyy;
    const double a2 = a * a, b2 = b * b, a3 = a2 * a, b3 = b2 * b, a4 = a2 * a2, b4 = b2 * b2;
    m->n = n;
    m->x = R10 - n * a;
    m->y = R01 - n * b;
    m->xx = R20 - 2.0 * a * R10 + n * a2;
    m->yy = R02 - 2.0 * b * R01 + n * b2;
    m->xy = R11 - a * R01 - b * R10 + n * a * b;
    m->xxx = R30 - 3.0 * a * R20 + 3.0 * a2 * R10 - n * a3;
    m->yyy = R03 - 3.0 * b * R02 + 3.0 * b2 * R01 - n * b3;
    m->xxy = R21 - b * R20 - 2.0 * a * R11 + 2.0 * a * b * R10 + a2 * R01 - n * a2 * b;
    m->xyy = R12 - a * R02 - 2.0 * b * R11 + 2.0 * a * b * R01 + b2 * R10 - n * a * b2;
    m->xxxx = R40 - 4.0 * a * R30 + 6.0 * a2 * R20 - 4.0 * a3 * R10 + n * a4;
    m->yyyy = R04 - 4.0 * b * R03 + 6.0 * b2 * R02 - 4.0 * b3 * R01 + n * b4;
    m->xxxy = R31 - b * R30 - 3.0 * a * R21 + 3.0 * a * b * R20 + 3.0 * a2 * R11 - 3.0 * a2 * b * R10 - a3 * R01 + n * a3 * b;
    m->xyyy = R13 - a * R03 - 3.0 * b * R12 + 3.0 * a * b * R02 + 3.0 * b2 * R11 - 3.0 * a * b2 * R01 - b3 * R10 + n * a * b3;
    m->xxyy = R22 - 2.0 * a * R12 + a2 * R02 - 2.0 * b * R21 + 4.0 * a * b * R11 - 2.0 * a2 * b * R01 + b2 * R20 - 2.0 * a * b2 * R10 +
              n * a2 * b2;
}

// The exact integer sums about (ox, oy) as doubles (exact below 2^53, correctly rounded above).
RMCV_HD void sums_to_moments(const ContourSums& c, Moments* R) {
    R->n = (double)c.n;
    R->x = (double)(c.sx - c.n * c.ox); R->y = (double)(c.sy - c.n * c.oy);
    R->xx = (double)c.xx; R->xy = (double)c.xy; R->yy = (double)c.yy;
    R->xxx = (double)c.xxx; R->xxy = (double)c.xxy; R->xyy = (double)c.xyy; R->yyy = (double)c.yyy;
    R->xxxx = (double)c.xxxx; R->xxxy = (double)c.xxxy; R->xxyy = (double)c.xxyy; R->xyyy = (double)c.xyyy; R->yyyy = (double)c.yyyy;
}

// Does this contour reach the fit at all?  (src/objdetect.cpp:64)
RMCV_HD bool contour_is_fitted(long long n, long long cross, const rmcv_params& prm) {
    const long long area2 = cross < 0 ? -cross : cross;
    const double area = (double)area2 * 0.5;
    return n >= 6 && area >= prm.area_min && area <= prm.area_max;
}

// Loop body of rm::filter_lightblobs (src/objdetect.cpp:62-84) once the sums of one contour are known:
//   n, sum_x, sum_y, cross  exact integers in absolute pixel coordinates,
//   R                       moment sums about the origin (Ox, Oy),
//   s                       L1 spread  sum |x-cx| + |y-cy|  about the double mean.
// kWarp (device only): all 32 lanes of a warp call it with identical arguments and the direct fit is the warp-cooperative
// twin; every lane returns the same result.
template <bool kWarp>
RMCV_HD void fit_from_moments_t(long long n_i, long long sum_x, long long sum_y, long long cross, const Moments& R, double Ox,
                                double Oy, double s, const rmcv_params& prm, int* status, int* branch, float* det0_out,
                                rmcv_rotated_rect* ell, rmcv_lightblob* blob) {
    *status = RMCV_CONTOUR_SKIPPED;
    *branch = RMCV_FIT_NONE;
    *det0_out = 0.f;
    ell->cx = ell->cy = ell->w = ell->h = ell->angle = 0.f;
    if (!contour_is_fitted(n_i, cross, prm)) return;
    const double n = (double)n_i;
    // ---- cv::fitEllipseDirect, first attempt: centre and L1 spread in double
    const double cx = (double)sum_x / n, cy = (double)sum_y / n;
    double scale = 100.0 / (s > RMCV_FLT_EPSILON ? s : RMCV_FLT_EPSILON);
    Moments m;
    shift_moments(R, cx - Ox, cy - Oy, &m);
    RMCV_FIT_MARK(1);
    double det = 0.0;
    bool ok;
#if defined(__CUDA_ARCH__)
    if (kWarp) ok = direct_fit_warp(m, scale, cx, cy, ell, &det);
    else
#endif
        ok = direct_fit(m, scale, cx, cy, ell, &det);
    *det0_out = (float)det;
    if (ok) {
        *branch = RMCV_FIT_DIRECT;
    } else {
        // ---- singular: the reference retries with RNG jitter and then returns cv::fitEllipseNoDirect, which keeps the
        // centre as Point2f.  x - c32 is exact in fp32 for every blob whose extent is below its centroid's binade, so
        // the float-centred sums are the same exact sums shifted to c32 (DESIGN.md "fallback centring").
        const float c32x = fdiv((float)sum_x, (float)n_i), c32y = fdiv((float)sum_y, (float)n_i);
        shift_moments(R, (double)c32x - Ox, (double)c32y - Oy, &m);
        nodirect_fit(m, scale, c32x, c32y, ell);
        // beyond 2^24 OpenCV's own float accumulation of the centre rounds point by point (order dependent): flagged
        *branch = (sum_x >= (1LL << 24) || sum_y >= (1LL << 24)) ? RMCV_FIT_FALLBACK_LONG : RMCV_FIT_FALLBACK;
    }
    *status = blob_gates(*ell, prm);
    RMCV_FIT_MARK(6);
    if (*status == RMCV_CONTOUR_POSITIVE) make_lightblob(*ell, prm.target, blob);
    RMCV_FIT_MARK(7);
}

RMCV_HD void fit_from_moments(long long n_i, long long sum_x, long long sum_y, long long cross, const Moments& R, double Ox,
                              double Oy, double s, const rmcv_params& prm, int* status, int* branch, float* det0_out,
                              rmcv_rotated_rect* ell, rmcv_lightblob* blob) {
    fit_from_moments_t<false>(n_i, sum_x, sum_y, cross, R, Ox, Oy, s, prm, status, branch, det0_out, ell, blob);
}

// Same, from the exact integer sums.
template <bool kWarp>
RMCV_HD void fit_contour_t(const ContourSums& c, const rmcv_params& prm, int* status, int* branch, float* det0_out,
                           rmcv_rotated_rect* ell, rmcv_lightblob* blob) {
    RMCV_FIT_MARK(0);
    Moments R;
    sums_to_moments(c, &R);
    const double s = c.n > 0 ? (double)c.s_int / (double)c.n : 0.0;
    fit_from_moments_t<kWarp>(c.n, c.sx, c.sy, c.cross, R, (double)c.ox, (double)c.oy, s, prm, status, branch, det0_out, ell, blob);
}
RMCV_HD void fit_contour(const ContourSums& c, const rmcv_params& prm, int* status, int* branch, float* det0_out,
                         rmcv_rotated_rect* ell, rmcv_lightblob* blob) {
    fit_contour_t<false>(c, prm, status, branch, det0_out, ell, blob);
}

}  // namespace rmcv
