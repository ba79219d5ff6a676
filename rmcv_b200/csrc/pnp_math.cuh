// rm::solve_PnP (reference: src/mobility.cpp:166-190) = cv::solvePnP(..., cv::SOLVEPNP_IPPE_SQUARE) on the four
// armour vertices, written once as host+device inline functions (the CUDA kernel in pnp.cu calls them; tests/hostmath
// compiles the same header for the CPU tests).  OpenCV's path, restated:
//   cv::undistortPoints          pixel -> normalised coordinates, 5 fixed-point iterations of the k1,k2,p1,p2,k3 model,
//                                result stored as float32 (the input is vector<Point2f>)
//   IPPE::PoseSolver::solveSquare  homography of the canonical square -> Jacobian at the origin -> the two rotations of
//                                Collins & Bartoli's "Infinitesimal Plane-based Pose Estimation" (2014) -> least-squares
//                                translation for each -> the pose with the smaller reprojection error first
//   cv::Rodrigues                rotation matrix -> rotation vector
// Everything in double; a numpy prototype of exactly these steps matched cv2.solvePnP to 1e-12 on random quadrilaterals.
#pragma once
#include "blob_math.cuh"

namespace rmcv {

// cv::undistortPoints for one point (no R, no P): returns normalised (x, y) rounded to float32 like OpenCV's output.
RMCV_HD void undistort_point(double u, double v, const double K[9], const double dist[5], double* xo, double* yo) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
    double x = (u - cx) / fx, y = (v - cy) / fy;
    const double x0 = x, y0 = y;
    for (int it = 0; it < 5; ++it) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
        const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    *xo = (double)(float)x;
    *yo = (double)(float)y;
}

RMCV_HD void mat3_mul(const double A[3][3], const double B[3][3], double C[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}

// cv::Rodrigues, matrix -> vector (OpenCV first re-orthonormalises R by SVD; R is orthonormal to rounding here).
RMCV_HD void rodrigues_to_vec(const double R[3][3], double r[3]) {
    double rx = R[2][1] - R[1][2], ry = R[0][2] - R[2][0], rz = R[1][0] - R[0][1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0][0] + R[1][1] + R[2][2] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
        double t = (R[0][0] + 1.0) * 0.5;
        rx = sqrt(t > 0 ? t : 0.0);
        t = (R[1][1] + 1.0) * 0.5;
        ry = sqrt(t > 0 ? t : 0.0) * (R[0][1] < 0 ? -1.0 : 1.0);
        t = (R[2][2] + 1.0) * 0.5;
        rz = sqrt(t > 0 ? t : 0.0) * (R[0][2] < 0 ? -1.0 : 1.0);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[1][2] > 0) != (ry * rz > 0))) rz = -rz;
        theta /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
    } else {
        const double vth = theta / (2.0 * s);
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

// Least-squares translation for a rotation (IPPE computeTranslation): object points (X, Y, 0), normalised image points.
RMCV_HD void ippe_translation(const double obj[4][2], const double img[4][2], const double R[3][3], double t[3]) {
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        const double X = obj[i][0], Y = obj[i][1], u = img[i][0], v = img[i][1];
        const double px = R[0][0] * X + R[0][1] * Y, py = R[1][0] * X + R[1][1] * Y, pz = R[2][0] * X + R[2][1] * Y;
        const double rows[2][3] = {{1.0, 0.0, -u}, {0.0, 1.0, -v}};
        const double rhs[2] = {u * pz - px, v * pz - py};
        for (int e = 0; e < 2; ++e)
            for (int a = 0; a < 3; ++a) {
                b[a] += rows[e][a] * rhs[e];
                for (int c = 0; c < 3; ++c) A[a][c] += rows[e][a] * rows[e][c];
            }
    }
    solve_n<3>(A, b, t);
}

RMCV_HD double ippe_reproj_err(const double obj[4][2], const double img[4][2], const double R[3][3], const double t[3]) {
    double e = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double X = obj[i][0], Y = obj[i][1];
        const double px = R[0][0] * X + R[0][1] * Y + t[0], py = R[1][0] * X + R[1][1] * Y + t[1],
                     pz = R[2][0] * X + R[2][1] * Y + t[2];
        const double du = px / pz - img[i][0], dv = py / pz - img[i][1];
        e += du * du + dv * dv;
    }
    return e;
}

struct PnpResult {
    double rvec[3], tvec[3], err;   // err = sum of squared reprojection errors in normalised coordinates
};

// rm::solve_PnP for one armour: points_image = armour.vertices (src/core.cpp:48 order), exact size in world units,
// ROI offset added to the image points (src/mobility.cpp:172,182-185).  Returns false for collinear image points.
RMCV_HD bool solve_pnp_square(const float points_image[4][2], const double K[9], const double dist[5], float exact_w,
                              float exact_h, float roi_x, float roi_y, PnpResult* out) {
    // object points (z = 0) and the image points in the reference's order: 1, 2, 3, 0
    const double hw = (double)fdiv(exact_w, 2.0f), hh = (double)fdiv(exact_h, 2.0f);
    const double obj[4][2] = {{-hw, hh}, {hw, hh}, {hw, -hh}, {-hw, -hh}};
    const int order[4] = {1, 2, 3, 0};
    double img[4][2];
    for (int i = 0; i < 4; ++i) {
        const float u = fadd(points_image[order[i]][0], roi_x), v = fadd(points_image[order[i]][1], roi_y);
        undistort_point((double)u, (double)v, K, dist, &img[i][0], &img[i][1]);
    }
    // homography object plane -> normalised image, H22 = 1 (exact for four points: 8x8 linear system)
    double A[8][8], b[8], h[8];
    for (int i = 0; i < 4; ++i) {
        const double X = obj[i][0], Y = obj[i][1], u = img[i][0], v = img[i][1];
        const double r0[8] = {X, Y, 1, 0, 0, 0, -u * X, -u * Y}, r1[8] = {0, 0, 0, X, Y, 1, -v * X, -v * Y};
        for (int j = 0; j < 8; ++j) { A[2 * i][j] = r0[j]; A[2 * i + 1][j] = r1[j]; }
        b[2 * i] = u; b[2 * i + 1] = v;
    }
    solve_n<8>(A, b, h);
    for (int j = 0; j < 8; ++j)
        if (!(h[j] == h[j]) || fabs(h[j]) > 1e300) return false;   // singular: collinear points
    const double v0 = h[2], v1 = h[5];
    const double j00 = h[0] - h[6] * v0, j01 = h[1] - h[7] * v0, j10 = h[3] - h[6] * v1, j11 = h[4] - h[7] * v1;
    // Rv rotates the optical axis onto the ray through the image of the object origin
    const double s = sqrt(v0 * v0 + v1 * v1 + 1.0), tt = sqrt(v0 * v0 + v1 * v1);
    const double costh = 1.0 / s, sinth = sqrt(1.0 - 1.0 / (s * s));
    double Rv[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (tt > 2.2204460492503131e-16) {
        const double Kc[3][3] = {{0, 0, v0 / tt}, {0, 0, v1 / tt}, {-v0 / tt, -v1 / tt, 0}};
        double K2[3][3];
        mat3_mul(Kc, Kc, K2);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rv[i][j] += sinth * Kc[i][j] + (1.0 - costh) * K2[i][j];
    }
    // B = ([I | -v] Rv)(:, 0:2);  A2 = B^-1 J
    const double b00 = Rv[0][0] - v0 * Rv[2][0], b01 = Rv[0][1] - v0 * Rv[2][1];
    const double b10 = Rv[1][0] - v1 * Rv[2][0], b11 = Rv[1][1] - v1 * Rv[2][1];
    const double dtinv = 1.0 / (b00 * b11 - b01 * b10);
    const double i00 = dtinv * b11, i01 = -dtinv * b01, i10 = -dtinv * b10, i11 = dtinv * b00;
    const double a00 = i00 * j00 + i01 * j10, a01 = i00 * j01 + i01 * j11, a10 = i10 * j00 + i11 * j10, a11 = i10 * j01 + i11 * j11;
    // largest singular value of A2
    const double ata00 = a00 * a00 + a01 * a01, ata01 = a00 * a10 + a01 * a11, ata11 = a10 * a10 + a11 * a11;
    const double gamma = sqrt(0.5 * (ata00 + ata11 + sqrt((ata00 - ata11) * (ata00 - ata11) + 4.0 * ata01 * ata01)));
    if (!(gamma > 0.0)) return false;
    const double r00 = a00 / gamma, r01 = a01 / gamma, r10 = a10 / gamma, r11 = a11 / gamma;
    const double m00 = 1.0 - (r00 * r00 + r10 * r10), m11 = 1.0 - (r01 * r01 + r11 * r11), m01 = -(r00 * r01 + r10 * r11);
    const double b0 = sqrt(m00 > 0 ? m00 : 0.0);
    double b1 = sqrt(m11 > 0 ? m11 : 0.0);
    if (m01 < 0) b1 = -b1;
    PnpResult best;
    best.err = 1e300;
    for (int k = 0; k < 2; ++k) {
        const double sg = k == 0 ? 1.0 : -1.0;
        const double c0[3] = {r00, r10, sg * b0}, c1[3] = {r01, r11, sg * b1};
        double c2[3];
        cross3(c0, c1, c2);
        const double Rt[3][3] = {{c0[0], c1[0], c2[0]}, {c0[1], c1[1], c2[1]}, {c0[2], c1[2], c2[2]}};
        double R[3][3], t[3];
        mat3_mul(Rv, Rt, R);
        ippe_translation(obj, img, R, t);
        const double err = ippe_reproj_err(obj, img, R, t);
        if (err < best.err) {   // strict: on a tie OpenCV keeps the second solution only if it is strictly better
            best.err = err;
            rodrigues_to_vec(R, best.rvec);
            best.tvec[0] = t[0]; best.tvec[1] = t[1]; best.tvec[2] = t[2];
        }
    }
    *out = best;
    return true;
}

}  // namespace rmcv
