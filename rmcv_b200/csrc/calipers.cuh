// cv::minAreaRect (reference: src/objdetect.cpp:16, legacy rm::MatchLightBlob(fitEllipse = false)) restated LITERALLY:
// OpenCV's float32 rotating calipers (modules/imgproc/src/rotcalipers.cpp, the 4.5.1+ version with rotated-vector
// comparisons), fed with the convex hull in the vertex order cv::convexHull(points, clockwise = false) returns, and
// the [-90, 0) angle normalisation of OpenCV 4.13.  The result depends on that order through the calipers' tie rules
// (`area <= minarea` keeps the LAST of several equal-area rectangles; the first vertex with the extreme coordinate starts
// each caliper), so an "any minimum-area rectangle" evaluation differs on a few per cent of pixel contours.  Pinned bit for
// bit against cv2.minAreaRect on ~10^4 contours per run in tests/test_hostmath.py (CPU) and tests/test_gpu_legacy.py.
//
// Host+device inline code (the same header compiles under g++ for tests/hostmath).  fp32 operations are spelled with the
// explicit-rounding helpers of blob_math.cuh: OpenCV's build has no FMA contraction in this file's arithmetic.
#pragma once
#include "blob_math.cuh"

namespace rmcv {

// next hull vertex candidate while gift-wrapping around `cur`: is p "more clockwise" than q (y down: negative cross), or
// equally oriented and farther?  (collinear points in between are skipped, like OpenCV's Sklansky scan drops them)
RMCV_HD bool hull_better_wrap(int cx, int cy, int qx, int qy, int px, int py) {
    const long long cr = (long long)(qx - cx) * (py - cy) - (long long)(qy - cy) * (px - cx);
    if (cr != 0) return cr < 0;
    const long long dq = (long long)(qx - cx) * (qx - cx) + (long long)(qy - cy) * (qy - cy);
    const long long dp = (long long)(px - cx) * (px - cx) + (long long)(py - cy) * (py - cy);
    return dp > dq;
}

// A hull of h >= 3 vertices as (x, y, index in the contour) triples, strictly convex, any start and orientation ->
// the order cv::convexHull(contour, hull, false, true) returns:
//   1. orientation with a positive shoelace sum in image coordinates;
//   2. starting at the last point of OpenCV's (x, then y) sort: largest x, then largest y;
//   3. the cyclic shift of modules/imgproc/src/convhull.cpp that makes the ORIGINAL indices ascend or descend when they
//      already do so cyclically (always the case for the hull of a traced contour without revisited hull vertices).
// tmp: scratch for h triples.
RMCV_HD void hull_to_cv_order(int32_t* hull, int h, int32_t* tmp) {
    long long area2 = 0;
    for (int k = 0; k < h; ++k) {
        const int k1 = k + 1 == h ? 0 : k + 1;
        area2 += (long long)hull[3 * k] * hull[3 * k1 + 1] - (long long)hull[3 * k1] * hull[3 * k + 1];
    }
    const bool rev = area2 < 0;
    int first = 0;
    for (int k = 1; k < h; ++k)
        if (hull[3 * k] > hull[3 * first] || (hull[3 * k] == hull[3 * first] && hull[3 * k + 1] > hull[3 * first + 1])) first = k;
    for (int i = 0; i < h; ++i) {
        int k = rev ? first - i : first + i;
        k %= h;
        if (k < 0) k += h;
        tmp[3 * i] = hull[3 * k]; tmp[3 * i + 1] = hull[3 * k + 1]; tmp[3 * i + 2] = hull[3 * k + 2];
    }
    // "try to make the convex hull indices form an ascending or descending sequence by the cyclic shift of the output"
    int start = 0;
    {
        int min_idx = 0, max_idx = 0, lt = 0;
        for (int i = 1; i < h; ++i) {
            const int idx = tmp[3 * i + 2];
            lt += tmp[3 * (i - 1) + 2] < idx;
            if (lt > 1 && lt <= i - 2) break;
            if (idx < tmp[3 * min_idx + 2]) min_idx = i;
            if (idx > tmp[3 * max_idx + 2]) max_idx = i;
        }
        const int mmdist = max_idx > min_idx ? max_idx - min_idx : min_idx - max_idx;
        if ((mmdist == 1 || mmdist == h - 1) && (lt <= 1 || lt >= h - 2)) {
            const bool ascending = (max_idx + 1) % h == min_idx;
            const int i0 = ascending ? min_idx : max_idx;
            if (i0 > 0) {
                int j = i0, i = 0;
                for (; i < h; ++i) {
                    const int curr = tmp[3 * j + 2];
                    const int nj = j + 1 < h ? j + 1 : 0;
                    const int next = tmp[3 * nj + 2];
                    if (i < h - 1 && (ascending != (curr < next))) break;
                    j = nj;
                }
                if (i == h) start = i0;
            }
        }
    }
    for (int i = 0; i < h; ++i) {
        const int k = (start + i) % h;
        hull[3 * i] = tmp[3 * k]; hull[3 * i + 1] = tmp[3 * k + 1]; hull[3 * i + 2] = tmp[3 * k + 2];
    }
}

// rotatingCalipers(points, n, CALIPERS_MINAREARECT, out) on the hull triples (n >= 3): out = corner + two edge vectors.
RMCV_HD void rotating_calipers_min_area(const int32_t* hull, int n, float out[6]) {
    auto PX = [&](int i) -> float { return (float)hull[3 * i]; };
    auto PY = [&](int i) -> float { return (float)hull[3 * i + 1]; };
    // vect[i] = points[i+1] - points[i] (exact: integer coordinates), inv_vect_length[i] = (float)(1 / sqrt(dx^2 + dy^2))
    auto VX = [&](int i) -> float { const int j = i + 1 == n ? 0 : i + 1; return (float)(hull[3 * j] - hull[3 * i]); };
    auto VY = [&](int i) -> float { const int j = i + 1 == n ? 0 : i + 1; return (float)(hull[3 * j + 1] - hull[3 * i + 1]); };
    auto INV = [&](int i) -> float {
        const double dx = (double)VX(i), dy = (double)VY(i);
        return (float)(1. / sqrt(dx * dx + dy * dy));
    };
    int left = 0, bottom = 0, right = 0, top = 0;
    {
        float left_x = PX(0), right_x = PX(0), top_y = PY(0), bottom_y = PY(0);
        for (int i = 0; i < n; ++i) {
            const float x = PX(i), y = PY(i);
            if (x < left_x) { left_x = x; left = i; }
            if (x > right_x) { right_x = x; right = i; }
            if (y > top_y) { top_y = y; top = i; }
            if (y < bottom_y) { bottom_y = y; bottom = i; }
        }
    }
    float orientation = 0.f;
    {
        double ax = (double)VX(n - 1), ay = (double)VY(n - 1);
        for (int i = 0; i < n; ++i) {
            const double bx = (double)VX(i), by = (double)VY(i);
            const double convexity = ax * by - ay * bx;
            if (convexity != 0) { orientation = convexity > 0 ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    float minarea = 3.402823466e+38f;
    int b_left = 0, b_bottom = 0;
    float b_a = 0.f, b_w = 0.f, b_b = 0.f, b_h = 0.f;
    for (int k = 0; k < n; ++k) {
        // the caliper edge that needs the smallest rotation: edge vectors turned into the frame of caliper 0
        float rx[4], ry[4];
        rx[0] = VX(seq[0]); ry[0] = VY(seq[0]);
        rx[1] = VY(seq[1]); ry[1] = -VX(seq[1]);      // rotate90CW
        rx[2] = -VX(seq[2]); ry[2] = -VY(seq[2]);     // rotate180
        rx[3] = -VY(seq[3]); ry[3] = VX(seq[3]);      // rotate90CCW
        int main_element = 0;
        for (int i = 1; i < 4; ++i) {
            // firstVecIsRight(rot[i], rot[main]): rotate90CW(rot[i]) . rot[main] < 0
            const float tx = ry[i], ty = -rx[i];
            if (fadd(fmul(tx, rx[main_element]), fmul(ty, ry[main_element])) < 0.f) main_element = i;
        }
        {
            const int pindex = seq[main_element];
            const float inv = INV(pindex);
            const float lead_x = fmul(VX(pindex), inv), lead_y = fmul(VY(pindex), inv);
            switch (main_element) {
                case 0: base_a = lead_x; base_b = lead_y; break;
                case 1: base_a = lead_y; base_b = -lead_x; break;
                case 2: base_a = -lead_x; base_b = -lead_y; break;
                default: base_a = -lead_y; base_b = lead_x; break;
            }
        }
        seq[main_element] += 1;
        if (seq[main_element] == n) seq[main_element] = 0;
        float dx = fsub(PX(seq[1]), PX(seq[3])), dy = fsub(PY(seq[1]), PY(seq[3]));
        const float width = fadd(fmul(dx, base_a), fmul(dy, base_b));
        dx = fsub(PX(seq[2]), PX(seq[0])); dy = fsub(PY(seq[2]), PY(seq[0]));
        const float height = fadd(fmul(-dx, base_b), fmul(dy, base_a));
        const float area = fmul(width, height);
        if (area <= minarea) {
            minarea = area;
            b_left = seq[3]; b_a = base_a; b_w = width; b_b = base_b; b_h = height; b_bottom = seq[0];
        }
    }
    const float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
    const float C1 = fadd(fmul(A1, PX(b_left)), fmul(PY(b_left), B1));
    const float C2 = fadd(fmul(A2, PX(b_bottom)), fmul(PY(b_bottom), B2));
    const float idet = fdiv(1.f, fsub(fmul(A1, B2), fmul(A2, B1)));
    out[0] = fmul(fsub(fmul(C1, B2), fmul(C2, B1)), idet);
    out[1] = fmul(fsub(fmul(A1, C2), fmul(A2, C1)), idet);
    out[2] = fmul(A1, b_w); out[3] = fmul(B1, b_w);
    out[4] = fmul(A2, b_h); out[5] = fmul(B2, b_h);
}

// cv::minAreaRect from the hull triples in cv::convexHull order (h = number of hull vertices, 1, 2 or >= 3).
RMCV_HD void min_area_rect_from_hull(const int32_t* hull, int h, rmcv_rotated_rect* box) {
    float cx = 0.f, cy = 0.f, w = 0.f, hh = 0.f;
    double angle = 0.0;
    if (h > 2) {
        float o[6];
        rotating_calipers_min_area(hull, h, o);
        cx = fadd(o[0], fmul(fadd(o[2], o[4]), 0.5f));
        cy = fadd(o[1], fmul(fadd(o[3], o[5]), 0.5f));
        w = (float)sqrt((double)o[2] * o[2] + (double)o[3] * o[3]);
        hh = (float)sqrt((double)o[4] * o[4] + (double)o[5] * o[5]);
        angle = atan2((double)o[3], (double)o[2]);
    } else if (h == 2) {
        cx = fmul(fadd((float)hull[0], (float)hull[3]), 0.5f);
        cy = fmul(fadd((float)hull[1], (float)hull[4]), 0.5f);
        const double dx = (double)(hull[3] - hull[0]), dy = (double)(hull[4] - hull[1]);
        w = (float)sqrt(dx * dx + dy * dy);
        hh = 0.f;
        angle = atan2(dy, dx);
    } else if (h == 1) {
        cx = (float)hull[0]; cy = (float)hull[1];
    }
    // degrees, brought into [-90, 0) in double (width and height swap with every quarter turn), rounded once
    angle = angle * 180 / RMCV_PI;
    while (angle >= 0) { angle -= 90; const float t = w; w = hh; hh = t; }
    while (angle < -90) { angle += 90; const float t = w; w = hh; hh = t; }
    box->cx = cx; box->cy = cy; box->w = w; box->h = hh; box->angle = (float)angle;
}

}  // namespace rmcv
