// Legacy rows of the path, standalone on caller-supplied contours / light blobs:
//   rm::MatchLightBlob   src/objdetect.cpp:9-28    size/area gate, cv::fitEllipseDirect, box = ellipse or cv::minAreaRect,
//                                                  ratio gate on the box, tilt gate on the ellipse
//   rm::FindLightBlobs   src/objdetect.cpp:30-53   MatchLightBlob + camp vote from the mean colour of the contour's
//                                                  bounding rect in the source image + rm::lightblob ctor
//   rm::LightBlobOverlap src/objdetect.cpp:89-112
//   cv::minAreaRect                                convex hull in cv::convexHull's vertex order + OpenCV's float32 rotating
//                                                  calipers restated literally, tie rules included (calipers.cuh)
// One warp per contour.  The hull is a warp-cooperative gift wrapping with exact integer cross products; one lane then
// runs the calipers (O(hull) steps of a few flops).  The camp vote compares exact integer channel sums
// over the bounding rect (the same decision as comparing the means).
#include "calipers.cuh"
#include "common.cuh"
#include "warp_fit.cuh"

namespace rmcv {

struct LegacyParams {
    const int32_t* xy; const int32_t* off; int n_contours;
    float min_ratio, max_ratio, tilt_angle, min_area, max_area;
    int fit_ellipse;          // box = ellipse (1) or minAreaRect (0); -1 = minAreaRect only, no gates (rmcv_min_area_rects)
    const uint8_t* src; size_t pitch; int W, H;   // BGR source image for the camp vote, or null
    int32_t* hull;            // scratch: 6 ints per contour point
    int32_t* matched; rmcv_rotated_rect* boxes; int32_t* camps; rmcv_lightblob* blobs;
};

// cv::minAreaRect of the n points at pts (warp-cooperative hull, then OpenCV's own float32 rotating calipers on one lane:
// calipers.cuh).  hull: scratch for 6 ints per point (hull triples x, y, contour index + the same again for the reorder).
// All lanes return the same box.  Degenerate hulls (a point or a segment) give (length, 0) like OpenCV.
__device__ void min_area_rect_warp(const int32_t* pts, int n, int32_t* hull, int lane, rmcv_rotated_rect* out) {
    out->cx = out->cy = out->w = out->h = out->angle = 0.f;
    if (n <= 0) return;
    // start vertex: lowest y, then lowest x, then lowest contour index (coincident points: a contour revisits pixels)
    long long key = 0x7fffffffffffffffLL;
    int kidx = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const long long k = ((long long)pts[2 * i + 1] << 32) | (unsigned)pts[2 * i];
        if (k < key) { key = k; kidx = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const long long k = __shfl_xor_sync(0xffffffffu, key, o);
        const int ki = __shfl_xor_sync(0xffffffffu, kidx, o);
        if (k < key || (k == key && ki < kidx)) { key = k; kidx = ki; }
    }
    const int sx = (int)(key & 0xffffffffLL), sy = (int)(key >> 32);
    int h = 0, cx = sx, cy = sy, ci = kidx;
    while (h < n) {
        if (lane == 0) { hull[3 * h] = cx; hull[3 * h + 1] = cy; hull[3 * h + 2] = ci; }
        ++h;
        int bx = cx, by = cy, bi = 0x7fffffff;  // candidate (none yet)
        for (int i = lane; i < n; i += 32) {
            const int px = pts[2 * i], py = pts[2 * i + 1];
            if (px == cx && py == cy) continue;
            if (bi == 0x7fffffff || ((px != bx || py != by) && hull_better_wrap(cx, cy, bx, by, px, py))) { bx = px; by = py; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const int ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            // deterministic choice on both sides of the exchange: the better wrap, coincident points by the lower index
            bool take = false;
            if (oi != 0x7fffffff) {
                if (bi == 0x7fffffff) take = true;
                else if (ox != bx || oy != by) take = hull_better_wrap(cx, cy, bx, by, ox, oy);
                else take = oi < bi;
            }
            if (take) { bx = ox; by = oy; bi = oi; }
        }
        if (bi == 0x7fffffff || (bx == sx && by == sy)) break;  // single point, or wrapped around
        cx = bx; cy = by; ci = bi;
    }
    __syncwarp();
    rmcv_rotated_rect box;
    box.cx = box.cy = box.w = box.h = box.angle = 0.f;
    if (lane == 0) {
        if (h >= 3) hull_to_cv_order(hull, h, hull + 3 * (size_t)n);
        else if (h == 2 && (hull[3] > hull[0] || (hull[3] == hull[0] && hull[4] > hull[1]))) {   // OpenCV: (max x, max y) first
            for (int q = 0; q < 3; ++q) { const int32_t t = hull[q]; hull[q] = hull[3 + q]; hull[3 + q] = t; }
        }
        min_area_rect_from_hull(hull, h, &box);
    }
    out->cx = __shfl_sync(0xffffffffu, box.cx, 0); out->cy = __shfl_sync(0xffffffffu, box.cy, 0);
    out->w = __shfl_sync(0xffffffffu, box.w, 0); out->h = __shfl_sync(0xffffffffu, box.h, 0);
    out->angle = __shfl_sync(0xffffffffu, box.angle, 0);
}

__global__ void __launch_bounds__(128) legacy_kernel(const LegacyParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= p.n_contours) return;
    const int p0 = p.off[gw], p1 = p.off[gw + 1], n = p1 - p0;
    const int32_t* pts = p.xy + 2 * (size_t)p0;
    int32_t* hull = p.hull + 6 * (size_t)p0;
    rmcv_rotated_rect box;
    if (p.fit_ellipse < 0) {  // cv::minAreaRect only
        min_area_rect_warp(pts, n, hull, lane, &box);
        if (lane == 0) p.boxes[gw] = box;
        return;
    }
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
    x0 = __reduce_min_sync(0xffffffffu, x0); y0 = __reduce_min_sync(0xffffffffu, y0);
    x1 = __reduce_max_sync(0xffffffffu, x1); y1 = __reduce_max_sync(0xffffffffu, y1);
    int ok = 1;
    memset(&box, 0, sizeof(box));
    rmcv_lightblob blob;
    memset(&blob, 0, sizeof(blob));
    int camp = RMCV_CAMP_NEUTRAL;
    const double area = (double)(cross < 0 ? -cross : cross) * 0.5;
    if (n < 6 || area < (double)p.min_area || area > (double)p.max_area) ok = 0;   // :11 (double vs float compare)
    if (ok) {
        rmcv_params prm;
        prm.target = RMCV_CAMP_NEUTRAL; prm.lower_bound = 0; prm.tilt_max = 360.f; prm.ratio_min = 0.f; prm.ratio_max = 3.4e38f;
        prm.area_min = -1.0; prm.area_max = 1e300;
        prm.angle_difference_max = prm.shear_max = prm.lenght_ratio_max = 0.f;
        int status, branch; float det0; rmcv_rotated_rect ell; rmcv_lightblob tmp;
        auto pass = [&](auto&& fn) { for (int i = lane; i < n; i += 32) fn(pts[2 * i], pts[2 * i + 1]); };
        fit_and_gate(n, sx, sy, cross, prm, pass, &status, &branch, &det0, &ell, &tmp);   // :15
        if (p.fit_ellipse) box = ell; else min_area_rect_warp(pts, n, hull, lane, &box);  // :16
        const float ratio = fdiv(fmaxf(box.w, box.h), fminf(box.w, box.h));               // :19
        if (ratio > p.max_ratio || ratio < p.min_ratio) ok = 0;
        const float angle = ell.angle > 90.f ? fsub(ell.angle, 90.f) : fadd(ell.angle, 90.f);  // :23
        if (fabsf(fsub(angle, 90.f)) > p.tilt_angle) ok = 0;
    }
    if (ok && p.src != nullptr) {  // camp vote: mean(source(boundingRect(contour))), :43-51, as exact channel sums
        unsigned long long sb = 0, sg = 0, sr = 0;
        const int bw = x1 - x0 + 1;
        const long long npx = (long long)bw * (y1 - y0 + 1);
        for (long long i = lane; i < npx; i += 32) {
            const int yy = y0 + (int)(i / bw), xx = x0 + (int)(i % bw);
            if (xx < 0 || yy < 0 || xx >= p.W || yy >= p.H) continue;
            const uint8_t* px = p.src + (size_t)yy * p.pitch + (size_t)xx * 3;
            sb += px[0]; sg += px[1]; sr += px[2];
        }
        sb = (unsigned long long)warp_sum((long long)sb); sg = (unsigned long long)warp_sum((long long)sg);
        sr = (unsigned long long)warp_sum((long long)sr);
        camp = (sg > sb && sg > sr) ? RMCV_CAMP_GUIDELIGHT : (sb > sr ? RMCV_CAMP_BLUE : RMCV_CAMP_RED);
    }
    if (lane == 0) {
        p.matched[gw] = ok;
        p.boxes[gw] = box;
        p.camps[gw] = camp;
        if (ok) make_lightblob(box, camp, &blob);
        p.blobs[gw] = blob;
    }
}

cudaError_t launch_legacy(const LegacyLaunch& L, cudaStream_t st, int64_t* launches) {
    if (L.n_contours <= 0) return cudaSuccess;
    LegacyParams p;
    p.xy = L.xy; p.off = L.off; p.n_contours = L.n_contours;
    p.min_ratio = L.min_ratio; p.max_ratio = L.max_ratio; p.tilt_angle = L.tilt_angle; p.min_area = L.min_area; p.max_area = L.max_area;
    p.fit_ellipse = L.fit_ellipse;
    p.src = L.src; p.pitch = L.pitch; p.W = L.W; p.H = L.H;
    p.hull = L.hull; p.matched = L.matched; p.boxes = L.boxes; p.camps = L.camps; p.blobs = L.blobs;
    const int warps_per_block = 4;
    legacy_kernel<<<(L.n_contours + warps_per_block - 1) / warps_per_block, 128, 0, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// rm::LightBlobOverlap (src/objdetect.cpp:89-112).  The reference's bound check admits right == size() (one past the
// end, undefined behaviour); here right >= n returns false.
__global__ void overlap_kernel(const rmcv_lightblob* b, int n, int left, int right, int32_t* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int res = 0;
    if (!(left < 0 || right >= n || right - left < 2) && b[left].target == b[right].target) {
        const float lowerY = fminf(fminf(b[left].vertices[1][1], b[left].vertices[2][1]), fminf(b[right].vertices[1][1], b[right].vertices[2][1]));
        const float upperY = fmaxf(fmaxf(b[left].vertices[0][1], b[left].vertices[3][1]), fmaxf(b[right].vertices[0][1], b[right].vertices[3][1]));
        for (int i = left; i < right; ++i) {
            if (b[i].target != b[left].target) continue;
            if (b[i].center[0] > b[left].center[0] && b[i].center[0] < b[right].center[0] && b[i].center[1] > lowerY &&
                b[i].center[1] < upperY) { res = 1; break; }
        }
    }
    *out = res;
}

cudaError_t launch_overlap(const rmcv_lightblob* d_blobs, int n, int left, int right, int32_t* d_out, cudaStream_t st, int64_t* launches) {
    overlap_kernel<<<1, 32, 0, st>>>(d_blobs, n, left, right, d_out);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
