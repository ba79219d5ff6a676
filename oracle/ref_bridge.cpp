// TEST INFRASTRUCTURE ONLY — C entry points over the REFERENCE'S OWN C++ for the hot path.
//
// This translation unit #includes the reference sources where they lie (REF_ROOT = /root/reference, passed by
// oracle/Makefile as -I$(REF_ROOT) -I$(REF_ROOT)/include), unmodified:
//     src/core.cpp       rm::lightblob / rm::armour ctors, tracking methods, rm::utils geometry helpers
//     src/objdetect.cpp  MatchLightBlob, FindLightBlobs, filter_lightblobs, LightBlobOverlap, filter_armours
//     src/imgproc.cpp    extract_color, affine_correction
//     src/mobility.cpp   solve_PnP
// against oracle/cvstub/opencv2 (types + trampolines into the real OpenCV of this image, see its header).  Nothing of
// the reference is copied into this repository; the resulting library goes to oracle/_ref/ (git-ignored).
// Used by oracle/ref_bridge.py to pin oracle/rm_oracle.py and to regenerate tests/golden/.
// rm::armour keeps its Kalman observer private (default access of `class`); the tests read its state back.  Access is
// widened for this translation unit only: `class` reads as `struct` while the reference's own declarations are parsed
// (standard headers and the stub are included first, so nothing else sees the macro; layout and code are unchanged).
#include <filesystem>
#include <thread>
#include <sys/stat.h>
#include <opencv2/opencv.hpp>
#include <opencv2/ml.hpp>
#define class struct
#include "src/core.cpp"
#include "src/objdetect.cpp"
#include "src/imgproc.cpp"
#include "src/mobility.cpp"
#undef class

extern "C" {
rmcv_ref_cvcall_t rmcv_ref_cvcall = nullptr;
rmcv_ref_release_t rmcv_ref_release = nullptr;
double rmcv_ref_tick_frequency = 1e9;   // cv::getTickFrequency() on Linux (std::chrono::steady_clock, ns)

struct ref_blob {      // public fields of rm::lightblob (include/core.h:92-96); same layout as rmcv_lightblob
    float angle; int32_t target; float center[2]; float vertices[4][2]; float size[2];
};
struct ref_armour {    // public geometry of rm::armour (include/core.h:110-112) + the pair that produced it
    float icon[4][2]; float vertices[4][2]; float bounding_box[4]; int32_t i, j;
};
}

namespace {
thread_local std::string g_err;

void to_pod(const rm::lightblob& b, ref_blob* o) {
    o->angle = b.angle; o->target = (int32_t)b.target;
    o->center[0] = b.center.x; o->center[1] = b.center.y;
    for (int k = 0; k < 4; ++k) { o->vertices[k][0] = b.vertices[k].x; o->vertices[k][1] = b.vertices[k].y; }
    o->size[0] = b.size.width; o->size[1] = b.size.height;
}
// rm::lightblob has no default ctor and its fields are public: copy a prototype, then overwrite every field
rm::lightblob from_pod(const ref_blob& p) {
    static const rm::lightblob proto(cv::RotatedRect(cv::Point2f(0, 0), cv::Size2f(1, 1), 0), rm::CAMP_NEUTRAL);
    rm::lightblob b = proto;
    b.angle = p.angle; b.target = (rm::camp)p.target;
    b.center = cv::Point2f(p.center[0], p.center[1]);
    for (int k = 0; k < 4; ++k) b.vertices[k] = cv::Point2f(p.vertices[k][0], p.vertices[k][1]);
    b.size = cv::Size2f(p.size[0], p.size[1]);
    return b;
}
void to_pod(const rm::armour& a, ref_armour* o) {
    for (int k = 0; k < 4; ++k) {
        o->icon[k][0] = a.icon[k].x; o->icon[k][1] = a.icon[k].y;
        o->vertices[k][0] = a.vertices[k].x; o->vertices[k][1] = a.vertices[k].y;
    }
    o->bounding_box[0] = a.bounding_box.x; o->bounding_box[1] = a.bounding_box.y;
    o->bounding_box[2] = a.bounding_box.width; o->bounding_box[3] = a.bounding_box.height;
    o->i = o->j = -1;
}
std::vector<rm::contour> contours_from(const int32_t* xy, const int32_t* off, int n) {
    std::vector<rm::contour> cs((size_t)n);
    for (int k = 0; k < n; ++k)
        for (int i = off[k]; i < off[k + 1]; ++i) cs[(size_t)k].emplace_back(xy[2 * i], xy[2 * i + 1]);
    return cs;
}
template <class F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown exception"; return -1; }
}
}  // namespace

extern "C" {

void rmcv_ref_set_callback(rmcv_ref_cvcall_t cb) { rmcv_ref_cvcall = cb; }
void rmcv_ref_set_release(rmcv_ref_release_t cb) { rmcv_ref_release = cb; }
void rmcv_ref_set_tick_frequency(double f) { rmcv_ref_tick_frequency = f; }
const char* rmcv_ref_last_error() { return g_err.c_str(); }
int rmcv_ref_with_math_h() {
#ifdef RMCV_CVSTUB_WITH_MATH_H
    return 1;
#else
    return 0;
#endif
}

// ---- a3: rm::lightblob::lightblob (src/core.cpp:9-19)
int rmcv_ref_make_lightblob(const float box[5], int camp, ref_blob* out) {
    return guarded([&] {
        rm::lightblob b(cv::RotatedRect(cv::Point2f(box[0], box[1]), cv::Size2f(box[2], box[3]), box[4]), (rm::camp)camp);
        to_pod(b, out);
    });
}

// ---- a5: rm::armour::armour (src/core.cpp:21-49) on one pair
int rmcv_ref_make_armour(const ref_blob* b0, const ref_blob* b1, ref_armour* out) {
    return guarded([&] {
        rm::armour a({from_pod(*b0), from_pod(*b1)});
        to_pod(a, out);
    });
}

// ---- a4: rm::filter_armours (src/objdetect.cpp:114-166).  The reference returns armours without their pair indices
// (i = j = -1 here); callers that need them use rmcv_ref_pair_passes per pair, which walks the same code.
// *n_out = number of armours (<= cap are written).
int rmcv_ref_filter_armours(const ref_blob* blobs, int n, float angle_difference_max, float shear_max, float lenght_ratio_max,
                            int enemy, ref_armour* out, int cap, int* n_out) {
    return guarded([&] {
        std::vector<rm::lightblob> v;
        for (int k = 0; k < n; ++k) v.push_back(from_pod(blobs[k]));
        const std::vector<rm::armour> r = rm::filter_armours(v, angle_difference_max, shear_max, lenght_ratio_max, (rm::camp)enemy);
        *n_out = (int)r.size();
        for (int k = 0; k < (int)r.size() && k < cap; ++k) to_pod(r[(size_t)k], out + k);
    });
}

// one pair through rm::filter_armours: 1 = the pair passes every gate (src/objdetect.cpp:131-159)
int rmcv_ref_pair_passes(const ref_blob* bi, const ref_blob* bj, float angle_difference_max, float shear_max, float lenght_ratio_max,
                         int enemy, int* pass) {
    return guarded([&] {
        std::vector<rm::lightblob> v{from_pod(*bi), from_pod(*bj)};
        *pass = (int)rm::filter_armours(v, angle_difference_max, shear_max, lenght_ratio_max, (rm::camp)enemy).size();
    });
}

// ---- a2: rm::filter_lightblobs (src/objdetect.cpp:55-87).  status[k]: 0 skipped, 1 positive, 2 negative (recovered from
// the order-preserving outputs: positives and negatives both keep contour order; a negative is the contour itself).
int rmcv_ref_filter_lightblobs(const int32_t* xy, const int32_t* off, int n, float tilt_max, float ratio_min, float ratio_max,
                               double area_min, double area_max, int enemy, ref_blob* positive, int* n_positive,
                               int32_t* negative_index, int* n_negative) {
    return guarded([&] {
        const std::vector<rm::contour> cs = contours_from(xy, off, n);
        auto [pos, neg] = rm::filter_lightblobs(cs, tilt_max, rm::range<float>(ratio_min, ratio_max),
                                                rm::range<double>(area_min, area_max), (rm::camp)enemy);
        *n_positive = (int)pos.size();
        for (size_t k = 0; k < pos.size(); ++k) to_pod(pos[k], positive + k);
        // negatives are copies of input contours in input order: match them back greedily
        *n_negative = (int)neg.size();
        size_t at = 0;
        for (size_t k = 0; k < neg.size(); ++k) {
            while (at < cs.size() && !(cs[at].size() == neg[k].size() &&
                                        std::equal(cs[at].begin(), cs[at].end(), neg[k].begin(),
                                                   [](const cv::Point& a, const cv::Point& b) { return a.x == b.x && a.y == b.y; })))
                ++at;
            negative_index[k] = at < cs.size() ? (int32_t)at : -1;
            ++at;
        }
    });
}

// ---- a6: rm::MatchLightBlob (src/objdetect.cpp:9-28) / rm::FindLightBlobs (:30-53)
int rmcv_ref_match_lightblob(const int32_t* xy, int n_points, float min_ratio, float max_ratio, float tilt_angle, float min_area,
                             float max_area, int fit_ellipse, int* matched, float box[5]) {
    return guarded([&] {
        rm::contour c;
        for (int i = 0; i < n_points; ++i) c.emplace_back(xy[2 * i], xy[2 * i + 1]);
        cv::RotatedRect r;
        *matched = rm::MatchLightBlob(c, min_ratio, max_ratio, tilt_angle, min_area, max_area, r, fit_ellipse != 0) ? 1 : 0;
        box[0] = r.center.x; box[1] = r.center.y; box[2] = r.size.width; box[3] = r.size.height; box[4] = r.angle;
    });
}
int rmcv_ref_find_lightblobs(const int32_t* xy, const int32_t* off, int n, float min_ratio, float max_ratio, float tilt_angle,
                             float min_area, float max_area, const uint8_t* bgr, int rows, int cols, int channels, int fit_ellipse,
                             ref_blob* out, int* n_out) {
    return guarded([&] {
        std::vector<rm::contour> cs = contours_from(xy, off, n);
        cv::Mat src(rows, cols, CV_MAKETYPE(CV_8U, channels), const_cast<uint8_t*>(bgr));
        std::vector<rm::lightblob> blobs;
        rm::FindLightBlobs(cs, blobs, min_ratio, max_ratio, tilt_angle, min_area, max_area, src, fit_ellipse != 0);
        *n_out = (int)blobs.size();
        for (size_t k = 0; k < blobs.size(); ++k) to_pod(blobs[k], out + k);
    });
}

// ---- a7: rm::LightBlobOverlap (src/objdetect.cpp:89-112).  right == n is admitted by the reference's bound check and
// reads one past the end (UB); the bridge refuses it so that the oracle's guard is what gets compared.
int rmcv_ref_lightblob_overlap(const ref_blob* blobs, int n, int left, int right, int* result) {
    return guarded([&] {
        if (right >= n) { *result = 0; return; }
        std::vector<rm::lightblob> v;
        for (int k = 0; k < n; ++k) v.push_back(from_pod(blobs[k]));
        *result = rm::LightBlobOverlap(v, left, right) ? 1 : 0;
    });
}

// ---- rm::utils helpers (src/core.cpp:265-404)
int rmcv_ref_point_distance(const float p1[2], const float p2[2], float* out) {
    return guarded([&] { *out = rm::utils::PointDistance(cv::Point2f(p1[0], p1[1]), cv::Point2f(p2[0], p2[1])); });
}
int rmcv_ref_extend_cord(const float p1[2], const float p2[2], float delta, float d1[2], float d2[2]) {
    return guarded([&] {
        cv::Point2f a, b;
        rm::utils::ExtendCord(cv::Point2f(p1[0], p1[1]), cv::Point2f(p2[0], p2[1]), delta, a, b);
        d1[0] = a.x; d1[1] = a.y; d2[0] = b.x; d2[1] = b.y;
    });
}
int rmcv_ref_calc_perspective(const float in[8], float out_ratio, float out[8]) {
    return guarded([&] {
        cv::Point2f i4[4], o4[4];
        for (int k = 0; k < 4; ++k) i4[k] = cv::Point2f(in[2 * k], in[2 * k + 1]);
        rm::utils::CalcPerspective(i4, o4, out_ratio);
        for (int k = 0; k < 4; ++k) { out[2 * k] = o4[k].x; out[2 * k + 1] = o4[k].y; }
    });
}
int rmcv_ref_line_center(const float p1[2], const float p2[2], float out[2]) {
    return guarded([&] {
        const cv::Point2f c = rm::utils::LineCenter(cv::Point2f(p1[0], p1[1]), cv::Point2f(p2[0], p2[1]));
        out[0] = c.x; out[1] = c.y;
    });
}

// ---- a1: rm::extract_color (src/imgproc.cpp:50-75).  Contours come back flattened: xy[2*total], off[n+1].
int rmcv_ref_extract_color(const uint8_t* bgr, int rows, int cols, int channels, int target, int lower_bound, uint8_t* binary,
                           int32_t* xy, int xy_cap_points, int32_t* off, int off_cap, int* n_contours) {
    return guarded([&] {
        cv::Mat img(rows, cols, CV_MAKETYPE(CV_8U, channels), const_cast<uint8_t*>(bgr));
        auto [contours, bin] = rm::extract_color(img, (rm::camp)target, lower_bound);
        for (int r = 0; r < rows; ++r) std::memcpy(binary + (size_t)r * cols, bin.ptr(r), (size_t)cols);
        *n_contours = (int)contours.size();
        int at = 0;
        for (size_t k = 0; k < contours.size() && (int)k < off_cap - 1; ++k) {
            off[k] = at;
            for (const cv::Point& p : contours[k]) {
                if (at < xy_cap_points) { xy[2 * at] = p.x; xy[2 * at + 1] = p.y; }
                ++at;
            }
            off[k + 1] = at;
        }
    });
}

// ---- the per-frame hot loop of the reference's only caller (executable/main.cpp:172-176): the three calls back to back
// with that call site's argument order, results kept in C++ like process_function keeps them.  counts = {contours,
// positive, negative, armours}.  This is what bench.py's CPU arm times (kind "reference").
int rmcv_ref_process_frame(const uint8_t* bgr, int rows, int cols, int target, int lower_bound, float tilt_max, float ratio_min,
                           float ratio_max, double area_min, double area_max, float angle_difference_max, float shear_max,
                           float lenght_ratio_max, int32_t counts[4]) {
    return guarded([&] {
        cv::Mat img(rows, cols, CV_8UC3, const_cast<uint8_t*>(bgr));
        auto [contours, binary] = rm::extract_color(img, (rm::camp)target, lower_bound);
        auto [positive, negative] = rm::filter_lightblobs(contours, tilt_max, rm::range<float>(ratio_min, ratio_max),
                                                          rm::range<double>(area_min, area_max), (rm::camp)target);
        auto armours = rm::filter_armours(positive, angle_difference_max, shear_max, lenght_ratio_max, (rm::camp)target);
        counts[0] = (int32_t)contours.size(); counts[1] = (int32_t)positive.size(); counts[2] = (int32_t)negative.size();
        counts[3] = (int32_t)armours.size();
    });
}

// ---- f2: rm::affine_correction (src/imgproc.cpp:9-35); vertices are clamped in place like the reference does
int rmcv_ref_affine_correction(const uint8_t* bgr, int rows, int cols, int channels, float vertices[8], int out_w, int out_h,
                               uint8_t* out) {
    return guarded([&] {
        cv::Mat img(rows, cols, CV_MAKETYPE(CV_8U, channels), const_cast<uint8_t*>(bgr));
        cv::Point2f v[4];
        for (int k = 0; k < 4; ++k) v[k] = cv::Point2f(vertices[2 * k], vertices[2 * k + 1]);
        const cv::Mat cal = rm::affine_correction(img, v, cv::Size(out_w, out_h));
        for (int k = 0; k < 4; ++k) { vertices[2 * k] = v[k].x; vertices[2 * k + 1] = v[k].y; }
        for (int r = 0; r < cal.rows; ++r) std::memcpy(out + (size_t)r * cal.cols * channels, cal.ptr(r), (size_t)cal.cols * channels);
    });
}

// ---- f1: rm::solve_PnP (src/mobility.cpp:166-190)
int rmcv_ref_solve_pnp(const float points_image[8], const double K[9], const double dist[5], float w, float h, int roi_x, int roi_y,
                       double rvec[3], double tvec[3]) {
    return guarded([&] {
        cv::Point2f p[4];
        for (int k = 0; k < 4; ++k) p[k] = cv::Point2f(points_image[2 * k], points_image[2 * k + 1]);
        cv::Mat Km(3, 3, CV_64F, const_cast<double*>(K)), Dm(1, 5, CV_64F, const_cast<double*>(dist));
        auto [r, t] = rm::solve_PnP(p, Km, Dm, cv::Size2f(w, h), cv::Rect(roi_x, roi_y, 0, 0));
        for (int k = 0; k < 3; ++k) { rvec[k] = r.at<double>(k); tvec[k] = t.at<double>(k); }
    });
}

// ---- f3: the tracking side of rm::armour (src/core.cpp:51-162) on heap objects
void* rmcv_ref_armour_new(const float bounding_box[4], const double position[3], int identity, long long timestamp) {
    try {
        rm::armour* a = new rm::armour(std::vector<rm::lightblob>{});   // not 2 blobs: the ctor returns early (src/core.cpp:23)
        a->bounding_box = cv::Rect2f(bounding_box[0], bounding_box[1], bounding_box[2], bounding_box[3]);
        a->position = cv::Point3d(position[0], position[1], position[2]);
        a->identity = identity; a->timestamp = timestamp;
        return a;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void rmcv_ref_armour_delete(void* a) { delete static_cast<rm::armour*>(a); }
int rmcv_ref_armour_reset(void* a, double q, double r, double err) {
    return guarded([&] { static_cast<rm::armour*>(a)->reset(q, r, err); });
}
int rmcv_ref_armour_update_obs(void* a, void* obs) {
    return guarded([&] { static_cast<rm::armour*>(a)->update(*static_cast<rm::armour*>(obs)); });
}
int rmcv_ref_armour_update_time(void* a, long long ts) {
    return guarded([&] { static_cast<rm::armour*>(a)->update((int64)ts); });
}
int rmcv_ref_armour_identity_max(void* a, int* id, double* prob) {
    return guarded([&] { auto [i, p] = static_cast<rm::armour*>(a)->identity_max(); *id = i; *prob = p; });
}
int rmcv_ref_armour_max_iou(void* a, void** others, int n, int* index, float* iou) {
    return guarded([&] {
        std::vector<rm::armour> v;
        for (int k = 0; k < n; ++k) v.push_back(*static_cast<rm::armour*>(others[k]));
        auto [i, m] = static_cast<rm::armour*>(a)->max_IoU(v);
        *index = i; *iou = m;
    });
}
int rmcv_ref_armour_get(void* a, long long* timestamp, int* lost_count) {
    return guarded([&] { *timestamp = static_cast<rm::armour*>(a)->timestamp; *lost_count = static_cast<rm::armour*>(a)->lost_count; });
}
int rmcv_ref_armour_state(void* a, double state_post[6], double cov_post[36], int* initialized) {
    return guarded([&] {
        const rm::armour* p = static_cast<rm::armour*>(a);
        for (int i = 0; i < 6; ++i) state_post[i] = p->observer.statePost.at<double>(i);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) cov_post[6 * i + j] = p->observer.errorCovPost.at<double>(i, j);
        *initialized = p->initialized ? 1 : 0;
    });
}
void rmcv_ref_armour_set_lost(void* a, int lost) { static_cast<rm::armour*>(a)->lost_count = lost; }

}  // extern "C"
