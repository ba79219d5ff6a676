// rm:: drop-in shim over the rmcv_b200 C ABI (include/rmcv_b200.h).
//
// Re-declares, with the reference's exact signatures, the three free functions the reference's only caller uses
// (executable/main.cpp:172-176):
//     rm::extract_color      include/imgproc.h:29       src/imgproc.cpp:50-75
//     rm::filter_lightblobs  include/objdetect.h:47-49  src/objdetect.cpp:55-87
//     rm::filter_armours     include/objdetect.h:70-71  src/objdetect.cpp:114-166
// and routes them to librmcv_b200.so.  Two build modes:
//
//   * RMCV_SHIM_WITH_REFERENCE — inside the reference tree: includes the reference's own "core.h" (OpenCV types,
//     rm::lightblob / rm::armour with their constructors).  Compile rm_shim.hpp's definitions INSTEAD OF the bodies in
//     src/imgproc.cpp:50-75 and src/objdetect.cpp:55-87,114-166.  Light blobs and armours are
//     rebuilt through the reference's public constructors from the GPU's cv::RotatedRect / blob pairs, so private
//     members (Kalman filter, history) stay valid.
//   * default (no OpenCV on this image) — a minimal `cv::`/`rm::` type set with the same public fields, enough to
//     compile and run the call site (tests/cpp/call_site.cpp).
//
// Header-only; link with -lrmcv_b200.  One rm::gpu::context per host thread (thread_local default provided).
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../rmcv_b200.h"

#if defined(RMCV_SHIM_WITH_REFERENCE)
#include "core.h"  // the reference's include/core.h (pulls in OpenCV)
#else
namespace cv {
struct Point { int x = 0, y = 0; };
struct Point2f { float x = 0, y = 0; };
struct Size2f { float width = 0, height = 0; };
struct Size { int width = 0, height = 0; };
struct Rect2f { float x = 0, y = 0, width = 0, height = 0; };
struct RotatedRect { Point2f center; Size2f size; float angle = 0; };
// 8-bit image view/owner with cv::Mat's field names for the members the shim touches
struct Mat {
    int rows = 0, cols = 0, chans = 1;
    size_t step = 0;
    uint8_t* data = nullptr;
    std::vector<uint8_t> storage;
    Mat() = default;
    Mat(int r, int c, int ch) : rows(r), cols(c), chans(ch), step((size_t)c * ch), storage((size_t)r * c * ch) { data = storage.data(); }
    Mat(int r, int c, int ch, uint8_t* ext, size_t st) : rows(r), cols(c), chans(ch), step(st), data(ext) {}
    // an owning Mat must re-point `data` at its own storage when copied or moved
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), chans(o.chans), step(o.step), data(o.data), storage(o.storage) {
        if (!o.storage.empty()) data = storage.data();
    }
    Mat(Mat&& o) noexcept : rows(o.rows), cols(o.cols), chans(o.chans), step(o.step), data(o.data) {
        const bool owning = !o.storage.empty();
        storage = std::move(o.storage);
        if (owning) data = storage.data();
        o.data = nullptr; o.rows = o.cols = 0;
    }
    Mat& operator=(Mat o) {
        rows = o.rows; cols = o.cols; chans = o.chans; step = o.step;
        const bool owning = !o.storage.empty();
        storage = std::move(o.storage);
        data = owning ? storage.data() : o.data;
        return *this;
    }
    int channels() const { return chans; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
};
using InputArray = const Mat&;
}  // namespace cv

namespace rm {
enum camp { CAMP_RED = 0, CAMP_BLUE = 1, CAMP_GUIDELIGHT = 2, CAMP_NEUTRAL = -1 };  // include/core.h:20-23
template <typename T>
struct range {  // include/core.h:30-44
    T lower_bound, upper_bound;
    range(T lower, T upper) : lower_bound(lower), upper_bound(upper) {}
    bool contains(T value) const { return value >= lower_bound && value <= upper_bound; }
};
typedef std::vector<cv::Point> contour;  // include/core.h:87
class lightblob {  // public fields of include/core.h:89-99
   public:
    float angle = 0;
    camp target = CAMP_NEUTRAL;
    cv::Point2f center;
    cv::Point2f vertices[4];
    cv::Size2f size;
};
class armour {  // public geometry of include/core.h:101-130
   public:
    cv::Point2f icon[4];
    cv::Point2f vertices[4];
    cv::Rect2f bounding_box;
    int64_t timestamp = 0;
    int lost_count = 0;
    int identity = -1;
};
}  // namespace rm
#endif

namespace rm {
namespace gpu {

struct error : std::runtime_error {
    int status;
    error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// What rm::extract_color leaves behind for the two calls that follow it at the reference's call site
// (executable/main.cpp:172-176): the device already ran the whole path on the frame, so rm::filter_lightblobs /
// rm::filter_armours hand back those results when they are called with the very contours / light blobs the previous
// call returned and with the parameters it ran with ("hidden handle").  Anything else — foreign contours,
// other thresholds — goes through the standalone kernels.
struct last_frame {
    bool valid = false;
    rmcv_params prm;                          // parameters the device ran with
    std::vector<int32_t> xy, off;             // the contours handed to the caller (exact copy)
    std::vector<rmcv_contour_info> infos;     // per contour: verdict, ellipse, blob index
    std::vector<rmcv_lightblob> blobs;        // positives in contour order
    std::vector<rmcv_armour> armours;
    bool blobs_handed_out = false;            // filter_lightblobs returned `blobs` unchanged
    long long reused_lightblobs = 0, reused_armours = 0;   // diagnostics (tests/cpp/call_site.cpp)
};

// Owns one rmcv_ctx sized for single-frame calls; grows on demand (frame size and per-frame capacities).
class context {
   public:
    context() = default;
    ~context() { if (ctx_) rmcv_ctx_destroy(ctx_); }
    context(const context&) = delete;
    context& operator=(const context&) = delete;

    rmcv_ctx* get(int width, int height) {
        if (!ctx_ || width > w_ || height > h_) {
            w_ = width > w_ ? width : w_;
            h_ = height > h_ ? height : h_;
            if (w_ < 1280) w_ = 1280;
            if (h_ < 1024) h_ = 1024;
            rebuild();
        }
        return ctx_;
    }
    // A frame overflowed a per-frame capacity (RMCV_FRAME_OVERFLOW_*): double what overflowed and rebuild the ctx.
    // false when the capacities are already at their ceilings (a frame cannot hold more runs than pixels / 2).
    bool grow(int flags) {
        const long long px = (long long)w_ * h_;
        bool grew = false;
        if (flags & (RMCV_FRAME_OVERFLOW_RUNS | RMCV_FRAME_OVERFLOW_POINTS)) {
            long long cur = runs_ > 0 ? runs_ : (px / 32 > 16384 ? px / 32 : 16384);
            if (cur < px / 2 + h_) { runs_ = (int)(cur * 4 < px / 2 + h_ ? cur * 4 : px / 2 + h_); grew = true; }
        }
        if (flags & RMCV_FRAME_OVERFLOW_BLOBS) {
            const int cur = blobs_ > 0 ? blobs_ : 512;
            if (cur < (1 << 15) - 1) { blobs_ = cur * 4 < (1 << 15) - 1 ? cur * 4 : (1 << 15) - 1; grew = true; }   // component ids are 16 bit
        }
        if (flags & RMCV_FRAME_OVERFLOW_ARMOURS) {
            const int cur = armours_ > 0 ? armours_ : 1024;
            if (cur < (1 << 22)) { armours_ = cur * 4; grew = true; }
        }
        if (grew) rebuild();
        return grew;
    }
    last_frame last;
    void check(int rc, const char* where) {
        if (rc != RMCV_OK) throw error(rc, std::string(where) + ": " + rmcv_status_string(rc) + " - " + (ctx_ ? rmcv_last_error(ctx_) : ""));
    }

   private:
    void rebuild() {
        if (ctx_) rmcv_ctx_destroy(ctx_);
        ctx_ = nullptr;
        last.valid = false;
        rmcv_config cfg;
        rmcv_default_config(&cfg);
        cfg.max_width = w_; cfg.max_height = h_; cfg.max_batch = 1;
        cfg.max_runs_per_frame = runs_; cfg.max_blobs_per_frame = blobs_; cfg.max_armours_per_frame = armours_;
        const int rc = rmcv_ctx_create(&cfg, &ctx_);
        if (rc != RMCV_OK) throw error(rc, std::string("rmcv_ctx_create: ") + rmcv_status_string(rc) + " (no CPU fallback exists)");
    }
    rmcv_ctx* ctx_ = nullptr;
    int w_ = 0, h_ = 0, runs_ = 0, blobs_ = 0, armours_ = 0;
};

inline context& default_context() {
    thread_local context c;
    return c;
}

inline lightblob to_lightblob(const rmcv_lightblob& b, const rmcv_rotated_rect* box) {
#if defined(RMCV_SHIM_WITH_REFERENCE)
    // rebuild through the reference's own ctor (src/core.cpp:9-19) from the GPU's ellipse
    (void)b;
    return lightblob(cv::RotatedRect(cv::Point2f(box->cx, box->cy), cv::Size2f(box->w, box->h), box->angle), static_cast<camp>(b.target));
#else
    (void)box;
    lightblob o;
    o.angle = b.angle; o.target = static_cast<camp>(b.target);
    o.center.x = b.center[0]; o.center.y = b.center[1];
    for (int i = 0; i < 4; ++i) { o.vertices[i].x = b.vertices[i][0]; o.vertices[i].y = b.vertices[i][1]; }
    o.size.width = b.size[0]; o.size.height = b.size[1];
    return o;
#endif
}

inline rmcv_lightblob from_lightblob(const lightblob& b) {
    rmcv_lightblob o;
    o.angle = b.angle; o.target = static_cast<int32_t>(b.target);
    o.center[0] = b.center.x; o.center[1] = b.center.y;
    for (int i = 0; i < 4; ++i) { o.vertices[i][0] = b.vertices[i].x; o.vertices[i][1] = b.vertices[i].y; }
    o.size[0] = b.size.width; o.size[1] = b.size.height;
    return o;
}

// The whole path in one call (not in the reference API; what a caller that does not need the contours should use).
struct detection {
    std::vector<lightblob> positive;
    std::vector<armour> armours;
    std::vector<rmcv_contour_info> contours;
};

}  // namespace gpu

// ---------------------------------------------------------------------------------------------------------------
// rm::extract_color — include/imgproc.h:29.  Returns the external contours (cv::findContours order, all points) and
// the binary mask, exactly like src/imgproc.cpp:50-75.
inline std::tuple<std::vector<contour>, cv::Mat> extract_color(cv::InputArray image_in, camp target, int lower_bound) {
#if defined(RMCV_SHIM_WITH_REFERENCE)
    cv::Mat image = image_in.getMat();
    cv::Mat binary(image.rows, image.cols, CV_8UC1);
    const size_t bstep = binary.step;
#else
    const cv::Mat& image = image_in;
    cv::Mat binary(image.rows, image.cols, 1);
    const size_t bstep = binary.step;
#endif
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(image.cols, image.rows);
    rmcv_params prm;
    rmcv_default_params(&prm);
    prm.target = static_cast<int32_t>(target);
    prm.lower_bound = lower_bound;
    rmcv_results res;
    int rc = RMCV_OK;
    for (;;) {
        rc = rmcv_detect_batch_host(ctx, image.data, image.step, image.step * (size_t)image.rows, image.cols, image.rows, 1, &prm,
                                    binary.data, bstep, bstep * (size_t)image.rows, &res);
        if (rc != RMCV_ERR_CAPACITY) break;
        // The reference returns EVERY external contour (src/imgproc.cpp:71-72).  A frame that overflows the ctx's run /
        // component / boundary-point capacities would come back truncated, so the ctx is rebuilt with larger ones and the
        // frame runs again; an armour overflow does not touch the contours (filter_armours below re-runs those pairs).
        const int flags = res.frames[0].flags;
        if (!(flags & (RMCV_FRAME_OVERFLOW_RUNS | RMCV_FRAME_OVERFLOW_BLOBS | RMCV_FRAME_OVERFLOW_POINTS))) break;
        if (!gc.grow(flags)) throw gpu::error(rc, "rm::extract_color: frame exceeds the largest per-frame capacities (contour list would be truncated)");
        ctx = gc.get(image.cols, image.rows);
    }
    if (rc != RMCV_OK && rc != RMCV_ERR_CAPACITY) gc.check(rc, "rmcv_detect_batch_host");
    const bool armour_overflow = rc == RMCV_ERR_CAPACITY;
    int nc = res.frames[0].n_contours, np = 0;       // sizes straight from the frame's records: one tracing call below
    for (int k = 0; k < nc; ++k) np += res.contours[res.frames[0].contour_offset + k].n_points;
    std::vector<int32_t> xy((size_t)np * 2 + 2), off((size_t)nc + 1);
    if (nc > 0) gc.check(rmcv_get_contours(ctx, 0, xy.data(), np, off.data(), nc, &nc, &np), "rmcv_get_contours");
    std::vector<contour> contours((size_t)nc);
    for (int k = 0; k < nc; ++k) {
        contours[k].resize((size_t)(off[k + 1] - off[k]));
        for (int i = off[k]; i < off[k + 1]; ++i) { contours[k][i - off[k]].x = xy[2 * i]; contours[k][i - off[k]].y = xy[2 * i + 1]; }
    }
    // keep what the device computed for this frame: the two calls that follow at the reference's call site reuse it
    gpu::last_frame& lf = gc.last;
    const rmcv_frame_info& fi = res.frames[0];
    lf.prm = prm;
    xy.resize((size_t)np * 2);
    lf.xy.swap(xy); lf.off.swap(off);
    lf.infos.assign(res.contours + fi.contour_offset, res.contours + fi.contour_offset + fi.n_contours);
    lf.blobs.assign(res.blobs + fi.blob_offset, res.blobs + fi.blob_offset + fi.n_positive);
    lf.armours.assign(res.armours + fi.armour_offset, res.armours + fi.armour_offset + fi.n_armours);
    lf.blobs_handed_out = false;
    lf.valid = !armour_overflow && (int)lf.infos.size() == nc;
    return {contours, binary};
}

// rm::filter_lightblobs — include/objdetect.h:47-49, src/objdetect.cpp:55-87.
inline auto filter_lightblobs(const std::vector<contour>& contours, const float tilt_max, const range<float> ratio_range,
                              const range<double> area_range, camp enemy)
    -> std::tuple<std::vector<lightblob>, std::vector<contour>> {
    std::vector<lightblob> positive;
    std::vector<contour> negative;
    if (contours.empty()) return {positive, negative};
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(1, 1);
    std::vector<int32_t> off(contours.size() + 1, 0);
    for (size_t k = 0; k < contours.size(); ++k) off[k + 1] = off[k] + (int32_t)contours[k].size();
    std::vector<int32_t> xy((size_t)off.back() * 2 + 2);
    for (size_t k = 0; k < contours.size(); ++k)
        for (size_t i = 0; i < contours[k].size(); ++i) { xy[2 * (off[k] + i)] = contours[k][i].x; xy[2 * (off[k] + i) + 1] = contours[k][i].y; }
    rmcv_params prm;
    rmcv_default_params(&prm);
    prm.target = static_cast<int32_t>(enemy);
    prm.tilt_max = tilt_max;
    prm.ratio_min = ratio_range.lower_bound; prm.ratio_max = ratio_range.upper_bound;
    prm.area_min = area_range.lower_bound; prm.area_max = area_range.upper_bound;
    {   // the contours rm::extract_color just returned, with the parameters the device ran with: its results are the answer
        gpu::last_frame& lf = gc.last;
        const size_t nxy = (size_t)off.back() * 2;
        if (lf.valid && lf.off == off && lf.xy.size() == nxy && std::memcmp(lf.xy.data(), xy.data(), nxy * sizeof(int32_t)) == 0 &&
            lf.prm.target == prm.target && lf.prm.tilt_max == prm.tilt_max && lf.prm.ratio_min == prm.ratio_min &&
            lf.prm.ratio_max == prm.ratio_max && lf.prm.area_min == prm.area_min && lf.prm.area_max == prm.area_max) {
            for (size_t k = 0; k < contours.size(); ++k) {
                if (lf.infos[k].status == RMCV_CONTOUR_NEGATIVE) negative.push_back(contours[k]);                       // :82
                else if (lf.infos[k].status == RMCV_CONTOUR_POSITIVE)
                    positive.push_back(gpu::to_lightblob(lf.blobs[(size_t)lf.infos[k].blob_index], &lf.infos[k].ellipse));   // :83
            }
            lf.blobs_handed_out = true;
            ++lf.reused_lightblobs;
            return {positive, negative};
        }
        lf.blobs_handed_out = false;
    }
    std::vector<rmcv_contour_info> infos(contours.size());
    std::vector<rmcv_lightblob> blobs(contours.size());
    int nb = 0;
    gc.check(rmcv_filter_lightblobs(ctx, xy.data(), off.data(), (int)contours.size(), &prm, infos.data(), blobs.data(), (int)blobs.size(), &nb),
             "rmcv_filter_lightblobs");
    for (size_t k = 0; k < contours.size(); ++k) {
        if (infos[k].status == RMCV_CONTOUR_NEGATIVE) negative.push_back(contours[k]);                       // :82
        else if (infos[k].status == RMCV_CONTOUR_POSITIVE) positive.push_back(gpu::to_lightblob(blobs[infos[k].blob_index], &infos[k].ellipse));  // :83
    }
    return {positive, negative};
}

// rm::filter_armours — include/objdetect.h:70-71, src/objdetect.cpp:114-166.
inline std::vector<armour> filter_armours(std::vector<lightblob>& lightblobs, const float angle_difference_max, const float shear_max,
                                          const float lenght_ratio_max, const camp enemy) {
    std::vector<armour> armours;
    if (lightblobs.size() < 2) return armours;  // :120
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(1, 1);
    std::vector<rmcv_lightblob> in(lightblobs.size());
    for (size_t i = 0; i < lightblobs.size(); ++i) in[i] = gpu::from_lightblob(lightblobs[i]);
    rmcv_params prm;
    rmcv_default_params(&prm);
    prm.target = static_cast<int32_t>(enemy);
    prm.angle_difference_max = angle_difference_max; prm.shear_max = shear_max; prm.lenght_ratio_max = lenght_ratio_max;
    int cap = 256, n = 0;
    std::vector<rmcv_armour> out((size_t)cap);
    int rc = RMCV_OK;
    gpu::last_frame& lf = gc.last;
    // the light blobs rm::filter_lightblobs just handed out (field for field), with the pair parameters the device ran with
    const bool reuse = lf.valid && lf.blobs_handed_out && lf.blobs.size() == in.size() && lf.prm.target == prm.target &&
                       lf.prm.angle_difference_max == prm.angle_difference_max && lf.prm.shear_max == prm.shear_max &&
                       lf.prm.lenght_ratio_max == prm.lenght_ratio_max &&
                       std::memcmp(lf.blobs.data(), in.data(), in.size() * sizeof(rmcv_lightblob)) == 0;
    if (reuse) {
        out = lf.armours;
        n = (int)out.size();
        ++lf.reused_armours;
    } else {
        rc = rmcv_filter_armours(ctx, in.data(), (int)in.size(), &prm, out.data(), cap, &n);
        if (rc == RMCV_ERR_CAPACITY) {
            cap = n; out.resize((size_t)cap);
            rc = rmcv_filter_armours(ctx, in.data(), (int)in.size(), &prm, out.data(), cap, &n);
        }
        gc.check(rc, "rmcv_filter_armours");
    }
    for (int k = 0; k < n; ++k) {
#if defined(RMCV_SHIM_WITH_REFERENCE)
        armours.push_back(armour({lightblobs[out[k].i], lightblobs[out[k].j]}));  // the reference's own ctor, :161
#else
        armour a;
        for (int i = 0; i < 4; ++i) {
            a.icon[i].x = out[k].icon[i][0]; a.icon[i].y = out[k].icon[i][1];
            a.vertices[i].x = out[k].vertices[i][0]; a.vertices[i].y = out[k].vertices[i][1];
        }
        a.bounding_box.x = out[k].bounding_box[0]; a.bounding_box.y = out[k].bounding_box[1];
        a.bounding_box.width = out[k].bounding_box[2]; a.bounding_box.height = out[k].bounding_box[3];
        armours.push_back(a);
#endif
    }
    return armours;
}

// ---------------------------------------------------------------------------------------------------------------
// Legacy functions (include/objdetect.h:22-37,62; src/objdetect.cpp:9-53,89-112), same signatures.
namespace gpu {
inline void pack_contours(const std::vector<contour>& contours, std::vector<int32_t>& xy, std::vector<int32_t>& off) {
    off.assign(contours.size() + 1, 0);
    for (size_t k = 0; k < contours.size(); ++k) off[k + 1] = off[k] + (int32_t)contours[k].size();
    xy.resize((size_t)off.back() * 2 + 2);
    for (size_t k = 0; k < contours.size(); ++k)
        for (size_t i = 0; i < contours[k].size(); ++i) { xy[2 * (off[k] + i)] = contours[k][i].x; xy[2 * (off[k] + i) + 1] = contours[k][i].y; }
}
inline cv::RotatedRect to_rotated_rect(const rmcv_rotated_rect& b) {
#if defined(RMCV_SHIM_WITH_REFERENCE)
    return cv::RotatedRect(cv::Point2f(b.cx, b.cy), cv::Size2f(b.w, b.h), b.angle);
#else
    cv::RotatedRect r; r.center.x = b.cx; r.center.y = b.cy; r.size.width = b.w; r.size.height = b.h; r.angle = b.angle;
    return r;
#endif
}
}  // namespace gpu

inline bool MatchLightBlob(const contour& c, float minRatio, float maxRatio, float tiltAngle, float minArea, float maxArea,
                           cv::RotatedRect& lightBlobBox, bool fitEllipse = true) {
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(1, 1);
    std::vector<int32_t> xy, off;
    gpu::pack_contours({c}, xy, off);
    int32_t matched = 0;
    rmcv_rotated_rect box;
    gc.check(rmcv_match_lightblobs(ctx, xy.data(), off.data(), 1, minRatio, maxRatio, tiltAngle, minArea, maxArea, fitEllipse ? 1 : 0,
                                   &matched, &box), "rmcv_match_lightblobs");
    if (matched) lightBlobBox = gpu::to_rotated_rect(box);
    return matched != 0;
}

inline void FindLightBlobs(std::vector<contour>& contours, std::vector<lightblob>& lightBlobs, float minRatio, float maxRatio,
                           float tiltAngle, float minArea, float maxArea, const cv::Mat& source, bool fitEllipse = true) {
    lightBlobs.clear();
    if (source.channels() != 3 || contours.empty()) return;   // src/objdetect.cpp:35
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(1, 1);
    std::vector<int32_t> xy, off;
    gpu::pack_contours(contours, xy, off);
    std::vector<rmcv_lightblob> out(contours.size());
    int nb = 0;
    gc.check(rmcv_find_lightblobs_legacy(ctx, xy.data(), off.data(), (int)contours.size(), minRatio, maxRatio, tiltAngle, minArea, maxArea,
                                         source.data, source.step, source.cols, source.rows, fitEllipse ? 1 : 0, out.data(),
                                         (int)out.size(), &nb), "rmcv_find_lightblobs_legacy");
    for (int k = 0; k < nb; ++k) {
#if defined(RMCV_SHIM_WITH_REFERENCE)
        // rebuild through the reference's ctor from the corner points' rectangle is not possible without the box; the
        // C ABI returns the finished light blob, copy its public fields into a default-constructed-from-box object
        lightblob lb(cv::RotatedRect(cv::Point2f(out[k].center[0], out[k].center[1]), cv::Size2f(out[k].size[0], out[k].size[1]), 0.f),
                     static_cast<camp>(out[k].target));
        lb.angle = out[k].angle;
        for (int i = 0; i < 4; ++i) lb.vertices[i] = cv::Point2f(out[k].vertices[i][0], out[k].vertices[i][1]);
        lb.size = cv::Size2f(out[k].size[0], out[k].size[1]);
        lightBlobs.push_back(lb);
#else
        lightBlobs.push_back(gpu::to_lightblob(out[k], nullptr));
#endif
    }
}

inline bool LightBlobOverlap(const std::vector<lightblob>& lightBlobs, int leftIndex, int rightIndex) {
    if (lightBlobs.empty()) return false;
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(1, 1);
    std::vector<rmcv_lightblob> in(lightBlobs.size());
    for (size_t i = 0; i < lightBlobs.size(); ++i) in[i] = gpu::from_lightblob(lightBlobs[i]);
    int res = 0;
    gc.check(rmcv_lightblob_overlap(ctx, in.data(), (int)in.size(), leftIndex, rightIndex, &res), "rmcv_lightblob_overlap");
    return res != 0;
}

// rm::affine_correction — include/imgproc.h:19, src/imgproc.cpp:9-35 (next row f2): the icon crop of one armour, bit-exact
// against OpenCV's getAffineTransform / warpAffine / resize.  The four vertices are clamped into the frame in place like the
// reference does.  The frame is uploaded for this call; rmcv_icon_batch / rmcv_identify_batch take a device-resident frame
// and all armours of it at once.
inline cv::Mat affine_correction(const cv::Mat& source, cv::Point2f vertices[4], const cv::Size outSize) {
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(source.cols, source.rows);
    rmcv_armour a;
    std::memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; ++i) { a.icon[i][0] = vertices[i].x; a.icon[i][1] = vertices[i].y; }
    const size_t bytes = (size_t)source.step * (size_t)source.rows;
    void* d_frame = nullptr;
    gc.check(rmcv_device_alloc(ctx, bytes, &d_frame), "rmcv_device_alloc");
    std::vector<uint8_t> icon((size_t)outSize.width * outSize.height * 3);
    int rc = rmcv_memcpy_h2d(ctx, d_frame, source.data, bytes);
    if (rc == RMCV_OK)
        rc = rmcv_icon_batch(ctx, static_cast<const uint8_t*>(d_frame), (size_t)source.step, source.cols, source.rows, &a, 1, outSize.width,
                             outSize.height, icon.data(), nullptr);
    rmcv_device_free(ctx, d_frame);
    gc.check(rc, "rmcv_icon_batch");
    for (int i = 0; i < 4; ++i) { vertices[i].x = a.icon[i][0]; vertices[i].y = a.icon[i][1]; }
#if defined(RMCV_SHIM_WITH_REFERENCE)
    cv::Mat out(outSize.height, outSize.width, CV_8UC3);
#else
    cv::Mat out(outSize.height, outSize.width, 3);
#endif
    for (int y = 0; y < outSize.height; ++y)
        std::memcpy(out.data + (size_t)y * out.step, icon.data() + (size_t)y * outSize.width * 3, (size_t)outSize.width * 3);
    return out;
}

#if defined(RMCV_SHIM_WITH_REFERENCE)
// rm::solve_PnP — include/mobility.h:106-108, src/mobility.cpp:166-190 (needs cv::Mat of doubles: reference build only).
inline std::tuple<cv::Mat, cv::Mat> solve_PnP(const cv::Point2f points_image[4], cv::InputArray cameraMatrix,
                                              cv::InputArray distortionFactor, const cv::Size2f& exactSize,
                                              const cv::Rect& ROI = {0, 0, 0, 0}) {
    gpu::context& gc = gpu::default_context();
    rmcv_ctx* ctx = gc.get(1, 1);
    rmcv_armour a;
    std::memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; ++i) { a.vertices[i][0] = points_image[i].x; a.vertices[i][1] = points_image[i].y; }
    cv::Mat K, D;
    cameraMatrix.getMat().convertTo(K, CV_64F);
    distortionFactor.getMat().convertTo(D, CV_64F);
    double dist[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5 && i < (int)D.total(); ++i) dist[i] = D.ptr<double>()[i];
    cv::Mat Kc = K.isContinuous() ? K : K.clone();
    rmcv_pose pose;
    gc.check(rmcv_solve_pnp(ctx, &a, 1, Kc.ptr<double>(), dist, exactSize.width, exactSize.height, (float)ROI.x, (float)ROI.y,
                            nullptr, &pose), "rmcv_solve_pnp");
    cv::Mat rvec = (cv::Mat_<double>(3, 1) << pose.rvec[0], pose.rvec[1], pose.rvec[2]);
    cv::Mat tvec = (cv::Mat_<double>(3, 1) << pose.tvec[0], pose.tvec[1], pose.tvec[2]);
    return {rvec, tvec};
}
#endif

namespace gpu {
// Fused single call: image -> positives + armours without materialising contours on the host.
inline detection detect(const cv::Mat& image, const rmcv_params& prm, cv::Mat* binary = nullptr) {
    context& gc = default_context();
    rmcv_ctx* ctx = gc.get(image.cols, image.rows);
    rmcv_results res;
    const size_t bstep = binary ? (size_t)binary->step : 0;
    gc.check(rmcv_detect_batch_host(ctx, image.data, image.step, image.step * (size_t)image.rows, image.cols, image.rows, 1, &prm,
                                    binary ? binary->data : nullptr, bstep, bstep * (size_t)image.rows, &res),
             "rmcv_detect_batch_host");
    detection d;
    const rmcv_frame_info& fi = res.frames[0];
    d.contours.assign(res.contours + fi.contour_offset, res.contours + fi.contour_offset + fi.n_contours);
    std::vector<const rmcv_rotated_rect*> boxes((size_t)fi.n_positive, nullptr);
    for (const auto& c : d.contours)
        if (c.blob_index >= 0) boxes[(size_t)c.blob_index] = &c.ellipse;
    for (int k = 0; k < fi.n_positive; ++k) d.positive.push_back(to_lightblob(res.blobs[fi.blob_offset + k], boxes[(size_t)k]));
    for (int k = 0; k < fi.n_armours; ++k) {
        const rmcv_armour& o = res.armours[fi.armour_offset + k];
#if defined(RMCV_SHIM_WITH_REFERENCE)
        d.armours.push_back(armour({d.positive[o.i], d.positive[o.j]}));
#else
        armour a;
        for (int i = 0; i < 4; ++i) {
            a.icon[i].x = o.icon[i][0]; a.icon[i].y = o.icon[i][1];
            a.vertices[i].x = o.vertices[i][0]; a.vertices[i].y = o.vertices[i][1];
        }
        a.bounding_box.x = o.bounding_box[0]; a.bounding_box.y = o.bounding_box[1];
        a.bounding_box.width = o.bounding_box[2]; a.bounding_box.height = o.bounding_box[3];
        d.armours.push_back(a);
#endif
    }
    return d;
}
}  // namespace gpu
}  // namespace rm
