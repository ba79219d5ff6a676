/*
 * rmcv_b200 — C ABI of the B200-native rmcv detection hot path.
 *
 * Drop-in boundary for the three free functions the reference calls back to back at
 * executable/main.cpp:172-176 (all citations relative to the reference tree):
 *
 *     rm::extract_color      include/imgproc.h:29       src/imgproc.cpp:50-75
 *     rm::filter_lightblobs  include/objdetect.h:47-49  src/objdetect.cpp:55-87
 *     rm::filter_armours     include/objdetect.h:70-71  src/objdetect.cpp:114-166
 *     rm::lightblob / rm::armour ctors   include/core.h:89-130   src/core.cpp:9-49
 *
 * The reference has no FFI of its own (it is a static C++ library over OpenCV), so this header
 * *is* the binding surface: plain pointers and sizes, POD structs, int status codes, no
 * exceptions, no OpenCV and no torch types.  include/rmcv_gpu/rm_shim.hpp rebuilds the rm::
 * signatures on top of it; INTEGRATION.md shows the reference-side wiring.
 *
 * Threading: one rmcv_ctx per host thread and GPU.  All device work of a ctx is ordered on the
 * ctx's own streams (pixel kernels, labelling kernels, write-out and staging copies each have
 * one); entry points documented "async" return before the GPU has finished — call rmcv_sync()
 * (or a fetch/host entry point, which waits) before reading outputs.  Up to three detect calls
 * may be in flight: a ctx rotates four result sets and rmcv_fetch_results() returns the oldest
 * unfetched call, so calls n+1 .. n+3 can be enqueued before call n is fetched.
 *
 * Ordering of a detect call against the pixel stream (rmcv_stream(), the "async, pixel stream"
 * helpers): everything enqueued on that stream BEFORE the call is waited for by the call, and
 * every helper of this library enqueued AFTER it runs after it.  A call of at most 16 frames
 * runs on an internal stream of its own (so that consecutive small calls overlap), therefore
 * (a) work the CALLER puts on rmcv_stream() after such a call is not ordered after it — fetch
 * or rmcv_sync() first; (b) calls in flight at the same time must not share an output buffer
 * (d_mask): which of them writes last is then unspecified.  With rmcv_config.stream set, all of
 * a small call runs on that stream and neither caveat applies.
 *
 * There is no CPU fallback anywhere behind this ABI: without a CUDA device every entry point
 * that needs one returns RMCV_ERR_CUDA / RMCV_ERR_NO_DEVICE.
 */
#ifndef RMCV_B200_H
#define RMCV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMCV_B200_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------------------------- */
enum {
    RMCV_OK = 0,
    RMCV_ERR_INVALID_ARG = -1, /* null pointer, non-positive size, size above the ctx maxima ...  */
    RMCV_ERR_CUDA = -2,        /* a CUDA runtime call failed; rmcv_last_error() has the text      */
    RMCV_ERR_CAPACITY = -3,    /* a per-frame capacity (runs / blobs / armours) overflowed;       */
                               /* the frames concerned carry RMCV_FRAME_OVERFLOW_* flags          */
    RMCV_ERR_NO_DEVICE = -4,
    RMCV_ERR_STATE = -5        /* call order violated (e.g. fetch before detect)                  */
};

/* rm::camp, include/core.h:20-23 */
enum { RMCV_CAMP_RED = 0, RMCV_CAMP_BLUE = 1, RMCV_CAMP_GUIDELIGHT = 2, RMCV_CAMP_NEUTRAL = -1 };

/* Daheng DX_PIXEL_COLOR_FILTER, hardware/include/daheng/DxImageProc.h:54-61 (the names give the
 * colours of the first two pixels of row 0).  daheng::capture passes 4, or 2 when mirrored
 * (hardware/src/daheng.cpp:81). */
enum { RMCV_BAYER_RG = 1, RMCV_BAYER_GB = 2, RMCV_BAYER_GR = 3, RMCV_BAYER_BG = 4 };

/* per-contour verdict of rm::filter_lightblobs (src/objdetect.cpp:64,82,83) */
enum { RMCV_CONTOUR_SKIPPED = 0, RMCV_CONTOUR_POSITIVE = 1, RMCV_CONTOUR_NEGATIVE = 2 };

/* which branch cv::fitEllipseDirect took for a contour */
enum {
    RMCV_FIT_NONE = 0, RMCV_FIT_DIRECT = 1, RMCV_FIT_FALLBACK = 2,
    /* the fallback (cv::fitEllipseNoDirect) for a contour whose coordinate sum reaches 2^24: OpenCV accumulates the centre
     * as a float Point2f point by point there, so ITS centre depends on the contour's point order by up to ~0.05 px; the
     * order-free sums used here give the exactly rounded centre instead (only long contours far from the origin:
     * n * x >= 16.7 M, e.g. 4100 points at x = 4090) */
    RMCV_FIT_FALLBACK_LONG = 3
};

enum {
    RMCV_FRAME_OVERFLOW_RUNS = 1,
    RMCV_FRAME_OVERFLOW_BLOBS = 2,
    RMCV_FRAME_OVERFLOW_ARMOURS = 4,
    RMCV_FRAME_OVERFLOW_POINTS = 8, /* boundary pixels above 4 * max_runs_per_frame */
    RMCV_FRAME_OVERFLOW_MOMENTS = 16 /* a fitted contour too large for the exact 64-bit moment sums (n * extent^4 > 2^62; only
                                        reachable with area_max far above the reference's 99999): its ellipse is not reliable */
};

/* ---- PODs --------------------------------------------------------------------------------- */

/* cv::RotatedRect as the reference consumes it (centre, size, angle in degrees), float32. */
typedef struct rmcv_rotated_rect {
    float cx, cy, w, h, angle;
} rmcv_rotated_rect;

/* rm::lightblob public fields, include/core.h:92-96.  56 bytes. */
typedef struct rmcv_lightblob {
    float angle;           /* 90 = upright                                                      */
    int32_t target;        /* rm::camp                                                          */
    float center[2];       /* x, y                                                              */
    float vertices[4][2];  /* left-down, left-up, right-up, right-down (src/core.cpp:276-280)   */
    float size[2];         /* width = short side, height = long side (src/core.cpp:18)          */
} rmcv_lightblob;

/* rm::armour public geometry, include/core.h:110-112, plus the pair that produced it and the
 * gate quantities of rm::filter_armours (the reference keeps no score).  112 bytes. */
typedef struct rmcv_armour {
    float icon[4][2];
    float vertices[4][2];
    float bounding_box[4]; /* x, y, width, height (cv::Rect2f)                                   */
    int32_t i, j;          /* indices into the frame's positive light-blob list, i < j           */
    float gates[6];        /* |Δangle|, shear_i, shear_j, min/max height, |Δcy|, |Δcx|           */
} rmcv_armour;

/* One external contour (= one 8-connected component not nested in a hole), in the order
 * cv::findContours returns them (reverse raster order of the first pixel).  72 bytes. */
typedef struct rmcv_contour_info {
    int32_t first_x, first_y;  /* raster-first pixel = contour[0]                                */
    int32_t n_points;          /* contour.size() with CHAIN_APPROX_NONE (revisits counted)       */
    int32_t status;            /* RMCV_CONTOUR_*                                                 */
    int64_t area2;             /* 2 * cv::contourArea, exact                                     */
    int32_t bbox[4];           /* x, y, width, height of the component                           */
    rmcv_rotated_rect ellipse; /* cv::fitEllipseDirect; zero when status == SKIPPED              */
    int32_t fit_branch;        /* RMCV_FIT_*                                                     */
    float det0;                /* |det M| of the first direct-fit attempt (diagnostic)           */
    int32_t blob_index;        /* index into the frame's positive list, or -1                    */
} rmcv_contour_info;

/* Per-frame counts and offsets into the dense result arrays of a batch.  32 bytes. */
typedef struct rmcv_frame_info {
    int32_t n_contours, n_positive, n_negative, n_armours;
    int32_t contour_offset, blob_offset, armour_offset;
    int32_t flags; /* RMCV_FRAME_* */
} rmcv_frame_info;

/* Parameters of the path; defaults = the literals at executable/main.cpp:172-176. */
typedef struct rmcv_params {
    int32_t target;      /* camp passed to extract_color / enemy passed to the filters           */
    int32_t lower_bound; /* inRange lower bound                                                  */
    float tilt_max;
    float ratio_min, ratio_max;
    double area_min, area_max;
    float angle_difference_max;
    float shear_max;
    float lenght_ratio_max; /* sic (include/objdetect.h:68); acts as a minimum, SURVEY B.2        */
} rmcv_params;

typedef struct rmcv_config {
    int32_t device;                /* CUDA ordinal                                               */
    int32_t max_width, max_height; /* largest frame this ctx will see                            */
    int32_t max_batch;             /* largest batch of one detect call                           */
    int32_t chunk_frames;          /* frames per internal pipeline step; 0 = default             */
    int32_t max_runs_per_frame;    /* 0 = default max(16384, W*H/32)                             */
    int32_t max_blobs_per_frame;   /* components per frame; 0 = default 1024                     */
    int32_t max_armours_per_frame; /* 0 = default 2048                                           */
    int32_t flags;                 /* reserved, 0                                                */
    void* stream;                  /* cudaStream_t to run the pixel kernels on, or NULL           */
} rmcv_config;

/* Pose of one armour: cv::solvePnP(SOLVEPNP_IPPE_SQUARE) on armour.vertices against the canonical square of the given
 * size (src/mobility.cpp:166-190), and tvec moved by the caller's 4x4 camera -> world transform (executable/main.cpp:
 * 186-192; position == tvec when no transform is given).  88 bytes. */
typedef struct rmcv_pose {
    double rvec[3];       /* Rodrigues rotation vector                                         */
    double tvec[3];       /* translation in the units of exact_w / exact_h                     */
    double position[3];   /* cam2world * [tvec; 1]                                             */
    double reproj_err;    /* sum of squared reprojection errors, normalised image coordinates  */
    int32_t ok;           /* 0: the four image points are collinear                            */
    int32_t pad;
} rmcv_pose;

/* View of the results of one detect call.  Pointers are ctx-owned pinned host memory, valid
 * until the fourth next detect call on the ctx (four result sets rotate, so up to three calls can
 * be in flight behind the one being fetched).  Dense arrays are indexed through frames[f].*_offset. */
typedef struct rmcv_results {
    int32_t batch;
    int32_t total_contours, total_blobs, total_armours;
    const rmcv_frame_info* frames;
    const rmcv_contour_info* contours;
    const rmcv_lightblob* blobs;
    const rmcv_armour* armours;
    const rmcv_pose* poses;   /* poses[k] belongs to armours[k]; NULL unless rmcv_set_camera() was called (f1) */
} rmcv_results;

typedef struct rmcv_ctx rmcv_ctx;

/* ---- lifetime ----------------------------------------------------------------------------- */
int rmcv_abi_version(void);
const char* rmcv_status_string(int status);
void rmcv_default_params(rmcv_params* p);                 /* main.cpp:172-176 literals */
void rmcv_default_config(rmcv_config* c);
int rmcv_device_count(int* count);
int rmcv_ctx_create(const rmcv_config* cfg, rmcv_ctx** out);
int rmcv_ctx_destroy(rmcv_ctx* ctx);
const char* rmcv_last_error(const rmcv_ctx* ctx);          /* text of the last failure on this ctx */
int rmcv_chunk_frames(const rmcv_ctx* ctx);                /* frames per internal pipeline step actually in use */

/* ---- memory + stream helpers (so C / ctypes hosts need no CUDA toolkit) ---------------------- */
int rmcv_device_alloc(rmcv_ctx* ctx, size_t bytes, void** dptr);
int rmcv_device_free(rmcv_ctx* ctx, void* dptr);
int rmcv_host_alloc(rmcv_ctx* ctx, size_t bytes, void** hptr);   /* pinned */
int rmcv_host_free(rmcv_ctx* ctx, void* hptr);
int rmcv_memcpy_h2d(rmcv_ctx* ctx, void* dst, const void* src, size_t bytes); /* async, pixel stream */
int rmcv_memcpy_d2h(rmcv_ctx* ctx, void* dst, const void* src, size_t bytes); /* async, pixel stream */
int rmcv_memset_d(rmcv_ctx* ctx, void* dst, int value, size_t bytes);         /* async, pixel stream */
int rmcv_sync(rmcv_ctx* ctx);                                                 /* all ctx streams */
void* rmcv_stream(rmcv_ctx* ctx);                                             /* the pixel-kernel cudaStream_t */

/* ---- a1: pixel stage of rm::extract_color (src/imgproc.cpp:52-69) --------------------------- */
/* BGR interleaved u8 frames -> binary mask {0,255}.  Device pointers, async.
 * frame f starts at d_bgr + f*frame_stride; rows are `pitch` bytes apart (cv::Mat::step).
 * d_mask may be NULL when only the bit-packed mask (kept inside the ctx) is wanted. */
int rmcv_extract_color_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride,
                             int width, int height, int batch, int target, int lower_bound,
                             uint8_t* d_mask, size_t mask_pitch, size_t mask_frame_stride);

/* a0+a1: 8-bit Bayer mosaic -> (bilinear B,R) -> same pixel stage.  Stands in for
 * DxRaw8toRGB24(RAW2RGB_NEIGHBOUR) of hardware/src/daheng.cpp:143-148 followed by extract_color. */
int rmcv_bayer_extract_color_batch(rmcv_ctx* ctx, const uint8_t* d_raw, size_t pitch, size_t frame_stride,
                                   int width, int height, int batch, int bayer_layout,
                                   int target, int lower_bound,
                                   uint8_t* d_mask, size_t mask_pitch, size_t mask_frame_stride);

/* ---- a1..a5: the whole path, frames resident on the device ---------------------------------- */
/* Async.  Results are read with rmcv_fetch_results().  d_mask may be NULL. */
int rmcv_detect_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride,
                      int width, int height, int batch, const rmcv_params* params,
                      uint8_t* d_mask, size_t mask_pitch, size_t mask_frame_stride);

int rmcv_bayer_detect_batch(rmcv_ctx* ctx, const uint8_t* d_raw, size_t pitch, size_t frame_stride,
                            int width, int height, int batch, int bayer_layout, const rmcv_params* params,
                            uint8_t* d_mask, size_t mask_pitch, size_t mask_frame_stride);

/* The whole path from HOST frames (what a cv::Mat caller has): host->device copies, kernels and
 * result read-back are pipelined chunk by chunk on the ctx's streams.  h_mask may be NULL.
 * Synchronous: results are ready on return. */
int rmcv_detect_batch_host(rmcv_ctx* ctx, const uint8_t* h_bgr, size_t pitch, size_t frame_stride,
                           int width, int height, int batch, const rmcv_params* params,
                           uint8_t* h_mask, size_t mask_pitch, size_t mask_frame_stride,
                           rmcv_results* out);

/* The same from raw 8-bit Bayer mosaics in HOST memory — what the camera actually delivers (hardware/src/daheng.cpp:
 * 74-89: GXGetImage hands back the raw frame; DxRaw8toRGB24 at :136-151 is what this path replaces).  1 B/px crosses PCIe
 * instead of 3.  Synchronous. */
int rmcv_bayer_detect_batch_host(rmcv_ctx* ctx, const uint8_t* h_raw, size_t pitch, size_t frame_stride,
                                 int width, int height, int batch, int bayer_layout, const rmcv_params* params,
                                 uint8_t* h_mask, size_t mask_pitch, size_t mask_frame_stride,
                                 rmcv_results* out);

/* Waits for the oldest detect call whose results have not been fetched yet and exposes them
 * (with one call in flight: the last call).  Without an unfetched call it re-exposes the
 * results fetched last.  An unfetched call is dropped when a fifth call is enqueued. */
int rmcv_fetch_results(rmcv_ctx* ctx, rmcv_results* out);

/* The on-demand getters below refer to the MOST RECENT detect call and wait for it.
 * Ordered contour of external contour `contour_index` of frame `frame` of the last detect call
 * (Suzuki border following, identical point sequence to cv::findContours).  xy receives up to
 * `cap` (x,y) pairs; *n_points receives the full length.  Host pointer, synchronous. */
int rmcv_get_contour(rmcv_ctx* ctx, int frame, int contour_index, int32_t* xy, int cap, int* n_points);

/* All external contours of frame `frame` in cv::findContours order, traced in one launch (one thread per contour).
 * offsets[k]..offsets[k+1] delimit contour k inside xy ((x,y) int32 pairs); offsets needs n_contours+1 entries.
 * *n_contours / *n_points always receive the full counts; RMCV_ERR_CAPACITY if a cap is too small. */
int rmcv_get_contours(rmcv_ctx* ctx, int frame, int32_t* xy, int cap_points, int32_t* offsets, int cap_contours,
                      int* n_contours, int* n_points);

/* int32 label map of frame `frame` of the last detect call: index of the external contour that
 * owns each pixel, -1 for background and nested components.  Host pointer, synchronous. */
int rmcv_get_label_map(rmcv_ctx* ctx, int frame, int32_t* labels, size_t pitch_elems);

/* Bit-packed mask (1 bit per pixel, LSB = lowest x, rows padded to 32-pixel words) of frame
 * `frame` of the last extract/detect call.  Host pointer, synchronous. */
int rmcv_get_bitmask(rmcv_ctx* ctx, int frame, uint32_t* words, int words_per_row);

/* ---- one batch across the GPUs of one box ------------------------------------ */
/* The reference's only caller (executable/main.cpp:163-209) processes frames one after the other on one thread and keeps no
 * cross-frame state, so a batch is partitioned BY FRAME: contiguous slices of batch / n_devices frames (+1 for the first
 * batch % n_devices slices), slice g on
 * device devices[g], one persistent host thread + one rmcv_ctx per device, no collective, results concatenated on the
 * host in frame order (identical, byte for byte, to one device running the whole batch).
 * cfg->max_batch is the largest WHOLE batch; cfg->device is ignored; devices == NULL with n_devices == 0 means every
 * visible device.  For full upload bandwidth the frames should sit in pinned host memory (rmcv_host_alloc allocates it
 * portable, i.e. usable from every device). */
typedef struct rmcv_multi rmcv_multi;
int rmcv_multi_create(const rmcv_config* cfg, const int* devices, int n_devices, rmcv_multi** out);
int rmcv_multi_destroy(rmcv_multi* m);
int rmcv_multi_device_count(const rmcv_multi* m);
const char* rmcv_multi_last_error(const rmcv_multi* m);
/* frames [*first, *first + *count) of a batch belong to device index g */
void rmcv_multi_slice(int batch, int n_devices, int g, int* first, int* count);
/* Synchronous.  Results (frame order, offsets into the merged dense arrays) stay valid until the next call on `m`. */
int rmcv_multi_detect_batch_host(rmcv_multi* m, const uint8_t* h_bgr, size_t pitch, size_t frame_stride,
                                 int width, int height, int batch, const rmcv_params* params,
                                 uint8_t* h_mask, size_t mask_pitch, size_t mask_frame_stride, rmcv_results* out);
int rmcv_multi_bayer_detect_batch_host(rmcv_multi* m, const uint8_t* h_raw, size_t pitch, size_t frame_stride,
                                       int width, int height, int batch, int bayer_layout, const rmcv_params* params,
                                       uint8_t* h_mask, size_t mask_pitch, size_t mask_frame_stride, rmcv_results* out);

/* ---- a2/a3 standalone: rm::filter_lightblobs on caller-supplied contours --------------------- */
/* xy = concatenated (x,y) int32 pairs, offsets[n_contours+1] in points.  Host pointers,
 * synchronous.  infos[n_contours] receives the verdicts in input order; blobs receives the
 * positives in input order (up to blob_cap). */
int rmcv_filter_lightblobs(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours,
                           const rmcv_params* params, rmcv_contour_info* infos,
                           rmcv_lightblob* blobs, int blob_cap, int* n_blobs);

/* ---- a4/a5 standalone: rm::filter_armours on caller-supplied light blobs ---------------------- */
int rmcv_filter_armours(rmcv_ctx* ctx, const rmcv_lightblob* blobs, int n_blobs, const rmcv_params* params,
                        rmcv_armour* armours, int armour_cap, int* n_armours);

/* a3 standalone: rm::lightblob ctor from a RotatedRect (src/core.cpp:9-19) */
int rmcv_make_lightblobs(rmcv_ctx* ctx, const rmcv_rotated_rect* boxes, int n, int target, rmcv_lightblob* out);

/* ---- a6 / a7 legacy rows, standalone ---------------------------------------------------------- */
/* rm::MatchLightBlob (src/objdetect.cpp:9-28) on every caller-supplied contour: size/area gate, cv::fitEllipseDirect,
 * box = the ellipse (fit_ellipse != 0) or cv::minAreaRect, ratio gate on the box, tilt gate on the ellipse.
 * matched[k] = 1/0, boxes[k] = lightBlobBox (valid where matched).  Host pointers, synchronous. */
int rmcv_match_lightblobs(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours,
                          float min_ratio, float max_ratio, float tilt_angle, float min_area, float max_area,
                          int fit_ellipse, int32_t* matched, rmcv_rotated_rect* boxes);

/* rm::FindLightBlobs (src/objdetect.cpp:30-53): MatchLightBlob + the camp vote from the mean colour of the contour's
 * bounding rect in the 3-channel source image (G > B && G > R: guide light; else B > R ? blue : red) + the rm::lightblob
 * ctor.  blobs receives the matches in input order (up to blob_cap); *n_blobs the full count. */
int rmcv_find_lightblobs_legacy(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours,
                                float min_ratio, float max_ratio, float tilt_angle, float min_area, float max_area,
                                const uint8_t* h_source_bgr, size_t pitch, int width, int height, int fit_ellipse,
                                rmcv_lightblob* blobs, int blob_cap, int* n_blobs);

/* cv::minAreaRect of every contour (convex hull + minimum-area enclosing rectangle; OpenCV 4.13 convention:
 * angle in [-90, 0), width = extent along that direction).  Used by MatchLightBlob(fitEllipse = false). */
int rmcv_min_area_rects(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours, rmcv_rotated_rect* boxes);

/* rm::LightBlobOverlap (src/objdetect.cpp:89-112).  *overlap = 1/0.  The reference's bound check admits
 * right == n (one past the end, undefined behaviour in C++); here right >= n yields 0. */
int rmcv_lightblob_overlap(rmcv_ctx* ctx, const rmcv_lightblob* blobs, int n_blobs, int left, int right, int* overlap);

/* ---- f4 (next row): camera front-end variants ------------------------------------------------- */
/* hardware/src/daheng.cpp:91-187 ahead of the demosaic: 10/12-bit samples in 16-bit containers are cut to bits 2..9 /
 * 4..11 (DxRaw16toRaw8 DX_BIT_2_9 / DX_BIT_4_11), DxImageMirror(HORIZONTAL_MIRROR) and the flip argument of
 * DxRaw8toRGB24 are folded into the index.  Device pointers, async (pixel stream).  bits: 8, 10 or 12; pitches and frame
 * strides in bytes.  Run rmcv_bayer_* on d_raw8 with rmcv_frontend_layout(layout, width, height, 0, flip): daheng::capture
 * already passes the post-mirror filter (2 instead of 4, daheng.cpp:81), so only the flip changes the layout. */
int rmcv_raw_frontend_batch(rmcv_ctx* ctx, const void* d_raw, size_t pitch, size_t frame_stride, int width, int height,
                            int batch, int bits, int mirror, int flip, uint8_t* d_raw8, size_t out_pitch,
                            size_t out_frame_stride);
/* Bayer layout (RMCV_BAYER_*) of a mosaic after a horizontal mirror and/or a vertical flip of a width x height frame. */
int rmcv_frontend_layout(int layout, int width, int height, int mirror, int flip);

/* ---- f1 (next row): rm::solve_PnP per armour -------------------------------------------------- */
/* Fused variant: with a camera set, every detect call also solves the pose of every armour it finds (one more small
 * kernel behind the write-out, executable/main.cpp:183-192 without the host round trip) and rmcv_results.poses is
 * filled.  exact_w must equal exact_h; dist_coeffs and cam2world may be NULL.  rmcv_clear_camera() turns it off. */
int rmcv_set_camera(rmcv_ctx* ctx, const double camera_matrix[9], const double dist_coeffs[5], float exact_w, float exact_h,
                    const double* cam2world);
int rmcv_clear_camera(rmcv_ctx* ctx);

/* camera_matrix: 3x3 row-major; dist_coeffs: k1, k2, p1, p2, k3 (NULL = none); exact_w must equal exact_h
 * (IPPE_SQUARE); roi_x/roi_y are added to the image points (the reference's ROI offset); cam2world: 4x4 row-major
 * or NULL.  Host pointers, synchronous. */
int rmcv_solve_pnp(rmcv_ctx* ctx, const rmcv_armour* armours, int n_armours, const double camera_matrix[9],
                   const double dist_coeffs[5], float exact_w, float exact_h, float roi_x, float roi_y,
                   const double* cam2world, rmcv_pose* poses);

/* ---- f2 (next row, first half): the icon crop that feeds the SVM ------------------------------------------------ */
/* rm::affine_correction (src/imgproc.cpp:9-35) + rm::utils::flatten_image (src/core.cpp:202-216) for every armour of a
 * frame that is resident in device memory (d_bgr: height x width x 3, interleaved, row pitch in bytes): the icon vertices
 * are clamped into the frame IN PLACE like the reference does (armours[k].icon is updated), the bounding box of the rounded
 * vertices is warped by cv::getAffineTransform / cv::warpAffine (bilinear, constant border 0) and resized (bilinear) to
 * out_w x out_h.  icons: n x out_h x out_w x 3 bytes (bit-exact against OpenCV); rows: the same values as float32, one
 * row of out_w * out_h * 3 per armour (the SVM's input, executable/main.cpp:180-181), or NULL.  armours / icons / rows are
 * host pointers; synchronous.  The SVM itself is not part of the library (the reference ships no svm.xml). */
int rmcv_icon_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, int width, int height, rmcv_armour* armours,
                    int n_armours, int out_w, int out_h, uint8_t* icons, float* rows);

/* ---- f2 (next row, second half): cv::ml::SVM::predict for the model the reference trains --------------------------- */
/* A trained C_SVC model with the LINEAR kernel (executable/svm/optimizer.cpp:18-21), as cv::ml::SVM holds it after
 * training / SVM::load: the (compressed) support vectors, and per one-vs-one decision function (class i against class j,
 * i < j, in that order) its rho and its (alpha, support-vector index) pairs.  From Python:
 * svm.getSupportVectors(), svm.getDecisionFunction(k) -> (rho, alpha, svidx).  Host pointers. */
typedef struct rmcv_svm_model {
    int32_t var_count;               /* features per sample (out_w * out_h * 3 = 1200 for the 20 x 20 icons)          */
    int32_t class_count;             /* <= 32                                                                         */
    int32_t sv_total;
    const float* support_vectors;    /* sv_total x var_count                                                          */
    const int32_t* class_labels;     /* class_count labels, ascending (the sorted distinct training responses)        */
    const double* rho;               /* class_count * (class_count - 1) / 2                                           */
    const int32_t* df_ofs;           /* decision function k uses entries df_ofs[k] .. df_ofs[k+1]-1 of the next two   */
    const double* df_alpha;
    const int32_t* df_index;
} rmcv_svm_model;
/* labels[s] = (int)svm->predict(rows[s]) for n rows of var_count floats (OpenCV's summation order and one-vs-one vote,
 * modules/ml/src/svm.cpp).  Host pointers, synchronous. */
int rmcv_svm_predict(rmcv_ctx* ctx, const rmcv_svm_model* model, const float* rows, int n, int32_t* labels);
/* executable/main.cpp:178-181 in one call: icon crop (out_w x out_h, out_w * out_h * 3 == model->var_count) and SVM
 * identity of every armour of a device-resident frame, without the host round trip between the two; armours[k].icon is
 * clamped in place as in rmcv_icon_batch. */
int rmcv_identify_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, int width, int height, rmcv_armour* armours,
                        int n_armours, int out_w, int out_h, const rmcv_svm_model* model, int32_t* identities);

/* ---- f3 (next row): armour tracking — IoU association + 6-state Kalman filter + identity vote ------------------ */
/* One tracked armour: the public tracking fields of rm::armour (include/core.h:103-122) and the state of its
 * cv::KalmanFilter(6, 6, 0, CV_64F) observer (src/core.cpp:21,51-69).  As in the reference, bbox / position / identity
 * are those of the armour that opened the track (rm::armour::update never refreshes them), lost_count is never
 * cleared, and the first correct() runs against a zero errorCovPre (so it leaves the state at zero). */
#define RMCV_TRACK_HIST 8
typedef struct rmcv_track {
    float bbox[4];                         /* bounding_box x, y, width, height                                   */
    double position[3];
    int64_t timestamp;                     /* ticks of the last update                                           */
    int32_t lost_count;
    int32_t identity;
    int32_t initialized;
    int32_t n_hist;                        /* identity_history (std::map<int,int>), sorted by identity           */
    int32_t hist_id[RMCV_TRACK_HIST], hist_count[RMCV_TRACK_HIST];
    double state_pre[6], state_post[6];    /* x, y, z, vx, vy, vz                                                */
    double cov_pre[36], cov_post[36];      /* errorCovPre / errorCovPost, row-major                              */
    double meas[6];                        /* measurement                                                        */
    double q, r;                           /* processNoiseCov / measurementNoiseCov diagonal                     */
} rmcv_track;

typedef struct rmcv_tracker rmcv_tracker;  /* device-resident track list of one camera stream */
int rmcv_tracker_create(rmcv_ctx* ctx, int capacity, rmcv_tracker** out);
int rmcv_tracker_destroy(rmcv_ctx* ctx, rmcv_tracker* tracker);
int rmcv_tracker_reset(rmcv_ctx* ctx, rmcv_tracker* tracker);
/* One iteration of the reference's tracking loop (executable/main.cpp:57-88) for the armours of one frame, in the
 * reference's order: for every track, rm::armour::max_IoU over the frame's remaining armours (src/core.cpp:145-161);
 * IoU > 0.5: rm::armour::update(observation) (:71-106) and the armour leaves the list; otherwise lost_count++ > 25
 * erases the track (and, like the reference's erase-in-a-for-loop, skips the one behind it), else update(timestamp)
 * (:108-121); the armours left over open new tracks (rm::armour::reset(process_noise, measurement_noise, error),
 * main.cpp:195).  An empty frame changes nothing; with no tracks the frame's armours become the tracks.
 * positions: n x 3 world positions (rmcv_pose.position); identities: n class ids or NULL (-1).  A track list or an
 * identity history that outgrows its capacity returns RMCV_ERR_CAPACITY (the list is then left as it was before the
 * call).  Host pointers, synchronous; the track list itself stays on the device. */
int rmcv_tracker_update(rmcv_ctx* ctx, rmcv_tracker* tracker, const rmcv_armour* armours, const double* positions,
                        const int32_t* identities, int n_armours, int64_t timestamp, double tick_frequency,
                        double process_noise, double measurement_noise, double error);
/* Copies the track list to the host (up to cap tracks; *n_tracks = tracks alive). */
int rmcv_tracker_read(rmcv_ctx* ctx, rmcv_tracker* tracker, rmcv_track* tracks, int cap, int* n_tracks);
/* rm::armour::identity_max (src/core.cpp:123-143): softmax over the identity history of one track. */
int rmcv_track_identity_max(const rmcv_track* track, int32_t* identity, double* probability);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* CUDA-event timing of the stages of detect/extract calls (on the streams that run them). */
enum {
    RMCV_STAGE_PIXEL = 0,   /* fused diff/threshold/close kernel -> byte mask + bit mask                       */
    RMCV_STAGE_EMIT = 1,    /* runs + boundary-pixel records from the bit mask                                 */
    RMCV_STAGE_LABEL = 2,   /* connected components (+ holes) of the runs                                      */
    RMCV_STAGE_CONTOUR = 3, /* per-component contour statistics (exact integer sums)                           */
    RMCV_STAGE_FIT = 4,     /* ellipse fits, light-blob gates                                                  */
    RMCV_STAGE_ORDER = 5,   /* cv::findContours order, armour pair gates, dense write-out                      */
    RMCV_STAGE_COUNT = 6
};
int rmcv_profile_enable(rmcv_ctx* ctx, int on);
/* accumulated milliseconds and launch counts per stage since the last reset */
int rmcv_profile_read(rmcv_ctx* ctx, double ms[RMCV_STAGE_COUNT], int64_t launches[RMCV_STAGE_COUNT], int reset);
/* Device-side stopwatch over ALL streams of the ctx: start records an event on every ctx stream, stop records again,
 * synchronises and returns last-stop minus first-start in milliseconds (CUDA events on the launching streams). */
int rmcv_timer_start(rmcv_ctx* ctx);
int rmcv_timer_stop(rmcv_ctx* ctx, double* ms);
/* number of kernels this library launched on the ctx since creation */
int64_t rmcv_kernel_launches(const rmcv_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RMCV_B200_H */
