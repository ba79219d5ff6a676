// Shared declarations of the rmcv_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rmcv_b200.h"

namespace rmcv {

// ------------------------------------------------------------------------------------------
// Per-frame scratch layout.  One "slot" owns scratch for chunk_frames frames and one stream;
// a detect call walks its batch chunk by chunk, alternating slots so that the copies of one
// chunk overlap the kernels of the other.
// ------------------------------------------------------------------------------------------
struct CompRec {            // one 8-connected component after the blob kernel
    int32_t firstkey;       // y*W + x of the raster-first pixel, -1 when not external
    int32_t n_points;       // contour.size()
    int64_t area2;          // 2*contourArea
    int32_t bbox[4];        // x0, y0, x1, y1 (inclusive)
    int32_t status;         // RMCV_CONTOUR_*, or -1 for nested / dropped components
    int32_t fit_branch;
    float det0;
    rmcv_rotated_rect ellipse;
    rmcv_lightblob blob;
};

struct CompAcc {            // exact integer sums over the contour point multiset of one component (frame kernel:
                            // written by the warp that owns the component, read by the thread that fits it)
    long long n, sx, sy;    // contour.size(), sum x, sum y (absolute pixel coordinates)
    long long cross;        // shoelace sum = +-2*contourArea
    long long xx, xy, yy, xxx, xxy, xyy, yyy, xxxx, xxxy, xxyy, xyyy, yyyy;  // moments about (ox, oy)
    long long s_int;        // n * L1 spread about the mean
    int32_t ox, oy;         // origin of the moments = first pixel of the root run
    int32_t bbox[4];        // x0, y0, x1, y1 (inclusive)
    int32_t firstkey;       // min over the component of y*W + xs = raster-first pixel
    int32_t fitted;         // passes the size / area gate of src/objdetect.cpp:64
};

struct FrameCounters {      // device-side, one per frame in the chunk (+1 trailing entry = chunk allocators)
    int32_t n_runs;         // atomically grown by the pixel kernel's run emission (may exceed R: overflow) ...
    int32_t n_recs;         // ... together with the boundary-pixel records (one packed 64-bit atomicAdd per band)
    int32_t n_comps;
    int32_t n_holes;        // number of holes (Euler relation on the run graph)
    int32_t flags;
    int32_t n_contours, n_positive, n_negative, n_armours;
    int32_t pad[3];
};

struct Geometry {           // frame geometry + derived sizes, shared by all kernels of a call
    int W, H, WB;           // WB = 32-pixel words per row
    int R;                  // run capacity per frame
    int C;                  // component capacity per frame
    int A;                  // armour capacity per frame
    int PC;                 // boundary-pixel record capacity per frame
    int SC;                 // entries per frame of the global bucket array (>= R+2 and >= PC)
};

// Scratch slots of a ctx: chunk n (counted over the life of the ctx) uses slot n % n_slots, so up to n_slots chunks are in
// flight between the start of their pixel kernel and the end of their write-out.
constexpr int kSlots = 8;          // upper bound; a ctx uses n_slots of them (default 3, RMCV_SLOTS)

struct SlotBuffers {
    // pixel stage outputs
    uint32_t* bits;         // [CF][H][WB]   final mask, bit-packed
    uint8_t* band_flags;    // [CF][ceil(H/8)] 1 = the band's own rows hold foreground (written by the BGR band kernel, read by emit)
    // runs (emitted by the pixel kernel; rows of one band are contiguous, bands land in arrival order)
    int2* rows;             // [CF][H]       (first run, one-past-last run) of each row
    uint32_t* run_x;        // [CF][R]       xs | xe<<16
    uint16_t* run_y;        // [CF][R]
    int32_t* parent;        // [CF][R]       flattened labels (root run index), written back by the frame kernel
    int32_t* gparent;       // [CF][R+2]     background-gap forest; only used when a frame does not fit in shared memory
    int16_t* run_cid;       // [CF][R]       component id per run; same remark
    uint2* recs;            // [CF][PC]      boundary pixels {x | y<<16, 8-neighbourhood | run<<8} in emission order (emit kernel)
    uint2* recs2;           // [CF][PC]      the same records bucketed by component (label kernel)
    int32_t* comp_start;    // [CF][C+1]     first record of each component in recs2
    int32_t* sorted;        // [CF][SC]      record indices bucketed by component (and the gap join flags before that) when
                            //               they do not fit in shared memory
    CompAcc* acc;           // [CF][C]       integer contour sums per component
    int32_t* comp_root;     // [CF][C]       root run of each component
    int32_t* comp_cnt;      // [CF][C]       Euler term of each component; after the label kernel: has holes of its own
    CompRec* comps;         // [CF][C]
    FrameCounters* counters;// [CF+1]        entry CF holds the chunk's dense-output allocators (n_runs,n_comps,n_holes)
    // ordered per-frame result slots (device) before dense write-out
    rmcv_contour_info* s_contours; // [CF][C]
    rmcv_lightblob* s_blobs;       // [CF][C]
    rmcv_armour* s_armours;        // [CF][A]
    int32_t* arm_offset;           // [CF][4] dense offsets of the frame's armours, contours, blobs in the result arrays
                                   //         (for the pose kernel and the grid-wide write-out of large frames)
    // staging for host-input calls
    uint8_t* frames;        // [CF][H][W*3] (allocated lazily)
    uint8_t* masks;         // [CF][H][W]   (allocated lazily)
    size_t frames_bytes, masks_bytes;
    // the events order the ctx streams (api.cu) around the slot's scratch
    cudaEvent_t ev_pix;     // pixel kernel done  (bits written; staging frames read)
    cudaEvent_t ev_fit;     // fit kernel done (the order kernel runs on its own stream behind it)
    cudaEvent_t ev_lab;     // labelling stages done (scratch free again)
    cudaEvent_t ev_h2d;     // staging upload done
    cudaEvent_t ev_d2h;     // mask download done
};

}  // namespace rmcv

struct rmcv_ctx {
    rmcv_config cfg;
    int device;
    int sm_count;
    int max_smem_optin;
    int CF;                 // chunk frames
    rmcv::Geometry cap;     // capacities at max_width x max_height
    rmcv::SlotBuffers slot[rmcv::kSlots];
    int n_slots;
    // pinned, device-mapped result arrays of the most recent detect call (one of the two sets kept in api.cu)
    rmcv_frame_info* h_frames;      // [max_batch]
    rmcv_contour_info* h_contours;  // [max_batch][C]  (chunk-dense)
    rmcv_lightblob* h_blobs;        // [max_batch][C]
    rmcv_armour* h_armours;         // [max_batch][A]
    rmcv_pose* h_poses;             // [max_batch][A], null unless a camera is set (rmcv_set_camera)
    // last call
    int last_batch, last_W, last_H;
    bool have_results;
    // profiling
    bool profiling;
    double prof_ms[RMCV_STAGE_COUNT];
    int64_t prof_launches[RMCV_STAGE_COUNT];
    int64_t kernel_launches;
    void* extra;            // C++ side state (event sets, allocation lists), see api.cu
    char err[512];
};

namespace rmcv {

#define RMCV_CUDA(ctx, call)                                                                      \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e_));                                                     \
            return RMCV_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

inline int set_err(rmcv_ctx* ctx, int code, const char* msg) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return code;
}

// ---- launchers implemented in the .cu files -----------------------------------------------
struct EmitLaunch {
    const uint32_t* bits; int W, H, batch;
    const uint8_t* band_flags = nullptr; int flag_bh = 0, flag_bands = 0;   // per-band "has foreground" flags of the pixel kernel, if it wrote them
    int2* rows; uint32_t* run_x; uint16_t* run_y; FrameCounters* counters; int R;
    uint2* recs; int PC;
};
struct PixelLaunch {
    const uint8_t* src; size_t pitch, frame_stride;
    uint8_t* mask; size_t mask_pitch, mask_frame_stride;
    uint32_t* bits;     // [batch][H][WB]
    int W, H, batch;
    int target, lower_bound;
    int bayer_layout;   // 0 = BGR input
    uint8_t* band_flags = nullptr;      // full calls: one byte per band, 1 = its own rows hold foreground
    int* flags_bh = nullptr;            // receives the band height the flags were written for (0 = not written)
    const EmitLaunch* emit = nullptr;   // full calls: where the labelling stage's runs / records go, if the pixel kernel can emit them
    int* emit_done = nullptr;           // set to 1 when it did (the separate emit launch is then skipped)
};
cudaError_t launch_pixel_stage(const PixelLaunch& p, int sm_count, cudaStream_t st, int64_t* launches);
// Bayer fast path (bayer_strip.cu); cudaErrorNotSupported when the call does not qualify for it
cudaError_t launch_bayer_strip(const PixelLaunch& p, int sm_count, cudaStream_t st, int64_t* launches);
// BGR alternative: TMA-fed bands consumed by register-resident lanes (bgr_bandstrip.cu, RMCV_BGR_STRIP=1), same convention
cudaError_t launch_bgr_bandstrip(const PixelLaunch& p, int sm_count, cudaStream_t st, int64_t* launches);

struct CameraSetup { double K[9], dist[5], M[16]; int has_M; float w, h; };

struct FrameLaunch {
    Geometry g; int frames; SlotBuffers* sb;
    int frame_base;               // index of the chunk's first frame in the batch
    cudaStream_t st_out;          // stream of the order/write-out kernel (null = same stream)
    rmcv_frame_info* o_frames;    // device-visible pointers of the pinned result arrays
    rmcv_contour_info* o_contours;
    rmcv_lightblob* o_blobs;
    rmcv_armour* o_armours;
    rmcv_pose* o_poses;           // null unless a camera is set
    const CameraSetup* camera;
    int emit_done = 0;            // the pixel kernel already emitted the runs / records (fused pixel+emit kernel)
    int flags_bh = 0;             // band height of sb->band_flags as written by the pixel kernel of this chunk (0 = none)
    int chained = 0;              // small chunk on ONE stream: every kernel is launched as a programmatic dependent of the one before
};
// everything after the pixel stage for one chunk (five launches: emit, label, contour sums, fits, order/pairs/write-out);
// stage_done(arg, RMCV_STAGE_*, stream) is called after each launch (profiling events), may be null
cudaError_t launch_frames(const FrameLaunch& p, const rmcv_params& prm, int max_smem_optin, cudaStream_t st, int64_t* launches,
                          void (*stage_done)(void*, int, cudaStream_t), void* stage_arg);

cudaError_t launch_emit(const EmitLaunch& p, cudaStream_t st, int64_t* launches, bool chained = false);

cudaError_t launch_trace_contour(const Geometry& g, const uint32_t* bits, int x0, int y0, int32_t* d_xy, int cap,
                                 int32_t* d_n, cudaStream_t st, int64_t* launches);
cudaError_t launch_trace_all(const Geometry& g, const uint32_t* bits, const int32_t* d_starts, const int32_t* d_offsets,
                             int n_contours, int32_t* d_xy, cudaStream_t st, int64_t* launches);
cudaError_t launch_label_map(const Geometry& g, SlotBuffers* sb, int frame, int32_t* d_labels, cudaStream_t st,
                             int64_t* launches);

cudaError_t launch_filter_lightblobs(const int32_t* d_xy, const int32_t* d_off, int n, const rmcv_params& prm,
                                     rmcv_contour_info* d_infos, rmcv_lightblob* d_blobs, cudaStream_t st,
                                     int64_t* launches);
cudaError_t launch_filter_armours(const rmcv_lightblob* d_blobs, int n, const rmcv_params& prm, rmcv_armour* d_out,
                                  int cap, int32_t* d_count, cudaStream_t st, int64_t* launches);
cudaError_t launch_make_lightblobs(const rmcv_rotated_rect* d_boxes, int n, int target, rmcv_lightblob* d_out,
                                   cudaStream_t st, int64_t* launches);

struct LegacyLaunch {   // rm::MatchLightBlob / rm::FindLightBlobs / cv::minAreaRect on caller-supplied contours (legacy.cu)
    const int32_t* xy; const int32_t* off; int n_contours;
    float min_ratio, max_ratio, tilt_angle, min_area, max_area;
    int fit_ellipse;          // 1 = box from the ellipse, 0 = box from cv::minAreaRect, -1 = cv::minAreaRect only (no gates)
    const uint8_t* src; size_t pitch; int W, H;   // device BGR image for the camp vote, or null
    int32_t* hull;            // scratch, 6 ints per contour point
    int32_t* matched; rmcv_rotated_rect* boxes; int32_t* camps; rmcv_lightblob* blobs;   // [n_contours] each
};
cudaError_t launch_legacy(const LegacyLaunch& p, cudaStream_t st, int64_t* launches);
cudaError_t launch_overlap(const rmcv_lightblob* d_blobs, int n, int left, int right, int32_t* d_out, cudaStream_t st,
                           int64_t* launches);

cudaError_t launch_pnp(const rmcv_armour* d_armours, int n, const double K[9], const double dist[5], float w, float h,
                       float roi_x, float roi_y, const double* cam2world, rmcv_pose* d_out, cudaStream_t st, int64_t* launches);

// pose of every armour of a chunk, written next to the armours in the dense result array (pnp.cu)
cudaError_t launch_chunk_poses(const SlotBuffers& sb, int frames, int A, rmcv_pose* o_poses, const CameraSetup& cam,
                               cudaStream_t st, int64_t* launches);
cudaError_t launch_frontend(const uint8_t* d_src, size_t pitch, size_t frame_stride, uint8_t* d_dst, size_t dpitch,
                            size_t dframe_stride, int W, int H, int batch, int bits, int mirror, int flip, cudaStream_t st,
                            int64_t* launches);

// f2: icon crops of the armours of one frame (icon.cu)
cudaError_t launch_icons(const uint8_t* d_bgr, size_t pitch, int W, int H, rmcv_armour* d_armours, int n, int ow, int oh,
                         uint8_t* d_icons, float* d_rows, cudaStream_t st, int64_t* launches);
cudaError_t launch_svm_predict(const float* d_rows, int n, const float* d_sv, int sv_total, int var_count, const double* d_rho,
                               const int32_t* d_df_ofs, const double* d_df_alpha, const int32_t* d_df_index, const int32_t* d_class_labels,
                               int class_count, float* d_kbuf, int32_t* d_labels, cudaStream_t st, int64_t* launches);
// f3: one iteration of the tracking loop (track.cu)
cudaError_t launch_track_update(rmcv_track* d_tracks, int32_t* d_n_tracks, int cap, rmcv_track* d_backup, const rmcv_armour* d_armours,
                                const double* d_positions, const int32_t* d_identities, int n, int32_t* d_remaining, long long timestamp,
                                double freq, double q, double r, double err, int32_t* d_status, cudaStream_t st, int64_t* launches);

void upload_luts();

// Tuning knobs for experiments and debugging.  The environment is read ONCE per process — when the first
// ctx is created — and never on a launch path; -1 = not set (the launcher's own default applies).
struct Tuning {
    int slots, prio, serial, lab_streams;                            // ctx: scratch slots, stream priorities, one stream
    int small_batch, frame_rs, label_minsmem, label_small, contour_gy, emit_bh;   // labelling stages
    int pix_bh, pix_rc, pix_s, pix_nt, pix_nobulk, pix_generic;      // BGR band kernel geometry
    int bgr_strip, bandstrip_rc, bayer_generic, strip_seg, strip_minb;   // alternative pixel kernels
    int staged_out;                                                  // result write-out through device staging: -1 auto, 0 never, 1 always
    int host_chunk;                                                  // frames per chunk of the host-input entry points
    int fused_emit, wide_label;                                      // emit inside the pixel kernel; cluster kernels for large frames
    int chained, fit_in_contour, warp_fit;                           // small chunks: chained launches (0 = off); fits on the contour kernel's warps up to n frames; warp-cooperative fit there (0 = lane 0 alone)
    int chain_pad;                                                   // experiment: pad the labelling kernels' dynamic shared memory to this many bytes per CTA
};
const Tuning& tuning();


// Programmatic dependent launch (small chunks, latency mode): a chunk's six kernels sit on one stream and each is launched
// with cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs become resident while its predecessor still runs and
// the 3-4 us launch gap between dependent kernels shrinks to the wake-up out of griddepcontrol.wait.  Every kernel of the
// chain calls chain_begin() first (lets ITS successor be scheduled) and chain_wait() before it touches global memory that a
// predecessor reads or writes.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void chain_begin() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// at most this many frames: per-frame kernels take their wide variants and the chunk runs chained on one stream
inline int small_batch_limit() { return tuning().small_batch >= 0 ? tuning().small_batch : 16; }

template <class Params>
inline cudaError_t launch_chained(void (*k)(Params), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool chained, const Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = chained ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, p);
}

// Kernel begin / end stamps on the global timer, for the launch-gap anatomy of scripts/phase_stamps.py: only in a
// -DRMCV_STAMPS build (the product library has none).  One array per translation unit (no relocatable device code).
#ifdef RMCV_STAMPS
#define RMCV_GSTAMP_ARRAY(name) __device__ unsigned long long name[8][2];
__device__ __forceinline__ unsigned long long gstamp_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define RMCV_GSTAMP_BEGIN(arr, k) do { if (threadIdx.x == 0) atomicMin(&arr[k][0], gstamp_now()); } while (0)
#define RMCV_GSTAMP_END(arr, k) do { __syncthreads(); if (threadIdx.x == 0) atomicMax(&arr[k][1], gstamp_now()); } while (0)
#define RMCV_GSTAMP_GETTER(fn, arr) extern "C" int fn(unsigned long long* out, int reset) { \
    cudaError_t e = cudaMemcpyFromSymbol(out, arr, sizeof(arr)); \
    if (reset) { unsigned long long z[8][2]; for (int i = 0; i < 8; ++i) { z[i][0] = ~0ull; z[i][1] = 0ull; } e = cudaMemcpyToSymbol(arr, z, sizeof(z)); } \
    return (int)e; }
#else
#define RMCV_GSTAMP_ARRAY(name)
#define RMCV_GSTAMP_BEGIN(arr, k) do { } while (0)
#define RMCV_GSTAMP_END(arr, k) do { } while (0)
#define RMCV_GSTAMP_GETTER(fn, arr)
#endif

}  // namespace rmcv
