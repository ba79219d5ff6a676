// C ABI of rmcv_b200 (include/rmcv_b200.h): context, memory helpers, the batched entry points and the
// chunked multi-slot pipeline that overlaps copies and kernels.  No CPU fallback: every compute entry point
// launches the CUDA kernels in pixel.cu / ccl.cu / blob.cu / armour.cu or fails with an error code.
#include <math.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "common.cuh"

using namespace rmcv;

namespace rmcv {
const Tuning& tuning();
}

namespace {

struct ProfSet {      // events of one chunk: pixel begin/end on the pixel stream, one mark per labelling stage on its stream
    cudaEvent_t pix[2];
    cudaEvent_t lab[RMCV_STAGE_COUNT];  // lab[0] = start of the labelling stages, lab[s] = end of stage s (s >= 1)
    bool rec, full;
};

struct ResultSet {  // pinned, device-mapped result arrays of one detect call (two sets: two calls may be in flight)
    rmcv_frame_info* frames = nullptr;      // [max_batch]
    rmcv_contour_info* contours = nullptr;  // [max_batch][C]  (chunk-dense)
    rmcv_lightblob* blobs = nullptr;        // [max_batch][C]
    rmcv_armour* armours = nullptr;         // [max_batch][A]
    rmcv_pose* poses = nullptr;             // [max_batch][A], allocated by the first rmcv_set_camera
    // Device-side mirrors for large batches: the write-out kernel fills these (dense, same layout) and the records travel
    // to the pinned arrays above as a few DMA copies sized from the per-frame counts once the call is fetched, instead of
    // as posted stores over PCIe from inside the kernel (which stall the kernel and, with eight GPUs behind one host
    // bridge, each other).  Small batches keep the zero-copy stores: lowest latency.
    rmcv_frame_info* d_frames = nullptr; rmcv_contour_info* d_contours = nullptr; rmcv_lightblob* d_blobs = nullptr;
    rmcv_armour* d_armours = nullptr;
    bool staged = false, materialised = true;
    int cf = 0;                             // frames per chunk of the call (dense regions are per chunk)
    bool with_poses = false;                // the call that filled this set had a camera
    int batch = 0;
    bool pending = false;                   // enqueued, not fetched yet
    long long call_id = -1;
    cudaEvent_t done[2] = {nullptr, nullptr};  // recorded on the write-out stream and on the pixel stream
};

// Result sets of a ctx: up to kResultSets - 1 detect calls can be in flight behind the one being fetched (short calls — a few
// frames, or one slice of a batch split over several GPUs — need more than one call ahead to keep the GPU busy).
constexpr int kResultSets = 4;

struct CtxExtra {  // C++ side of the ctx (kept out of the POD part)
    ResultSet rs[kResultSets];
    long long n_calls = 0;
    int last_fetched = -1;
    std::vector<ProfSet> prof;
    size_t prof_used = 0;
    std::vector<void*> dev_allocs, host_allocs;
    // scratch for the standalone entry points
    void* tmp_dev = nullptr; size_t tmp_dev_bytes = 0;
    void* tmp_host = nullptr; size_t tmp_host_bytes = 0;
    int last_nchunks = 0;
    int last_cf = 0;     // frames per chunk of the last call (host-input calls use shorter chunks)
    int last_kind = 0;  // 0 none, 1 extract, 2 detect
    cudaEvent_t t_start[4 + kSlots] = {}, t_stop[4 + kSlots] = {};
    // Streams of the ctx.  The pixel kernels of consecutive chunks run back to back on `pix`; the labelling kernels of
    // chunk i run on the high-priority stream `lab` behind an event, so that they overlap the pixel kernel of chunk
    // i+1; host<->device staging copies have their own streams (both copy engines stay busy).
    cudaStream_t pix = nullptr, lab = nullptr, out = nullptr, h2d = nullptr, d2h = nullptr;
    cudaStream_t labs[kSlots] = {};   // labelling stream of each slot (labs[0] == lab)
    long long chunk_counter = 0;      // chunks enqueued over the life of the ctx
    int last_first_slot = 0;          // slot of chunk 0 of the last call
    cudaStream_t chain = nullptr;     // stream of the current call's chained chunk (small chunks), else null
    cudaEvent_t ev_order = nullptr;   // "everything on the pixel stream so far", waited for by a chained chunk
    bool slot_chained[kSlots] = {};   // the slot's last chunk ran chained on its own stream (its ev_lab marks the chain's end)
    unsigned call_slots = 0;          // slots whose streams carried chained chunks of the call being enqueued
    bool own_pix = false;
    // where the write-out kernels of the current call put their records (pinned host arrays or the device mirrors)
    rmcv_frame_info* o_frames = nullptr; rmcv_contour_info* o_contours = nullptr; rmcv_lightblob* o_blobs = nullptr;
    rmcv_armour* o_armours = nullptr;
    CameraSetup camera;      // f1 fused: pose of every armour behind the write-out kernel
    bool have_camera = false;
};

CtxExtra* extra(rmcv_ctx* c) { return static_cast<CtxExtra*>(c->extra); }

int env_or(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
Tuning read_tuning() {
    Tuning t;
    t.slots = env_or("RMCV_SLOTS", -1); t.prio = env_or("RMCV_PRIO", -1); t.serial = env_or("RMCV_SERIAL", -1) > 0 ? 1 : 0;
    t.lab_streams = env_or("RMCV_LAB_STREAMS", -1);
    t.small_batch = env_or("RMCV_SMALL_BATCH", -1); t.frame_rs = env_or("RMCV_FRAME_RS", -1);
    t.label_minsmem = env_or("RMCV_LABEL_MINSMEM", -1); t.label_small = env_or("RMCV_LABEL_SMALL", -1) > 0 ? 1 : 0;
    t.contour_gy = env_or("RMCV_CONTOUR_GY", -1); t.emit_bh = env_or("RMCV_EMIT_BH", -1);
    t.pix_bh = env_or("RMCV_PIX_BH", -1); t.pix_rc = env_or("RMCV_PIX_RC", -1); t.pix_s = env_or("RMCV_PIX_S", -1);
    t.pix_nt = env_or("RMCV_PIX_NT", -1); t.pix_nobulk = env_or("RMCV_PIX_NOBULK", 0); t.pix_generic = env_or("RMCV_PIX_GENERIC", 0);
    t.bgr_strip = env_or("RMCV_BGR_STRIP", 0); t.bandstrip_rc = env_or("RMCV_BANDSTRIP_RC", -1);
    t.bayer_generic = env_or("RMCV_BAYER_GENERIC", 0); t.strip_seg = env_or("RMCV_STRIP_SEG", -1); t.strip_minb = env_or("RMCV_STRIP_MINB", -1);
    t.host_chunk = env_or("RMCV_HOST_CHUNK", -1); t.staged_out = env_or("RMCV_STAGED_OUT", -1);
    t.fused_emit = env_or("RMCV_FUSED_EMIT", -1); t.wide_label = env_or("RMCV_WIDE_LABEL", -1); t.chained = env_or("RMCV_CHAINED", -1); t.fit_in_contour = env_or("RMCV_FIT_IN_CONTOUR", -1); t.warp_fit = env_or("RMCV_WARP_FIT", -1);
    t.chain_pad = env_or("RMCV_CHAIN_PAD", -1);
    return t;
}

template <class T>
cudaError_t dalloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)); }

int ensure_tmp(rmcv_ctx* ctx, size_t dev_bytes, size_t host_bytes) {
    CtxExtra* ex = extra(ctx);
    if (dev_bytes > ex->tmp_dev_bytes) {
        if (ex->tmp_dev) cudaFree(ex->tmp_dev);
        ex->tmp_dev = nullptr; ex->tmp_dev_bytes = 0;
        RMCV_CUDA(ctx, cudaMalloc(&ex->tmp_dev, dev_bytes));
        ex->tmp_dev_bytes = dev_bytes;
    }
    if (host_bytes > ex->tmp_host_bytes) {
        if (ex->tmp_host) cudaFreeHost(ex->tmp_host);
        ex->tmp_host = nullptr; ex->tmp_host_bytes = 0;
        RMCV_CUDA(ctx, cudaMallocHost(&ex->tmp_host, host_bytes));
        ex->tmp_host_bytes = host_bytes;
    }
    return RMCV_OK;
}

int alloc_slot(rmcv_ctx* ctx, SlotBuffers& sb, bool first) {
    const Geometry& g = ctx->cap;
    const size_t CF = ctx->CF, H = g.H, WB = g.WB, R = g.R, C = g.C, A = g.A, PC = g.PC, SC = g.SC;
    memset(&sb, 0, sizeof(sb));
    RMCV_CUDA(ctx, dalloc(&sb.bits, CF * H * WB));
    RMCV_CUDA(ctx, dalloc(&sb.band_flags, CF * ((H + 7) / 8)));
    RMCV_CUDA(ctx, dalloc(&sb.rows, CF * H));
    RMCV_CUDA(ctx, dalloc(&sb.run_x, CF * R));
    RMCV_CUDA(ctx, dalloc(&sb.run_y, CF * R));
    RMCV_CUDA(ctx, dalloc(&sb.parent, CF * R));
    RMCV_CUDA(ctx, dalloc(&sb.gparent, CF * (R + 2)));
    RMCV_CUDA(ctx, dalloc(&sb.run_cid, CF * R));
    RMCV_CUDA(ctx, dalloc(&sb.sorted, CF * SC));
    RMCV_CUDA(ctx, dalloc(&sb.recs, CF * PC));
    RMCV_CUDA(ctx, dalloc(&sb.recs2, CF * PC));
    RMCV_CUDA(ctx, dalloc(&sb.comp_start, CF * (C + 1)));
    RMCV_CUDA(ctx, dalloc(&sb.acc, CF * C));
    RMCV_CUDA(ctx, dalloc(&sb.comp_root, CF * C));
    RMCV_CUDA(ctx, dalloc(&sb.comp_cnt, CF * C));
    RMCV_CUDA(ctx, dalloc(&sb.comps, CF * C));
    RMCV_CUDA(ctx, dalloc(&sb.counters, CF + 1));
    RMCV_CUDA(ctx, cudaMemset(sb.counters, 0, (CF + 1) * sizeof(FrameCounters)));
    RMCV_CUDA(ctx, dalloc(&sb.s_contours, CF * C));
    RMCV_CUDA(ctx, dalloc(&sb.s_blobs, CF * C));
    RMCV_CUDA(ctx, dalloc(&sb.s_armours, CF * A));
    RMCV_CUDA(ctx, dalloc(&sb.arm_offset, CF * 4));
    (void)first;
    RMCV_CUDA(ctx, cudaEventCreateWithFlags(&sb.ev_pix, cudaEventDisableTiming));
    RMCV_CUDA(ctx, cudaEventCreateWithFlags(&sb.ev_lab, cudaEventDisableTiming));
    RMCV_CUDA(ctx, cudaEventCreateWithFlags(&sb.ev_fit, cudaEventDisableTiming));
    RMCV_CUDA(ctx, cudaEventCreateWithFlags(&sb.ev_h2d, cudaEventDisableTiming));
    RMCV_CUDA(ctx, cudaEventCreateWithFlags(&sb.ev_d2h, cudaEventDisableTiming));
    return RMCV_OK;
}

void free_slot(SlotBuffers& sb) {
    cudaFree(sb.bits); cudaFree(sb.band_flags); cudaFree(sb.rows); cudaFree(sb.run_x); cudaFree(sb.run_y);
    cudaFree(sb.parent); cudaFree(sb.gparent); cudaFree(sb.run_cid); cudaFree(sb.sorted); cudaFree(sb.recs); cudaFree(sb.recs2); cudaFree(sb.comp_start); cudaFree(sb.acc); cudaFree(sb.comp_root); cudaFree(sb.comp_cnt); cudaFree(sb.comps);
    cudaFree(sb.counters); cudaFree(sb.s_contours); cudaFree(sb.s_blobs); cudaFree(sb.s_armours); cudaFree(sb.arm_offset);
    if (sb.frames) cudaFree(sb.frames);
    if (sb.masks) cudaFree(sb.masks);
    if (sb.ev_pix) cudaEventDestroy(sb.ev_pix);
    if (sb.ev_lab) cudaEventDestroy(sb.ev_lab);
    if (sb.ev_fit) cudaEventDestroy(sb.ev_fit);
    if (sb.ev_h2d) cudaEventDestroy(sb.ev_h2d);
    if (sb.ev_d2h) cudaEventDestroy(sb.ev_d2h);
    memset(&sb, 0, sizeof(sb));
}

int check_geometry(rmcv_ctx* ctx, int width, int height, int batch) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    if (width <= 0 || height <= 0 || batch <= 0) return set_err(ctx, RMCV_ERR_INVALID_ARG, "width, height and batch must be positive");
    if (width > ctx->cfg.max_width || height > ctx->cfg.max_height || batch > ctx->cfg.max_batch)
        return set_err(ctx, RMCV_ERR_INVALID_ARG, "frame size or batch exceeds the ctx maxima");
    if (width > 32767 || height > 32767) return set_err(ctx, RMCV_ERR_INVALID_ARG, "frames above 32767 px per side are not supported");
    return RMCV_OK;
}

Geometry call_geometry(const rmcv_ctx* ctx, int W, int H) {
    Geometry g = ctx->cap;
    g.W = W; g.H = H; g.WB = (W + 31) / 32;
    return g;
}

ProfSet* prof_begin(rmcv_ctx* ctx) {
    if (!ctx->profiling) return nullptr;
    CtxExtra* ex = extra(ctx);
    if (ex->prof_used == ex->prof.size()) {
        ProfSet ps;
        for (int i = 0; i < 2; ++i) cudaEventCreate(&ps.pix[i]);
        for (int i = 0; i < RMCV_STAGE_COUNT; ++i) cudaEventCreate(&ps.lab[i]);
        ps.rec = false; ps.full = false;
        ex->prof.push_back(ps);
    }
    ProfSet* ps = &ex->prof[ex->prof_used++];
    ps->rec = true; ps->full = false;
    return ps;
}
void prof_collect(rmcv_ctx* ctx) {  // after a sync
    CtxExtra* ex = extra(ctx);
    for (size_t i = 0; i < ex->prof_used; ++i) {
        ProfSet& ps = ex->prof[i];
        if (!ps.rec) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ps.pix[0], ps.pix[1]) == cudaSuccess) ctx->prof_ms[RMCV_STAGE_PIXEL] += ms;
        for (int s = 1; ps.full && s < RMCV_STAGE_COUNT; ++s)
            if (cudaEventElapsedTime(&ms, ps.lab[s - 1], ps.lab[s]) == cudaSuccess) ctx->prof_ms[s] += ms;
        ps.rec = false;
    }
    ex->prof_used = 0;
    cudaGetLastError();
}

// Small chunks run on their slot's stream, not on the pixel stream (enqueue_chunk).  Whatever the library itself enqueues on
// the pixel stream afterwards - helper copies, a front-end pass, the pixel kernels of a large call - comes after them:
// a mask is complete before it is copied, an input frame is not overwritten before it has been read.
int order_pix_after_chains(rmcv_ctx* ctx) {
    CtxExtra* ex = extra(ctx);
    for (int i = 0; i < ctx->n_slots && i < kSlots; ++i)
        if (ex->slot_chained[i]) {
            RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->pix, ctx->slot[i].ev_lab, 0));
            ex->slot_chained[i] = false;
        }
    return RMCV_OK;
}

// Enqueue all stages for `frames` frames whose pixels are at `src`, using the scratch of slot `sb`.
int enqueue_chunk(rmcv_ctx* ctx, SlotBuffers& sb, const uint8_t* src, size_t pitch, size_t frame_stride, int W, int H,
                  int frames, int frame_base, int bayer_layout, const rmcv_params& prm, uint8_t* mask, size_t mask_pitch,
                  size_t mask_frame_stride, bool full, bool host_path = false) {
    CtxExtra* ex = extra(ctx);
    // each slot has its own labelling stream: the labelling kernels of consecutive chunks are latency-bound and overlap
    // each other as well as the pixel kernels
    // small chunks (latency mode; also the thread-block-cluster kernels of large frames): all kernels on one stream, each a
    // programmatic dependent of the one before (common.cuh: chain_begin / chain_wait) - no cross-stream hops, launch gaps hidden
    // (the slot's own stream, so that consecutive small calls still overlap; a caller-supplied stream keeps everything)
    const bool chained = full && frames <= small_batch_limit() && tuning().chained != 0;
    cudaStream_t sl = ex->labs[&sb - ctx->slot];
    cudaStream_t sp = chained && ex->own_pix ? sl : ex->pix;
    if (chained) {
        sl = sp;
        ex->chain = sp;
        if (sp != ex->pix) {
            // everything enqueued on the pixel stream before this call comes first: the waits of begin_call (the result set's
            // previous call), the frame upload of the host path, helper copies / a front-end pass, the caller's own work on
            // rmcv_stream()
            if (!ex->ev_order) RMCV_CUDA(ctx, cudaEventCreateWithFlags(&ex->ev_order, cudaEventDisableTiming));
            RMCV_CUDA(ctx, cudaEventRecord(ex->ev_order, ex->pix));
            RMCV_CUDA(ctx, cudaStreamWaitEvent(sp, ex->ev_order, 0));
            ex->slot_chained[&sb - ctx->slot] = true;
            ex->call_slots |= 1u << (unsigned)(&sb - ctx->slot);
        }
    } else {
        const int rc_o = order_pix_after_chains(ctx);
        if (rc_o != RMCV_OK) return rc_o;
    }
    // the slot's scratch is free once the labelling stages (and the mask download) of its previous chunk are done
    RMCV_CUDA(ctx, cudaStreamWaitEvent(sp, sb.ev_lab, 0));
    RMCV_CUDA(ctx, cudaStreamWaitEvent(sp, sb.ev_d2h, 0));
    if (full) RMCV_CUDA(ctx, cudaMemsetAsync(sb.counters, 0, (size_t)(frames + 1) * sizeof(FrameCounters), sp));
    ProfSet* ps = prof_begin(ctx);
    if (ps) cudaEventRecord(ps->pix[0], sp);
    const int64_t l0 = ctx->kernel_launches;
    PixelLaunch pl;
    pl.src = src; pl.pitch = pitch; pl.frame_stride = frame_stride;
    pl.mask = mask; pl.mask_pitch = mask_pitch; pl.mask_frame_stride = mask_frame_stride;
    pl.bits = sb.bits; pl.W = W; pl.H = H; pl.batch = frames;
    pl.target = prm.target; pl.lower_bound = prm.lower_bound; pl.bayer_layout = bayer_layout;
    EmitLaunch el;
    int emit_done = 0, flags_bh = 0;
    if (full) {   // the fixed-geometry BGR kernel can cut the runs / boundary records itself
        const Geometry cg = call_geometry(ctx, W, H);
        el.bits = sb.bits; el.W = W; el.H = H; el.batch = frames;
        el.rows = sb.rows; el.run_x = sb.run_x; el.run_y = sb.run_y; el.counters = sb.counters; el.R = cg.R;
        el.recs = sb.recs; el.PC = cg.PC;
        pl.emit = &el; pl.emit_done = &emit_done;
        pl.band_flags = sb.band_flags; pl.flags_bh = &flags_bh;
    }
    RMCV_CUDA(ctx, launch_pixel_stage(pl, ctx->sm_count, sp, &ctx->kernel_launches));
    if (ps) cudaEventRecord(ps->pix[1], sp);
    // (only the host path waits for ev_pix; an event between two kernels would undo the chained launch of the second)
    if (!chained || host_path) RMCV_CUDA(ctx, cudaEventRecord(sb.ev_pix, sp));
    ctx->prof_launches[RMCV_STAGE_PIXEL] += ctx->kernel_launches - l0;
    if (!full) return RMCV_OK;
    if (sl != sp) RMCV_CUDA(ctx, cudaStreamWaitEvent(sl, sb.ev_pix, 0));
    if (ps) { ps->full = true; cudaEventRecord(ps->lab[0], sl); }
    FrameLaunch fl;
    fl.g = call_geometry(ctx, W, H); fl.frames = frames; fl.sb = &sb; fl.frame_base = frame_base;
    fl.st_out = chained ? nullptr : ex->out;
    fl.chained = chained ? 1 : 0;
    fl.emit_done = emit_done;
    fl.flags_bh = flags_bh;
    fl.o_frames = ex->o_frames; fl.o_contours = ex->o_contours; fl.o_blobs = ex->o_blobs; fl.o_armours = ex->o_armours;
    fl.o_poses = ex->have_camera ? ctx->h_poses : nullptr; fl.camera = ex->have_camera ? &ex->camera : nullptr;
    struct Mark { ProfSet* ps; rmcv_ctx* ctx; };
    Mark mk{ps, ctx};
    auto stage_done = [](void* arg, int stage, cudaStream_t s2) {
        Mark* m = static_cast<Mark*>(arg);
        if (m->ps) cudaEventRecord(m->ps->lab[stage], s2);
        m->ctx->prof_launches[stage] += 1;
    };
    RMCV_CUDA(ctx, launch_frames(fl, prm, ctx->max_smem_optin, sl, &ctx->kernel_launches, stage_done, &mk));
    RMCV_CUDA(ctx, cudaEventRecord(sb.ev_lab, chained ? sp : ex->out));
    return RMCV_OK;
}

// A detect call writes into result set (call number & 1); the previous call's set stays readable, so the host can fetch
// call n while call n+1 is already running (rmcv_fetch_results returns the oldest unfetched call).
cudaError_t alloc_result_set(rmcv_ctx* ctx, ResultSet& r) {
    const size_t B = ctx->cfg.max_batch, C = ctx->cap.C, A = ctx->cap.A;
    const unsigned hflags = cudaHostAllocMapped | cudaHostAllocPortable;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&r.frames), B * sizeof(rmcv_frame_info), hflags);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&r.contours), B * C * sizeof(rmcv_contour_info), hflags);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&r.blobs), B * C * sizeof(rmcv_lightblob), hflags);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&r.armours), B * A * sizeof(rmcv_armour), hflags);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.done[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.done[1], cudaEventDisableTiming);
    if (e == cudaSuccess) memset(r.frames, 0, B * sizeof(rmcv_frame_info));
    return e;
}

cudaError_t alloc_result_mirror(rmcv_ctx* ctx, ResultSet& r) {
    const size_t B = ctx->cfg.max_batch, C = ctx->cap.C, A = ctx->cap.A;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&r.d_frames), B * sizeof(rmcv_frame_info));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r.d_contours), B * C * sizeof(rmcv_contour_info));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r.d_blobs), B * C * sizeof(rmcv_lightblob));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&r.d_armours), B * A * sizeof(rmcv_armour));
    return e;
}

int begin_call(rmcv_ctx* ctx, int batch = 0, int cf = 0) {
    CtxExtra* ex = extra(ctx);
    ResultSet& r = ex->rs[ex->n_calls % kResultSets];
    if (!r.frames) RMCV_CUDA(ctx, alloc_result_set(ctx, r));
    if (ex->have_camera && !r.poses)
        RMCV_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&r.poses), (size_t)ctx->cfg.max_batch * ctx->cap.A * sizeof(rmcv_pose),
                                     cudaHostAllocMapped | cudaHostAllocPortable));
    r.pending = false;  // an unfetched call kResultSets calls back is dropped
    // that call wrote this set from the write-out stream or, chained, from a slot stream: this call's work comes after it
    RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->pix, r.done[0], 0));
    RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->pix, r.done[1], 0));
    ctx->h_frames = r.frames; ctx->h_contours = r.contours; ctx->h_blobs = r.blobs; ctx->h_armours = r.armours;
    ctx->h_poses = r.poses;
    r.with_poses = ex->have_camera && r.poses;
    const int mode = tuning().staged_out;      // -1 auto, 0 never, 1 always
    r.staged = mode > 0 || (mode < 0 && batch > 64);
    r.materialised = !r.staged;
    r.cf = cf;
    if (r.staged && !r.d_frames) RMCV_CUDA(ctx, alloc_result_mirror(ctx, r));
    ex->o_frames = r.staged ? r.d_frames : r.frames; ex->o_contours = r.staged ? r.d_contours : r.contours;
    ex->o_blobs = r.staged ? r.d_blobs : r.blobs; ex->o_armours = r.staged ? r.d_armours : r.armours;
    return RMCV_OK;
}
int end_call(rmcv_ctx* ctx, int batch) {
    CtxExtra* ex = extra(ctx);
    ResultSet& r = ex->rs[ex->n_calls % kResultSets];
    if (r.staged)   // the per-frame counts and offsets travel first; the dense records follow when the call is fetched
        RMCV_CUDA(ctx, cudaMemcpyAsync(r.frames, r.d_frames, (size_t)batch * sizeof(rmcv_frame_info), cudaMemcpyDeviceToHost, ex->out));
    // a call of several small chunks ran them on several slot streams: the write-out stream, whose event ends the call, waits
    // for the end of each (ev_lab: recorded behind a chunk's last kernel)
    for (int i = 0; i < ctx->n_slots && i < kSlots; ++i)
        if (ex->call_slots & (1u << i)) RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->out, ctx->slot[i].ev_lab, 0));
    ex->call_slots = 0;
    RMCV_CUDA(ctx, cudaEventRecord(r.done[0], ex->out));
    RMCV_CUDA(ctx, cudaEventRecord(r.done[1], ex->chain ? ex->chain : ex->pix));
    ex->chain = nullptr;
    r.batch = batch; r.pending = true; r.call_id = ex->n_calls++;
    return RMCV_OK;
}

int run_device_batch(rmcv_ctx* ctx, const uint8_t* d_src, size_t pitch, size_t frame_stride, int W, int H, int batch,
                     int bayer_layout, const rmcv_params& prm, uint8_t* d_mask, size_t mask_pitch, size_t mask_frame_stride,
                     bool full) {
    int rc = check_geometry(ctx, W, H, batch);
    if (rc != RMCV_OK) return rc;
    if (!d_src) return set_err(ctx, RMCV_ERR_INVALID_ARG, "null frame pointer");
    const size_t rowbytes = bayer_layout ? (size_t)W : (size_t)W * 3;
    if (pitch < rowbytes) return set_err(ctx, RMCV_ERR_INVALID_ARG, "pitch smaller than a row");
    if (d_mask && mask_pitch < (size_t)W) return set_err(ctx, RMCV_ERR_INVALID_ARG, "mask pitch smaller than a row");
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const int CF = ctx->CF;
    if (full) { rc = begin_call(ctx, batch, CF); if (rc != RMCV_OK) return rc; }
    int nchunks = 0;
    extra(ctx)->last_first_slot = (int)(extra(ctx)->chunk_counter % ctx->n_slots);
    for (int f0 = 0; f0 < batch; f0 += CF, ++nchunks) {
        SlotBuffers& sb = ctx->slot[extra(ctx)->chunk_counter++ % ctx->n_slots];
        const int frames = batch - f0 < CF ? batch - f0 : CF;
        rc = enqueue_chunk(ctx, sb, d_src + (size_t)f0 * frame_stride, pitch, frame_stride, W, H, frames, f0, bayer_layout,
                           prm, d_mask ? d_mask + (size_t)f0 * mask_frame_stride : nullptr, mask_pitch, mask_frame_stride, full);
        if (rc != RMCV_OK) return rc;
    }
    ctx->last_batch = batch; ctx->last_W = W; ctx->last_H = H;
    ctx->have_results = full;
    extra(ctx)->last_nchunks = nchunks;
    extra(ctx)->last_cf = CF;
    extra(ctx)->last_kind = full ? 2 : 1;
    if (full) return end_call(ctx, batch);
    return RMCV_OK;
}

// Brings the dense records of a staged call into its pinned arrays: one DMA copy per array and chunk, sized from the
// per-frame counts (which came over with the frame infos).  Idempotent.
int materialise(rmcv_ctx* ctx, ResultSet& r) {
    if (!r.staged || r.materialised || r.batch <= 0) return RMCV_OK;
    CtxExtra* ex = extra(ctx);
    RMCV_CUDA(ctx, cudaEventSynchronize(r.done[0]));
    const size_t C = ctx->cap.C, A = ctx->cap.A;
    const int cf = r.cf > 0 ? r.cf : r.batch;
    for (int f0 = 0; f0 < r.batch; f0 += cf) {
        const int f1 = f0 + cf < r.batch ? f0 + cf : r.batch;
        size_t nc = 0, nb = 0, na = 0;
        for (int f = f0; f < f1; ++f) { nc += (size_t)r.frames[f].n_contours; nb += (size_t)r.frames[f].n_positive; na += (size_t)r.frames[f].n_armours; }
        if (nc) RMCV_CUDA(ctx, cudaMemcpyAsync(r.contours + (size_t)f0 * C, r.d_contours + (size_t)f0 * C, nc * sizeof(rmcv_contour_info), cudaMemcpyDeviceToHost, ex->d2h));
        if (nb) RMCV_CUDA(ctx, cudaMemcpyAsync(r.blobs + (size_t)f0 * C, r.d_blobs + (size_t)f0 * C, nb * sizeof(rmcv_lightblob), cudaMemcpyDeviceToHost, ex->d2h));
        if (na) RMCV_CUDA(ctx, cudaMemcpyAsync(r.armours + (size_t)f0 * A, r.d_armours + (size_t)f0 * A, na * sizeof(rmcv_armour), cudaMemcpyDeviceToHost, ex->d2h));
    }
    RMCV_CUDA(ctx, cudaStreamSynchronize(ex->d2h));
    r.materialised = true;
    return RMCV_OK;
}

int sync_all(rmcv_ctx* ctx) {
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    CtxExtra* ex = extra(ctx);
    RMCV_CUDA(ctx, cudaStreamSynchronize(ex->h2d));
    RMCV_CUDA(ctx, cudaStreamSynchronize(ex->pix));
    for (int i = 0; i < kSlots; ++i) RMCV_CUDA(ctx, cudaStreamSynchronize(ex->labs[i]));
    RMCV_CUDA(ctx, cudaStreamSynchronize(ex->out));
    RMCV_CUDA(ctx, cudaStreamSynchronize(ex->d2h));
    prof_collect(ctx);
    if (ex->n_calls > 0) {   // the on-demand getters read the most recent call's records on the host
        ResultSet& r = ex->rs[(ex->n_calls - 1) % kResultSets];
        if (r.call_id == ex->n_calls - 1) { const int rc = materialise(ctx, r); if (rc != RMCV_OK) return rc; }
    }
    return RMCV_OK;
}

int fill_results(rmcv_ctx* ctx, const ResultSet& r, rmcv_results* out) {
    int flags = 0;
    long long tc = 0, tb = 0, ta = 0;
    for (int f = 0; f < r.batch; ++f) {
        const rmcv_frame_info& fi = r.frames[f];
        flags |= fi.flags; tc += fi.n_contours; tb += fi.n_positive; ta += fi.n_armours;
    }
    if (out) {
        out->batch = r.batch;
        out->total_contours = (int32_t)tc; out->total_blobs = (int32_t)tb; out->total_armours = (int32_t)ta;
        out->frames = r.frames; out->contours = r.contours; out->blobs = r.blobs; out->armours = r.armours;
        out->poses = r.with_poses ? r.poses : nullptr;
    }
    if (flags) return set_err(ctx, RMCV_ERR_CAPACITY, "a per-frame capacity overflowed; see rmcv_frame_info.flags");
    return RMCV_OK;
}

// Waits for the oldest unfetched detect call (or re-exposes the last fetched one) and fills `out`.
int fetch_oldest(rmcv_ctx* ctx, rmcv_results* out) {
    CtxExtra* ex = extra(ctx);
    int pick = -1;
    for (int i = 0; i < kResultSets; ++i)
        if (ex->rs[i].pending && (pick < 0 || ex->rs[i].call_id < ex->rs[pick].call_id)) pick = i;
    if (pick < 0) {
        if (ex->last_fetched < 0) return set_err(ctx, RMCV_ERR_STATE, "no detect call to fetch results from");
        return fill_results(ctx, ex->rs[ex->last_fetched], out);
    }
    ResultSet& r = ex->rs[pick];
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    RMCV_CUDA(ctx, cudaEventSynchronize(r.done[0]));
    RMCV_CUDA(ctx, cudaEventSynchronize(r.done[1]));
    if (int mrc = materialise(ctx, r)) return mrc;
    r.pending = false;
    ex->last_fetched = pick;
    // profiling events of finished chunks are collected here too: an async detect/fetch loop never reaches sync_all, and
    // without this the event pool would grow with every call
    bool any_pending = false;
    for (int i = 0; i < kResultSets; ++i) any_pending |= ex->rs[i].pending;
    if (ctx->profiling && !any_pending) prof_collect(ctx);
    else if (ctx->profiling && ex->prof_used > 4096) { ex->prof_used = 0; }   // bounded: drop the oldest marks
    return fill_results(ctx, r, out);
}

// The on-demand getters refer to the most recent detect call: wait for THAT call (its two done events) and bring its records
// to the host, instead of synchronising every stream of the ctx (a dozen cudaStreamSynchronize calls, ~50 us when idle).
int wait_latest_call(rmcv_ctx* ctx) {
    CtxExtra* ex = extra(ctx);
    if (ex->n_calls <= 0 || ex->last_kind != 2) return sync_all(ctx);
    ResultSet& r = ex->rs[(ex->n_calls - 1) % kResultSets];
    if (r.call_id != ex->n_calls - 1) return sync_all(ctx);
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (r.pending) {
        RMCV_CUDA(ctx, cudaEventSynchronize(r.done[0]));
        RMCV_CUDA(ctx, cudaEventSynchronize(r.done[1]));
    }
    return materialise(ctx, r);
}

// slot and local index of frame f of the last call, or null when its scratch has been recycled
SlotBuffers* resident_slot(rmcv_ctx* ctx, int frame, int* local) {
    if (frame < 0 || frame >= ctx->last_batch) return nullptr;
    const int cf = extra(ctx)->last_cf > 0 ? extra(ctx)->last_cf : ctx->CF;
    const int chunk = frame / cf;
    if (chunk + ctx->n_slots < extra(ctx)->last_nchunks) return nullptr;
    *local = frame - chunk * cf;
    return &ctx->slot[(extra(ctx)->last_first_slot + chunk) % ctx->n_slots];
}

}  // namespace

namespace rmcv {
const Tuning& tuning() {
    static const Tuning t = read_tuning();   // thread-safe one-time initialisation
    return t;
}
}  // namespace rmcv

// =============================================================================================== ABI
extern "C" {

int rmcv_abi_version(void) { return RMCV_B200_ABI_VERSION; }

const char* rmcv_status_string(int s) {
    switch (s) {
        case RMCV_OK: return "ok";
        case RMCV_ERR_INVALID_ARG: return "invalid argument";
        case RMCV_ERR_CUDA: return "CUDA error";
        case RMCV_ERR_CAPACITY: return "per-frame capacity overflow";
        case RMCV_ERR_NO_DEVICE: return "no CUDA device";
        case RMCV_ERR_STATE: return "invalid call order";
        default: return "unknown status";
    }
}

void rmcv_default_params(rmcv_params* p) {  // executable/main.cpp:172-176
    if (!p) return;
    p->target = RMCV_CAMP_BLUE; p->lower_bound = 80; p->tilt_max = 70.f;
    p->ratio_min = 1.5f; p->ratio_max = 80.f; p->area_min = 10.0; p->area_max = 99999.0;
    p->angle_difference_max = 12.f; p->shear_max = 22.f; p->lenght_ratio_max = 0.4f;
}

void rmcv_default_config(rmcv_config* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->device = 0; c->max_width = 1280; c->max_height = 1024; c->max_batch = 64;
}

int rmcv_device_count(int* count) {
    if (!count) return RMCV_ERR_INVALID_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    *count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess) { cudaGetLastError(); return RMCV_ERR_NO_DEVICE; }
    return RMCV_OK;
}

int rmcv_ctx_create(const rmcv_config* cfg, rmcv_ctx** out) {
    if (!cfg || !out) return RMCV_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->max_width <= 0 || cfg->max_height <= 0 || cfg->max_batch <= 0 || cfg->max_width > 32767 || cfg->max_height > 32767) return RMCV_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return RMCV_ERR_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) return RMCV_ERR_INVALID_ARG;
    rmcv_ctx* ctx = static_cast<rmcv_ctx*>(calloc(1, sizeof(rmcv_ctx)));
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    ctx->extra = new (std::nothrow) CtxExtra();
    if (!ctx->extra) { free(ctx); return RMCV_ERR_INVALID_ARG; }
    const Tuning& tune = tuning();
    auto fail = [&](int code) {
        static thread_local char keep[512];
        snprintf(keep, sizeof(keep), "%s", ctx->err);
        fprintf(stderr, "rmcv_ctx_create failed: %s\n", keep);
        rmcv_ctx_destroy(ctx);
        return code;
    };
    if (cudaSetDevice(ctx->device) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "cudaSetDevice failed"); return fail(RMCV_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "cudaGetDeviceProperties failed"); return fail(RMCV_ERR_CUDA); }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    Geometry& g = ctx->cap;
    g.W = cfg->max_width; g.H = cfg->max_height; g.WB = (g.W + 31) / 32;
    const long long px = (long long)g.W * g.H;
    long long R = cfg->max_runs_per_frame > 0 ? cfg->max_runs_per_frame : (px / 32 > 16384 ? px / 32 : 16384);
    if (R > px / 2 + g.H) R = px / 2 + g.H;
    g.R = (int)R;
    g.PC = (int)(4 * R > px ? px : 4 * R);  // a boundary pixel is a foreground pixel
    if (g.PC < 64) g.PC = 64;
    g.SC = g.PC > g.R + 2 ? g.PC : g.R + 2;
    g.C = cfg->max_blobs_per_frame > 0 ? cfg->max_blobs_per_frame : 512;
    g.A = cfg->max_armours_per_frame > 0 ? cfg->max_armours_per_frame : 1024;
    int CF = cfg->chunk_frames;
    if (CF <= 0) {
        // A chunk is one launch of each kernel.  The pixel kernel and the labelling kernels alternate on the SMs rather than
        // overlap, and every kernel boundary costs a ramp-up and a drain, so the fewer and larger the launches
        // the better: a whole call is one chunk whenever its scratch fits (sweep on 1024 / 2048 frames of 1280x1024: one
        // chunk 1.19 / 2.34 ms, chunks of 592 frames 1.23 / 2.45 ms); calls in flight overlap through the scratch slots.
        // Bound: about 4 bytes of scratch per pixel and slot (budgeted at 4.5) -> at most 6 GB per slot.
        const long long c = (6144LL << 20) / (px * 9 / 2);
        CF = (int)(c < 1 ? 1 : (c > 65535 ? 65535 : c));
    }
    if (CF > cfg->max_batch) CF = cfg->max_batch;
    ctx->CF = CF;
    {
        const int ns = tune.slots > 0 ? tune.slots : 3;
        ctx->n_slots = ns < 2 ? 2 : (ns > kSlots ? kSlots : ns);
    }
    {
        CtxExtra* ex = extra(ctx);
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (tune.prio == 1) greatest = least;          // experiment: 1 = equal priorities, 2 = pixel stream first
        else if (tune.prio == 2) { const int t = least; least = greatest; greatest = t; }
        cudaError_t se = cudaSuccess;
        if (cfg->stream) { ex->pix = reinterpret_cast<cudaStream_t>(cfg->stream); ex->own_pix = false; }
        else { se = cudaStreamCreateWithPriority(&ex->pix, cudaStreamNonBlocking, least); ex->own_pix = true; }
        if (tune.serial) {   // debug aid: every kernel on one stream
            ex->lab = ex->pix; ex->out = ex->pix;
            for (int i = 0; i < kSlots; ++i) ex->labs[i] = ex->pix;
        } else {
            if (se == cudaSuccess) se = cudaStreamCreateWithPriority(&ex->lab, cudaStreamNonBlocking, greatest);
            ex->labs[0] = ex->lab;
            for (int i = 1; i < kSlots; ++i) {
                if (tune.lab_streams == 1) ex->labs[i] = ex->lab;   // all slots share one labelling stream (experiments)
                else if (se == cudaSuccess) se = cudaStreamCreateWithPriority(&ex->labs[i], cudaStreamNonBlocking, greatest);
            }
            if (se == cudaSuccess) se = cudaStreamCreateWithPriority(&ex->out, cudaStreamNonBlocking, greatest);
        }
        if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&ex->h2d, cudaStreamNonBlocking);
        if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&ex->d2h, cudaStreamNonBlocking);
        if (se != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "stream creation failed: %s", cudaGetErrorString(se)); return fail(RMCV_ERR_CUDA); }
    }
    int rc = RMCV_OK;
    for (int i = 0; i < ctx->n_slots && rc == RMCV_OK; ++i) rc = alloc_slot(ctx, ctx->slot[i], i == 0);
    if (rc != RMCV_OK) return fail(rc);
    cudaError_t e = cudaSuccess;
    // every result set (and, for ctxs that will see batches above 64 frames, its device mirror) up front: an allocation
    // inside a detect call would stall that call by tens of milliseconds
    for (int i = 0; i < kResultSets && e == cudaSuccess; ++i) {
        ResultSet& r = extra(ctx)->rs[i];
        e = alloc_result_set(ctx, r);
        const int mode = tuning().staged_out;
        if (e == cudaSuccess && (mode > 0 || (mode < 0 && cfg->max_batch > 64))) e = alloc_result_mirror(ctx, r);
    }
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "pinned result allocation failed: %s", cudaGetErrorString(e));
        return fail(RMCV_ERR_CUDA);
    }
    begin_call(ctx);
    upload_luts();
    if (cudaDeviceSynchronize() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "device sync after init failed"); return fail(RMCV_ERR_CUDA); }
    *out = ctx;
    return RMCV_OK;
}

int rmcv_ctx_destroy(rmcv_ctx* ctx) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < ctx->n_slots; ++i) free_slot(ctx->slot[i]);
    CtxExtra* ex = extra(ctx);
    if (ex) {
        for (int i = 0; i < kResultSets; ++i) {
            ResultSet& r = ex->rs[i];
            if (r.frames) cudaFreeHost(r.frames);
            if (r.contours) cudaFreeHost(r.contours);
            if (r.blobs) cudaFreeHost(r.blobs);
            if (r.armours) cudaFreeHost(r.armours);
            if (r.poses) cudaFreeHost(r.poses);
            if (r.d_frames) cudaFree(r.d_frames);
            if (r.d_contours) cudaFree(r.d_contours);
            if (r.d_blobs) cudaFree(r.d_blobs);
            if (r.d_armours) cudaFree(r.d_armours);
            for (int k = 0; k < 2; ++k) if (r.done[k]) cudaEventDestroy(r.done[k]);
        }
        for (auto& ps : ex->prof) {
            for (int i = 0; i < 2; ++i) cudaEventDestroy(ps.pix[i]);
            for (int i = 0; i < RMCV_STAGE_COUNT; ++i) cudaEventDestroy(ps.lab[i]);
        }
        for (int i = 0; i < 4 + kSlots; ++i) { if (ex->t_start[i]) cudaEventDestroy(ex->t_start[i]); if (ex->t_stop[i]) cudaEventDestroy(ex->t_stop[i]); }
        for (int i = 1; i < kSlots; ++i) if (ex->labs[i] && ex->labs[i] != ex->pix && ex->labs[i] != ex->lab) cudaStreamDestroy(ex->labs[i]);
        if (ex->lab && ex->lab != ex->pix) cudaStreamDestroy(ex->lab);
        if (ex->out && ex->out != ex->pix) cudaStreamDestroy(ex->out);
        if (ex->ev_order) cudaEventDestroy(ex->ev_order);
        if (ex->pix && ex->own_pix) cudaStreamDestroy(ex->pix);
        if (ex->h2d) cudaStreamDestroy(ex->h2d);
        if (ex->d2h) cudaStreamDestroy(ex->d2h);
        for (void* p : ex->dev_allocs) cudaFree(p);
        for (void* p : ex->host_allocs) cudaFreeHost(p);
        if (ex->tmp_dev) cudaFree(ex->tmp_dev);
        if (ex->tmp_host) cudaFreeHost(ex->tmp_host);
        delete ex;
    }
    cudaGetLastError();
    free(ctx);
    return RMCV_OK;
}

const char* rmcv_last_error(const rmcv_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
int rmcv_chunk_frames(const rmcv_ctx* ctx) { return ctx ? ctx->CF : 0; }

int rmcv_device_alloc(rmcv_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return RMCV_ERR_INVALID_ARG;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    RMCV_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    extra(ctx)->dev_allocs.push_back(*dptr);
    return RMCV_OK;
}
int rmcv_device_free(rmcv_ctx* ctx, void* dptr) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    auto& v = extra(ctx)->dev_allocs;
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i] == dptr) { v.erase(v.begin() + i); break; }
    RMCV_CUDA(ctx, cudaFree(dptr));
    return RMCV_OK;
}
int rmcv_host_alloc(rmcv_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return RMCV_ERR_INVALID_ARG;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    RMCV_CUDA(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocPortable));
    extra(ctx)->host_allocs.push_back(*hptr);
    return RMCV_OK;
}
int rmcv_host_free(rmcv_ctx* ctx, void* hptr) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    auto& v = extra(ctx)->host_allocs;
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i] == hptr) { v.erase(v.begin() + i); break; }
    RMCV_CUDA(ctx, cudaFreeHost(hptr));
    return RMCV_OK;
}
int rmcv_memcpy_h2d(rmcv_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    if (order_pix_after_chains(ctx) != RMCV_OK) return RMCV_ERR_CUDA;
    RMCV_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, extra(ctx)->pix));
    return RMCV_OK;
}
int rmcv_memcpy_d2h(rmcv_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    if (order_pix_after_chains(ctx) != RMCV_OK) return RMCV_ERR_CUDA;
    RMCV_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, extra(ctx)->pix));
    return RMCV_OK;
}
int rmcv_memset_d(rmcv_ctx* ctx, void* dst, int value, size_t bytes) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    if (order_pix_after_chains(ctx) != RMCV_OK) return RMCV_ERR_CUDA;
    RMCV_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, extra(ctx)->pix));
    return RMCV_OK;
}
int rmcv_sync(rmcv_ctx* ctx) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    return sync_all(ctx);
}
void* rmcv_stream(rmcv_ctx* ctx) { return ctx ? reinterpret_cast<void*>(extra(ctx)->pix) : nullptr; }

int rmcv_extract_color_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride, int width, int height,
                             int batch, int target, int lower_bound, uint8_t* d_mask, size_t mask_pitch,
                             size_t mask_frame_stride) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    rmcv_params prm;
    rmcv_default_params(&prm);
    prm.target = target; prm.lower_bound = lower_bound;
    return run_device_batch(ctx, d_bgr, pitch, frame_stride, width, height, batch, 0, prm, d_mask, mask_pitch, mask_frame_stride, false);
}

int rmcv_bayer_extract_color_batch(rmcv_ctx* ctx, const uint8_t* d_raw, size_t pitch, size_t frame_stride, int width,
                                   int height, int batch, int bayer_layout, int target, int lower_bound, uint8_t* d_mask,
                                   size_t mask_pitch, size_t mask_frame_stride) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    if (bayer_layout < RMCV_BAYER_RG || bayer_layout > RMCV_BAYER_BG) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad bayer layout");
    if (width < 3 || height < 3) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bayer frames must be at least 3x3");
    rmcv_params prm;
    rmcv_default_params(&prm);
    prm.target = target; prm.lower_bound = lower_bound;
    return run_device_batch(ctx, d_raw, pitch, frame_stride, width, height, batch, bayer_layout, prm, d_mask, mask_pitch,
                            mask_frame_stride, false);
}

int rmcv_detect_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride, int width, int height, int batch,
                      const rmcv_params* params, uint8_t* d_mask, size_t mask_pitch, size_t mask_frame_stride) {
    if (!ctx || !params) return RMCV_ERR_INVALID_ARG;
    return run_device_batch(ctx, d_bgr, pitch, frame_stride, width, height, batch, 0, *params, d_mask, mask_pitch, mask_frame_stride, true);
}

int rmcv_bayer_detect_batch(rmcv_ctx* ctx, const uint8_t* d_raw, size_t pitch, size_t frame_stride, int width, int height,
                            int batch, int bayer_layout, const rmcv_params* params, uint8_t* d_mask, size_t mask_pitch,
                            size_t mask_frame_stride) {
    if (!ctx || !params) return RMCV_ERR_INVALID_ARG;
    if (bayer_layout < RMCV_BAYER_RG || bayer_layout > RMCV_BAYER_BG) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad bayer layout");
    if (width < 3 || height < 3) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bayer frames must be at least 3x3");
    return run_device_batch(ctx, d_raw, pitch, frame_stride, width, height, batch, bayer_layout, *params, d_mask, mask_pitch,
                            mask_frame_stride, true);
}

// Host-buffer entry points: frames are staged chunk by chunk (upload on the copy stream, kernels, mask download), so the
// copies of one chunk overlap the kernels of the other.  bayer_layout == 0: interleaved BGR, else a raw 8-bit mosaic.
static int run_host_batch(rmcv_ctx* ctx, const uint8_t* h_bgr, size_t pitch, size_t frame_stride, int width, int height,
                          int batch, int bayer_layout, const rmcv_params* params, uint8_t* h_mask, size_t mask_pitch,
                          size_t mask_frame_stride, rmcv_results* out) {
    if (!ctx || !params || !h_bgr) return RMCV_ERR_INVALID_ARG;
    int rc = check_geometry(ctx, width, height, batch);
    if (rc != RMCV_OK) return rc;
    const size_t rowbytes = bayer_layout ? (size_t)width : (size_t)width * 3;
    if (pitch < rowbytes) return set_err(ctx, RMCV_ERR_INVALID_ARG, "pitch smaller than a row");
    if (h_mask && mask_pitch < (size_t)width) return set_err(ctx, RMCV_ERR_INVALID_ARG, "mask pitch smaller than a row");
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    // Host-input calls are bound by the PCIe copies, not by the kernels: short chunks keep both copy engines busy (the
    // upload of chunk i+1 overlaps the mask download of chunk i) and leave only a short download tail behind the last upload.
    int CF = ctx->CF;
    {
        const int hc = tuning().host_chunk > 0 ? tuning().host_chunk : 64;
        if (CF > hc) CF = hc;
    }
    rc = begin_call(ctx, batch, CF);
    if (rc != RMCV_OK) return rc;
    const size_t dev_frame = (size_t)height * rowbytes, dev_mask = (size_t)height * width;
    int nchunks = 0;
    extra(ctx)->last_first_slot = (int)(extra(ctx)->chunk_counter % ctx->n_slots);
    for (int f0 = 0; f0 < batch; f0 += CF, ++nchunks) {
        SlotBuffers& sb = ctx->slot[extra(ctx)->chunk_counter++ % ctx->n_slots];
        const int frames = batch - f0 < CF ? batch - f0 : CF;
        const size_t need_px = (size_t)(ctx->CF < 64 ? ctx->CF : 64) * (size_t)ctx->cfg.max_height * ctx->cfg.max_width;
        const size_t need = bayer_layout ? need_px : need_px * 3;   // a ctx that only ever sees mosaics stages 1 B/px
        if (sb.frames_bytes < need) {
            if (sb.frames) cudaFree(sb.frames);
            sb.frames = nullptr; sb.frames_bytes = 0;
            RMCV_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&sb.frames), need));
            sb.frames_bytes = need;
        }
        if (h_mask && sb.masks_bytes < need_px) {
            if (sb.masks) cudaFree(sb.masks);
            sb.masks = nullptr; sb.masks_bytes = 0;
            RMCV_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&sb.masks), need_px));
            sb.masks_bytes = need_px;
        }
        const uint8_t* hsrc = h_bgr + (size_t)f0 * frame_stride;
        CtxExtra* ex = extra(ctx);
        // upload on the copy stream once the pixel kernel of the slot's previous chunk has read the staging buffer
        RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->h2d, sb.ev_pix, 0));
        if (pitch == rowbytes && frame_stride == dev_frame) {
            RMCV_CUDA(ctx, cudaMemcpyAsync(sb.frames, hsrc, (size_t)frames * dev_frame, cudaMemcpyHostToDevice, ex->h2d));
        } else {
            for (int f = 0; f < frames; ++f)
                RMCV_CUDA(ctx, cudaMemcpy2DAsync(sb.frames + (size_t)f * dev_frame, rowbytes, hsrc + (size_t)f * frame_stride, pitch,
                                                 rowbytes, height, cudaMemcpyHostToDevice, ex->h2d));
        }
        RMCV_CUDA(ctx, cudaEventRecord(sb.ev_h2d, ex->h2d));
        RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->pix, sb.ev_h2d, 0));
        rc = enqueue_chunk(ctx, sb, sb.frames, rowbytes, dev_frame, width, height, frames, f0, bayer_layout, *params,
                           h_mask ? sb.masks : nullptr, width, dev_mask, true, true);
        if (rc != RMCV_OK) return rc;
        if (h_mask) {
            uint8_t* hdst = h_mask + (size_t)f0 * mask_frame_stride;
            RMCV_CUDA(ctx, cudaStreamWaitEvent(ex->d2h, sb.ev_pix, 0));
            if (mask_pitch == (size_t)width && mask_frame_stride == dev_mask) {
                RMCV_CUDA(ctx, cudaMemcpyAsync(hdst, sb.masks, (size_t)frames * dev_mask, cudaMemcpyDeviceToHost, ex->d2h));
            } else {
                for (int f = 0; f < frames; ++f)
                    RMCV_CUDA(ctx, cudaMemcpy2DAsync(hdst + (size_t)f * mask_frame_stride, mask_pitch, sb.masks + (size_t)f * dev_mask,
                                                     width, width, height, cudaMemcpyDeviceToHost, ex->d2h));
            }
            RMCV_CUDA(ctx, cudaEventRecord(sb.ev_d2h, ex->d2h));
        }
    }
    ctx->last_batch = batch; ctx->last_W = width; ctx->last_H = height;
    ctx->have_results = true;
    extra(ctx)->last_nchunks = nchunks;
    extra(ctx)->last_cf = CF;
    extra(ctx)->last_kind = 2;
    const int set = (int)(extra(ctx)->n_calls % kResultSets);
    rc = end_call(ctx, batch);
    if (rc != RMCV_OK) return rc;
    rc = sync_all(ctx);
    if (rc != RMCV_OK) return rc;
    extra(ctx)->rs[set].pending = false;
    extra(ctx)->last_fetched = set;
    return fill_results(ctx, extra(ctx)->rs[set], out);
}

int rmcv_detect_batch_host(rmcv_ctx* ctx, const uint8_t* h_bgr, size_t pitch, size_t frame_stride, int width, int height,
                           int batch, const rmcv_params* params, uint8_t* h_mask, size_t mask_pitch, size_t mask_frame_stride,
                           rmcv_results* out) {
    return run_host_batch(ctx, h_bgr, pitch, frame_stride, width, height, batch, 0, params, h_mask, mask_pitch, mask_frame_stride, out);
}

int rmcv_bayer_detect_batch_host(rmcv_ctx* ctx, const uint8_t* h_raw, size_t pitch, size_t frame_stride, int width, int height,
                                 int batch, int bayer_layout, const rmcv_params* params, uint8_t* h_mask, size_t mask_pitch,
                                 size_t mask_frame_stride, rmcv_results* out) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    if (bayer_layout < RMCV_BAYER_RG || bayer_layout > RMCV_BAYER_BG) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad bayer layout");
    if (width < 3 || height < 3) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bayer frames must be at least 3x3");
    return run_host_batch(ctx, h_raw, pitch, frame_stride, width, height, batch, bayer_layout, params, h_mask, mask_pitch,
                          mask_frame_stride, out);
}

int rmcv_fetch_results(rmcv_ctx* ctx, rmcv_results* out) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    return fetch_oldest(ctx, out);
}

int rmcv_get_contour(rmcv_ctx* ctx, int frame, int contour_index, int32_t* xy, int cap, int* n_points) {
    if (!ctx || !n_points || cap < 0 || (cap > 0 && !xy)) return RMCV_ERR_INVALID_ARG;
    int rc = wait_latest_call(ctx);
    if (rc != RMCV_OK) return rc;
    if (!ctx->have_results) return set_err(ctx, RMCV_ERR_STATE, "rmcv_get_contour needs a detect call first");
    int local = 0;
    SlotBuffers* sb = resident_slot(ctx, frame, &local);
    if (!sb) return set_err(ctx, RMCV_ERR_STATE, "frame scratch no longer resident (only the last two chunks are kept)");
    const rmcv_frame_info& fi = ctx->h_frames[frame];
    if (contour_index < 0 || contour_index >= fi.n_contours) return set_err(ctx, RMCV_ERR_INVALID_ARG, "contour index out of range");
    const rmcv_contour_info& ci = ctx->h_contours[fi.contour_offset + contour_index];
    const size_t bytes = (size_t)(cap > 0 ? cap : 1) * 2 * sizeof(int32_t) + 16;
    rc = ensure_tmp(ctx, bytes, bytes);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    int32_t* d_n = reinterpret_cast<int32_t*>(ex->tmp_dev);
    int32_t* d_xy = d_n + 4;
    Geometry g = call_geometry(ctx, ctx->last_W, ctx->last_H);
    cudaStream_t st = extra(ctx)->lab;
    RMCV_CUDA(ctx, launch_trace_contour(g, sb->bits + (size_t)local * g.H * g.WB, ci.first_x, ci.first_y, d_xy, cap, d_n, st,
                                        &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, ex->tmp_dev, bytes, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    const int32_t* h = reinterpret_cast<const int32_t*>(ex->tmp_host);
    *n_points = h[0];
    const int ncopy = h[0] < cap ? h[0] : cap;
    if (ncopy > 0) memcpy(xy, h + 4, (size_t)ncopy * 2 * sizeof(int32_t));
    return RMCV_OK;
}

int rmcv_get_contours(rmcv_ctx* ctx, int frame, int32_t* xy, int cap_points, int32_t* offsets, int cap_contours, int* n_contours,
                      int* n_points) {
    if (!ctx || !n_contours || !n_points || cap_points < 0 || cap_contours < 0) return RMCV_ERR_INVALID_ARG;
    int rc = wait_latest_call(ctx);
    if (rc != RMCV_OK) return rc;
    if (!ctx->have_results) return set_err(ctx, RMCV_ERR_STATE, "rmcv_get_contours needs a detect call first");
    int local = 0;
    SlotBuffers* sb = resident_slot(ctx, frame, &local);
    if (!sb) return set_err(ctx, RMCV_ERR_STATE, "frame scratch no longer resident (only the last two chunks are kept)");
    const rmcv_frame_info& fi = ctx->h_frames[frame];
    const int nc = fi.n_contours;
    long long total = 0;
    for (int k = 0; k < nc; ++k) total += ctx->h_contours[fi.contour_offset + k].n_points;
    *n_contours = nc;
    *n_points = (int)total;
    if (nc > cap_contours || total > cap_points || (nc > 0 && (!offsets || !xy)))
        return set_err(ctx, RMCV_ERR_CAPACITY, "rmcv_get_contours: caller buffers too small");
    if (offsets) offsets[0] = 0;
    if (nc == 0) return RMCV_OK;
    const size_t b_meta = (((size_t)nc * 3 + 1) * 4 + 15) & ~(size_t)15, b_xy = (size_t)total * 8;
    rc = ensure_tmp(ctx, b_meta + b_xy + 16, b_meta + b_xy + 16);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    int32_t* h = reinterpret_cast<int32_t*>(ex->tmp_host);  // [offsets nc+1 | starts 2nc]
    h[0] = 0;
    for (int k = 0; k < nc; ++k) {
        const rmcv_contour_info& ci = ctx->h_contours[fi.contour_offset + k];
        h[k + 1] = h[k] + ci.n_points;
        h[nc + 1 + 2 * k] = ci.first_x;
        h[nc + 1 + 2 * k + 1] = ci.first_y;
    }
    for (int k = 0; k <= nc; ++k) offsets[k] = h[k];
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = extra(ctx)->lab;
    Geometry g = call_geometry(ctx, ctx->last_W, ctx->last_H);
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, h, ((size_t)nc * 3 + 1) * 4, cudaMemcpyHostToDevice, st));
    const int32_t* d_off = reinterpret_cast<const int32_t*>(d);
    RMCV_CUDA(ctx, launch_trace_all(g, sb->bits + (size_t)local * g.H * g.WB, d_off + nc + 1, d_off, nc,
                                    reinterpret_cast<int32_t*>(d + b_meta), st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(xy, d + b_meta, b_xy, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    return RMCV_OK;
}

int rmcv_get_label_map(rmcv_ctx* ctx, int frame, int32_t* labels, size_t pitch_elems) {
    if (!ctx || !labels) return RMCV_ERR_INVALID_ARG;
    int rc = wait_latest_call(ctx);
    if (rc != RMCV_OK) return rc;
    if (!ctx->have_results) return set_err(ctx, RMCV_ERR_STATE, "rmcv_get_label_map needs a detect call first");
    int local = 0;
    SlotBuffers* sb = resident_slot(ctx, frame, &local);
    if (!sb) return set_err(ctx, RMCV_ERR_STATE, "frame scratch no longer resident (only the last two chunks are kept)");
    Geometry g = call_geometry(ctx, ctx->last_W, ctx->last_H);
    if (pitch_elems < (size_t)g.W) return set_err(ctx, RMCV_ERR_INVALID_ARG, "label pitch smaller than a row");
    const size_t bytes = (size_t)g.W * g.H * sizeof(int32_t);
    rc = ensure_tmp(ctx, bytes, 0);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    cudaStream_t st = extra(ctx)->lab;
    RMCV_CUDA(ctx, launch_label_map(g, sb, local, reinterpret_cast<int32_t*>(ex->tmp_dev), st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpy2DAsync(labels, pitch_elems * sizeof(int32_t), ex->tmp_dev, (size_t)g.W * sizeof(int32_t),
                                     (size_t)g.W * sizeof(int32_t), g.H, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    return RMCV_OK;
}

int rmcv_get_bitmask(rmcv_ctx* ctx, int frame, uint32_t* words, int words_per_row) {
    if (!ctx || !words) return RMCV_ERR_INVALID_ARG;
    int rc = sync_all(ctx);
    if (rc != RMCV_OK) return rc;
    if (extra(ctx)->last_kind == 0) return set_err(ctx, RMCV_ERR_STATE, "no extract/detect call yet");
    int local = 0;
    SlotBuffers* sb = resident_slot(ctx, frame, &local);
    if (!sb) return set_err(ctx, RMCV_ERR_STATE, "frame scratch no longer resident (only the last two chunks are kept)");
    Geometry g = call_geometry(ctx, ctx->last_W, ctx->last_H);
    if (words_per_row < g.WB) return set_err(ctx, RMCV_ERR_INVALID_ARG, "words_per_row too small");
    RMCV_CUDA(ctx, cudaMemcpy2DAsync(words, (size_t)words_per_row * 4, sb->bits + (size_t)local * g.H * g.WB, (size_t)g.WB * 4,
                                     (size_t)g.WB * 4, g.H, cudaMemcpyDeviceToHost, extra(ctx)->lab));
    RMCV_CUDA(ctx, cudaStreamSynchronize(extra(ctx)->lab));
    return RMCV_OK;
}

// offsets of caller-supplied contours: offsets[0] == 0, non-decreasing (the kernels index xy and their scratch with them)
static int check_offsets(rmcv_ctx* ctx, const int32_t* offsets, int n) {
    if (offsets[0] != 0) return set_err(ctx, RMCV_ERR_INVALID_ARG, "contour offsets must start at 0");
    for (int k = 0; k < n; ++k)
        if (offsets[k + 1] < offsets[k]) return set_err(ctx, RMCV_ERR_INVALID_ARG, "contour offsets must be non-decreasing");
    return RMCV_OK;
}

int rmcv_filter_lightblobs(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours, const rmcv_params* params,
                           rmcv_contour_info* infos, rmcv_lightblob* blobs, int blob_cap, int* n_blobs) {
    if (!ctx || !params || !offsets || n_contours < 0 || !n_blobs || (n_contours > 0 && !infos)) return RMCV_ERR_INVALID_ARG;
    *n_blobs = 0;
    if (n_contours == 0) return RMCV_OK;
    if (int orc = check_offsets(ctx, offsets, n_contours)) return orc;
    const size_t npts = (size_t)offsets[n_contours];
    if (npts > 0 && !xy) return RMCV_ERR_INVALID_ARG;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    // device layout: xy | offsets | infos | blobs
    const size_t b_xy = ((npts * 2 * 4) + 15) & ~(size_t)15, b_off = (((size_t)n_contours + 1) * 4 + 15) & ~(size_t)15;
    const size_t b_info = (size_t)n_contours * sizeof(rmcv_contour_info), b_blob = (size_t)n_contours * sizeof(rmcv_lightblob);
    int rc = ensure_tmp(ctx, b_xy + b_off + b_info + b_blob + 64, b_info + b_blob + 64);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = extra(ctx)->pix;
    if (npts) RMCV_CUDA(ctx, cudaMemcpyAsync(d, xy, npts * 2 * 4, cudaMemcpyHostToDevice, st));
    RMCV_CUDA(ctx, cudaMemcpyAsync(d + b_xy, offsets, ((size_t)n_contours + 1) * 4, cudaMemcpyHostToDevice, st));
    rmcv_contour_info* d_info = reinterpret_cast<rmcv_contour_info*>(d + b_xy + b_off);
    rmcv_lightblob* d_blob = reinterpret_cast<rmcv_lightblob*>(d + b_xy + b_off + b_info);
    RMCV_CUDA(ctx, launch_filter_lightblobs(reinterpret_cast<const int32_t*>(d), reinterpret_cast<const int32_t*>(d + b_xy), n_contours,
                                            *params, d_info, d_blob, st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, d_info, b_info + b_blob, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    const rmcv_contour_info* h_info = reinterpret_cast<const rmcv_contour_info*>(ex->tmp_host);
    const rmcv_lightblob* h_blob = reinterpret_cast<const rmcv_lightblob*>(reinterpret_cast<const uint8_t*>(ex->tmp_host) + b_info);
    int nb = 0;
    for (int i = 0; i < n_contours; ++i) {
        infos[i] = h_info[i];
        if (h_info[i].status == RMCV_CONTOUR_POSITIVE) {
            infos[i].blob_index = nb;
            if (blobs && nb < blob_cap) blobs[nb] = h_blob[i];
            ++nb;
        }
    }
    *n_blobs = nb;
    return (blobs && nb > blob_cap) ? set_err(ctx, RMCV_ERR_CAPACITY, "blob_cap too small") : RMCV_OK;
}

int rmcv_filter_armours(rmcv_ctx* ctx, const rmcv_lightblob* blobs, int n_blobs, const rmcv_params* params, rmcv_armour* armours,
                        int armour_cap, int* n_armours) {
    if (!ctx || !params || n_blobs < 0 || !n_armours || armour_cap < 0 || (n_blobs > 0 && !blobs)) return RMCV_ERR_INVALID_ARG;
    *n_armours = 0;
    if (n_blobs < 2) return RMCV_OK;  // src/objdetect.cpp:120
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b_blob = ((size_t)n_blobs * sizeof(rmcv_lightblob) + 15) & ~(size_t)15;
    const size_t b_arm = (size_t)(armour_cap > 0 ? armour_cap : 1) * sizeof(rmcv_armour);
    int rc = ensure_tmp(ctx, b_blob + b_arm + 64, b_arm + 64);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = extra(ctx)->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, blobs, (size_t)n_blobs * sizeof(rmcv_lightblob), cudaMemcpyHostToDevice, st));
    int32_t* d_count = reinterpret_cast<int32_t*>(d + b_blob);
    rmcv_armour* d_arm = reinterpret_cast<rmcv_armour*>(d + b_blob + 16);
    RMCV_CUDA(ctx, launch_filter_armours(reinterpret_cast<const rmcv_lightblob*>(d), n_blobs, *params, d_arm, armour_cap, d_count, st,
                                         &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, d + b_blob, 16 + b_arm, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    const int total = *reinterpret_cast<const int32_t*>(ex->tmp_host);
    *n_armours = total;
    const int ncopy = total < armour_cap ? total : armour_cap;
    if (ncopy > 0 && armours) memcpy(armours, reinterpret_cast<const uint8_t*>(ex->tmp_host) + 16, (size_t)ncopy * sizeof(rmcv_armour));
    return total > armour_cap ? set_err(ctx, RMCV_ERR_CAPACITY, "armour_cap too small") : RMCV_OK;
}

int rmcv_make_lightblobs(rmcv_ctx* ctx, const rmcv_rotated_rect* boxes, int n, int target, rmcv_lightblob* out) {
    if (!ctx || n < 0 || (n > 0 && (!boxes || !out))) return RMCV_ERR_INVALID_ARG;
    if (n == 0) return RMCV_OK;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b_in = ((size_t)n * sizeof(rmcv_rotated_rect) + 15) & ~(size_t)15, b_out = (size_t)n * sizeof(rmcv_lightblob);
    int rc = ensure_tmp(ctx, b_in + b_out, b_out);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = extra(ctx)->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, boxes, (size_t)n * sizeof(rmcv_rotated_rect), cudaMemcpyHostToDevice, st));
    RMCV_CUDA(ctx, launch_make_lightblobs(reinterpret_cast<const rmcv_rotated_rect*>(d), n, target, reinterpret_cast<rmcv_lightblob*>(d + b_in),
                                          st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, d + b_in, b_out, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(out, ex->tmp_host, b_out);
    return RMCV_OK;
}

// shared by the three legacy contour entry points: upload, one launch, download
static int run_legacy(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n, float min_ratio, float max_ratio,
                      float tilt_angle, float min_area, float max_area, int fit_ellipse, const uint8_t* h_src, size_t pitch,
                      int width, int height, int32_t* matched, rmcv_rotated_rect* boxes, int32_t* camps, rmcv_lightblob* blobs) {
    if (!ctx || !offsets || n < 0) return RMCV_ERR_INVALID_ARG;
    if (n == 0) return RMCV_OK;
    if (int orc = check_offsets(ctx, offsets, n)) return orc;
    const size_t npts = (size_t)offsets[n];
    if (npts > 0 && !xy) return RMCV_ERR_INVALID_ARG;
    if (h_src && (width <= 0 || height <= 0 || pitch < (size_t)width * 3)) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad source image geometry");
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t b_xy = up16(npts * 8 + 8), b_off = up16(((size_t)n + 1) * 4), b_hull = up16(npts * 24 + 24);
    const size_t b_m = up16((size_t)n * 4), b_box = up16((size_t)n * sizeof(rmcv_rotated_rect)), b_c = up16((size_t)n * 4);
    const size_t b_blob = up16((size_t)n * sizeof(rmcv_lightblob));
    const size_t b_src = h_src ? up16((size_t)height * width * 3) : 0;
    const size_t b_out = b_m + b_box + b_c + b_blob;
    int rc = ensure_tmp(ctx, b_xy + b_off + b_hull + b_out + b_src + 64, b_out + 64);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = ex->pix;
    if (npts) RMCV_CUDA(ctx, cudaMemcpyAsync(d, xy, npts * 8, cudaMemcpyHostToDevice, st));
    RMCV_CUDA(ctx, cudaMemcpyAsync(d + b_xy, offsets, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, st));
    uint8_t* d_out = d + b_xy + b_off + b_hull;
    uint8_t* d_src = d_out + b_out;
    if (h_src)
        RMCV_CUDA(ctx, cudaMemcpy2DAsync(d_src, (size_t)width * 3, h_src, pitch, (size_t)width * 3, height, cudaMemcpyHostToDevice, st));
    LegacyLaunch L;
    L.xy = reinterpret_cast<const int32_t*>(d); L.off = reinterpret_cast<const int32_t*>(d + b_xy); L.n_contours = n;
    L.min_ratio = min_ratio; L.max_ratio = max_ratio; L.tilt_angle = tilt_angle; L.min_area = min_area; L.max_area = max_area;
    L.fit_ellipse = fit_ellipse;
    L.src = h_src ? d_src : nullptr; L.pitch = (size_t)width * 3; L.W = width; L.H = height;
    L.hull = reinterpret_cast<int32_t*>(d + b_xy + b_off);
    L.matched = reinterpret_cast<int32_t*>(d_out);
    L.boxes = reinterpret_cast<rmcv_rotated_rect*>(d_out + b_m);
    L.camps = reinterpret_cast<int32_t*>(d_out + b_m + b_box);
    L.blobs = reinterpret_cast<rmcv_lightblob*>(d_out + b_m + b_box + b_c);
    if (fit_ellipse < 0) RMCV_CUDA(ctx, cudaMemsetAsync(d_out, 0, b_out, st));
    RMCV_CUDA(ctx, launch_legacy(L, st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, d_out, b_out, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    const uint8_t* h = reinterpret_cast<const uint8_t*>(ex->tmp_host);
    if (matched) memcpy(matched, h, (size_t)n * 4);
    if (boxes) memcpy(boxes, h + b_m, (size_t)n * sizeof(rmcv_rotated_rect));
    if (camps) memcpy(camps, h + b_m + b_box, (size_t)n * 4);
    if (blobs) memcpy(blobs, h + b_m + b_box + b_c, (size_t)n * sizeof(rmcv_lightblob));
    return RMCV_OK;
}

int rmcv_match_lightblobs(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours, float min_ratio,
                          float max_ratio, float tilt_angle, float min_area, float max_area, int fit_ellipse, int32_t* matched,
                          rmcv_rotated_rect* boxes) {
    if (!matched || !boxes) return RMCV_ERR_INVALID_ARG;
    return run_legacy(ctx, xy, offsets, n_contours, min_ratio, max_ratio, tilt_angle, min_area, max_area, fit_ellipse ? 1 : 0,
                      nullptr, 0, 0, 0, matched, boxes, nullptr, nullptr);
}

int rmcv_find_lightblobs_legacy(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours, float min_ratio,
                                float max_ratio, float tilt_angle, float min_area, float max_area, const uint8_t* h_source_bgr,
                                size_t pitch, int width, int height, int fit_ellipse, rmcv_lightblob* blobs, int blob_cap,
                                int* n_blobs) {
    if (!ctx || !n_blobs || !h_source_bgr || blob_cap < 0) return RMCV_ERR_INVALID_ARG;
    *n_blobs = 0;
    if (n_contours <= 0) return n_contours < 0 ? RMCV_ERR_INVALID_ARG : RMCV_OK;
    std::vector<int32_t> matched((size_t)n_contours);
    std::vector<rmcv_lightblob> all((size_t)n_contours);
    int rc = run_legacy(ctx, xy, offsets, n_contours, min_ratio, max_ratio, tilt_angle, min_area, max_area, fit_ellipse ? 1 : 0,
                        h_source_bgr, pitch, width, height, matched.data(), nullptr, nullptr, all.data());
    if (rc != RMCV_OK) return rc;
    int nb = 0;
    for (int k = 0; k < n_contours; ++k)
        if (matched[(size_t)k]) { if (blobs && nb < blob_cap) blobs[nb] = all[(size_t)k]; ++nb; }
    *n_blobs = nb;
    return (blobs && nb > blob_cap) ? set_err(ctx, RMCV_ERR_CAPACITY, "blob_cap too small") : RMCV_OK;
}

int rmcv_min_area_rects(rmcv_ctx* ctx, const int32_t* xy, const int32_t* offsets, int n_contours, rmcv_rotated_rect* boxes) {
    if (!boxes) return RMCV_ERR_INVALID_ARG;
    return run_legacy(ctx, xy, offsets, n_contours, 0.f, 0.f, 0.f, 0.f, 0.f, -1, nullptr, 0, 0, 0, nullptr, boxes, nullptr, nullptr);
}

int rmcv_lightblob_overlap(rmcv_ctx* ctx, const rmcv_lightblob* blobs, int n_blobs, int left, int right, int* overlap) {
    if (!ctx || !overlap || n_blobs < 0 || (n_blobs > 0 && !blobs)) return RMCV_ERR_INVALID_ARG;
    *overlap = 0;
    if (n_blobs == 0) return RMCV_OK;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b_blob = ((size_t)n_blobs * sizeof(rmcv_lightblob) + 15) & ~(size_t)15;
    int rc = ensure_tmp(ctx, b_blob + 64, 64);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = ex->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, blobs, (size_t)n_blobs * sizeof(rmcv_lightblob), cudaMemcpyHostToDevice, st));
    int32_t* d_out = reinterpret_cast<int32_t*>(d + b_blob);
    RMCV_CUDA(ctx, launch_overlap(reinterpret_cast<const rmcv_lightblob*>(d), n_blobs, left, right, d_out, st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, d_out, 4, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    *overlap = *reinterpret_cast<const int32_t*>(ex->tmp_host);
    return RMCV_OK;
}

int rmcv_raw_frontend_batch(rmcv_ctx* ctx, const void* d_raw, size_t pitch, size_t frame_stride, int width, int height, int batch,
                            int bits, int mirror, int flip, uint8_t* d_raw8, size_t out_pitch, size_t out_frame_stride) {
    if (!ctx || !d_raw || !d_raw8) return RMCV_ERR_INVALID_ARG;
    if (width <= 0 || height <= 0 || batch <= 0) return set_err(ctx, RMCV_ERR_INVALID_ARG, "width, height and batch must be positive");
    if (bits != 8 && bits != 10 && bits != 12) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bits must be 8, 10 or 12");
    const size_t rowbytes = (size_t)width * (bits > 8 ? 2 : 1);
    if (pitch < rowbytes || out_pitch < (size_t)width) return set_err(ctx, RMCV_ERR_INVALID_ARG, "pitch smaller than a row");
    if (order_pix_after_chains(ctx) != RMCV_OK) return RMCV_ERR_CUDA;   // d_raw8 may be the input of a small call still in flight
    if (bits > 8 && ((pitch | frame_stride | reinterpret_cast<size_t>(d_raw)) & 1)) return set_err(ctx, RMCV_ERR_INVALID_ARG, "16-bit rows must be 2-byte aligned");
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    RMCV_CUDA(ctx, launch_frontend(static_cast<const uint8_t*>(d_raw), pitch, frame_stride, d_raw8, out_pitch, out_frame_stride, width,
                                   height, batch, bits, mirror ? 1 : 0, flip ? 1 : 0, extra(ctx)->pix, &ctx->kernel_launches));
    return RMCV_OK;
}

int rmcv_frontend_layout(int layout, int width, int height, int mirror, int flip) {
    // colour at (row parity, column parity): RG = [[R,G],[G,B]], GB = [[G,B],[R,G]], GR = [[G,R],[B,G]], BG = [[B,G],[G,R]]
    if (layout < RMCV_BAYER_RG || layout > RMCV_BAYER_BG) return RMCV_ERR_INVALID_ARG;
    // a mirror of an even-width frame swaps the two columns of the 2x2 cell, a flip of an even-height frame its rows
    const bool swap_cols = mirror && (width % 2 == 0), swap_rows = flip && (height % 2 == 0);
    static const int col_swapped[5] = {0, RMCV_BAYER_GR, RMCV_BAYER_BG, RMCV_BAYER_RG, RMCV_BAYER_GB};
    static const int row_swapped[5] = {0, RMCV_BAYER_GB, RMCV_BAYER_RG, RMCV_BAYER_BG, RMCV_BAYER_GR};
    if (swap_cols) layout = col_swapped[layout];
    if (swap_rows) layout = row_swapped[layout];
    return layout;
}

int rmcv_set_camera(rmcv_ctx* ctx, const double camera_matrix[9], const double dist_coeffs[5], float exact_w, float exact_h,
                    const double* cam2world) {
    if (!ctx || !camera_matrix) return RMCV_ERR_INVALID_ARG;
    if (exact_w != exact_h || !(exact_w > 0.f)) return set_err(ctx, RMCV_ERR_INVALID_ARG, "IPPE_SQUARE needs a square object of positive size");
    CtxExtra* ex = extra(ctx);
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < kResultSets; ++i) {   // the pose arrays exist only for callers that ask for poses
        ResultSet& r = ex->rs[i];
        if (r.frames && !r.poses)
            RMCV_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&r.poses), (size_t)ctx->cfg.max_batch * ctx->cap.A * sizeof(rmcv_pose),
                                         cudaHostAllocMapped | cudaHostAllocPortable));
    }
    CameraSetup& c = ex->camera;
    for (int i = 0; i < 9; ++i) c.K[i] = camera_matrix[i];
    for (int i = 0; i < 5; ++i) c.dist[i] = dist_coeffs ? dist_coeffs[i] : 0.0;
    c.has_M = cam2world != nullptr;
    for (int i = 0; i < 16; ++i) c.M[i] = cam2world ? cam2world[i] : 0.0;
    c.w = exact_w; c.h = exact_h;
    ex->have_camera = true;
    return RMCV_OK;
}

int rmcv_clear_camera(rmcv_ctx* ctx) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    extra(ctx)->have_camera = false;
    return RMCV_OK;
}

int rmcv_solve_pnp(rmcv_ctx* ctx, const rmcv_armour* armours, int n_armours, const double camera_matrix[9],
                   const double dist_coeffs[5], float exact_w, float exact_h, float roi_x, float roi_y, const double* cam2world,
                   rmcv_pose* poses) {
    if (!ctx || n_armours < 0 || !camera_matrix || (n_armours > 0 && (!armours || !poses))) return RMCV_ERR_INVALID_ARG;
    if (exact_w != exact_h || !(exact_w > 0.f)) return set_err(ctx, RMCV_ERR_INVALID_ARG, "IPPE_SQUARE needs a square object of positive size");
    if (n_armours == 0) return RMCV_OK;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b_in = ((size_t)n_armours * sizeof(rmcv_armour) + 15) & ~(size_t)15, b_out = (size_t)n_armours * sizeof(rmcv_pose);
    int rc = ensure_tmp(ctx, b_in + b_out, b_out);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    cudaStream_t st = ex->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, armours, (size_t)n_armours * sizeof(rmcv_armour), cudaMemcpyHostToDevice, st));
    RMCV_CUDA(ctx, launch_pnp(reinterpret_cast<const rmcv_armour*>(d), n_armours, camera_matrix, dist_coeffs, exact_w, exact_h, roi_x,
                              roi_y, cam2world, reinterpret_cast<rmcv_pose*>(d + b_in), st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(ex->tmp_host, d + b_in, b_out, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(poses, ex->tmp_host, b_out);
    return RMCV_OK;
}

int rmcv_icon_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, int width, int height, rmcv_armour* armours, int n_armours,
                    int out_w, int out_h, uint8_t* icons, float* rows) {
    if (!ctx || !d_bgr || n_armours < 0 || (n_armours > 0 && (!armours || !icons))) return RMCV_ERR_INVALID_ARG;
    if (width <= 0 || height <= 0 || out_w <= 0 || out_h <= 0 || out_w > 4096 || out_h > 4096 || pitch < (size_t)width * 3)
        return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad frame or icon geometry");
    if (n_armours == 0) return RMCV_OK;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_armours, npx = (size_t)out_w * out_h * 3;
    const size_t b_arm = (n * sizeof(rmcv_armour) + 15) & ~(size_t)15, b_icon = (n * npx + 15) & ~(size_t)15,
                 b_rows = rows ? n * npx * sizeof(float) : 0;
    int rc = ensure_tmp(ctx, b_arm + b_icon + b_rows, b_arm + b_icon + b_rows);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* h = reinterpret_cast<uint8_t*>(ex->tmp_host);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    memcpy(h, armours, n * sizeof(rmcv_armour));
    cudaStream_t st = ex->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, h, n * sizeof(rmcv_armour), cudaMemcpyHostToDevice, st));
    RMCV_CUDA(ctx, launch_icons(d_bgr, pitch, width, height, reinterpret_cast<rmcv_armour*>(d), n_armours, out_w, out_h, d + b_arm,
                                rows ? reinterpret_cast<float*>(d + b_arm + b_icon) : nullptr, st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(h, d, b_arm + b_icon + b_rows, cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(armours, h, n * sizeof(rmcv_armour));
    memcpy(icons, h + b_arm, n * npx);
    if (rows) memcpy(rows, h + b_arm + b_icon, n * npx * sizeof(float));
    return RMCV_OK;
}

namespace {
struct SvmDevice {   // model arrays inside one device allocation
    const float* sv; const double* rho; const int32_t* df_ofs; const double* df_alpha; const int32_t* df_index; const int32_t* labels;
    size_t bytes;
};
int check_svm_model(rmcv_ctx* ctx, const rmcv_svm_model* m) {
    if (!m || !m->support_vectors || !m->class_labels || !m->rho || !m->df_ofs || !m->df_alpha || !m->df_index)
        return set_err(ctx, RMCV_ERR_INVALID_ARG, "incomplete svm model");
    if (m->var_count <= 0 || m->class_count < 2 || m->class_count > 32 || m->sv_total <= 0)
        return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad svm model sizes");
    const int ndf = m->class_count * (m->class_count - 1) / 2;
    for (int k = 0; k < ndf; ++k) if (m->df_ofs[k + 1] < m->df_ofs[k] || m->df_ofs[k] < 0) return set_err(ctx, RMCV_ERR_INVALID_ARG, "bad svm decision function offsets");
    for (int k = m->df_ofs[0]; k < m->df_ofs[ndf]; ++k)
        if (m->df_index[k] < 0 || m->df_index[k] >= m->sv_total) return set_err(ctx, RMCV_ERR_INVALID_ARG, "svm support vector index out of range");
    return RMCV_OK;
}
// lays the model out at d (device) through the staging area h (pinned host); returns the device views
SvmDevice stage_svm_model(const rmcv_svm_model* m, uint8_t* h, uint8_t* d, bool copy) {
    const int ndf = m->class_count * (m->class_count - 1) / 2, nk = m->df_ofs[ndf];
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0;
    SvmDevice v;
    const size_t b_sv = (size_t)m->sv_total * m->var_count * sizeof(float);
    v.sv = reinterpret_cast<const float*>(d + o); if (copy) memcpy(h + o, m->support_vectors, b_sv); o = al(o + b_sv);
    v.rho = reinterpret_cast<const double*>(d + o); if (copy) memcpy(h + o, m->rho, ndf * sizeof(double)); o = al(o + ndf * sizeof(double));
    v.df_alpha = reinterpret_cast<const double*>(d + o); if (copy) memcpy(h + o, m->df_alpha, nk * sizeof(double)); o = al(o + nk * sizeof(double));
    v.df_ofs = reinterpret_cast<const int32_t*>(d + o); if (copy) memcpy(h + o, m->df_ofs, (ndf + 1) * sizeof(int32_t)); o = al(o + (ndf + 1) * sizeof(int32_t));
    v.df_index = reinterpret_cast<const int32_t*>(d + o); if (copy) memcpy(h + o, m->df_index, nk * sizeof(int32_t)); o = al(o + nk * sizeof(int32_t));
    v.labels = reinterpret_cast<const int32_t*>(d + o); if (copy) memcpy(h + o, m->class_labels, m->class_count * sizeof(int32_t)); o = al(o + m->class_count * sizeof(int32_t));
    v.bytes = o;
    return v;
}
}  // namespace

int rmcv_svm_predict(rmcv_ctx* ctx, const rmcv_svm_model* model, const float* rows, int n, int32_t* labels) {
    if (!ctx || n < 0 || (n > 0 && (!rows || !labels))) return RMCV_ERR_INVALID_ARG;
    int rc = check_svm_model(ctx, model);
    if (rc != RMCV_OK) return rc;
    if (n == 0) return RMCV_OK;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b_model = stage_svm_model(model, nullptr, nullptr, false).bytes;
    const size_t b_rows = ((size_t)n * model->var_count * sizeof(float) + 15) & ~(size_t)15;
    const size_t b_k = ((size_t)n * model->sv_total * sizeof(float) + 15) & ~(size_t)15, b_lab = ((size_t)n * sizeof(int32_t) + 15) & ~(size_t)15;
    rc = ensure_tmp(ctx, b_model + b_rows + b_k + b_lab, b_model + b_rows + b_lab);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* h = reinterpret_cast<uint8_t*>(ex->tmp_host);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    const SvmDevice v = stage_svm_model(model, h, d, true);
    memcpy(h + b_model, rows, (size_t)n * model->var_count * sizeof(float));
    cudaStream_t st = ex->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, h, b_model + b_rows, cudaMemcpyHostToDevice, st));
    int32_t* d_lab = reinterpret_cast<int32_t*>(d + b_model + b_rows + b_k);
    RMCV_CUDA(ctx, launch_svm_predict(reinterpret_cast<const float*>(d + b_model), n, v.sv, model->sv_total, model->var_count, v.rho, v.df_ofs,
                                      v.df_alpha, v.df_index, v.labels, model->class_count, reinterpret_cast<float*>(d + b_model + b_rows),
                                      d_lab, st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(h, d_lab, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(labels, h, (size_t)n * sizeof(int32_t));
    return RMCV_OK;
}

int rmcv_identify_batch(rmcv_ctx* ctx, const uint8_t* d_bgr, size_t pitch, int width, int height, rmcv_armour* armours, int n_armours,
                        int out_w, int out_h, const rmcv_svm_model* model, int32_t* identities) {
    if (!ctx || !d_bgr || n_armours < 0 || (n_armours > 0 && (!armours || !identities))) return RMCV_ERR_INVALID_ARG;
    int rc = check_svm_model(ctx, model);
    if (rc != RMCV_OK) return rc;
    if (width <= 0 || height <= 0 || out_w <= 0 || out_h <= 0 || pitch < (size_t)width * 3 || (long long)out_w * out_h * 3 != model->var_count)
        return set_err(ctx, RMCV_ERR_INVALID_ARG, "icon size does not match the model's var_count");
    if (n_armours == 0) return RMCV_OK;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_armours, npx = (size_t)model->var_count;
    const size_t b_model = stage_svm_model(model, nullptr, nullptr, false).bytes;
    const size_t b_arm = (n * sizeof(rmcv_armour) + 15) & ~(size_t)15, b_icon = (n * npx + 15) & ~(size_t)15, b_rows = n * npx * sizeof(float);
    const size_t b_k = (n * model->sv_total * sizeof(float) + 15) & ~(size_t)15, b_lab = (n * sizeof(int32_t) + 15) & ~(size_t)15;
    rc = ensure_tmp(ctx, b_model + b_arm + b_icon + b_rows + b_k + b_lab, b_model + b_arm + b_lab);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* h = reinterpret_cast<uint8_t*>(ex->tmp_host);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    const SvmDevice v = stage_svm_model(model, h, d, true);
    memcpy(h + b_model, armours, n * sizeof(rmcv_armour));
    cudaStream_t st = ex->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, h, b_model + b_arm, cudaMemcpyHostToDevice, st));
    rmcv_armour* d_arm = reinterpret_cast<rmcv_armour*>(d + b_model);
    float* d_rows = reinterpret_cast<float*>(d + b_model + b_arm + b_icon);
    int32_t* d_lab = reinterpret_cast<int32_t*>(d + b_model + b_arm + b_icon + b_rows + b_k);
    RMCV_CUDA(ctx, launch_icons(d_bgr, pitch, width, height, d_arm, n_armours, out_w, out_h, d + b_model + b_arm, d_rows, st, &ctx->kernel_launches));
    RMCV_CUDA(ctx, launch_svm_predict(d_rows, n_armours, v.sv, model->sv_total, model->var_count, v.rho, v.df_ofs, v.df_alpha, v.df_index, v.labels,
                                      model->class_count, reinterpret_cast<float*>(d + b_model + b_arm + b_icon + b_rows), d_lab, st,
                                      &ctx->kernel_launches));
    RMCV_CUDA(ctx, cudaMemcpyAsync(h, d_arm, n * sizeof(rmcv_armour), cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaMemcpyAsync(h + b_arm, d_lab, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(armours, h, n * sizeof(rmcv_armour));
    memcpy(identities, h + b_arm, n * sizeof(int32_t));
    return RMCV_OK;
}

// ---- f3: tracking ----------------------------------------------------------------------------------------------------
}  // extern "C"
struct rmcv_tracker {
    int cap;
    rmcv_track* tracks; rmcv_track* backup;   // [cap] each, device
    int32_t* n_tracks; int32_t* status;       // device
};
extern "C" {

int rmcv_tracker_create(rmcv_ctx* ctx, int capacity, rmcv_tracker** out) {
    if (!ctx || !out || capacity <= 0) return RMCV_ERR_INVALID_ARG;
    *out = nullptr;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    rmcv_tracker* t = new (std::nothrow) rmcv_tracker();
    if (!t) return set_err(ctx, RMCV_ERR_CUDA, "out of host memory");
    t->cap = capacity; t->tracks = t->backup = nullptr; t->n_tracks = t->status = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&t->tracks), (size_t)capacity * sizeof(rmcv_track));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->backup), (size_t)capacity * sizeof(rmcv_track));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->n_tracks), 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(t->n_tracks, 0, 2 * sizeof(int32_t));
    if (e != cudaSuccess) {
        cudaFree(t->tracks); cudaFree(t->backup); cudaFree(t->n_tracks);
        delete t;
        snprintf(ctx->err, sizeof(ctx->err), "tracker allocation failed: %s", cudaGetErrorString(e));
        return RMCV_ERR_CUDA;
    }
    t->status = t->n_tracks + 1;
    *out = t;
    return RMCV_OK;
}

int rmcv_tracker_destroy(rmcv_ctx* ctx, rmcv_tracker* t) {
    if (!ctx || !t) return RMCV_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(extra(ctx)->pix);
    cudaFree(t->tracks); cudaFree(t->backup); cudaFree(t->n_tracks);
    delete t;
    return RMCV_OK;
}

int rmcv_tracker_reset(rmcv_ctx* ctx, rmcv_tracker* t) {
    if (!ctx || !t) return RMCV_ERR_INVALID_ARG;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    RMCV_CUDA(ctx, cudaMemsetAsync(t->n_tracks, 0, 2 * sizeof(int32_t), extra(ctx)->pix));
    RMCV_CUDA(ctx, cudaStreamSynchronize(extra(ctx)->pix));
    return RMCV_OK;
}

int rmcv_tracker_update(rmcv_ctx* ctx, rmcv_tracker* t, const rmcv_armour* armours, const double* positions, const int32_t* identities,
                        int n_armours, int64_t timestamp, double tick_frequency, double process_noise, double measurement_noise,
                        double error) {
    if (!ctx || !t || n_armours < 0 || (n_armours > 0 && (!armours || !positions))) return RMCV_ERR_INVALID_ARG;
    if (!(tick_frequency > 0.0)) return set_err(ctx, RMCV_ERR_INVALID_ARG, "tick_frequency must be positive");
    if (n_armours == 0) return RMCV_OK;   // executable/main.cpp:63: an empty frame changes nothing
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_armours;
    const size_t b_arm = (n * sizeof(rmcv_armour) + 15) & ~(size_t)15, b_pos = (n * 3 * sizeof(double) + 15) & ~(size_t)15,
                 b_id = (n * sizeof(int32_t) + 15) & ~(size_t)15;
    int rc = ensure_tmp(ctx, b_arm + b_pos + 2 * b_id, b_arm + b_pos + b_id + 16);
    if (rc != RMCV_OK) return rc;
    CtxExtra* ex = extra(ctx);
    uint8_t* h = reinterpret_cast<uint8_t*>(ex->tmp_host);
    uint8_t* d = reinterpret_cast<uint8_t*>(ex->tmp_dev);
    memcpy(h, armours, n * sizeof(rmcv_armour));
    memcpy(h + b_arm, positions, n * 3 * sizeof(double));
    if (identities) memcpy(h + b_arm + b_pos, identities, n * sizeof(int32_t));
    cudaStream_t st = ex->pix;
    RMCV_CUDA(ctx, cudaMemcpyAsync(d, h, b_arm + b_pos + b_id, cudaMemcpyHostToDevice, st));
    RMCV_CUDA(ctx, launch_track_update(t->tracks, t->n_tracks, t->cap, t->backup, reinterpret_cast<const rmcv_armour*>(d),
                                       reinterpret_cast<const double*>(d + b_arm),
                                       identities ? reinterpret_cast<const int32_t*>(d + b_arm + b_pos) : nullptr, n_armours,
                                       reinterpret_cast<int32_t*>(d + b_arm + b_pos + b_id), (long long)timestamp, tick_frequency,
                                       process_noise, measurement_noise, error, t->status, st, &ctx->kernel_launches));
    int32_t* hs = reinterpret_cast<int32_t*>(h + b_arm + b_pos + b_id);
    RMCV_CUDA(ctx, cudaMemcpyAsync(hs, t->status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    if (*hs == 1) return set_err(ctx, RMCV_ERR_CAPACITY, "track list capacity exceeded");
    if (*hs == 2) return set_err(ctx, RMCV_ERR_CAPACITY, "identity history of a track exceeded RMCV_TRACK_HIST");
    return RMCV_OK;
}

int rmcv_tracker_read(rmcv_ctx* ctx, rmcv_tracker* t, rmcv_track* tracks, int cap, int* n_tracks) {
    if (!ctx || !t || !n_tracks || cap < 0 || (cap > 0 && !tracks)) return RMCV_ERR_INVALID_ARG;
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = extra(ctx)->pix;
    int32_t n = 0;
    RMCV_CUDA(ctx, cudaMemcpyAsync(&n, t->n_tracks, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    *n_tracks = n;
    const int m = n < cap ? n : cap;
    if (m > 0) {
        RMCV_CUDA(ctx, cudaMemcpyAsync(tracks, t->tracks, (size_t)m * sizeof(rmcv_track), cudaMemcpyDeviceToHost, st));
        RMCV_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return RMCV_OK;
}

int rmcv_track_identity_max(const rmcv_track* track, int32_t* identity, double* probability) {
    if (!track || !identity || !probability) return RMCV_ERR_INVALID_ARG;
    double sum = 0.0;                                            // src/core.cpp:123-143
    for (int i = 0; i < track->n_hist; ++i) sum += exp((double)track->hist_count[i]);
    double best = 0.0; int32_t id = -1;
    for (int i = 0; i < track->n_hist; ++i) {
        const double prob = exp((double)track->hist_count[i]) / sum;
        if (prob > best) { best = prob; id = track->hist_id[i]; }
    }
    *identity = id; *probability = best;
    return RMCV_OK;
}

int rmcv_profile_enable(rmcv_ctx* ctx, int on) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    ctx->profiling = on != 0;
    return RMCV_OK;
}

int rmcv_profile_read(rmcv_ctx* ctx, double ms[RMCV_STAGE_COUNT], int64_t launches[RMCV_STAGE_COUNT], int reset) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    const int rc0 = sync_all(ctx);  // collects the stage events of every finished call
    if (rc0 != RMCV_OK) return rc0;
    for (int s = 0; s < RMCV_STAGE_COUNT; ++s) {
        if (ms) ms[s] = ctx->prof_ms[s];
        if (launches) launches[s] = ctx->prof_launches[s];
        if (reset) { ctx->prof_ms[s] = 0.0; ctx->prof_launches[s] = 0; }
    }
    return RMCV_OK;
}

int rmcv_timer_start(rmcv_ctx* ctx) {
    if (!ctx) return RMCV_ERR_INVALID_ARG;
    CtxExtra* ex = extra(ctx);
    RMCV_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st[4 + kSlots] = {ex->pix, ex->out, ex->h2d, ex->d2h};
    for (int i = 0; i < kSlots; ++i) st[4 + i] = ex->labs[i];
    for (int i = 0; i < 4 + kSlots; ++i) {
        if (!ex->t_start[i]) { RMCV_CUDA(ctx, cudaEventCreate(&ex->t_start[i])); RMCV_CUDA(ctx, cudaEventCreate(&ex->t_stop[i])); }
        RMCV_CUDA(ctx, cudaEventRecord(ex->t_start[i], st[i]));
    }
    return RMCV_OK;
}

int rmcv_timer_stop(rmcv_ctx* ctx, double* ms) {
    if (!ctx || !ms) return RMCV_ERR_INVALID_ARG;
    CtxExtra* ex = extra(ctx);
    if (!ex->t_start[0]) return set_err(ctx, RMCV_ERR_STATE, "rmcv_timer_stop without rmcv_timer_start");
    cudaStream_t st[4 + kSlots] = {ex->pix, ex->out, ex->h2d, ex->d2h};
    for (int i = 0; i < kSlots; ++i) st[4 + i] = ex->labs[i];
    for (int i = 0; i < 4 + kSlots; ++i) RMCV_CUDA(ctx, cudaEventRecord(ex->t_stop[i], st[i]));
    int rc = sync_all(ctx);
    if (rc != RMCV_OK) return rc;
    float best = 0.f;
    for (int i = 0; i < 4 + kSlots; ++i)
        for (int j = 0; j < 4 + kSlots; ++j) {
            float t = 0.f;
            RMCV_CUDA(ctx, cudaEventElapsedTime(&t, ex->t_start[i], ex->t_stop[j]));
            if (t > best) best = t;
        }
    *ms = (double)best;
    return RMCV_OK;
}

int64_t rmcv_kernel_launches(const rmcv_ctx* ctx) { return ctx ? ctx->kernel_launches : 0; }

}  // extern "C"
