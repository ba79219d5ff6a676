// One batch partitioned by frame across the GPUs of one box (BASELINE north_star).
//
// Frames are independent units of the reference's hot loop (executable/main.cpp:163-209 handles one frame per iteration
// and keeps no cross-frame state), so a batch of B frames is cut into G contiguous slices of ceil(B/G) frames; slice g is
// staged into GPU g's HBM by a host thread that owns one rmcv_ctx on that device; no collective, no peer traffic.  The
// only cross-GPU step is the host-side concatenation of the per-slice results (a few KB per frame) in frame order.
// Workers are persistent threads: a call posts one job descriptor, every worker runs the whole path on its slice through
// the single-GPU host entry point (api.cu: chunked upload / kernels / download pipeline), the caller's thread merges.
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" void rmcv_multi_slice(int batch, int n_devices, int g, int* first, int* count) {
    // contiguous slices whose sizes differ by at most one frame (the first batch % G slices hold ceil(B/G) frames); the
    // same rule as rmcv_b200/shard.py::frame_slice, which the one-process-per-GPU launch uses
    if (n_devices <= 0 || g < 0 || g >= n_devices || batch < 0) { if (first) *first = 0; if (count) *count = 0; return; }
    const int base = batch / n_devices, rem = batch % n_devices;
    if (first) *first = g * base + (g < rem ? g : rem);
    if (count) *count = base + (g < rem ? 1 : 0);
}

struct rmcv_multi {
    struct Job {
        const uint8_t* h_src = nullptr; size_t pitch = 0, frame_stride = 0;
        int width = 0, height = 0, batch = 0, bayer_layout = 0;
        rmcv_params params;
        uint8_t* h_mask = nullptr; size_t mask_pitch = 0, mask_frame_stride = 0;
    };
    struct Worker {
        std::thread th;
        int device = 0;
        rmcv_ctx* ctx = nullptr;
        int create_rc = RMCV_OK;
        int rc = RMCV_OK;            // status of the last job on this slice
        rmcv_results res;            // view into the worker's ctx (valid until its next call)
        int first = 0, count = 0;
        double ms = 0.0;             // wall time of the slice's call (diagnostic)
    };
    rmcv_config cfg;
    std::vector<Worker> w;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    long long job_id = 0;            // bumped for every job; workers run job k once
    int n_done = 0;
    bool stop = false;
    Job job;
    // merged results of the last call (frame order)
    std::vector<rmcv_frame_info> frames;
    std::vector<rmcv_contour_info> contours;
    std::vector<rmcv_lightblob> blobs;
    std::vector<rmcv_armour> armours;
    char err[512];
};

namespace {

void worker_main(rmcv_multi* m, int g) {
    rmcv_multi::Worker& me = m->w[(size_t)g];
    rmcv_config cfg = m->cfg;
    cfg.device = me.device;
    const int G = (int)m->w.size();
    cfg.max_batch = (m->cfg.max_batch + G - 1) / G;
    cfg.stream = nullptr;
    me.create_rc = rmcv_ctx_create(&cfg, &me.ctx);
    long long seen = 0;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        ++m->n_done;
    }
    m->cv_done.notify_all();
    while (true) {
        rmcv_multi::Job job;
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_job.wait(lk, [&] { return m->stop || m->job_id != seen; });
            if (m->stop) break;
            seen = m->job_id;
            job = m->job;
        }
        rmcv_multi_slice(job.batch, G, g, &me.first, &me.count);
        me.rc = RMCV_OK;
        memset(&me.res, 0, sizeof(me.res));
        if (me.count > 0) {
            if (!me.ctx) me.rc = me.create_rc != RMCV_OK ? me.create_rc : RMCV_ERR_STATE;
            else {
                const uint8_t* src = job.h_src + (size_t)me.first * job.frame_stride;
                uint8_t* mask = job.h_mask ? job.h_mask + (size_t)me.first * job.mask_frame_stride : nullptr;
                if (job.bayer_layout)
                    me.rc = rmcv_bayer_detect_batch_host(me.ctx, src, job.pitch, job.frame_stride, job.width, job.height, me.count,
                                                         job.bayer_layout, &job.params, mask, job.mask_pitch, job.mask_frame_stride, &me.res);
                else
                    me.rc = rmcv_detect_batch_host(me.ctx, src, job.pitch, job.frame_stride, job.width, job.height, me.count, &job.params,
                                                   mask, job.mask_pitch, job.mask_frame_stride, &me.res);
            }
        }
        {
            std::lock_guard<std::mutex> lk(m->mu);
            ++m->n_done;
        }
        m->cv_done.notify_all();
    }
    if (me.ctx) rmcv_ctx_destroy(me.ctx);
    me.ctx = nullptr;
}

int run_job(rmcv_multi* m, const rmcv_multi::Job& job, rmcv_results* out) {
    const int G = (int)m->w.size();
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->job = job;
        m->n_done = 0;
        ++m->job_id;
    }
    m->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(m->mu);
        m->cv_done.wait(lk, [&] { return m->n_done == G; });
    }
    // merge in frame order: slices are contiguous and ordered by g, dense arrays are concatenated and the per-frame
    // offsets re-based
    int status = RMCV_OK;
    size_t nc = 0, nb = 0, na = 0;
    for (const auto& w : m->w) {
        if (w.count == 0) continue;
        if (w.rc != RMCV_OK && w.rc != RMCV_ERR_CAPACITY) {
            snprintf(m->err, sizeof(m->err), "device %d: %s (%s)", w.device, rmcv_status_string(w.rc), w.ctx ? rmcv_last_error(w.ctx) : "no ctx");
            return w.rc;
        }
        if (w.rc == RMCV_ERR_CAPACITY) status = RMCV_ERR_CAPACITY;
        nc += (size_t)w.res.total_contours; nb += (size_t)w.res.total_blobs; na += (size_t)w.res.total_armours;
    }
    m->frames.resize((size_t)job.batch);
    m->contours.resize(nc); m->blobs.resize(nb); m->armours.resize(na);
    size_t oc = 0, ob = 0, oa = 0;
    for (const auto& w : m->w) {
        if (w.count == 0) continue;
        for (int f = 0; f < w.count; ++f) {
            const rmcv_frame_info& src = w.res.frames[f];
            rmcv_frame_info fi = src;
            fi.contour_offset = (int32_t)oc; fi.blob_offset = (int32_t)ob; fi.armour_offset = (int32_t)oa;
            if (src.n_contours) memcpy(&m->contours[oc], w.res.contours + src.contour_offset, (size_t)src.n_contours * sizeof(rmcv_contour_info));
            if (src.n_positive) memcpy(&m->blobs[ob], w.res.blobs + src.blob_offset, (size_t)src.n_positive * sizeof(rmcv_lightblob));
            if (src.n_armours) memcpy(&m->armours[oa], w.res.armours + src.armour_offset, (size_t)src.n_armours * sizeof(rmcv_armour));
            oc += (size_t)src.n_contours; ob += (size_t)src.n_positive; oa += (size_t)src.n_armours;
            m->frames[(size_t)(w.first + f)] = fi;
        }
    }
    if (out) {
        out->batch = job.batch;
        out->total_contours = (int32_t)oc; out->total_blobs = (int32_t)ob; out->total_armours = (int32_t)oa;
        out->frames = m->frames.data(); out->contours = m->contours.data(); out->blobs = m->blobs.data(); out->armours = m->armours.data();
        out->poses = nullptr;
    }
    if (status != RMCV_OK) snprintf(m->err, sizeof(m->err), "a per-frame capacity overflowed; see rmcv_frame_info.flags");
    return status;
}

int check_call(rmcv_multi* m, const uint8_t* h_src, int width, int height, int batch, const rmcv_params* params) {
    if (!m || !h_src || !params) return RMCV_ERR_INVALID_ARG;
    if (width <= 0 || height <= 0 || batch <= 0 || width > m->cfg.max_width || height > m->cfg.max_height || batch > m->cfg.max_batch) {
        snprintf(m->err, sizeof(m->err), "frame size or batch exceeds the maxima given to rmcv_multi_create");
        return RMCV_ERR_INVALID_ARG;
    }
    return RMCV_OK;
}

}  // namespace

extern "C" {

int rmcv_multi_create(const rmcv_config* cfg, const int* devices, int n_devices, rmcv_multi** out) {
    if (!cfg || !out || n_devices < 0) return RMCV_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return RMCV_ERR_NO_DEVICE; }
    if (n_devices == 0) n_devices = ndev;      // all visible devices
    if (n_devices > ndev && !devices) return RMCV_ERR_INVALID_ARG;
    rmcv_multi* m = new (std::nothrow) rmcv_multi();
    if (!m) return RMCV_ERR_INVALID_ARG;
    m->cfg = *cfg;
    m->err[0] = 0;
    m->w.resize((size_t)n_devices);
    for (int g = 0; g < n_devices; ++g) {
        m->w[(size_t)g].device = devices ? devices[g] : g;
        if (m->w[(size_t)g].device < 0 || m->w[(size_t)g].device >= ndev) { delete m; return RMCV_ERR_INVALID_ARG; }
    }
    for (int g = 0; g < n_devices; ++g) m->w[(size_t)g].th = std::thread(worker_main, m, g);
    {
        std::unique_lock<std::mutex> lk(m->mu);
        m->cv_done.wait(lk, [&] { return m->n_done == n_devices; });
    }
    for (const auto& w : m->w)
        if (w.create_rc != RMCV_OK) {
            const int rc = w.create_rc;
            rmcv_multi_destroy(m);
            return rc;
        }
    *out = m;
    return RMCV_OK;
}

int rmcv_multi_destroy(rmcv_multi* m) {
    if (!m) return RMCV_ERR_INVALID_ARG;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->stop = true;
    }
    m->cv_job.notify_all();
    for (auto& w : m->w)
        if (w.th.joinable()) w.th.join();
    delete m;
    return RMCV_OK;
}

int rmcv_multi_device_count(const rmcv_multi* m) { return m ? (int)m->w.size() : 0; }
const char* rmcv_multi_last_error(const rmcv_multi* m) { return m ? m->err : "null rmcv_multi"; }

int rmcv_multi_detect_batch_host(rmcv_multi* m, const uint8_t* h_bgr, size_t pitch, size_t frame_stride, int width, int height,
                                 int batch, const rmcv_params* params, uint8_t* h_mask, size_t mask_pitch, size_t mask_frame_stride,
                                 rmcv_results* out) {
    if (int rc = check_call(m, h_bgr, width, height, batch, params)) return rc;
    rmcv_multi::Job job;
    job.h_src = h_bgr; job.pitch = pitch; job.frame_stride = frame_stride; job.width = width; job.height = height; job.batch = batch;
    job.bayer_layout = 0; job.params = *params; job.h_mask = h_mask; job.mask_pitch = mask_pitch; job.mask_frame_stride = mask_frame_stride;
    return run_job(m, job, out);
}

int rmcv_multi_bayer_detect_batch_host(rmcv_multi* m, const uint8_t* h_raw, size_t pitch, size_t frame_stride, int width, int height,
                                       int batch, int bayer_layout, const rmcv_params* params, uint8_t* h_mask, size_t mask_pitch,
                                       size_t mask_frame_stride, rmcv_results* out) {
    if (int rc = check_call(m, h_raw, width, height, batch, params)) return rc;
    if (bayer_layout < RMCV_BAYER_RG || bayer_layout > RMCV_BAYER_BG) {
        snprintf(m->err, sizeof(m->err), "bad bayer layout");
        return RMCV_ERR_INVALID_ARG;
    }
    rmcv_multi::Job job;
    job.h_src = h_raw; job.pitch = pitch; job.frame_stride = frame_stride; job.width = width; job.height = height; job.batch = batch;
    job.bayer_layout = bayer_layout; job.params = *params; job.h_mask = h_mask; job.mask_pitch = mask_pitch;
    job.mask_frame_stride = mask_frame_stride;
    return run_job(m, job, out);
}

}  // extern "C"
