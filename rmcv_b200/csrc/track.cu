// f3 (next row of SURVEY 8f) — armour tracking: the reference's tracking loop (executable/main.cpp:57-88) with
// rm::armour::max_IoU / update / reset (src/core.cpp:51-161) and cv::KalmanFilter(6, 6, 0, CV_64F)::predict / correct.
// The step is tiny and sequential (a matched armour leaves the list the next track sees; an erased track skips the one
// behind it), so it runs as ONE thread: what it buys is that a detect -> pose -> track chain never has to leave the
// device.  fp64 throughout, like the reference; the 6x6 solve of correct() is Gauss-Jordan with partial pivoting
// (OpenCV: DECOMP_SVD; the systems are diagonally dominant, results agree to ~1e-13 relative).
#include <math.h>

#include "common.cuh"

namespace rmcv {

namespace {

struct TrackParams {
    rmcv_track* tracks; int32_t* n_tracks; int cap;
    rmcv_track* backup;                 // copy of the list taken before the update (restored on overflow)
    const rmcv_armour* armours; const double* positions; const int32_t* identities; int n;
    int32_t* remaining;                 // scratch: indices of the armours not yet matched
    long long timestamp; double freq, q, r, err;
    int32_t* status;                    // 0 ok, 1 track list overflow, 2 identity history overflow
};

// intersection.area() / (a.area() + b.area() - intersection.area()) of two cv::Rect2f (src/core.cpp:151-154), in float
// like the reference; the intersection follows cv::Rect_::operator&= of OpenCV 4.5+ (modules/core/.../types.hpp)
__device__ float iou_of(const float* a, const float* b) {
    float iw = 0.f, ih = 0.f;
    const bool a_empty = a[2] <= 0.f || a[3] <= 0.f, b_empty = b[2] <= 0.f || b[3] <= 0.f;
    if (!a_empty && !b_empty) {
        const float* xmin = a[0] < b[0] ? a : b; const float* xmax = a[0] < b[0] ? b : a;
        const float* ymin = a[1] < b[1] ? a : b; const float* ymax = a[1] < b[1] ? b : a;
        const bool apart = (xmin[0] < 0.f && __fadd_rn(xmin[0], xmin[2]) < xmax[0]) || (ymin[1] < 0.f && __fadd_rn(ymin[1], ymin[3]) < ymax[1]);
        if (!apart) {
            iw = fminf(__fsub_rn(xmin[2], __fsub_rn(xmax[0], xmin[0])), xmax[2]);
            ih = fminf(__fsub_rn(ymin[3], __fsub_rn(ymax[1], ymin[1])), ymax[3]);
            if (iw <= 0.f || ih <= 0.f) { iw = 0.f; ih = 0.f; }
        }
    }
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a[2], a[3]), __fmul_rn(b[2], b[3])), inter);
    return __fdiv_rn(inter, uni);
}

// rm::armour::reset + the tracking fields of a fresh armour (main.cpp:180-195)
__device__ void open_track(rmcv_track& t, const rmcv_armour& a, const double* pos, int identity, long long ts, double q, double r,
                           double err) {
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < 4; ++i) t.bbox[i] = a.bounding_box[i];
    for (int i = 0; i < 3; ++i) t.position[i] = pos[i];
    t.timestamp = ts; t.identity = identity;
    for (int i = 0; i < 6; ++i) t.cov_post[7 * i] = err;
    t.q = q; t.r = r;
}

// cv::KalmanFilter::predict with transitionMatrix = I + dt on (0,3), (1,4), (2,5); processNoiseCov = q I
__device__ void kf_predict(rmcv_track& t, double dt) {
    for (int i = 0; i < 3; ++i) { t.state_pre[i] = t.state_post[i] + dt * t.state_post[i + 3]; t.state_pre[i + 3] = t.state_post[i + 3]; }
    double FP[36];   // F * P
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) FP[6 * i + j] = t.cov_post[6 * i + j] + (i < 3 ? dt * t.cov_post[6 * (i + 3) + j] : 0.0);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j)   // (F P) F^T + Q
            t.cov_pre[6 * i + j] = FP[6 * i + j] + (j < 3 ? dt * FP[6 * i + j + 3] : 0.0) + (i == j ? t.q : 0.0);
    for (int i = 0; i < 6; ++i) t.state_post[i] = t.state_pre[i];
    for (int i = 0; i < 36; ++i) t.cov_post[i] = t.cov_pre[i];
}

// cv::KalmanFilter::correct with measurementMatrix = I, measurementNoiseCov = r I
__device__ void kf_correct(rmcv_track& t) {
    double S[36], X[36];   // S = P' + R;  X = S^-1 P'  (gain = X^T)
    for (int i = 0; i < 36; ++i) { S[i] = t.cov_pre[i]; X[i] = t.cov_pre[i]; }
    for (int i = 0; i < 6; ++i) S[7 * i] += t.r;
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int i = c + 1; i < 6; ++i) if (fabs(S[6 * i + c]) > fabs(S[6 * piv + c])) piv = i;
        if (piv != c)
            for (int j = 0; j < 6; ++j) {
                double tmp = S[6 * c + j]; S[6 * c + j] = S[6 * piv + j]; S[6 * piv + j] = tmp;
                tmp = X[6 * c + j]; X[6 * c + j] = X[6 * piv + j]; X[6 * piv + j] = tmp;
            }
        const double d = 1.0 / S[7 * c];
        for (int j = 0; j < 6; ++j) { S[6 * c + j] *= d; X[6 * c + j] *= d; }
        for (int i = 0; i < 6; ++i) {
            if (i == c) continue;
            const double f = S[6 * i + c];
            if (f == 0.0) continue;
            for (int j = 0; j < 6; ++j) { S[6 * i + j] -= f * S[6 * c + j]; X[6 * i + j] -= f * X[6 * c + j]; }
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = t.meas[i] - t.state_pre[i];
    for (int i = 0; i < 6; ++i) {
        double acc = 0.0;
        for (int j = 0; j < 6; ++j) acc += X[6 * j + i] * y[j];          // gain(i, j) = X(j, i)
        t.state_post[i] = t.state_pre[i] + acc;
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += X[6 * k + i] * t.cov_pre[6 * k + j];   // gain * (H P')
            t.cov_post[6 * i + j] = t.cov_pre[6 * i + j] - acc;
        }
}

// rm::armour::update(const armour&), src/core.cpp:71-106
__device__ bool update_observed(rmcv_track& t, const double* pos, int identity, long long ts, double freq) {
    int k = 0;
    while (k < t.n_hist && t.hist_id[k] < identity) ++k;
    if (k < t.n_hist && t.hist_id[k] == identity) ++t.hist_count[k];
    else {
        if (t.n_hist >= RMCV_TRACK_HIST) return false;
        for (int m = t.n_hist; m > k; --m) { t.hist_id[m] = t.hist_id[m - 1]; t.hist_count[m] = t.hist_count[m - 1]; }
        t.hist_id[k] = identity; t.hist_count[k] = 1; ++t.n_hist;
    }
    if (t.initialized) {
        const double dt = (double)(ts - t.timestamp) / freq;
        kf_predict(t, dt);
        for (int i = 0; i < 3; ++i) t.meas[i + 3] = (pos[i] - t.meas[i]) / dt;
        for (int i = 0; i < 3; ++i) t.meas[i] = pos[i];
        kf_correct(t);
    } else {
        for (int i = 0; i < 3; ++i) t.meas[i] = pos[i];
        kf_correct(t);
        t.initialized = 1;
    }
    t.timestamp = ts;
    return true;
}

__global__ void track_kernel(const TrackParams p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    *p.status = 0;
    int nt = *p.n_tracks;
    if (p.n == 0) return;                                            // main.cpp:63
    for (int i = 0; i < nt; ++i) p.backup[i] = p.tracks[i];
    const int nt0 = nt;
    auto fail = [&](int code) {
        for (int i = 0; i < nt0; ++i) p.tracks[i] = p.backup[i];
        *p.n_tracks = nt0; *p.status = code;
    };
    int nrem = p.n;
    for (int i = 0; i < nrem; ++i) p.remaining[i] = i;
    if (nt > 0) {                                                    // main.cpp:71-82 (an empty list: the frame becomes the list)
        for (int i = 0; i < nt; ++i) {
            rmcv_track& t = p.tracks[i];
            int index = -1; float best = 0.f;                        // rm::armour::max_IoU, src/core.cpp:145-161
            for (int k = 0; k < nrem; ++k) {
                const float v = iou_of(t.bbox, p.armours[p.remaining[k]].bounding_box);
                if (v > best) { best = v; index = k; }
            }
            if (best > 0.5f) {
                const int a = p.remaining[index];
                if (!update_observed(t, p.positions + 3 * a, p.identities ? p.identities[a] : -1, p.timestamp, p.freq)) { fail(2); return; }
                for (int k = index; k + 1 < nrem; ++k) p.remaining[k] = p.remaining[k + 1];
                --nrem;
            } else if (t.lost_count++ > 25) {
                for (int k = i; k + 1 < nt; ++k) p.tracks[k] = p.tracks[k + 1];   // erase(begin + i); the loop's ++i skips the next track
                --nt;
            } else if (t.initialized) {                              // update(timestamp) with its own timestamp: dt = 0
                kf_predict(t, 0.0);
            }
        }
    }
    if (nt + nrem > p.cap) { fail(1); return; }
    for (int k = 0; k < nrem; ++k) {
        const int a = p.remaining[k];
        open_track(p.tracks[nt++], p.armours[a], p.positions + 3 * a, p.identities ? p.identities[a] : -1, p.timestamp, p.q, p.r, p.err);
    }
    *p.n_tracks = nt;
}

}  // namespace

cudaError_t launch_track_update(rmcv_track* d_tracks, int32_t* d_n_tracks, int cap, rmcv_track* d_backup, const rmcv_armour* d_armours,
                                const double* d_positions, const int32_t* d_identities, int n, int32_t* d_remaining, long long timestamp,
                                double freq, double q, double r, double err, int32_t* d_status, cudaStream_t st, int64_t* launches) {
    TrackParams p;
    p.tracks = d_tracks; p.n_tracks = d_n_tracks; p.cap = cap; p.backup = d_backup;
    p.armours = d_armours; p.positions = d_positions; p.identities = d_identities; p.n = n; p.remaining = d_remaining;
    p.timestamp = timestamp; p.freq = freq; p.q = q; p.r = r; p.err = err; p.status = d_status;
    track_kernel<<<1, 32, 0, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
