// f2 (next row of SURVEY 8f, first half) — the icon crop that feeds the SVM: rm::affine_correction (src/imgproc.cpp:9-35)
// followed by rm::utils::flatten_image (src/core.cpp:202-216), per armour, bit-exact against OpenCV:
//   clamp the icon vertices into the frame (in place, like the reference)          imgproc.cpp:11-15
//   box = cv::boundingRect of the vertices rounded to int (cvRound)                 :17
//   warp = cv::getAffineTransform (6x6 LU, partial pivoting, double)                :18-28
//   cv::warpAffine(source(box), warp), INTER_LINEAR, BORDER_CONSTANT 0              :31   (10-bit fixed-point coordinates,
//        5-bit sub-pixel positions, 15-bit bilinear weights)
//   cv::resize to out_w x out_h, INTER_LINEAR                                       :32   (11-bit coefficients)
//   reshape(1, 1) + convertTo(CV_32FC1)                                             core.cpp:212-213
// resize only ever reads two columns and two rows of the warped image per output pixel, so the warped image is never
// materialised: one thread per output pixel evaluates its (up to) four warped pixels straight from the frame.  All
// double / float steps use the _rn intrinsics (no FMA contraction), like the reference's x86 build.
#include <float.h>

#include "common.cuh"

namespace rmcv {

namespace {

struct IconParams {
    const uint8_t* src; size_t pitch; int W, H;
    rmcv_armour* armours; int n;
    int ow, oh;
    uint8_t* icons;      // [n][oh][ow][3]
    float* rows;         // [n][oh*ow*3] or null
};

struct IconGeom {        // per armour, built by thread 0
    double M[6];         // inverse map of warpAffine
    int bx, by, bw, bh;  // source(box)
};

__device__ __forceinline__ int cv_round_d(double v) { return __double2int_rn(v); }

// cv::solve(A, b, x, DECOMP_LU) for 6x6 (hal::LU64f: partial pivoting, eps = DBL_EPSILON * 100); false when singular
__device__ bool lu_solve6(double a[6][6], double b[6]) {
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j) if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (fabs(a[k][i]) < DBL_EPSILON * 100) return false;
        if (k != i) {
            for (int j = i; j < 6; ++j) { const double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = __ddiv_rn(-1.0, a[i][i]);
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = __dmul_rn(a[j][i], d);
            for (int c = i + 1; c < 6; ++c) a[j][c] = __dadd_rn(a[j][c], __dmul_rn(alpha, a[i][c]));
            b[j] = __dadd_rn(b[j], __dmul_rn(alpha, b[i]));
        }
    }
    for (int i = 5; i >= 0; --i) {
        double s = b[i];
        for (int c = i + 1; c < 6; ++c) s = __dsub_rn(s, __dmul_rn(a[i][c], b[c]));
        b[i] = __ddiv_rn(s, a[i][i]);
    }
    return true;
}

// one pixel (3 channels) of cv::warpAffine(source(box), M) at (x, y) of the warped image
__device__ void warped_pixel(const IconParams& p, const IconGeom& g, int x, int y, int out[3]) {
    const int X0 = cv_round_d(__dmul_rn(__dadd_rn(__dmul_rn(g.M[1], (double)y), g.M[2]), 1024.0)) + 16;
    const int Y0 = cv_round_d(__dmul_rn(__dadd_rn(__dmul_rn(g.M[4], (double)y), g.M[5]), 1024.0)) + 16;
    const int X = (X0 + cv_round_d(__dmul_rn(__dmul_rn(g.M[0], (double)x), 1024.0))) >> 5;
    const int Y = (Y0 + cv_round_d(__dmul_rn(__dmul_rn(g.M[3], (double)x), 1024.0))) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int fx = X & 31, fy = Y & 31;
    out[0] = out[1] = out[2] = 0;
    if (sx >= g.bw || sx + 1 < 0 || sy >= g.bh || sy + 1 < 0) return;          // BORDER_CONSTANT, value 0
    const int w[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int xx = sx + (t & 1), yy = sy + (t >> 1);
        if (w[t] == 0 || xx < 0 || xx >= g.bw || yy < 0 || yy >= g.bh) continue;
        const uint8_t* s = p.src + (size_t)(g.by + yy) * p.pitch + (size_t)(g.bx + xx) * 3;
        for (int c = 0; c < 3; ++c) acc[c] += (int)s[c] * w[t];
    }
    if (fx == 0 && fy == 0) {                                                  // weight 32768 does not fit a short: the
        for (int c = 0; c < 3; ++c) out[c] = acc[c] >> 15;                     // table holds 32767 (+1 elsewhere), result = tap
    } else {
        for (int c = 0; c < 3; ++c) out[c] = (acc[c] + (1 << 14)) >> 15;
    }
}

// cv::resize INTER_LINEAR source index and 11-bit coefficients of destination index d (size dn) over a source of sn
__device__ void resize_coef(int d, int dn, int sn, int* s0, int* c0, int* c1) {
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dn, (double)sn));
    float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    *s0 = s;
    *c0 = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    *c1 = (int)rintf(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(128) icon_kernel(const IconParams p) {
    __shared__ IconGeom g;
    const int k = blockIdx.x;
    rmcv_armour& arm = p.armours[k];
    if (threadIdx.x == 0) {
        float vx[4], vy[4];
        int ix[4], iy[4];
        for (int i = 0; i < 4; ++i) {                                           // imgproc.cpp:11-15
            vx[i] = fmaxf(0.0f, fminf(arm.icon[i][0], (float)p.W - 1.f));
            vy[i] = fmaxf(0.0f, fminf(arm.icon[i][1], (float)p.H - 1.f));
            arm.icon[i][0] = vx[i]; arm.icon[i][1] = vy[i];
            ix[i] = __float2int_rn(vx[i]); iy[i] = __float2int_rn(vy[i]);      // cv::Point(Point2f): cvRound
        }
        const int x0 = min(min(ix[0], ix[1]), min(ix[2], ix[3])), x1 = max(max(ix[0], ix[1]), max(ix[2], ix[3]));
        const int y0 = min(min(iy[0], iy[1]), min(iy[2], iy[3])), y1 = max(max(iy[0], iy[1]), max(iy[2], iy[3]));
        g.bx = x0; g.by = y0; g.bw = x1 - x0 + 1; g.bh = y1 - y0 + 1;
        const int order[3] = {1, 2, 0};                                         // srcPts = vertices[1], [2], [0] minus the box origin
        const float dstx[3] = {0.f, (float)g.bw, 0.f}, dsty[3] = {0.f, 0.f, (float)g.bh};
        double a[6][6], b[6];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) a[i][j] = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double sx = (double)__fsub_rn(vx[order[i]], (float)g.bx), sy = (double)__fsub_rn(vy[order[i]], (float)g.by);
            a[2 * i][0] = sx; a[2 * i][1] = sy; a[2 * i][2] = 1.0;
            a[2 * i + 1][3] = sx; a[2 * i + 1][4] = sy; a[2 * i + 1][5] = 1.0;
            b[2 * i] = (double)dstx[i]; b[2 * i + 1] = (double)dsty[i];
        }
        double M[6];
        if (lu_solve6(a, b)) { for (int i = 0; i < 6; ++i) M[i] = b[i]; }
        else { for (int i = 0; i < 6; ++i) M[i] = 0.0; }                        // cv::solve: a singular system leaves zeros
        // cv::warpAffine inverts the map
        double D = __dsub_rn(__dmul_rn(M[0], M[4]), __dmul_rn(M[1], M[3]));
        D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
        const double A11 = __dmul_rn(M[4], D), A22 = __dmul_rn(M[0], D);
        M[0] = A11; M[1] = __dmul_rn(M[1], -D); M[3] = __dmul_rn(M[3], -D); M[4] = A22;
        const double b1 = __dsub_rn(__dmul_rn(-M[0], M[2]), __dmul_rn(M[1], M[5]));
        const double b2 = __dsub_rn(__dmul_rn(-M[3], M[2]), __dmul_rn(M[4], M[5]));
        M[2] = b1; M[5] = b2;
        for (int i = 0; i < 6; ++i) g.M[i] = M[i];
    }
    __syncthreads();
    const int npx = p.ow * p.oh;
    for (int o = threadIdx.x; o < npx; o += blockDim.x) {
        const int dy = o / p.ow, dx = o - dy * p.ow;
        int out[3];
        if (g.bw == p.ow && g.bh == p.oh) {                                     // cv::resize to the same size copies
            warped_pixel(p, g, dx, dy, out);
        } else {
            int sx, a0, a1, sy, b0, b1;
            resize_coef(dx, p.ow, g.bw, &sx, &a0, &a1);
            if (sx < 0) { sx = 0; a0 = 2048; a1 = 0; }
            if (sx + 1 >= g.bw) { sx = g.bw - 1; a0 = 2048; a1 = 0; }
            resize_coef(dy, p.oh, g.bh, &sy, &b0, &b1);
            const int yr[2] = {min(max(sy, 0), g.bh - 1), min(max(sy + 1, 0), g.bh - 1)};
            int h[2][3];
            for (int r = 0; r < 2; ++r) {
                int pa[3], pb[3] = {0, 0, 0};
                warped_pixel(p, g, sx, yr[r], pa);
                if (a1 != 0) warped_pixel(p, g, sx + 1, yr[r], pb);
                for (int c = 0; c < 3; ++c) h[r][c] = pa[c] * a0 + pb[c] * a1;
            }
            for (int c = 0; c < 3; ++c) out[c] = ((((b0 * (h[0][c] >> 4)) >> 16) + ((b1 * (h[1][c] >> 4)) >> 16) + 2) >> 2) & 255;
        }
        uint8_t* d = p.icons + ((size_t)k * npx + o) * 3;
        for (int c = 0; c < 3; ++c) d[c] = (uint8_t)out[c];
        if (p.rows) for (int c = 0; c < 3; ++c) p.rows[((size_t)k * npx + o) * 3 + c] = (float)out[c];
    }
}

// cv::ml::SVM::predict for a C_SVC model with the LINEAR kernel (executable/svm/optimizer.cpp:18-21 trains exactly that;
// OpenCV modules/ml/src/svm.cpp: calc_non_rbf_base + the one-vs-one vote of PredictBody).  One thread per (sample, support
// vector) keeps OpenCV's summation order: float products, groups of four added in float, accumulated in double.
struct SvmParams {
    const float* rows; int n;
    const float* sv; int sv_total, var_count;
    const double* rho; const int32_t* df_ofs; const double* df_alpha; const int32_t* df_index;
    const int32_t* class_labels; int class_count;
    float* kbuf;          // [n][sv_total] kernel values
    int32_t* labels;      // [n]
};

__global__ void svm_kernel_values(const SvmParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n * p.sv_total) return;
    const int smp = i / p.sv_total, j = i - smp * p.sv_total;
    const float* sample = p.sv + (size_t)j * p.var_count;
    const float* another = p.rows + (size_t)smp * p.var_count;
    double s = 0.0;
    int k = 0;
    for (; k <= p.var_count - 4; k += 4) {
        const float t = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(sample[k], another[k]), __fmul_rn(sample[k + 1], another[k + 1])),
                                            __fmul_rn(sample[k + 2], another[k + 2])), __fmul_rn(sample[k + 3], another[k + 3]));
        s = __dadd_rn(s, (double)t);
    }
    for (; k < p.var_count; ++k) s = __dadd_rn(s, (double)__fmul_rn(sample[k], another[k]));
    p.kbuf[i] = (float)s;
}

__global__ void svm_vote(const SvmParams p) {
    const int smp = blockIdx.x * blockDim.x + threadIdx.x;
    if (smp >= p.n) return;
    const float* buffer = p.kbuf + (size_t)smp * p.sv_total;
    int vote[32];
    for (int i = 0; i < p.class_count; ++i) vote[i] = 0;
    int dfi = 0;
    for (int i = 0; i < p.class_count; ++i)
        for (int j = i + 1; j < p.class_count; ++j, ++dfi) {
            double sum = -p.rho[dfi];
            for (int k = p.df_ofs[dfi]; k < p.df_ofs[dfi + 1]; ++k) sum = __dadd_rn(sum, __dmul_rn(p.df_alpha[k], (double)buffer[p.df_index[k]]));
            vote[sum > 0 ? i : j]++;
        }
    int best = 0;
    for (int i = 1; i < p.class_count; ++i) if (vote[i] > vote[best]) best = i;
    p.labels[smp] = p.class_labels[best];
}

}  // namespace

cudaError_t launch_svm_predict(const float* d_rows, int n, const float* d_sv, int sv_total, int var_count, const double* d_rho,
                               const int32_t* d_df_ofs, const double* d_df_alpha, const int32_t* d_df_index, const int32_t* d_class_labels,
                               int class_count, float* d_kbuf, int32_t* d_labels, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    SvmParams p;
    p.rows = d_rows; p.n = n; p.sv = d_sv; p.sv_total = sv_total; p.var_count = var_count; p.rho = d_rho; p.df_ofs = d_df_ofs;
    p.df_alpha = d_df_alpha; p.df_index = d_df_index; p.class_labels = d_class_labels; p.class_count = class_count;
    p.kbuf = d_kbuf; p.labels = d_labels;
    const int total = n * sv_total;
    svm_kernel_values<<<(total + 63) / 64, 64, 0, st>>>(p);
    svm_vote<<<(n + 63) / 64, 64, 0, st>>>(p);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_icons(const uint8_t* d_bgr, size_t pitch, int W, int H, rmcv_armour* d_armours, int n, int ow, int oh,
                         uint8_t* d_icons, float* d_rows, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    IconParams p;
    p.src = d_bgr; p.pitch = pitch; p.W = W; p.H = H; p.armours = d_armours; p.n = n; p.ow = ow; p.oh = oh;
    p.icons = d_icons; p.rows = d_rows;
    icon_kernel<<<n, 128, 0, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
