// The reference's call site (executable/main.cpp:172-176) compiled against include/rmcv_gpu/rm_shim.hpp in its
// RMCV_SHIM_WITH_REFERENCE mode: rm::lightblob / rm::armour / rm::camp / rm::range are the REFERENCE'S OWN declarations
// (include/core.h) and the objects are rebuilt through the reference's own constructors (src/core.cpp, included below
// unmodified from /root/reference) from the GPU's ellipses and pair indices — the wiring INTEGRATION.md describes.
// The image has no OpenCV C++, so <opencv2/opencv.hpp> is oracle/cvstub's types-and-trampolines header; the two cv::
// functions the constructors call (RotatedRect::points, boundingRect of float points) are served natively below.
// TEST INFRASTRUCTURE: built by oracle/Makefile into oracle/_ref/ (needs /root/reference), run by tests/test_gpu_shim.py.
#define RMCV_SHIM_WITH_REFERENCE 1
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <thread>
#include <sys/stat.h>
#include <opencv2/opencv.hpp>
#include <opencv2/ml.hpp>

#include "src/core.cpp"                       // the reference's own rm::lightblob / rm::armour / rm::utils definitions
#include "rmcv_gpu/rm_shim.hpp"               // rm::extract_color / filter_lightblobs / filter_armours over the C ABI
#include "../../rmcv_b200/csrc/blob_math.cuh" // host build of rotated_rect_points (= cv::RotatedRect::points)

extern "C" {
rmcv_ref_release_t rmcv_ref_release = nullptr;
double rmcv_ref_tick_frequency = 1e9;
static float g_out[16];
static int32_t g_rect[4];
// the cv:: calls of the two constructors, natively (no Python in this binary)
static int native_cvcall(const char* op, const rmcv_ref_arr* in, int n_in, const double*, int, rmcv_ref_arr* out, int n_out) {
    if (n_in < 1 || n_out < 1) return 1;
    if (!std::strcmp(op, "boxPoints")) {                      // cv::RotatedRect::points
        const float* b = static_cast<const float*>(in[0].data);
        rmcv_rotated_rect r{b[0], b[1], b[2], b[3], b[4]};
        float pt[4][2];
        rmcv::rotated_rect_points(r, pt);
        for (int i = 0; i < 4; ++i) { g_out[2 * i] = pt[i][0]; g_out[2 * i + 1] = pt[i][1]; }
        out[0].data = g_out; out[0].rows = 4; out[0].cols = 2; out[0].type = CV_32F; out[0].step = 8; out[0].owner = 0;
        return 0;
    }
    if (!std::strcmp(op, "boundingRect") && CV_MAT_DEPTH(in[0].type) == CV_32F) {   // float points: floor rule
        const int n = in[0].rows;
        float minx = 0, miny = 0, maxx = 0, maxy = 0;
        for (int i = 0; i < n; ++i) {
            const float* p = reinterpret_cast<const float*>(static_cast<const uchar*>(in[0].data) + (size_t)i * in[0].step);
            if (i == 0) { minx = maxx = p[0]; miny = maxy = p[1]; }
            minx = std::min(minx, p[0]); maxx = std::max(maxx, p[0]); miny = std::min(miny, p[1]); maxy = std::max(maxy, p[1]);
        }
        const int x0 = cv::cvFloor(minx), y0 = cv::cvFloor(miny), x1 = cv::cvFloor(maxx), y1 = cv::cvFloor(maxy);
        g_rect[0] = x0; g_rect[1] = y0; g_rect[2] = x1 - x0 + 1; g_rect[3] = y1 - y0 + 1;
        out[0].data = g_rect; out[0].rows = 1; out[0].cols = 4; out[0].type = CV_32S; out[0].step = 16; out[0].owner = 0;
        return 0;
    }
    return 1;
}
rmcv_ref_cvcall_t rmcv_ref_cvcall = native_cvcall;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: call_site_ref frame.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    int wh[2];
    if (fread(wh, sizeof(int), 2, f) != 2) return 2;
    cv::Mat image(wh[1], wh[0], CV_8UC3);
    if (fread(image.data, 1, (size_t)wh[0] * wh[1] * 3, f) != (size_t)wh[0] * wh[1] * 3) return 2;
    fclose(f);
    try {
        // executable/main.cpp:172-176, verbatim
        auto [contours, binary] = rm::extract_color(image, rm::CAMP_BLUE, 80);
        auto [positive, negtive] = rm::filter_lightblobs(contours, 70, {1.5, 80}, {10, 99999}, rm::CAMP_BLUE);
        auto armours = rm::filter_armours(positive, 12, 22, 0.4, rm::CAMP_BLUE);
        unsigned long long fg = 0;
        for (int y = 0; y < binary.rows; ++y)
            for (int x = 0; x < binary.cols; ++x) fg += binary.data[(size_t)y * binary.step + x] == 255;
        printf("{\"n_contours\": %zu, \"n_positive\": %zu, \"n_negative\": %zu, \"n_armours\": %zu, \"mask_fg\": %llu,\n \"blobs\": [",
               contours.size(), positive.size(), negtive.size(), armours.size(), fg);
        for (size_t k = 0; k < positive.size(); ++k) {
            const rm::lightblob& b = positive[k];
            printf("%s[%.9g,%d,%.9g,%.9g,%.9g,%.9g", k ? "," : "", b.angle, (int)b.target, b.center.x, b.center.y, b.size.width, b.size.height);
            for (int i = 0; i < 4; ++i) printf(",%.9g,%.9g", b.vertices[i].x, b.vertices[i].y);
            printf("]");
        }
        printf("],\n \"armours\": [");
        for (size_t k = 0; k < armours.size(); ++k) {
            const rm::armour& a = armours[k];
            printf("%s[%.9g,%.9g,%.9g,%.9g", k ? "," : "", a.bounding_box.x, a.bounding_box.y, a.bounding_box.width, a.bounding_box.height);
            for (int i = 0; i < 4; ++i) printf(",%.9g,%.9g", a.icon[i].x, a.icon[i].y);
            for (int i = 0; i < 4; ++i) printf(",%.9g,%.9g", a.vertices[i].x, a.vertices[i].y);
            printf("]");
        }
        printf("],\n \"identity0\": %d, \"lost0\": %d}\n", armours.empty() ? -1 : armours[0].identity, armours.empty() ? 0 : armours[0].lost_count);
    } catch (const rm::gpu::error& e) {
        fprintf(stderr, "rm::gpu::error %d: %s\n", e.status, e.what());
        return 3;
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 4;
    }
    return 0;
}
