// The reference's call site (executable/main.cpp:172-176) compiled against include/rmcv_gpu/rm_shim.hpp in its
// OpenCV-free mode.  Reads a raw BGR frame (W H then W*H*3 bytes) from argv[1], runs the three rm:: calls with the
// reference's literal parameters and prints the results as JSON for tests/test_gpu_shim.py to compare with the oracle.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/rmcv_gpu/rm_shim.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: call_site frame.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    int wh[2];
    if (fread(wh, sizeof(int), 2, f) != 2) return 2;
    cv::Mat image(wh[1], wh[0], 3);
    if (fread(image.data, 1, (size_t)wh[0] * wh[1] * 3, f) != (size_t)wh[0] * wh[1] * 3) return 2;
    fclose(f);
    try {
        auto [contours, binary] = rm::extract_color(image, rm::CAMP_BLUE, 80);
        auto [positive, negtive] = rm::filter_lightblobs(contours, 70, {1.5, 80}, {10, 99999}, rm::CAMP_BLUE);
        auto armours = rm::filter_armours(positive, 12, 22, 0.4, rm::CAMP_BLUE);
        // fused variant must agree with the three-call variant
        rmcv_params prm;
        rmcv_default_params(&prm);
        auto det = rm::gpu::detect(image, prm);
        // the three-call path reuses the device results of extract_color (SURVEY 8(b) "hidden handle"); a call with other
        // parameters or foreign contours must not
        const long long reused_lb = rm::gpu::default_context().last.reused_lightblobs, reused_ar = rm::gpu::default_context().last.reused_armours;
        auto [positive2, negative2] = rm::filter_lightblobs(contours, 60, {1.5, 80}, {10, 99999}, rm::CAMP_BLUE);   // other tilt_max: standalone kernels
        const long long reused_lb2 = rm::gpu::default_context().last.reused_lightblobs;
        auto t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < 20; ++rep) {
            auto [c3, b3] = rm::extract_color(image, rm::CAMP_BLUE, 80);
            auto [p3, n3] = rm::filter_lightblobs(c3, 70, {1.5, 80}, {10, 99999}, rm::CAMP_BLUE);
            auto a3 = rm::filter_armours(p3, 12, 22, 0.4, rm::CAMP_BLUE);
        }
        const double three_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / 20;
        t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < 20; ++rep) {
            cv::Mat bin2(image.rows, image.cols, 1);   // a fresh mask per frame, like rm::extract_color has to return
            rm::gpu::detect(image, prm, &bin2);
        }
        const double fused_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / 20;
        // legacy entry points (include/objdetect.h:22-37,62) over the same contours
        std::vector<rm::lightblob> legacy;
        rm::FindLightBlobs(contours, legacy, 1.5f, 80.f, 70.f, 10.f, 99999.f, image, false);   // box from cv::minAreaRect
        cv::RotatedRect one;
        const bool matched0 = !contours.empty() && rm::MatchLightBlob(contours[0], 1.5f, 80.f, 70.f, 10.f, 99999.f, one, true);
        const bool overlap = legacy.size() >= 3 && rm::LightBlobOverlap(legacy, 0, (int)legacy.size() - 1);
        // next row f2: the icon crop of every armour (src/imgproc.cpp:9-35) through the shim; FNV-1a of the 20x20x3 bytes
        std::vector<unsigned long long> icon_hash;
        for (auto& arm : armours) {
            cv::Mat icon = rm::affine_correction(image, arm.icon, cv::Size{20, 20});
            unsigned long long h = 1469598103934665603ull;
            for (int y = 0; y < icon.rows; ++y)
                for (int x = 0; x < icon.cols * 3; ++x) { h ^= icon.data[(size_t)y * icon.step + x]; h *= 1099511628211ull; }
            icon_hash.push_back(h);
        }
        unsigned long long fg = 0;
        for (int y = 0; y < binary.rows; ++y)
            for (int x = 0; x < binary.cols; ++x) fg += binary.data[(size_t)y * binary.step + x] == 255;
        printf("{\"n_contours\": %zu, \"n_positive\": %zu, \"n_negative\": %zu, \"n_armours\": %zu, \"mask_fg\": %llu, "
               "\"fused_positive\": %zu, \"fused_armours\": %zu, \"legacy_count\": %zu, \"legacy_blue\": %d, \"matched0\": %d, \"overlap\": %d, "
               "\"reused_lightblobs\": %lld, \"reused_armours\": %lld, \"reused_after_foreign_params\": %lld, \"n_positive_tilt60\": %zu, "
               "\"three_call_ms\": %.4f, \"fused_ms\": %.4f,\n \"contour_sizes\": [",
               contours.size(), positive.size(), negtive.size(), armours.size(), fg, det.positive.size(), det.armours.size(), legacy.size(),
               (int)(legacy.empty() ? 0 : legacy[0].target == rm::CAMP_BLUE), (int)matched0, (int)overlap, reused_lb, reused_ar,
               reused_lb2 - reused_lb, positive2.size(), three_ms, fused_ms);
        for (size_t k = 0; k < contours.size(); ++k) printf("%s%zu", k ? "," : "", contours[k].size());
        printf("],\n \"first_points\": [");
        for (size_t k = 0; k < contours.size(); ++k) printf("%s[%d,%d]", k ? "," : "", contours[k][0].x, contours[k][0].y);
        printf("],\n \"blob_centers\": [");
        for (size_t k = 0; k < positive.size(); ++k) printf("%s[%.6f,%.6f,%.6f]", k ? "," : "", positive[k].center.x, positive[k].center.y, positive[k].angle);
        printf("],\n \"armour_boxes\": [");
        for (size_t k = 0; k < armours.size(); ++k)
            printf("%s[%.1f,%.1f,%.1f,%.1f]", k ? "," : "", armours[k].bounding_box.x, armours[k].bounding_box.y, armours[k].bounding_box.width,
                   armours[k].bounding_box.height);
        printf("],\n \"icon_hashes\": [");
        for (size_t k = 0; k < icon_hash.size(); ++k) printf("%s\"%llu\"", k ? "," : "", icon_hash[k]);
        printf("]}\n");
    } catch (const rm::gpu::error& e) {
        fprintf(stderr, "rm::gpu::error %d: %s\n", e.status, e.what());
        return 3;
    }
    return 0;
}
