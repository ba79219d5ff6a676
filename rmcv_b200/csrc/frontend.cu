// f4: camera front-end variants of hardware/src/daheng.cpp:91-187 ahead of the Bayer pixel kernel:
//   DxRaw16toRaw8(DX_BIT_2_9 / DX_BIT_4_11)   10/12-bit samples in 16-bit containers -> bits 2..9 / 4..11 (daheng.cpp:111,130)
//   DxImageMirror(HORIZONTAL_MIRROR)           folded into the column index                              (daheng.cpp:108)
//   the bFlip argument of DxRaw8toRGB24        folded into the row index                                 (daheng.cpp:112)
// One streaming pass raw -> canonical 8-bit mosaic (2 or 1 B/px read, 1 B/px written); the Bayer kernel then runs on it
// with the layout rmcv_frontend_layout() gives (a mirror shifts the column phase, a flip the row phase of the mosaic).
// The Daheng SDK is closed source: like row a0 this restates the documented behaviour of those calls (parity unpinned).
#include "common.cuh"

namespace rmcv {

struct FrontParams {
    const uint8_t* src; size_t pitch, frame_stride;   // bytes
    uint8_t* dst; size_t dpitch, dframe_stride;
    int W, H, shift, bytes_per_px, mirror, flip;
};

__global__ void __launch_bounds__(256) frontend_kernel(const FrontParams p) {
    const int y = blockIdx.y, f = blockIdx.z;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;   // four output pixels per thread
    if (x4 >= p.W) return;
    const int sy = p.flip ? p.H - 1 - y : y;
    const uint8_t* row = p.src + (size_t)f * p.frame_stride + (size_t)sy * p.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x4 + i;
        if (x >= p.W) break;
        const int sx = p.mirror ? p.W - 1 - x : x;
        uint32_t v;
        if (p.bytes_per_px == 2) v = (__ldg(reinterpret_cast<const uint16_t*>(row) + sx) >> p.shift) & 0xffu;
        else v = __ldg(row + sx);
        out |= v << (8 * i);
    }
    uint8_t* d = p.dst + (size_t)f * p.dframe_stride + (size_t)y * p.dpitch + x4;
    if (x4 + 4 <= p.W && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = out;
    else for (int i = 0; i < 4 && x4 + i < p.W; ++i) d[i] = (uint8_t)(out >> (8 * i));
}

cudaError_t launch_frontend(const uint8_t* d_src, size_t pitch, size_t frame_stride, uint8_t* d_dst, size_t dpitch,
                            size_t dframe_stride, int W, int H, int batch, int bits, int mirror, int flip, cudaStream_t st,
                            int64_t* launches) {
    FrontParams p;
    p.src = d_src; p.pitch = pitch; p.frame_stride = frame_stride;
    p.dst = d_dst; p.dpitch = dpitch; p.dframe_stride = dframe_stride;
    p.W = W; p.H = H; p.mirror = mirror; p.flip = flip;
    p.bytes_per_px = bits > 8 ? 2 : 1;
    p.shift = bits == 12 ? 4 : (bits == 10 ? 2 : 0);
    if (H > 65535 || batch > 65535) return cudaErrorInvalidValue;
    dim3 grid(((W + 3) / 4 + 255) / 256, H, batch);
    frontend_kernel<<<grid, 256, 0, st>>>(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace rmcv
