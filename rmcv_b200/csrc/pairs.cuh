// Block-level helpers shared by the frame kernel and the standalone rm::filter_armours kernel.
#pragma once
#include <cuda_runtime.h>

namespace rmcv {

// Exclusive block scan of one int per thread (blockDim.x <= 1024).  sh: 33 ints of shared scratch.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    __syncthreads();  // protect sh from the previous use
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarps ? sh[lane] : 0;
        int i2 = w;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) i2 += u;
        }
        sh[lane] = i2 - w;
        if (lane == 31) sh[32] = i2;
    }
    __syncthreads();
    *total = sh[32];
    return sh[warp] + incl - v;
}

// Same for a 64-bit value (two packed 32-bit counters).  sh: 34 long longs (272 bytes) of shared scratch.
__device__ __forceinline__ long long block_excl_scan64(long long v, long long* total, long long* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < nwarps ? sh[lane] : 0;
        long long i2 = w;
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) i2 += u;
        }
        __syncwarp();
        sh[lane] = i2 - w;
        if (lane == 31) sh[32] = i2;
    }
    __syncthreads();
    *total = sh[32];
    return sh[warp] + incl - v;
}

// pair index k (row-major over i<j) -> (i, j): the order of the reference's double loop (src/objdetect.cpp:122-126)
__device__ __forceinline__ void pair_from_index(long long k, int P, int* pi, int* pj) {
    const double b = 2.0 * P - 1.0;
    int i = (int)floor((b - sqrt(b * b - 8.0 * (double)k)) * 0.5);
    if (i < 0) i = 0;
    if (i > P - 2) i = P - 2;
    auto offs = [P](long long ii) { return ii * (2LL * P - ii - 1) / 2; };
    while (i > 0 && offs(i) > k) --i;
    while (i < P - 2 && offs(i + 1) <= k) ++i;
    *pi = i;
    *pj = (int)(k - offs(i)) + i + 1;
}

__device__ __forceinline__ void copy_words(void* dst, const void* src, size_t bytes, int tid, int nt) {
    // records are multiples of 8 bytes, 8-byte aligned; 16-byte words when both ends allow (fewer, larger PCIe writes)
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15) == 0) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        const size_t n = bytes / 16;
        for (size_t i = tid; i < n; i += nt) d[i] = s[i];
        return;
    }
    const uint64_t* s = reinterpret_cast<const uint64_t*>(src);
    uint64_t* d = reinterpret_cast<uint64_t*>(dst);
    const size_t n = bytes / 8;
    for (size_t i = tid; i < n; i += nt) d[i] = s[i];
}

}  // namespace rmcv
