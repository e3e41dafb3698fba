// red_peak.cu -- measured peaks for the two L2-bound inner loops of the hash field on B200:
//   * scatter (k_hash_field_bwd / k_hash_scatter): random red.global.add into an L2-resident gradient table
//       f32 scalar / .v2.f32 / .v4.f32 into 42 MB (the fp32 accumulator) and .noftz.f16x2 scalar / .v2 / .v4 into 21 MB (an fp16
//       accumulator, the reference's own precision, gridencoder.cu:299-305)
//   * gather (k_hash_field_fwd): random 4-byte / 8-byte / 16-byte / 32-byte ld.global.nc from a 21 MB L2-resident table
// Addresses are a hash of (thread, iteration): uniform over the buffer, no index loads, every lane of a warp in a different sector
// (the worst case the hashed levels produce).  Output: one JSON line, operations per second and the "sector rate" (one operation
// touches one 32-byte sector) -- the denominators DESIGN.md quotes the scatter / gather kernels against.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/micro/red_peak.cu -o scripts/micro/red_peak && scripts/micro/red_peak
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE 0: red f32 x1, 1: red v2.f32, 2: red v4.f32, 3: red f16x2 x1, 4: red v2.f16x2, 5: red v4.f16x2
// MODE 10: ld 4 B, 11: ld 8 B, 12: ld 16 B, 13: ld 32 B (two 16-byte loads of one sector)
template <int MODE>
__global__ void __launch_bounds__(256) k_probe(uint8_t* __restrict__ buf, uint32_t slots /* power of two, in units of the access */, uint32_t iters,
                                               float* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    constexpr uint32_t BYTES = (MODE == 0 || MODE == 3 || MODE == 10) ? 4u : (MODE == 1 || MODE == 4 || MODE == 11) ? 8u
                               : (MODE == 2 || MODE == 5 || MODE == 12) ? 16u : 32u;
#pragma unroll 4
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t s = mix(tid * 0x9e3779b9u + it * 0x85ebca6bu) & (slots - 1u);
        uint8_t* p = buf + (size_t)s * BYTES;
        if constexpr (MODE == 0) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
        } else if constexpr (MODE == 1) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(2.0f) : "memory");
        } else if constexpr (MODE == 2) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
        } else if constexpr (MODE == 3) {
            asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(p), "r"(0x3c003c00u) : "memory");
        } else if constexpr (MODE == 4) {
            asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(0x3c003c00u), "r"(0x3c003c00u) : "memory");
        } else if constexpr (MODE == 5) {
            asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u),
                         "r"(0x3c003c00u)
                         : "memory");
        } else if constexpr (MODE == 10) {
            uint32_t v;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
            acc += __uint_as_float(v);
        } else if constexpr (MODE == 11) {
            uint32_t a, b;
            asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
            acc += __uint_as_float(a ^ b);
        } else if constexpr (MODE == 12) {
            uint32_t a, b, c, d;
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
            acc += __uint_as_float(a ^ b ^ c ^ d);
        } else {
            uint32_t a, b, c, d, e, f, g, h;
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
            acc += __uint_as_float(a ^ b ^ c ^ d ^ e ^ f ^ g ^ h);
        }
    }
    if (acc == 123.456f) sink[0] = acc;   // keeps the loads alive
}

template <int MODE>
static double run(uint8_t* buf, size_t buf_bytes, uint32_t access_bytes, int blocks, uint32_t iters, float* sink) {
    uint32_t slots = 1;
    while ((size_t)slots * 2 * access_bytes <= buf_bytes) slots *= 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<float> ts;
    for (int rep = 0; rep < 7; ++rep) {
        cudaEventRecord(a);
        k_probe<MODE><<<blocks, 256>>>(buf, slots, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    const double ops = (double)blocks * 256.0 * iters;
    return ops / (ts[ts.size() / 2] * 1e-3);   // operations per second
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t big = 64ull << 20;    // slots are cut to a power of two: 32 MiB of f32 / 16 MiB of f16 addressed -> L2-resident
    uint8_t* buf;
    float* sink;
    CK(cudaMalloc(&buf, big));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, big));
    const int blocks = sms * 8;
    const uint32_t iters = 32;   // 148 * 8 * 256 * 32 = 9.7 M operations per launch: the scale of one training step's scatter
    printf("{\"sms\": %d, \"ops_per_launch\": %.0f", sms, (double)blocks * 256.0 * iters);
    // reductions: fp32 accumulator table of 32 MiB, fp16 accumulator table of 16 MiB (the nearest powers of two to 42 / 21 MB)
    const size_t t32 = 32ull << 20, t16 = 16ull << 20;
    double r;
    r = run<0>(buf, t32, 4, blocks, iters, sink);  printf(", \"red_f32_x1_Gops\": %.2f", r / 1e9);
    r = run<1>(buf, t32, 8, blocks, iters, sink);  printf(", \"red_f32_v2_Gops\": %.2f", r / 1e9);
    r = run<2>(buf, t32, 16, blocks, iters, sink); printf(", \"red_f32_v4_Gops\": %.2f", r / 1e9);
    CK(cudaMemset(buf, 0, big));
    r = run<3>(buf, t16, 4, blocks, iters, sink);  printf(", \"red_f16x2_x1_Gops\": %.2f", r / 1e9);
    r = run<4>(buf, t16, 8, blocks, iters, sink);  printf(", \"red_f16x2_v2_Gops\": %.2f", r / 1e9);
    r = run<5>(buf, t16, 16, blocks, iters, sink); printf(", \"red_f16x2_v4_Gops\": %.2f", r / 1e9);
    r = run<10>(buf, t16, 4, blocks, iters, sink);  printf(", \"ld_4B_Gops\": %.2f", r / 1e9);
    r = run<11>(buf, t16, 8, blocks, iters, sink);  printf(", \"ld_8B_Gops\": %.2f", r / 1e9);
    r = run<12>(buf, t16, 16, blocks, iters, sink); printf(", \"ld_16B_Gops\": %.2f", r / 1e9);
    r = run<13>(buf, t16, 32, blocks, iters, sink); printf(", \"ld_32B_Gops\": %.2f", r / 1e9);
    // occupancy sensitivity of the fp32 v2 reduction (what the scatter issues most): 2 and 16 CTAs of 256 threads per SM
    r = run<1>(buf, t32, 8, sms * 2, iters * 4, sink);  printf(", \"red_f32_v2_Gops_2cta\": %.2f", r / 1e9);
    r = run<1>(buf, t32, 8, sms * 16, iters / 2, sink); printf(", \"red_f32_v2_Gops_16cta\": %.2f", r / 1e9);
    printf("}\n");
    CK(cudaGetLastError());
    return 0;
}
