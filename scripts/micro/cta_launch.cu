// Micro-benchmark: when do the CTAs of a short kernel start?  (diagnostic for the launch ramp seen in k_hash_field_fwd)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cta_launch cta_launch.cu && ./cta_launch
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void k_prev(float* p) { p[blockIdx.x * blockDim.x + threadIdx.x] += 1.0f; }

template <int REGS_HINT>
__global__ void __launch_bounds__(128) k_probe(unsigned long long* out, int spin_us, int use_tmem, int smem_bytes) {
    extern __shared__ unsigned char smem[];
    __shared__ unsigned int tmem_base;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        out[3 * blockIdx.x] = t0;
        out[3 * blockIdx.x + 2] = smid;
    }
    if (use_tmem && threadIdx.x < 32) {
        unsigned int a = (unsigned int)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (smem_bytes > 0) smem[threadIdx.x] = (unsigned char)threadIdx.x;
    __syncthreads();
    unsigned long long t = t0;
    while (t - t0 < (unsigned long long)spin_us * 1000ull) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    __syncthreads();
    if (use_tmem && threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
    if (threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        out[3 * blockIdx.x + 1] = t;
    }
}

static void run(const char* name, int smem, int carve, int use_tmem, int prev_kernel) {
    const int grid = 576;
    unsigned long long* d;
    float* p;
    cudaMalloc(&d, grid * 3 * 8);
    cudaMalloc(&p, 1 << 22);
    cudaMemset(p, 0, 1 << 22);
    cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (carve >= 0) cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    std::vector<unsigned long long> h(grid * 3);
    for (int rep = 0; rep < 3; ++rep) {
        if (prev_kernel) k_prev<<<1024, 256>>>(p);
        k_probe<0><<<grid, 128, smem>>>(d, 10, use_tmem, smem);
        cudaDeviceSynchronize();
    }
    cudaError_t e = cudaGetLastError();
    cudaMemcpy(h.data(), d, grid * 3 * 8, cudaMemcpyDeviceToHost);
    unsigned long long g0 = ~0ull;
    for (int i = 0; i < grid; ++i) g0 = std::min(g0, h[3 * i]);
    std::vector<double> st;
    for (int i = 0; i < grid; ++i) st.push_back((h[3 * i] - g0) / 1e3);
    std::sort(st.begin(), st.end());
    printf("%-46s smem %6d carve %3d tmem %d prev %d : start p25 %.2f p50 %.2f p75 %.2f max %.2f us  (%s)\n", name, smem, carve, use_tmem,
           prev_kernel, st[grid / 4], st[grid / 2], st[3 * grid / 4], st[grid - 1], cudaGetErrorString(e));
    // per-SM start list for SM 0
    if (e != cudaSuccess) return;
    printf("    SM0 starts:");
    for (int i = 0; i < grid; ++i)
        if (h[3 * i + 2] == 0) printf(" %.2f", (h[3 * i] - g0) / 1e3);
    printf("\n");
    cudaFree(d);
    cudaFree(p);
}

int main() {
    run("no smem", 0, -1, 0, 1);
    run("45 KB smem", 45056, -1, 0, 1);
    run("45 KB smem, no previous kernel", 45056, -1, 0, 0);
    run("45 KB smem, carveout 100", 45056, 100, 0, 1);
    run("45 KB smem + tmem alloc 128", 45056, -1, 1, 1);
    run("45 KB smem + tmem alloc 128, carveout 100", 45056, 100, 1, 1);
    run("16 KB smem", 16384, -1, 0, 1);
    run("16 KB smem + tmem", 16384, -1, 1, 1);
    return 0;
}
