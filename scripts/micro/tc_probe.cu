// tc_probe.cu -- self-test of the tcgen05 building blocks (tc5.cuh) used by the fused field kernels:
// one CTA stages two fp16 matrices into chunk-layout shared-memory tiles, issues tcgen05.mma with the
// descriptor forms the field kernels rely on, and dumps the raw TMEM accumulator (128 lanes x N columns).
//   mode 0: D[128,N]  = A[128,K] * B[N,K]^T            (A K-major, B K-major)       -- forward layer
//   mode 1: D[128,N]  = A[128,K] * W[K,N]               (A K-major, B MN-major)      -- data gradient
//   mode 2: D[64,N]   = sum_s T1[s,0:64]^T T2[s,0:N]    (both MN-major, M = 64)      -- weight gradient
//   mode 3: as mode 2 with M = 128 (T1 has 128 columns)
// A stand-alone diagnostic (scripts/micro/tc_probe.py builds and runs it); not part of libpvd_b200.so or its ABI.
#include "common.cuh"
#include "tc5.cuh"

namespace pvd {

__global__ void __launch_bounds__(128) k_tc_probe(int mode, const __half* __restrict__ Ag, uint32_t RA, uint32_t CA,
                                                 const __half* __restrict__ Bg, uint32_t RB, uint32_t CB, float* __restrict__ out,
                                                 uint32_t N, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* tileA = smem;                        // RA x CA halves, chunk layout
    uint8_t* tileB = smem + 128 * 128 * 2;        // up to 32 KB reserved for A
    const uint32_t tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t e = tid; e < RA * (CA / 8); e += blockDim.x) {
        const uint32_t r = e % RA, j = e / RA;
        *reinterpret_cast<uint4*>(tileA + tc5::chunk_off(RA, r, j)) = *reinterpret_cast<const uint4*>(Ag + (size_t)r * CA + 8 * j);
    }
    for (uint32_t e = tid; e < RB * (CB / 8); e += blockDim.x) {
        const uint32_t r = e % RB, j = e / RB;
        *reinterpret_cast<uint4*>(tileB + tc5::chunk_off(RB, r, j)) = *reinterpret_cast<const uint4*>(Bg + (size_t)r * CB + 8 * j);
    }
    if (tid == 0) {
        tc5::mbar_init(&bar, 1);
        tc5::mbar_fence_init();
    }
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, 128);
    tc5::fence_async_smem();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tbase = tmem_base_s;

    if (tid == 0) {
        const uint32_t sa = tc5::smem_u32(tileA), sb = tc5::smem_u32(tileB);
        if (mode == 0) {
            const uint32_t idesc = tc5::instr_desc_f16(128, N, 0, 0);
            for (uint32_t k0 = 0; k0 < CA; k0 += 16)
                tc5::mma_f16_ss(tbase, tc5::desc_kmajor(sa, RA, k0), tc5::desc_kmajor(sb, RB, k0), idesc, k0 > 0);
        } else if (mode == 1) {
            const uint32_t idesc = tc5::instr_desc_f16(128, N, 0, 1);
            for (uint32_t k0 = 0; k0 < CA; k0 += 16)
                tc5::mma_f16_ss(tbase, tc5::desc_kmajor(sa, RA, k0), tc5::desc_mnmajor(sb, RB, k0, 0), idesc, k0 > 0);
        } else {
            const uint32_t idesc = tc5::instr_desc_f16(mode == 2 ? 64 : 128, N, 1, 1);
            for (uint32_t k0 = 0; k0 < RA; k0 += 16)
                tc5::mma_f16_ss(tbase, tc5::desc_mnmajor(sa, RA, k0, 0), tc5::desc_mnmajor(sb, RB, k0, 0), idesc, k0 > 0);
        }
        tc5::mma_commit(&bar);
    }
    const bool ok = tc5::mbar_wait(&bar, 0);
    tc5::fence_after_sync();
    if (!ok && tid == 0) atomicExch(status, 1);
    if (ok) {
        for (uint32_t c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc5::tmem_ld16(tc5::tmem_addr(tbase, warp * 32, c0), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(size_t)tid * N + c0 + i] = v[i];
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tbase, 128);
}

}  // namespace pvd

extern "C" int pvd_tc_probe(int mode, const void* A, uint32_t RA, uint32_t CA, const void* B, uint32_t RB, uint32_t CB,
                            float* out, uint32_t N, int* status, void* stream) {
    PVD_REQUIRE(A && B && out && status);
    PVD_REQUIRE(RA <= 128 && CA <= 128 && RB <= 128 && CB <= 128 && CA % 8 == 0 && CB % 8 == 0 && N % 8 == 0 && N <= 128);
    const size_t smem = 2 * 128 * 128 * 2;
    cudaError_t e = cudaFuncSetAttribute(pvd::k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    pvd::k_tc_probe<<<1, 128, smem, (cudaStream_t)stream>>>(mode, (const __half*)A, RA, CA, (const __half*)B, RB, CB, out, N,
                                                            status);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}
