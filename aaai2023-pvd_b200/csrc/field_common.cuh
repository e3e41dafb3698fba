// field_common.cuh -- pieces shared by the fused field kernels (field_hash.cu, field_vm.cu): the tcgen05 issue helpers over
// chunk-layout tiles, the MMA <-> thread hand-off, per-row epilogues, and the TMEM accumulator flush.
#pragma once
#include "common.cuh"
#include "tc5.cuh"

namespace pvd {

constexpr uint32_t kTile = 128;

struct Pipe {
    uint64_t* bar;
    uint32_t phase;
    uint32_t tmem;
    int32_t* status;
    uint32_t trec = 0;  // timeline record of this CTA (PVD_TRACE builds only)
    uint64_t* wbar = nullptr;  // non-null while the TMA copy of the weight blob may still be in flight (thread 0 waits once)
    uint32_t wphase = 0;       // parity of `wbar` to wait for
    uint32_t team = 0;         // 0: the whole CTA works on one tile (barrier 0, leader = thread 0); n > 0: a 128-thread warpgroup
                               // works on its own tile (named barrier n, leader = the warpgroup's first thread)
    bool dead = false;         // a bounded wait of this thread has timed out: later waits poll once instead of 2^22 times, so a wedged
                               // pipeline costs one long wait per thread, not one per layer
};

__device__ __forceinline__ bool team_leader(const Pipe& p) { return p.team ? ((threadIdx.x & 127u) == 0u) : (threadIdx.x == 0u); }

// Every thread: publish shared-memory operand writes to the async proxy and order prior TMEM reads, then barrier.
__device__ __forceinline__ void operands_ready() {
    tc5::fence_async_smem();
    tc5::fence_before_sync();
    __syncthreads();
}
// The same for the team of `p` (see Pipe::team).
__device__ __forceinline__ void operands_ready(const Pipe& p) {
    tc5::fence_async_smem();
    tc5::fence_before_sync();
    if (p.team == 0u) __syncthreads();
    else asm volatile("bar.sync %0, 128;" ::"r"(p.team) : "memory");
}
// Every thread: wait for the MMAs committed by thread 0.
__device__ __forceinline__ void mma_wait(Pipe& p) {
    if (!tc5::mbar_wait(p.bar, p.phase, p.dead ? 1u : (1u << 22))) {
        p.dead = true;
        atomicExch(p.status, 1);
    }
    p.phase ^= 1u;
    tc5::fence_after_sync();
}

// D[128 x N] (=|+=) A[128 x K] * B[N x K]^T, both K-major chunk tiles (forward layer)
__device__ __forceinline__ void issue_fwd(uint32_t d_tmem, uint32_t a_tile, uint32_t K, uint32_t b_tile, uint32_t b_rows, uint32_t N) {
    const uint32_t idesc = tc5::instr_desc_f16(128, N, 0, 0);
    for (uint32_t k0 = 0; k0 < K; k0 += 16)
        tc5::mma_f16_ss(d_tmem, tc5::desc_kmajor(a_tile, kTile, k0), tc5::desc_kmajor(b_tile, b_rows, k0), idesc, k0 > 0);
}
// D[128 x N] = G[128 x K] * W[K x N] with W stored as the forward operand tile [K rows(out) x N cols(in)] (data gradient)
__device__ __forceinline__ void issue_dgrad(uint32_t d_tmem, uint32_t g_tile, uint32_t K, uint32_t w_tile, uint32_t w_rows, uint32_t N) {
    const uint32_t idesc = tc5::instr_desc_f16(128, N, 0, 1);
    for (uint32_t k0 = 0; k0 < K; k0 += 16)
        tc5::mma_f16_ss(d_tmem, tc5::desc_kmajor(g_tile, kTile, k0), tc5::desc_mnmajor(w_tile, w_rows, k0, 0), idesc, k0 > 0);
}
// D[64 x N] (+)= P[128 x 64]^T * Q[128 x N]  (weight gradient: reduction over the 128 samples of the tile)
__device__ __forceinline__ void issue_wgrad(uint32_t d_tmem, uint32_t p_tile, uint32_t q_tile, uint32_t N, bool first) {
    const uint32_t idesc = tc5::instr_desc_f16(64, N, 1, 1);
    for (uint32_t s0 = 0; s0 < kTile; s0 += 16)
        tc5::mma_f16_ss(d_tmem, tc5::desc_mnmajor(p_tile, kTile, s0, 0), tc5::desc_mnmajor(q_tile, kTile, s0, 0), idesc,
                        !(first && s0 == 0));
}

// this thread's row of a [128 x 16*NC16] TMEM accumulator -> ReLU -> fp16 chunk tile
template <int NC16>
__device__ __forceinline__ void relu_to_tile(uint32_t tmem_row, uint8_t* tile, uint32_t row) {
#pragma unroll
    for (int c = 0; c < NC16; ++c) {
        float v[16];
        tc5::tmem_ld16(tmem_row + 16 * c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
        *reinterpret_cast<uint4*>(tile + tc5::chunk_off(kTile, row, 2 * c)) = tc5::pack8(v);
        *reinterpret_cast<uint4*>(tile + tc5::chunk_off(kTile, row, 2 * c + 1)) = tc5::pack8(v + 8);
    }
}

// this thread's row of a 64-wide data gradient, masked by the sign of the saved activation, written IN PLACE over it
__device__ __forceinline__ void mask_grad_in_place(uint32_t tmem_row, uint8_t* tile, uint32_t row) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v[16];
        tc5::tmem_ld16(tmem_row + 16 * c, v);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint4* p = reinterpret_cast<uint4*>(tile + tc5::chunk_off(kTile, row, 2 * c + h));
            const uint4 a = *p;
            const __half2* ah = reinterpret_cast<const __half2*>(&a);
            float g[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 af = __half22float2(ah[i]);
                g[2 * i] = af.x > 0.0f ? v[8 * h + 2 * i] : 0.0f;
                g[2 * i + 1] = af.y > 0.0f ? v[8 * h + 2 * i + 1] : 0.0f;
            }
            *p = tc5::pack8(g);
        }
    }
}

__device__ __forceinline__ void flush_acc(uint32_t tmem_base, uint32_t col, uint32_t ncols, float* __restrict__ dst, uint32_t max_rows = 64) {
    // M = 64 accumulator: row m lives in TMEM lane (m/16)*32 + m%16 (verified on B200, profiles/r01_tcgen05_probe.log)
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t m = warp * 16 + lane;
    for (uint32_t c = 0; c < ncols; c += 16) {
        float v[16];
        tc5::tmem_ld16(tc5::tmem_addr(tmem_base, warp * 32, col + c), v);
        if (lane < 16 && m < max_rows) {
            float* d = dst + (size_t)m * ncols + c;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                             : "memory");
        }
    }
}

// Thread 0, before its first MMA that reads the weight tiles: wait for the bulk copy started by stage_blob_async.  The tensor
// core reads shared memory through the async proxy, the same proxy TMA wrote it through: no proxy fence is needed.
__device__ __forceinline__ void weights_ready(Pipe& p) {
    if (p.wbar != nullptr) {
        if (!tc5::mbar_wait(p.wbar, p.wphase, p.dead ? 1u : (1u << 22))) {
            p.dead = true;
            atomicExch(p.status, 1);
        }
        p.wbar = nullptr;
    }
}
// Thread 0: start ONE TMA bulk copy of the packed weight tiles into shared memory (mbarrier `wbar`, initialised with count 1,
// completes when the bytes have landed).  Nothing waits here: the copy overlaps the gather / the first tile's loads.
__device__ __forceinline__ void stage_blob_async(uint8_t* smw, const uint8_t* __restrict__ blob, uint32_t bytes, uint64_t* wbar) {
    tc5::mbar_expect_tx(wbar, bytes);
    tc5::bulk_g2s(tc5::smem_u32(smw), blob, bytes, wbar);
}

// copy `bytes` (multiple of 16) of packed weight tiles from global to shared memory
__device__ __forceinline__ void stage_blob(uint8_t* smw, const uint8_t* __restrict__ blob, uint32_t bytes) {
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smw)[i] = __ldg(reinterpret_cast<const uint4*>(blob) + i);
}

// pack one fp32 [out, in] matrix into an fp16 chunk tile of R rows x K cols (zero padded); the whole GRID strides over the tile
// (the pack sits on the critical path of a step whose parameters an optimizer has just changed: one CTA took 12 us)
__device__ __forceinline__ void pack_matrix(const float* __restrict__ w, uint32_t out, uint32_t in, uint8_t* tile, uint32_t R,
                                            uint32_t K) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < R * K; e += gridDim.x * blockDim.x) {
        const uint32_t r = e / K, k = e - r * K;
        const float v = (r < out && k < in) ? w[(size_t)r * in + k] : 0.0f;
        *reinterpret_cast<__half*>(tile + tc5::chunk_off(R, r, k >> 3) + (k & 7u) * 2) = __float2half_rn(v);
    }
}

inline int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace pvd
