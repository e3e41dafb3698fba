// field_mlp.cu -- fused "mlp" (NeRF) field FORWARD for sm_100a: frequency encoding -> 8 x 256 ReLU MLP with a skip
// connection -> sigma_net / color_net tail, one kernel.  This is the teacher of BASELINE config 5 (mlp -> hash distillation);
// it is evaluated under torch.no_grad (distill_mutual/utils.py:1008-1018), so only the forward is fused.
//
// Replaces NeRFNetwork.forward for model_type "mlp" (distill_mutual/network.py:56-70,324-333,413-437 and
// tools/encoding.py:6-49): 21 sin/cos/cat kernels, 8 cuBLAS GEMMs with bias + ReLU + cat kernels between them, then the
// same 5-GEMM tail as the hash model -- 0.87 MFLOP per sample, the one part of the path where FLOPs dominate.
//
// Structure per 128-sample tile (one persistent CTA per SM, 128 threads, thread t owns sample row t):
//   * PE(10): 63 features (x, then sin/cos(2^k x), k = 0..9) -> fp16 operand tile X0 [128 x 64].
//   * every 256-wide layer is  D[128 x 256] (TMEM, 256 columns)  +=  A[128 x K] * W[256 x K]^T  streamed in K-chunks of 64:
//     a chunk of W is a 32 KB fp16 operand tile (pre-packed on the host side), double-buffered in shared memory and moved by
//     the TMA engine: ONE thread arms a "full" mbarrier with the byte count (mbarrier.arrive.expect_tx) and issues one
//     cp.async.bulk.shared::cluster.global (UBLKCP) per chunk; the same thread waits for it, issues the four tcgen05.mma
//     (N = 256, K = 16) of the chunk and commits them to the buffer's "empty" mbarrier, so the copy of chunk c+1 overlaps
//     the MMAs of chunk c and no other thread takes part in weight streaming (no block-wide barrier per chunk).
//   * the skip layer is two accumulating GEMMs: in_pts (X0, K = 64) and the 256 hidden units -- the concat never exists.
//   * layer epilogue: each thread pulls its own row from TMEM 16 columns at a time, adds the bias, applies ReLU, and writes
//     the next layer's A operand (fp16, chunk layout) over the previous one.
//   * the last layer (256 -> 28) feeds the sigma_net / color_net tail of field_tail.cuh unchanged.
#include <stdlib.h>
#include "common.cuh"
PVD_TRACE_TU(pvd_debug_trace_field_mlp)
#include "field_tail.cuh"

namespace pvd {

constexpr uint32_t kMlpChunkBytes = 256 * 64 * 2;   // one K-chunk of a 256-row weight matrix
constexpr uint32_t kMlpChunks = 26;                 // 1 (L0) + 3*4 (L1-3) + 1+4 (L4: in_pts part, hidden part) + 2*4 (L5-6)
constexpr uint32_t kMlpL7Bytes = 4 * (32 * 64 * 2); // 256 -> 28 as four [32 x 64] chunks
constexpr uint32_t kMlpBiasOff = kMlpChunks * kMlpChunkBytes + kMlpL7Bytes;
constexpr uint32_t kMlpBiasBytes = 8 * 256 * 4;
static_assert(kMlpBiasOff + kMlpBiasBytes == PVD_MLP_WBLOB_BYTES, "mlp blob size");

struct MlpArgs {
    const uint8_t* wblob;      // PVD_MLP_WBLOB_BYTES
    const uint8_t* tail_blob;  // PVD_FIELD_WBLOB_BYTES (sigma_net / color_net)
    float clip_min, clip_max, density_scale;
    uint32_t diag;   // timing diagnostics only (PVD_MLP_DIAG): 1 no bias loads, 2 no operand stores, 4 no TMEM loads, 8 no PE
};

// layer schedule: for each of the 26 streamed chunks, which layer it belongs to and where its A operand comes from
struct ChunkDesc {
    uint8_t layer;     // 0..6
    uint8_t from_x0;   // A = X0 tile (in_pts) instead of the activation tile
    uint8_t k_chunk;   // which 64-column slice of the activation tile
    uint8_t last;      // last chunk of its layer
};
__constant__ ChunkDesc kSchedule[kMlpChunks] = {
    {0, 1, 0, 1},
    {1, 0, 0, 0}, {1, 0, 1, 0}, {1, 0, 2, 0}, {1, 0, 3, 1},
    {2, 0, 0, 0}, {2, 0, 1, 0}, {2, 0, 2, 0}, {2, 0, 3, 1},
    {3, 0, 0, 0}, {3, 0, 1, 0}, {3, 0, 2, 0}, {3, 0, 3, 1},
    {4, 1, 0, 0}, {4, 0, 0, 0}, {4, 0, 1, 0}, {4, 0, 2, 0}, {4, 0, 3, 1},
    {5, 0, 0, 0}, {5, 0, 1, 0}, {5, 0, 2, 0}, {5, 0, 3, 1},
    {6, 0, 0, 0}, {6, 0, 1, 0}, {6, 0, 2, 0}, {6, 0, 3, 1},
};

// =============================================================================================== v2: warp-specialised, two tiles per CTA
// One persistent CTA per SM works on PAIRS of 128-sample tiles with ONE weight stream:
//   warps 0-3 (warpgroup 0) own the rows of tile 0, warps 4-7 (warpgroup 1) the rows of tile 1: PE, layer epilogues (TMEM -> bias ->
//     ReLU -> fp16 operand tile, in place), the sigma/colour tail and the outputs of their tile;
//   warp 8, one lane: TMA producer.  The packed weights are streamed in 16 KB pieces ([256 x 32] fp16 = half of a v1 chunk, same
//     blob) through a FOUR-stage ring; loads run up to four pieces ahead of the tensor core, across layers, tiles pairs and the
//     tail, so their latency (the bound of v1: a two-stage ring serialised copy -> MMA -> copy) is off the critical path;
//   warp 9, one lane: MMA issuer.  Every piece feeds 2 + 2 tcgen05.mma (tile 0, tile 1; N = 256, K = 16), so L2 -> shared weight
//     traffic per sample is halved; accumulators: tile t in TMEM columns [256 t, 256 t + 256) -- all 512 columns of the SM.
// Hand-offs are mbarriers only (no CTA-wide barrier in the steady state):
//   ring_full[s] (TMA bytes) / ring_empty[s] (tcgen05.commit)        producer <-> issuer
//   acc_full[t]  (tcgen05.commit after a layer's last MMA)           issuer   ->  warpgroup t: accumulator complete, operand tile dead
//   act_ready[t] (128 arrivals)                                      warpgroup t -> issuer: next operand tile written, TMEM drained
// Shared memory: 2 x 64 KB operand tiles + 2 x 16 KB PE tiles + 4 x 16 KB ring = 224 KB.  The 20 KB tail weights are copied (TMA)
// into the dead upper half of a tile's own operand buffer when its layer 7 has completed; biases are read through the constant-
// like __ldg path (all threads of a warp read the same address).
constexpr uint32_t kPiece = 16384;                    // bytes per streamed piece
constexpr uint32_t kPieces = 2 * kMlpChunks + 1;      // 52 half-chunks of layers 0-6 + layer 7
constexpr uint32_t kStages = 4;
constexpr uint32_t kV2Threads = 320;
constexpr uint32_t kTailWOff = 32768;
#ifndef PVD_MLP_STAGGER
#define PVD_MLP_STAGGER 1
#endif
constexpr bool kStaggerDefault = PVD_MLP_STAGGER != 0;  // 1: per layer, tile 0's pass then tile 1's pass (weights streamed twice, MMA of one tile under
                                                 // the epilogue of the other); 0: both tiles consume every piece (weights streamed once)                 // tail weights inside a tile's operand buffer

struct V2Wait {  // bounded waits that stop costing time after the first failure (a wrong barrier must not hang the GPU)
    int32_t* status;
    bool dead = false;
    __device__ __forceinline__ void operator()(uint64_t* bar, uint32_t parity) {
        if (!tc5::mbar_wait(bar, parity, dead ? 1u : (1u << 22))) {
            dead = true;
            atomicExch(status, 2);
        }
    }
};

__global__ void __launch_bounds__(kV2Threads, 1) k_mlp_field_fwd(MlpArgs a, const float* __restrict__ xyzs, const float* __restrict__ dirs,
                                                                uint32_t M, float* __restrict__ sigmas, float* __restrict__ rgbs,
                                                                float* __restrict__ feat16, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t ring_full[kStages], ring_empty[kStages], acc_full[2], act_ready[2], tail_bar[2], tail_w[2];
    __shared__ uint32_t tmem_base_s;
    uint8_t* const ring = smem + 2 * 65536 + 2 * 16384;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;

    if (tid == 0) {
        for (uint32_t s = 0; s < kStages; ++s) {
            tc5::mbar_init(&ring_full[s], 1);
            tc5::mbar_init(&ring_empty[s], 1);
        }
        for (uint32_t t = 0; t < 2; ++t) {
            tc5::mbar_init(&acc_full[t], 1);
            tc5::mbar_init(&act_ready[t], 128);
            tc5::mbar_init(&tail_bar[t], 1);
            tc5::mbar_init(&tail_w[t], 1);
        }
        tc5::mbar_fence_init();
    }
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    const uint32_t n_pairs = (n_tiles + 1) / 2;
    V2Wait wait{status};
    const bool kStagger = kStaggerDefault != ((a.diag & 16u) != 0u);   // diag bit 16 flips the schedule

    if (warp == 8) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t pc = 0;  // pieces issued so far (ring position)
            auto load = [&](uint32_t q) {
                const uint32_t s = pc % kStages;
                if (pc >= kStages) wait(&ring_empty[s], ((pc / kStages) - 1u) & 1u);
                tc5::mbar_expect_tx(&ring_full[s], kPiece);
                tc5::bulk_g2s(tc5::smem_u32(ring + s * kPiece), a.wblob + (size_t)q * kPiece, kPiece, &ring_full[s]);
                ++pc;
            };
            for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                if (kStagger) {   // per layer: the pieces for tile 0's pass, then the same pieces again for tile 1's pass
                    uint32_t q0 = 0;
                    for (uint32_t layer = 0; layer < 8; ++layer) {
                        const uint32_t n_pieces = (layer == 0) ? 2u : (layer == 4 ? 10u : (layer == 7 ? 1u : 8u));
                        for (uint32_t t = 0; t < 2; ++t)
                            for (uint32_t j = 0; j < n_pieces; ++j) load(q0 + j);
                        q0 += n_pieces;
                    }
                } else {
                    for (uint32_t q = 0; q < kPieces; ++q) load(q);
                }
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t pc = 0, acts = 0;  // pieces consumed; act_ready phases consumed (same count for both tiles)
            const uint32_t idesc = tc5::instr_desc_f16(128, 256, 0, 0);
            const uint32_t idesc7 = tc5::instr_desc_f16(128, 32, 0, 0);
            for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                uint32_t q0 = 0;
                for (uint32_t layer = 0; layer < 7; ++layer) {
                    const uint32_t n_pieces = (layer == 0) ? 2u : (layer == 4 ? 10u : 8u);
                    if (kStagger) {
                        // tile 0's whole layer, then tile 1's: while the tensor core runs tile 1, warpgroup 0 is already in its epilogue
                        for (uint32_t t = 0; t < 2; ++t) {
                            wait(&act_ready[t], acts & 1u);  // operand tile of this layer written, accumulator drained
                            for (uint32_t j = 0; j < n_pieces; ++j, ++pc) {
                                const uint32_t q = q0 + j;
                                const ChunkDesc cd = kSchedule[q >> 1];
                                const uint32_t s = pc % kStages;
                                wait(&ring_full[s], (pc / kStages) & 1u);
                                tc5::fence_after_sync();
                                const uint32_t b_tile = tc5::smem_u32(ring + s * kPiece);
                                const uint32_t a_base = cd.from_x0 ? tc5::smem_u32(smem + 2 * 65536 + t * 16384)
                                                                   : tc5::smem_u32(smem + t * 65536) + (uint32_t)cd.k_chunk * 8u * (kTile * 16u);
                                const uint32_t a_tile = a_base + (q & 1u) * 4u * (kTile * 16u);
#pragma unroll
                                for (uint32_t k0 = 0; k0 < 32; k0 += 16)
                                    tc5::mma_f16_ss(tmem + 256u * t, tc5::desc_kmajor(a_tile, kTile, k0), tc5::desc_kmajor(b_tile, 256, k0), idesc,
                                                    !(j == 0 && k0 == 0));
                                tc5::mma_commit(&ring_empty[s]);
                            }
                            tc5::mma_commit(&acc_full[t]);
                        }
                    } else {
                        for (uint32_t j = 0; j < n_pieces; ++j, ++pc) {
                            const uint32_t q = q0 + j;
                            const ChunkDesc cd = kSchedule[q >> 1];
                            const uint32_t s = pc % kStages;
                            wait(&ring_full[s], (pc / kStages) & 1u);
                            const uint32_t b_tile = tc5::smem_u32(ring + s * kPiece);
                            for (uint32_t t = 0; t < 2; ++t) {
                                if (j == 0) wait(&act_ready[t], acts & 1u);
                                tc5::fence_after_sync();
                                const uint32_t a_base = cd.from_x0 ? tc5::smem_u32(smem + 2 * 65536 + t * 16384)
                                                                   : tc5::smem_u32(smem + t * 65536) + (uint32_t)cd.k_chunk * 8u * (kTile * 16u);
                                const uint32_t a_tile = a_base + (q & 1u) * 4u * (kTile * 16u);   // which 32-column half of the 64-column slice
#pragma unroll
                                for (uint32_t k0 = 0; k0 < 32; k0 += 16)
                                    tc5::mma_f16_ss(tmem + 256u * t, tc5::desc_kmajor(a_tile, kTile, k0), tc5::desc_kmajor(b_tile, 256, k0), idesc,
                                                    !(j == 0 && k0 == 0));
                                if (j + 1 == n_pieces) tc5::mma_commit(&acc_full[t]);
                            }
                            tc5::mma_commit(&ring_empty[s]);
                        }
                    }
                    q0 += n_pieces;
                    ++acts;
                }
                // layer 7: 256 -> 28 (N = 32); the piece holds four [32 x 64] operand tiles
                for (uint32_t t = 0; t < 2; ++t) {
                    const uint32_t s = pc % kStages;
                    if (kStagger || t == 0) wait(&ring_full[s], (pc / kStages) & 1u);
                    const uint32_t b_base = tc5::smem_u32(ring + s * kPiece);
                    wait(&act_ready[t], acts & 1u);
                    tc5::fence_after_sync();
                    for (uint32_t c = 0; c < 4; ++c)
#pragma unroll
                        for (uint32_t k0 = 0; k0 < 64; k0 += 16)
                            tc5::mma_f16_ss(tmem + 256u * t, tc5::desc_kmajor(tc5::smem_u32(smem + t * 65536) + c * 8u * (kTile * 16u), kTile, k0),
                                            tc5::desc_kmajor(b_base + c * (32u * 64u * 2u), 32, k0), idesc7, !(c == 0 && k0 == 0));
                    tc5::mma_commit(&acc_full[t]);
                    if (kStagger || t == 1) {
                        tc5::mma_commit(&ring_empty[s]);
                        ++pc;
                    }
                }
                ++acts;
            }
        }
    } else {
        // ------------------------------------------------------------------ the two warpgroups: one tile each
        const uint32_t t = warp >> 2;          // tile of the pair / warpgroup
        const uint32_t r = tid & 127u;         // row inside the tile
        uint8_t* const A = smem + t * 65536;
        uint8_t* const X0 = smem + 2 * 65536 + t * 16384;
        const uint32_t trow = tc5::tmem_addr(tmem + 256u * t, (warp & 3u) * 32u, 0);
        const float* __restrict__ bias = reinterpret_cast<const float*>(a.wblob + kMlpBiasOff);
        Pipe p{&tail_bar[t], 0u, tmem + 256u * t, status};
        p.team = 1u + t;
        uint32_t accs = 0, pairs_done = 0;
        for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pairs_done) {
            const uint32_t row = (2u * pair + t) * kTile + r;
            const bool live = row < M;
            float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
                    dir[d] = __ldg(dirs + 3 * (size_t)row + d);
                }
            }
            {   // FreqEncoder (tools/encoding.py:36-49): [x, sin(f0 x), cos(f0 x), sin(f1 x), ...], f_k = 2^k, k = 0..9
                float f[64];
                f[0] = pos[0]; f[1] = pos[1]; f[2] = pos[2];
                float freq = 1.0f;
#pragma unroll
                for (int k = 0; k < 10; ++k) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        float sn = 0.f, cs = 0.f;
                        if (!(a.diag & 8u)) sincosf(pos[d] * freq, &sn, &cs);
                        f[3 + 6 * k + d] = sn;
                        f[3 + 6 * k + 3 + d] = cs;
                    }
                    freq *= 2.0f;
                }
                f[63] = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(X0 + tc5::chunk_off(kTile, r, j)) = tc5::pack8(f + 8 * j);
            }
            tc5::fence_async_smem();
            tc5::fence_before_sync();
            tc5::mbar_arrive(&act_ready[t]);
            // ---- layers 0..6: wait for the accumulator, bias + ReLU, rewrite the operand tile in place
            for (uint32_t layer = 0; layer < 7; ++layer, ++accs) {
                wait(&acc_full[t], accs & 1u);
                tc5::fence_after_sync();
                const float4* __restrict__ bl = reinterpret_cast<const float4*>(bias + 256 * layer);
                uint32_t buf[2][32];
                const bool no_bias = a.diag & 1u, no_store = a.diag & 2u, no_ld = a.diag & 4u;
                if (!no_ld) tc5::tmem_ld32_issue(trow, buf[0]);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (!no_ld) {
                        tc5::tmem_ld_wait();
                        if (c + 1 < 8) tc5::tmem_ld32_issue(trow + 32 * (c + 1), buf[(c + 1) & 1]);
                    }
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = no_bias ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(bl + 8 * c + i);
                        v[4 * i + 0] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 0]) + b4.x, 0.0f);
                        v[4 * i + 1] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 1]) + b4.y, 0.0f);
                        v[4 * i + 2] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 2]) + b4.z, 0.0f);
                        v[4 * i + 3] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 3]) + b4.w, 0.0f);
                    }
                    if (!no_store) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(A + tc5::chunk_off(kTile, r, 4 * c + i)) = tc5::pack8(v + 8 * i);
                    }
                }
                tc5::fence_async_smem();
                tc5::fence_before_sync();
                tc5::mbar_arrive(&act_ready[t]);
            }
            // ---- layer 7 -> x28; then the sigma / colour tail on this warpgroup's tile
            wait(&acc_full[t], accs & 1u);
            ++accs;
            tc5::fence_after_sync();
            if (r == 0) {  // the operand tile is dead: fetch the tail weights into its upper half while x28 is formed
                tc5::mbar_expect_tx(&tail_w[t], PVD_FIELD_WBLOB_BYTES);
                tc5::bulk_g2s(tc5::smem_u32(A + kTailWOff), a.tail_blob, PVD_FIELD_WBLOB_BYTES, &tail_w[t]);
            }
            uint8_t* X = A;                 // 8192
            uint8_t* CIN = A;               // aliases X (dead after the first tail layer)
            uint8_t* H = A + 8192;          // 16384 : H1, H3, H4
            {
                float v[32];
                tc5::tmem_ld16(trow, *reinterpret_cast<float(*)[16]>(&v[0]));
                tc5::tmem_ld16(trow + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
                const float* b7 = bias + 256 * 7;
#pragma unroll
                for (int i = 0; i < 28; ++i) v[i] += __ldg(b7 + i);
#pragma unroll
                for (int i = 28; i < 32; ++i) v[i] = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(X + tc5::chunk_off(kTile, r, j)) = tc5::pack8(v + 8 * j);
            }
            FieldArgs fa;
            fa.clip_min = a.clip_min; fa.clip_max = a.clip_max; fa.density_scale = a.density_scale;
            float sigma, o16[16];
            FwdRegs fr;
            p.wbar = &tail_w[t];
            p.wphase = pairs_done & 1u;
            mlp_forward(p, fa, A + kTailWOff, X, H, CIN, H, H, dir, r, sigma, o16, fr);
            if (live) {
                sigmas[row] = sigma;
                rgbs[3 * (size_t)row] = fr.rgb[0];
                rgbs[3 * (size_t)row + 1] = fr.rgb[1];
                rgbs[3 * (size_t)row + 2] = fr.rgb[2];
                if (feat16) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(feat16 + 16 * (size_t)row + 4 * q) =
                            make_float4(o16[4 * q], o16[4 * q + 1], o16[4 * q + 2], o16[4 * q + 3]);
                }
            }
            // the tail's last TMEM read / its tiles must be done before the next pair's PE arrival lets layer 0 overwrite them
            tc5::fence_before_sync();
            asm volatile("bar.sync %0, 128;" ::"r"(1u + t) : "memory");
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem, 512);
}

constexpr size_t kMlpSmemV2 = 2 * 65536 + 2 * 16384 + kStages * kPiece;  // 229376

// pack nerf_mlp weights/biases (fp32, nn.Linear layout) into the streamed blob
__global__ void k_mlp_pack(const float* const* __restrict__ w, const float* const* __restrict__ b, uint8_t* __restrict__ blob) {
    // chunk index -> (layer, first input column of the slice).  Layer 4's input is cat([in_pts(63), hidden(256)]) (network.py:331-332)
    const uint32_t ci = blockIdx.x;
    if (ci < kMlpChunks) {
        const ChunkDesc cd = kSchedule[ci];
        const uint32_t in_dim = (cd.layer == 0) ? 63u : (cd.layer == 4 ? 319u : 256u);
        uint32_t col0, ncols;
        if (cd.layer == 0) { col0 = 0; ncols = 63; }
        else if (cd.layer == 4 && cd.from_x0) { col0 = 0; ncols = 63; }
        else if (cd.layer == 4) { col0 = 63 + 64 * cd.k_chunk; ncols = 64; }
        else { col0 = 64 * cd.k_chunk; ncols = 64; }
        uint8_t* tile = blob + (size_t)ci * kMlpChunkBytes;
        const float* W = w[cd.layer];
        for (uint32_t e = threadIdx.x; e < 256 * 64; e += blockDim.x) {
            const uint32_t r = e / 64, k = e - r * 64;
            const float v = (k < ncols) ? W[(size_t)r * in_dim + col0 + k] : 0.0f;
            *reinterpret_cast<__half*>(tile + tc5::chunk_off(256, r, k >> 3) + (k & 7u) * 2) = __float2half_rn(v);
        }
    } else if (ci < kMlpChunks + 4) {  // layer 7: [28 x 256] as four [32 x 64] chunks
        const uint32_t c = ci - kMlpChunks;
        uint8_t* tile = blob + (size_t)kMlpChunks * kMlpChunkBytes + c * (32 * 64 * 2);
        const float* W = w[7];
        for (uint32_t e = threadIdx.x; e < 32 * 64; e += blockDim.x) {
            const uint32_t r = e / 64, k = e - r * 64;
            const float v = (r < 28) ? W[(size_t)r * 256 + 64 * c + k] : 0.0f;
            *reinterpret_cast<__half*>(tile + tc5::chunk_off(32, r, k >> 3) + (k & 7u) * 2) = __float2half_rn(v);
        }
    } else {  // biases: 8 x 256 floats (layer 7 has 28)
        float* bb = reinterpret_cast<float*>(blob + kMlpBiasOff);
        for (uint32_t e = threadIdx.x; e < 8 * 256; e += blockDim.x) {
            const uint32_t l = e / 256, i = e - l * 256;
            bb[e] = (l < 7 || i < 28) ? b[l][i] : 0.0f;
        }
    }
}

constexpr size_t kMlpSmem = 16384 + 65536 + 2 * kMlpChunkBytes + PVD_FIELD_WBLOB_BYTES + kMlpBiasBytes;  // 176128

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_mlp_pack_weights(const float* const* weights8, const float* const* biases8, void* wblob, void* stream) {
    PVD_REQUIRE(weights8 && biases8 && wblob);
    k_mlp_pack<<<kMlpChunks + 5, 256, 0, (cudaStream_t)stream>>>(weights8, biases8, (uint8_t*)wblob);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_mlp_field_forward(const PvdMlpField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                          float* feat16, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && f->wblob && f->tail_wblob && xyzs && dirs && sigmas && rgbs && status);
    MlpArgs a;
    a.wblob = reinterpret_cast<const uint8_t*>(f->wblob);
    a.tail_blob = reinterpret_cast<const uint8_t*>(f->tail_wblob);
    a.clip_min = f->sigma_clip_min; a.clip_max = f->sigma_clip_max; a.density_scale = f->density_scale;
    static const uint32_t diag = []() { const char* v = getenv("PVD_MLP_DIAG"); return v ? (uint32_t)atoi(v) : 0u; }();
    a.diag = diag;
    const uint32_t tiles = (M + kTile - 1) / kTile;
    const uint32_t pairs = (tiles + 1) / 2;
    const uint32_t grid2 = min(pairs, (uint32_t)sm_count());
    cudaError_t e2 = cudaFuncSetAttribute(k_mlp_field_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMlpSmemV2);
    if (e2 != cudaSuccess) return (int)e2;
    k_mlp_field_fwd<<<grid2, kV2Threads, kMlpSmemV2, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, feat16, status);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
