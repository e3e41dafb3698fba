// field_mlp.cu -- fused "mlp" (NeRF) field FORWARD for sm_100a: frequency encoding -> 8 x 256 ReLU MLP with a skip
// connection -> sigma_net / color_net tail, one kernel.  This is the teacher of BASELINE config 5 (mlp -> hash distillation),
// evaluated under torch.no_grad (distill_mutual/utils.py:1008-1018), and -- with SAVE -- the forward of a TRAINED mlp model
// (main_just_train_tea.py --model_type mlp), whose backward is field_mlp_bwd.cu.
//
// Replaces NeRFNetwork.forward for model_type "mlp" (distill_mutual/network.py:56-70,324-333,413-437 and
// tools/encoding.py:6-49): 21 sin/cos/cat kernels, 8 cuBLAS GEMMs with bias + ReLU + cat kernels between them, then the
// same 5-GEMM tail as the hash model -- 0.87 MFLOP per sample, the one part of the path where FLOPs dominate.
//
// One persistent CTA per SM works on PAIRS of 128-sample tiles with ONE weight stream (19 warps):
//   warps 0-7 own tile 0, warps 8-15 tile 1.  Inside a tile, warp w serves the TMEM lane quadrant w % 4 (= sample rows
//     32 (w % 4) .. +31: a warp may only touch its own quadrant) and the COLUMN HALF (w / 4) % 2 of the 256-wide layer: two threads
//     per sample row, 128 columns each.  PE, layer epilogues (TMEM -> bias -> ReLU -> fp16 operand tile, in place) are split that
//     way; the sigma/colour tail and the outputs of a tile belong to its column-half-0 warpgroup.  Round 1 ran ONE thread per row
//     (8 epilogue warps per SM, two per scheduler): the epilogue's dependent chain tcgen05.ld -> bias -> ReLU -> pack -> st.shared
//     was latency-bound at ~3 us per layer against 1.1 us of MMAs; 16 warps halve the per-thread chain and double the warps a
//     scheduler can switch between.
//   warp 16, one lane: TMA producer.  The packed weights are streamed in 16 KB pieces ([256 x 32] fp16) through a FOUR-stage ring
//     of cp.async.bulk copies; loads run up to four pieces ahead of the tensor core, across layers, tile pairs and the tail.
//   warps 17 / 18, one lane each: MMA issuers of tile 0 / tile 1, both consuming every ring piece (a stage is released when both
//     have committed).  Accumulators: tile t in TMEM columns [256 t, 256 t + 256) -- all 512 columns of the SM.
// Hand-offs are mbarriers only (no CTA-wide barrier in the steady state):
//   ring_full[s] (TMA bytes) / ring_empty[s] (2 x tcgen05.commit)    producer <-> issuers
//   acc_full[t]  (tcgen05.commit after a layer's last MMA)           issuer   ->  the 8 warps of tile t: accumulator complete
//   act_ready[t] (256 arrivals)                                      tile t's warps -> issuer: next operand tile written, TMEM drained
// Shared memory: 2 x 64 KB operand tiles + 2 x 16 KB PE tiles + 4 x 16 KB ring = 224 KB.  The 20 KB tail weights are copied (TMA)
// into the dead upper half of a tile's own operand buffer when its layer 7 has completed; biases are read through the constant-
// like __ldg path (all threads of a warp read the same address).
// SAVE (training): every operand tile the tensor core consumed (PE tile, act_1..act_7; fp16 chunk layout, 464 KB per tile) and
// the 28-wide trunk output (the tail's "encoding", [M,32] fp16) also go to global memory -- warp-wide 512-byte row stores -- for
// the backward kernels.
#include <stdlib.h>
#include "field_mlp.cuh"
PVD_TRACE_TU(pvd_debug_trace_field_mlp)

namespace pvd {

constexpr uint32_t kEpiWarps = 16;
constexpr uint32_t kV2Threads = 32 * (kEpiWarps + 3);   // 608: 16 epilogue warps, TMA producer, two MMA issuers
constexpr uint32_t kTailWOff = 32768;                   // tail weights inside a tile's operand buffer
// FreqEncoder (tools/encoding.py:36-49): [x, sin(f0 x), cos(f0 x), sin(f1 x), ...], f_k = 2^k, k = 0..9 -> 63 features (+1 zero pad);
// this thread's 32 of them (features 32 HALF .. 32 HALF + 31).  Feature 3 + 6k + d = sin(2^k x_d), 3 + 6k + 3 + d = cos(2^k x_d).
template <int HALF>
__device__ __forceinline__ void pe_half(const float (&pos)[3], float (&f)[32], bool skip) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = 0.0f;
    if (HALF == 0) { f[0] = pos[0]; f[1] = pos[1]; f[2] = pos[2]; }
    constexpr int k_lo = HALF == 0 ? 0 : 4, k_hi = HALF == 0 ? 4 : 9, lo = 32 * HALF, hi = 32 * HALF + 32;
#pragma unroll
    for (int k = k_lo; k <= k_hi; ++k) {
        const float freq = (float)(1 << k);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int gs = 3 + 6 * k + d, gc = gs + 3;
            const bool need_s = gs >= lo && gs < hi, need_c = gc >= lo && gc < hi;
            if (need_s || need_c) {
                float sn = 0.f, cs = 0.f;
                if (!skip) sincosf(pos[d] * freq, &sn, &cs);
                if (need_s) f[gs - lo] = sn;
                if (need_c) f[gc - lo] = cs;
            }
        }
    }
}

template <bool SAVE>
__global__ void __launch_bounds__(kV2Threads, 1) k_mlp_field_fwd(MlpArgs a, const float* __restrict__ xyzs, const float* __restrict__ dirs,
                                                                uint32_t M, float* __restrict__ sigmas, float* __restrict__ rgbs,
                                                                float* __restrict__ feat16, uint8_t* __restrict__ save,
                                                                __half* __restrict__ enc_out, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t ring_full[kStages], ring_empty[kStages], acc_full[2], act_ready[2], tail_bar[2], tail_w[2];
    __shared__ uint32_t tmem_base_s;
    uint8_t* const ring = smem + 2 * 65536 + 2 * 16384;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;

    if (tid == 0) {
        for (uint32_t s = 0; s < kStages; ++s) {
            tc5::mbar_init(&ring_full[s], 1);
            tc5::mbar_init(&ring_empty[s], 2);
        }
        for (uint32_t t = 0; t < 2; ++t) {
            tc5::mbar_init(&acc_full[t], 1);
            tc5::mbar_init(&act_ready[t], 256);
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
    a.wblob += (size_t)(blockIdx.x % a.replicas) * PVD_MLP_WBLOB_BYTES;

    if (warp == kEpiWarps) {
        // ------------------------------------------------------------------ TMA producer: ONE weight stream for both tiles
        if (lane == 0) {
            uint32_t pc = 0;  // pieces issued so far (ring position)
            // diag (timing experiments only): bit 32 = no copies (the full barrier is just arrived on)
            for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
                for (uint32_t q = 0; q < kPieces; ++q, ++pc) {
                    const uint32_t s = pc % kStages;
                    if (pc >= kStages) wait(&ring_empty[s], ((pc / kStages) - 1u) & 1u);   // BOTH issuers' MMAs on this stage are complete
                    if (a.diag & 32u) {
                        tc5::mbar_arrive(&ring_full[s]);
                    } else {
                        tc5::mbar_expect_tx(&ring_full[s], kPiece);
                        tc5::bulk_g2s(tc5::smem_u32(ring + s * kPiece), a.wblob + (size_t)q * kPiece, kPiece, &ring_full[s]);
                    }
                }
        }
    } else if (warp > kEpiWarps) {
        // ------------------------------------------------------------------ MMA issuers: warp 17 for tile 0, warp 18 for tile 1
        // Both consume the SAME ring pieces (a stage is released when both have committed: ring_empty counts 2).  Measured with the
        // kernel's diag bits (no copies, no MMAs, empty epilogues): one issuer lane spent ~450 cycles per piece in waits, fences and
        // commits -- more than the 256 cycles of tensor-core time its two MMAs per piece are worth -- so ONE issuer could not keep the
        // pipe half busy whatever the epilogue did.  Two issuer lanes halve the pieces per MMA and double the issue rate.
        if (lane == 0) {
            const uint32_t t = warp - (kEpiWarps + 1u);
            uint32_t pc = 0, acts = 0;  // pieces consumed; act_ready phases consumed
            const uint32_t idesc = tc5::instr_desc_f16(128, 256, 0, 0);
            const uint32_t idesc7 = tc5::instr_desc_f16(128, 32, 0, 0);
            const bool no_mma = (a.diag & 64u) != 0u;   // timing experiments only: commits without MMAs
            const uint32_t a_tile0 = tc5::smem_u32(smem + t * 65536), x0_tile = tc5::smem_u32(smem + 2 * 65536 + t * 16384);
            const uint32_t d_tmem = tmem + 256u * t;
            for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (uint32_t layer = 0; layer < 7; ++layer, ++acts) {
                    // pieces of this layer: [256 x 32] slices of its K range; layer 0 = the PE tile (2 pieces), layer 4 = PE tile
                    // (2 pieces) then the 256 hidden units (8 pieces), the others 8 pieces of the activation tile
                    const uint32_t n_x0 = (layer == 0u || layer == 4u) ? 2u : 0u, n_pieces = (layer == 0u) ? 2u : n_x0 + 8u;
                    wait(&act_ready[t], acts & 1u);   // operand tile of this layer written, accumulator drained
                    for (uint32_t j = 0; j < n_pieces; ++j, ++pc) {
                        const uint32_t s = pc % kStages;
                        wait(&ring_full[s], (pc / kStages) & 1u);
                        tc5::fence_after_sync();
                        const uint32_t b_tile = tc5::smem_u32(ring + s * kPiece);
                        const uint32_t a_tile = (j < n_x0) ? x0_tile + j * 4u * (kTile * 16u) : a_tile0 + (j - n_x0) * 4u * (kTile * 16u);
                        if (!no_mma) {
#pragma unroll
                            for (uint32_t k0 = 0; k0 < 32; k0 += 16)
                                tc5::mma_f16_ss(d_tmem, tc5::desc_kmajor(a_tile, kTile, k0), tc5::desc_kmajor(b_tile, 256, k0), idesc, !(j == 0 && k0 == 0));
                        }
                        tc5::mma_commit(&ring_empty[s]);
                    }
                    tc5::mma_commit(&acc_full[t]);
                }
                // layer 7: 256 -> 28 (N = 32); the piece holds four [32 x 64] operand tiles
                {
                    const uint32_t s = pc % kStages;
                    wait(&act_ready[t], acts & 1u);
                    wait(&ring_full[s], (pc / kStages) & 1u);
                    tc5::fence_after_sync();
                    const uint32_t b_base = tc5::smem_u32(ring + s * kPiece);
                    if (!no_mma) {
                        for (uint32_t c = 0; c < 4; ++c)
#pragma unroll
                            for (uint32_t k0 = 0; k0 < 64; k0 += 16)
                                tc5::mma_f16_ss(d_tmem, tc5::desc_kmajor(a_tile0 + c * 8u * (kTile * 16u), kTile, k0),
                                                tc5::desc_kmajor(b_base + c * (32u * 64u * 2u), 32, k0), idesc7, !(c == 0 && k0 == 0));
                    }
                    tc5::mma_commit(&acc_full[t]);
                    tc5::mma_commit(&ring_empty[s]);
                    ++pc;
                    ++acts;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ the epilogue warps: 8 per tile, two threads per row
        const uint32_t t = warp >> 3;                    // tile of the pair
        const uint32_t half = (warp >> 2) & 1u;          // column half of the 256-wide layers
        const uint32_t r = (warp & 3u) * 32u + lane;     // row inside the tile (TMEM lane)
        uint8_t* const A = smem + t * 65536;
        uint8_t* const X0 = smem + 2 * 65536 + t * 16384;
        const uint32_t trow = tc5::tmem_addr(tmem + 256u * t, (warp & 3u) * 32u, 0);
        const float* __restrict__ bias = reinterpret_cast<const float*>(a.wblob + kMlpBiasOff);
        Pipe p{&tail_bar[t], 0u, tmem + 256u * t, status};
        p.team = 1u + t;
        uint32_t accs = 0, pairs_done = 0;
        for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pairs_done) {
            const uint32_t tile = 2u * pair + t;
            const uint32_t row = tile * kTile + r;
            const bool live = row < M;
            uint8_t* const sv = (SAVE && tile < n_tiles) ? save + (size_t)tile * kSaveTileBytes : nullptr;   // whole tiles: rows past M hold the image of x = 0
            float pos[3] = {0.f, 0.f, 0.f};
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; ++d) pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
            }
            {
                float f[32];
                if (half == 0) pe_half<0>(pos, f, (a.diag & 8u) != 0u);
                else pe_half<1>(pos, f, (a.diag & 8u) != 0u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 u = tc5::pack8(f + 8 * j);
                    const uint32_t off = tc5::chunk_off(kTile, r, 4 * half + j);
                    *reinterpret_cast<uint4*>(X0 + off) = u;
                    if (SAVE && sv) *reinterpret_cast<uint4*>(sv + kSavePe + off) = u;
                }
            }
            tc5::fence_async_smem();
            tc5::fence_before_sync();
            tc5::mbar_arrive(&act_ready[t]);
            // ---- layers 0..6: wait for the accumulator, bias + ReLU, rewrite this thread's half of the operand row in place
            for (uint32_t layer = 0; layer < 7; ++layer, ++accs) {
                wait(&acc_full[t], accs & 1u);
                tc5::fence_after_sync();
                const float4* __restrict__ bl = reinterpret_cast<const float4*>(bias + 256 * layer + 128 * half);
                const uint32_t tcol = trow + 128u * half;
                uint8_t* const svl = (SAVE && sv) ? sv + kSaveAct + layer * 65536u : nullptr;
                // 16 columns per tcgen05.ld, the next load in flight under bias + ReLU + pack + 2 x st.shared of the current one
                // (18 warps = 5 on one scheduler: 96 registers per thread, so the x32 double buffer of the 8-warp version does not fit)
                uint32_t buf[2][16];
                const bool no_bias = a.diag & 1u, no_store = a.diag & 2u, no_ld = a.diag & 4u;
                if (!no_ld) tc5::tmem_ld16_issue(tcol, buf[0]);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (!no_ld) {
                        tc5::tmem_ld_wait();
                        if (c + 1 < 8) tc5::tmem_ld16_issue(tcol + 16 * (c + 1), buf[(c + 1) & 1]);
                    }
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b4 = no_bias ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(bl + 4 * c + i);
                        v[4 * i + 0] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 0]) + b4.x, 0.0f);
                        v[4 * i + 1] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 1]) + b4.y, 0.0f);
                        v[4 * i + 2] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 2]) + b4.z, 0.0f);
                        v[4 * i + 3] = fmaxf(__uint_as_float(buf[c & 1][4 * i + 3]) + b4.w, 0.0f);
                    }
                    if (!no_store) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint4 u = tc5::pack8(v + 8 * i);
                            const uint32_t off = tc5::chunk_off(kTile, r, 16 * half + 2 * c + i);
                            *reinterpret_cast<uint4*>(A + off) = u;
                            if (SAVE && svl) *reinterpret_cast<uint4*>(svl + off) = u;
                        }
                    }
                }
                tc5::fence_async_smem();
                tc5::fence_before_sync();
                tc5::mbar_arrive(&act_ready[t]);
            }
            // ---- layer 7 -> x28; then the sigma / colour tail on this tile's column-half-0 warpgroup
            wait(&acc_full[t], accs & 1u);
            ++accs;
            if (half != 0u) continue;   // the other warpgroup goes on to the next pair's PE (X0 is free: layer 4 of this pair is complete)
            tc5::fence_after_sync();
            float dir[3] = {0.f, 0.f, 0.f};
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; ++d) dir[d] = __ldg(dirs + 3 * (size_t)row + d);
            }
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
                for (int j = 0; j < 4; ++j) {
                    const uint4 u = tc5::pack8(v + 8 * j);
                    *reinterpret_cast<uint4*>(X + tc5::chunk_off(kTile, r, j)) = u;
                    if (SAVE && live) *reinterpret_cast<uint4*>(enc_out + (size_t)row * PVD_FIELD_ENC_STRIDE + 8 * j) = u;
                }
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
            // the tail's last TMEM read / its tiles must be done before this warpgroup's next PE arrival lets layer 0 overwrite them
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

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_mlp_pack_weights(const float* const* weights8, const float* const* biases8, void* wblob, void* stream) {
    PVD_REQUIRE(weights8 && biases8 && wblob);
    k_mlp_pack<<<kMlpChunks + 5, 256, 0, (cudaStream_t)stream>>>(weights8, biases8, (uint8_t*)wblob);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_mlp_replicate_weights(void* wblob, uint64_t bytes, uint32_t replicas, void* stream) {
    PVD_REQUIRE(wblob && bytes);
    for (uint32_t r = 1; r < replicas; ++r) {
        cudaError_t e = cudaMemcpyAsync((uint8_t*)wblob + (size_t)r * bytes, wblob, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    return PVD_OK;
}

static int mlp_forward_launch(const PvdMlpField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                              float* feat16, void* save, void* enc, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && f->wblob && f->tail_wblob && xyzs && dirs && sigmas && rgbs && status);
    MlpArgs a;
    a.wblob = reinterpret_cast<const uint8_t*>(f->wblob);
    a.tail_blob = reinterpret_cast<const uint8_t*>(f->tail_wblob);
    a.clip_min = f->sigma_clip_min; a.clip_max = f->sigma_clip_max; a.density_scale = f->density_scale;
    a.replicas = f->replicas ? f->replicas : 1u;
    static const uint32_t diag = []() { const char* v = getenv("PVD_MLP_DIAG"); return v ? (uint32_t)atoi(v) : 0u; }();
    a.diag = diag;
    const uint32_t tiles = (M + kTile - 1) / kTile;
    const uint32_t pairs = (tiles + 1) / 2;
    const uint32_t grid2 = min(pairs, (uint32_t)sm_count());
    cudaError_t e2;
    if (save != nullptr) {
        e2 = cudaFuncSetAttribute(k_mlp_field_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMlpSmemV2);
        if (e2 != cudaSuccess) return (int)e2;
        k_mlp_field_fwd<true><<<grid2, kV2Threads, kMlpSmemV2, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, feat16, (uint8_t*)save,
                                                                                        (__half*)enc, status);
    } else {
        e2 = cudaFuncSetAttribute(k_mlp_field_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMlpSmemV2);
        if (e2 != cudaSuccess) return (int)e2;
        k_mlp_field_fwd<false><<<grid2, kV2Threads, kMlpSmemV2, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, feat16, nullptr, nullptr, status);
    }
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_mlp_field_forward(const PvdMlpField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                          float* feat16, int32_t* status, void* stream) {
    return mlp_forward_launch(f, xyzs, dirs, M, sigmas, rgbs, feat16, nullptr, nullptr, status, stream);
}

int pvd_mlp_field_forward_train(const PvdMlpField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                                float* feat16, void* save_ws, void* enc, int32_t* status, void* stream) {
    PVD_REQUIRE(save_ws && enc);
    return mlp_forward_launch(f, xyzs, dirs, M, sigmas, rgbs, feat16, save_ws, enc, status, stream);
}

}  // extern "C"
