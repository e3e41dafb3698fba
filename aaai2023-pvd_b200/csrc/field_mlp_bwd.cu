// field_mlp_bwd.cu -- backward of the fused NeRF-MLP trunk (8 x 256 with a skip, distill_mutual/network.py:56-70,324-333) for sm_100a:
// what autograd does in the reference when an mlp model is TRAINED (main_just_train_tea.py --model_type mlp): 8 cuBLAS dgrad
// GEMMs + 8 wgrad GEMMs + 8 bias reductions + the ReLU / cat / slice mirrors between them.
//
// The sigma/colour tail is back-propagated by the hash model's tcgen05 tail kernel (field_hash.cu, dx_ws mode), which leaves
// d(loss)/d(x28) [M,32] fp16.  From there, two kernels:
//
// k_mlp_trunk_bwd  -- DATA gradients, the forward's structure run backwards.  One persistent CTA per SM, pairs of 128-sample
//   tiles, 16 epilogue warps (two threads per sample row), one TMA producer lane, two MMA issuer lanes (one per tile, same ring pieces).  Round rho = 0..6 computes
//   d(act_{7-rho}) [128 x 256] = G_{7-rho} [128 x K] * W_{7-rho} [K x 256] on the tensor core (accumulator in TMEM), the weights
//   streamed as the TRANSPOSED operand pieces of pvd_mlp_pack_weights_t ([256 in x 32 out] fp16, 16 KB, four-stage ring -- the
//   forward's machinery with the roles of in and out swapped; layer 4 contributes only its hidden part, d(in_pts) has no consumer),
//   and the epilogue forms G_{6-rho} = d(act) where the SAVED activation is positive (mask straight from the fp16 operand tile the
//   forward stored: 16-byte loads, compared two halves at a time), writes it as the next round's operand tile in place AND to
//   global memory for the weight-gradient kernel.
//
// k_mlp_wgrad  -- WEIGHT gradients: dW_l = G_l^T act_l is a reduction over ALL samples, [256 x 256] fp32 = the whole TMEM of an SM,
//   so it cannot ride along in the per-tile kernel.  One (layer, sample-split) job per CTA: the job's two operand streams (G_l tile
//   64 KB, double-buffered; act_l in 16 KB pieces of 64 columns through a four-stage ring) are TMA bulk copies of the saved tiles,
//   both consumed as MN-major operands (the reduction index = the 128 samples = the ROWS of the chunk layout, tc5.cuh), and the
//   accumulators D_h [128 out x N in], h = 0, 1, stay in TMEM over all the job's tiles (512 columns); they leave the SM once, as
//   vector reductions into the fp32 workspace.  While the tensor core works, the CTA's four otherwise idle warps form the bias
//   gradient: column sums of the G_l tile that sits in shared memory anyway (bank-conflict-free skewed row order).
//   HBM-bound by construction: 1.0 MB per tile over the nine jobs against 17 MFLOP per job-tile.
#include <stdlib.h>
#include "field_mlp.cuh"

namespace pvd {

constexpr uint32_t kBwdPieces = 49;                     // layer 7^T (1) + layers 6..1 (8 each)
static_assert(kBwdPieces * kPiece == PVD_MLP_WBLOB_T_BYTES, "transposed blob size");
constexpr uint32_t kEpiWarpsB = 16;
constexpr uint32_t kBwdThreads = 32 * (kEpiWarpsB + 3);  // 608: 16 epilogue warps, TMA producer, two MMA issuers
constexpr size_t kTrunkSmem = 2 * 65536 + kStages * kPiece;  // 196608

// piece q of the transposed stream: B operand [256 rows = layer INPUT index] x [32 cols = 32 of the layer's OUTPUTS], K-major
__global__ void k_mlp_pack_t(const float* const* __restrict__ w, uint8_t* __restrict__ blob) {
    const uint32_t q = blockIdx.x;
    uint32_t layer, o0, col0, in_dim, n_out;
    if (q == 0) { layer = 7; o0 = 0; col0 = 0; in_dim = 256; n_out = 28; }
    else {
        layer = 6u - (q - 1u) / 8u;
        o0 = 32u * ((q - 1u) % 8u);
        col0 = (layer == 4u) ? 63u : 0u;          // hidden part of cat([in_pts, h]) (network.py:331-332)
        in_dim = (layer == 4u) ? 319u : 256u;
        n_out = 256;
    }
    const float* W = w[layer];
    uint8_t* tile = blob + (size_t)q * kPiece;
    for (uint32_t e = threadIdx.x; e < 256 * 32; e += blockDim.x) {
        const uint32_t i = e >> 5, k = e & 31u;    // row = input index, column = output index inside the piece
        const uint32_t o = o0 + k;
        const float v = (o < n_out) ? W[(size_t)o * in_dim + col0 + i] : 0.0f;
        *reinterpret_cast<__half*>(tile + tc5::chunk_off(256, i, k >> 3) + (k & 7u) * 2) = __float2half_rn(v);
    }
}

__device__ __forceinline__ uint32_t pos_mask2(uint32_t act2) {   // 0xFFFF per half of `act2` that is > 0
    return __hgt2_mask(*reinterpret_cast<const __half2*>(&act2), __float2half2_rn(0.0f));
}

__global__ void __launch_bounds__(kBwdThreads, 1) k_mlp_trunk_bwd(const uint8_t* __restrict__ wblob_t, uint32_t replicas, const uint8_t* __restrict__ save,
                                                const __half* __restrict__ d_x28, uint32_t M, const int32_t* __restrict__ n_valid_p,
                                                uint8_t* __restrict__ grad_ws, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t ring_full[kStages], ring_empty[kStages], acc_full[2], act_ready[2];
    __shared__ uint32_t tmem_base_s;
    uint8_t* const ring = smem + 2 * 65536;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    if (tid == 0) {
        for (uint32_t s = 0; s < kStages; ++s) {
            tc5::mbar_init(&ring_full[s], 1);
            tc5::mbar_init(&ring_empty[s], 2);
        }
        for (uint32_t t = 0; t < 2; ++t) {
            tc5::mbar_init(&acc_full[t], 1);
            tc5::mbar_init(&act_ready[t], 256);
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
    // rows >= *n_valid are padding (the tail kernel did not write their d_x28): they get zero gradients
    const uint32_t m_end = n_valid_p ? min((uint32_t)max(*n_valid_p, 0), M) : M;
    V2Wait wait{status};
    wblob_t += (size_t)(blockIdx.x % replicas) * PVD_MLP_WBLOB_T_BYTES;

    if (warp == kEpiWarpsB) {
        // ------------------------------------------------------------------ TMA producer: ONE transposed weight stream for both tiles
        if (lane == 0) {
            uint32_t pc = 0;
            for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
                for (uint32_t q = 0; q < kBwdPieces; ++q, ++pc) {
                    const uint32_t s = pc % kStages;
                    if (pc >= kStages) wait(&ring_empty[s], ((pc / kStages) - 1u) & 1u);
                    tc5::mbar_expect_tx(&ring_full[s], kPiece);
                    tc5::bulk_g2s(tc5::smem_u32(ring + s * kPiece), wblob_t + (size_t)q * kPiece, kPiece, &ring_full[s]);
                }
        }
    } else if (warp > kEpiWarpsB) {
        // ------------------------------------------------------------------ MMA issuers: one lane per tile, both on the same ring pieces
        // (see k_mlp_field_fwd: one issuer lane cannot keep the tensor pipe busy at two MMAs per piece)
        if (lane == 0) {
            const uint32_t t = warp - (kEpiWarpsB + 1u);
            uint32_t pc = 0, acts = 0;
            const uint32_t idesc = tc5::instr_desc_f16(128, 256, 0, 0);
            const uint32_t a_tile0 = tc5::smem_u32(smem + t * 65536), d_tmem = tmem + 256u * t;
            for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
                for (uint32_t rho = 0; rho < 7; ++rho, ++acts) {
                    const uint32_t n_pieces = rho == 0 ? 1u : 8u;
                    wait(&act_ready[t], acts & 1u);   // G operand tile written, accumulator drained
                    for (uint32_t j = 0; j < n_pieces; ++j, ++pc) {
                        const uint32_t s = pc % kStages;
                        wait(&ring_full[s], (pc / kStages) & 1u);
                        tc5::fence_after_sync();
                        const uint32_t b_tile = tc5::smem_u32(ring + s * kPiece);
                        const uint32_t a_tile = a_tile0 + j * 4u * (kTile * 16u);   // G columns 32 j .. 32 j + 31
#pragma unroll
                        for (uint32_t k0 = 0; k0 < 32; k0 += 16)
                            tc5::mma_f16_ss(d_tmem, tc5::desc_kmajor(a_tile, kTile, k0), tc5::desc_kmajor(b_tile, 256, k0), idesc, !(j == 0 && k0 == 0));
                        tc5::mma_commit(&ring_empty[s]);
                    }
                    tc5::mma_commit(&acc_full[t]);
                }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps: 8 per tile, two threads per row
        const uint32_t t = warp >> 3, half = (warp >> 2) & 1u, r = (warp & 3u) * 32u + lane;
        uint8_t* const A = smem + t * 65536;
        const uint32_t tcol = tc5::tmem_addr(tmem + 256u * t + 128u * half, (warp & 3u) * 32u, 0);
        uint32_t accs = 0;
        for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const uint32_t tile = 2u * pair + t;
            const bool valid = tile < n_tiles;            // the odd tile count's phantom partner touches no global memory
            const uint32_t row = tile * kTile + r;
            const uint8_t* const sv = save + (size_t)tile * kSaveTileBytes;
            uint8_t* const gv = grad_ws + (size_t)tile * kGradTileBytes;
            if (half == 0u) {   // G7 = d(loss)/d(x28): operand columns 0..31 of the tile, and its chunk-tile copy for the wgrad kernel
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    uint4 u = make_uint4(0, 0, 0, 0);
                    if (valid && row < m_end) u = __ldg(reinterpret_cast<const uint4*>(d_x28 + (size_t)row * PVD_FIELD_ENC_STRIDE + 8 * j));
                    *reinterpret_cast<uint4*>(A + tc5::chunk_off(kTile, r, j)) = u;
                    if (valid) *reinterpret_cast<uint4*>(gv + kGradG7 + tc5::chunk_off(kTile, r, j)) = u;
                }
            }
            tc5::fence_async_smem();
            tc5::fence_before_sync();
            tc5::mbar_arrive(&act_ready[t]);
            for (uint32_t rho = 0; rho < 7; ++rho, ++accs) {
                const uint32_t slot = 6u - rho;             // produces G_slot; the mask is act_{slot+1}, saved in slot `slot`
                const uint8_t* const am = sv + kSaveAct + slot * 65536u;
                uint8_t* const go = gv + kGradG + slot * 65536u;
                uint4 mk[2][2];
                auto ldmask = [&](int c, uint4 (&m)[2]) {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        m[i] = valid ? __ldg(reinterpret_cast<const uint4*>(am + tc5::chunk_off(kTile, r, 16 * half + 2 * c + i))) : make_uint4(0, 0, 0, 0);
                };
                ldmask(0, mk[0]);
                wait(&acc_full[t], accs & 1u);
                tc5::fence_after_sync();
                uint32_t buf[2][16];
                tc5::tmem_ld16_issue(tcol, buf[0]);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    tc5::tmem_ld_wait();
                    if (c + 1 < 8) {
                        tc5::tmem_ld16_issue(tcol + 16 * (c + 1), buf[(c + 1) & 1]);
                        ldmask(c + 1, mk[(c + 1) & 1]);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(buf[c & 1][8 * i + e]);
                        uint4 u = tc5::pack8(v);
                        const uint4 m = mk[c & 1][i];
                        u.x &= pos_mask2(m.x); u.y &= pos_mask2(m.y); u.z &= pos_mask2(m.z); u.w &= pos_mask2(m.w);
                        const uint32_t off = tc5::chunk_off(kTile, r, 16 * half + 2 * c + i);
                        *reinterpret_cast<uint4*>(A + off) = u;
                        if (valid) *reinterpret_cast<uint4*>(go + off) = u;
                    }
                }
                tc5::fence_before_sync();
                if (rho < 6u) {   // the last round's G_0 has no consumer here: the next arrival is the next pair's G7
                    tc5::fence_async_smem();
                    tc5::mbar_arrive(&act_ready[t]);
                }
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem, 512);
}

// =============================================================================================== weight gradients
constexpr uint32_t kJobs = 9;
struct WgradJob {
    uint32_t a_grad;      // 1: the M-side operand tile comes from grad_ws, 0: from save_ws
    uint32_t a_off;       // byte offset of the [128 x 256] M-side tile inside its per-tile record
    uint32_t b_grad, b_off;   // the N-side tile
    uint32_t n_pieces;    // N-side pieces per tile
    uint32_t piece_bytes; // 16384 ([128 x 64]) or 8192 ([128 x 32])
    uint32_t n;           // columns per piece
    uint32_t gw_off;      // float offset of the job's output inside gw_ws
    uint32_t gw_pitch, gw_col0;
    uint32_t bias_off;    // float offset of the bias gradient, or 0xFFFFFFFF
    uint32_t bias_from_b; // 1: bias = column sums of the N-side piece (layer 7), 0: of the M-side tile
};
struct WgradPlan {
    WgradJob job[kJobs];
    uint32_t first_cta[kJobs + 1];
};
constexpr uint32_t kGwL = 256u * 320u;                     // floats per layer 0..6 in gw_ws
constexpr uint32_t kGwL7 = 7u * kGwL;
constexpr uint32_t kGwBias = kGwL7 + 256u * 32u;
static_assert(kGwBias + 8u * 256u == PVD_MLP_GW_FLOATS, "gw size");
constexpr uint32_t kWgThreads = 192;
constexpr size_t kWgradSmem = 2 * 65536 + 4 * 16384;       // 196608

__global__ void __launch_bounds__(kWgThreads, 1) k_mlp_wgrad(WgradPlan plan, const uint8_t* __restrict__ save, const uint8_t* __restrict__ grad_ws,
                                                            uint32_t n_tiles, float* __restrict__ gw, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t a_full[2], a_empty[2], b_full[4], b_empty[4], done_bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* const ring = smem + 2 * 65536;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    uint32_t j = 0;
    while (j + 1 < kJobs && blockIdx.x >= plan.first_cta[j + 1]) ++j;
    const WgradJob job = plan.job[j];
    const uint32_t split = blockIdx.x - plan.first_cta[j], n_split = plan.first_cta[j + 1] - plan.first_cta[j];
    const bool bias_a = job.bias_off != 0xFFFFFFFFu && !job.bias_from_b, bias_b = job.bias_off != 0xFFFFFFFFu && job.bias_from_b;
    if (tid == 0) {
        for (uint32_t s = 0; s < 2; ++s) {
            tc5::mbar_init(&a_full[s], 1);
            tc5::mbar_init(&a_empty[s], bias_a ? 129u : 1u);
        }
        for (uint32_t s = 0; s < 4; ++s) {
            tc5::mbar_init(&b_full[s], 1);
            tc5::mbar_init(&b_empty[s], bias_b ? 33u : 1u);
        }
        tc5::mbar_init(&done_bar, 1);
        tc5::mbar_fence_init();
    }
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    V2Wait wait{status};
    const uint32_t my_tiles = (n_tiles > split) ? (n_tiles - split + n_split - 1) / n_split : 0u;

    if (warp == 4) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t pc = 0;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t tile = split + it * n_split;
                const uint8_t* rec_a = job.a_grad ? grad_ws + (size_t)tile * kGradTileBytes : save + (size_t)tile * kSaveTileBytes;
                const uint8_t* rec_b = job.b_grad ? grad_ws + (size_t)tile * kGradTileBytes : save + (size_t)tile * kSaveTileBytes;
                const uint32_t sa = it & 1u;
                if (it >= 2u) wait(&a_empty[sa], ((it >> 1) - 1u) & 1u);
                tc5::mbar_expect_tx(&a_full[sa], 65536u);
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q)
                    tc5::bulk_g2s(tc5::smem_u32(smem + sa * 65536u + q * 16384u), rec_a + job.a_off + q * 16384u, 16384u, &a_full[sa]);
                for (uint32_t p = 0; p < job.n_pieces; ++p, ++pc) {
                    const uint32_t s = pc & 3u;
                    if (pc >= 4u) wait(&b_empty[s], ((pc >> 2) - 1u) & 1u);
                    tc5::mbar_expect_tx(&b_full[s], job.piece_bytes);
                    tc5::bulk_g2s(tc5::smem_u32(ring + s * 16384u), rec_b + job.b_off + p * job.piece_bytes, job.piece_bytes, &b_full[s]);
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer: D_h[128 x n] += Atile[:, 128 h ..]^T Bpiece
        if (lane == 0) {
            uint32_t pc = 0;
            const uint32_t idesc = tc5::instr_desc_f16(128, job.n, 1, 1);
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t sa = it & 1u;
                wait(&a_full[sa], (it >> 1) & 1u);
                const uint32_t a_tile = tc5::smem_u32(smem + sa * 65536u);
                for (uint32_t p = 0; p < job.n_pieces; ++p, ++pc) {
                    const uint32_t s = pc & 3u;
                    wait(&b_full[s], (pc >> 2) & 1u);
                    tc5::fence_after_sync();
                    const uint32_t b_tile = tc5::smem_u32(ring + s * 16384u);
                    for (uint32_t h = 0; h < 2; ++h)
#pragma unroll
                        for (uint32_t s0 = 0; s0 < kTile; s0 += 16)
                            tc5::mma_f16_ss(tmem + 256u * h + job.n * p, tc5::desc_mnmajor(a_tile, kTile, s0, 128u * h),
                                            tc5::desc_mnmajor(b_tile, kTile, s0, 0), idesc, !(it == 0 && s0 == 0));
                    tc5::mma_commit(&b_empty[s]);
                }
                tc5::mma_commit(&a_empty[sa]);
            }
            tc5::mma_commit(&done_bar);
        }
    } else {
        // ------------------------------------------------------------------ warps 0-3: bias gradient under the MMAs, then the flush
        // thread u sums columns 2u, 2u+1 of the [128 x 256] tile: one 32-bit word per row; rows visited in an order skewed by
        // lane / 4 so that the 32 lanes of a warp hit 32 different banks
        float s0 = 0.0f, s1 = 0.0f;
        if (bias_a) {
            const uint32_t cj = tid >> 2, cw = (tid & 3u) * 4u, skew = lane >> 2;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t sa = it & 1u;
                wait(&a_full[sa], (it >> 1) & 1u);
                const uint8_t* col = smem + sa * 65536u + cj * (kTile * 16u) + cw;
#pragma unroll 8
                for (uint32_t rr = 0; rr < kTile; ++rr) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(col + ((rr + skew) & 127u) * 16u);
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
                    s0 += f.x;
                    s1 += f.y;
                }
                tc5::mbar_arrive(&a_empty[sa]);
            }
            if (my_tiles) {
                atomicAdd(gw + job.bias_off + 2u * tid, s0);
                atomicAdd(gw + job.bias_off + 2u * tid + 1u, s1);
            }
        } else if (bias_b && warp == 0) {
            // layer 7: G7 piece [128 x 32]; lane u sums column u (one half per row; two lanes share a word: broadcast, no conflict
            // beyond the 2-way one of neighbouring chunks)
            const uint32_t cj = lane >> 3, ch = (lane & 7u) * 2u;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t s = it & 3u;
                wait(&b_full[s], (it >> 2) & 1u);
                const uint8_t* col = ring + s * 16384u + cj * (kTile * 16u) + ch;
#pragma unroll 8
                for (uint32_t rr = 0; rr < kTile; ++rr) s0 += __half2float(*reinterpret_cast<const __half*>(col + rr * 16u));
                tc5::mbar_arrive(&b_empty[s]);
            }
            if (my_tiles) atomicAdd(gw + job.bias_off + lane, s0);
        }
        // ---- flush: TMEM lane = accumulator row (M = 128)
        if (my_tiles) {
            wait(&done_bar, 0u);
            tc5::fence_after_sync();
            const uint32_t ncols = job.n * job.n_pieces;
            for (uint32_t h = 0; h < 2; ++h) {
                float* dst = gw + job.gw_off + (size_t)(128u * h + warp * 32u + lane) * job.gw_pitch + job.gw_col0;
                for (uint32_t c = 0; c < ncols; c += 16) {
                    float v[16];
                    tc5::tmem_ld16(tc5::tmem_addr(tmem + 256u * h, warp * 32u, c), v);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                     "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                                     : "memory");
                }
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem, 512);
}

// gw_ws -> parameter-shaped gradient buffers (accumulating)
__global__ void k_mlp_unpack_wgrads(const float* __restrict__ gw, float* const* __restrict__ gwt, float* const* __restrict__ gbs) {
    const uint32_t l = blockIdx.y;
    const uint32_t in_dim = (l == 0) ? 63u : (l == 4 ? 319u : 256u), out = (l == 7) ? 28u : 256u;
    float* W = gwt[l];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < out * in_dim; e += gridDim.x * blockDim.x) {
        const uint32_t o = e / in_dim, i = e - o * in_dim;
        float g;
        if (l == 7) g = gw[kGwL7 + i * 32u + o];
        else if (l == 4) g = gw[4u * kGwL + o * 320u + (i < 63u ? i : i + 1u)];     // hidden part starts at column 64
        else g = gw[l * kGwL + o * 320u + i];
        W[e] += g;
    }
    if (blockIdx.x == 0) {
        float* B = gbs[l];
        for (uint32_t o = threadIdx.x; o < out; o += blockDim.x) B[o] += gw[kGwBias + 256u * l + o];
    }
}

static WgradPlan make_plan(uint32_t grid) {
    WgradPlan p;
    uint32_t weight[kJobs];
    auto big = [&](uint32_t idx, uint32_t l) {     // dW_l = G_l^T act_l, l in {1,2,3,4 (hidden part),5,6}
        p.job[idx] = WgradJob{1u, kGradG + l * 65536u, 0u, kSaveAct + (l - 1u) * 65536u, 4u, 16384u, 64u, l * kGwL, 320u, (l == 4u) ? 64u : 0u,
                              kGwBias + 256u * l, 0u};
        weight[idx] = 128;
    };
    big(0, 1); big(1, 2); big(2, 3); big(3, 4); big(4, 5); big(5, 6);
    // layer 0 and the in_pts part of layer 4: N side = the PE tile
    p.job[6] = WgradJob{1u, kGradG + 0u * 65536u, 0u, kSavePe, 1u, 16384u, 64u, 0u * kGwL, 320u, 0u, kGwBias + 0u, 0u};
    p.job[7] = WgradJob{1u, kGradG + 4u * 65536u, 0u, kSavePe, 1u, 16384u, 64u, 4u * kGwL, 320u, 0u, 0xFFFFFFFFu, 0u};
    weight[6] = weight[7] = 80;
    // layer 7, transposed: dW7^T [256 in][32 out] = act_7^T G7
    p.job[8] = WgradJob{0u, kSaveAct + 6u * 65536u, 1u, kGradG7, 1u, 8192u, 32u, kGwL7, 32u, 0u, kGwBias + 256u * 7u, 1u};
    weight[8] = 72;
    // CTAs per job proportional to the bytes a tile costs it; every job gets at least one
    uint32_t total = 0, n[kJobs], used = 0;
    for (uint32_t j = 0; j < kJobs; ++j) total += weight[j];
    if (grid < kJobs) grid = kJobs;
    for (uint32_t j = 0; j < kJobs; ++j) {
        n[j] = (uint32_t)((uint64_t)grid * weight[j] / total);
        if (n[j] == 0) n[j] = 1;
        used += n[j];
    }
    for (uint32_t j = 0; used < grid; j = (j + 1) % kJobs) { ++n[j]; ++used; }     // leftovers to the heavy jobs first
    for (uint32_t j = kJobs; used > grid;) { j = (j == 0) ? kJobs - 1 : j - 1; if (n[j] > 1) { --n[j]; --used; } }
    p.first_cta[0] = 0;
    for (uint32_t j = 0; j < kJobs; ++j) p.first_cta[j + 1] = p.first_cta[j] + n[j];
    return p;
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_mlp_pack_weights_t(const float* const* weights8, void* wblob_t, void* stream) {
    PVD_REQUIRE(weights8 && wblob_t);
    k_mlp_pack_t<<<kBwdPieces, 256, 0, (cudaStream_t)stream>>>(weights8, (uint8_t*)wblob_t);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_mlp_trunk_backward(const void* wblob_t, uint32_t replicas, const void* save_ws, const void* d_x28, uint32_t M, const int32_t* n_valid,
                           void* grad_ws, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(wblob_t && save_ws && d_x28 && grad_ws && status);
    const uint32_t tiles = (M + kTile - 1) / kTile, pairs = (tiles + 1) / 2;
    const uint32_t grid = min(pairs, (uint32_t)sm_count());
    cudaError_t e = cudaFuncSetAttribute(k_mlp_trunk_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrunkSmem);
    if (e != cudaSuccess) return (int)e;
    k_mlp_trunk_bwd<<<grid, kBwdThreads, kTrunkSmem, (cudaStream_t)stream>>>((const uint8_t*)wblob_t, replicas ? replicas : 1u, (const uint8_t*)save_ws, (const __half*)d_x28, M,
                                                                            n_valid, (uint8_t*)grad_ws, status);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_mlp_weight_grads(const void* save_ws, const void* grad_ws, uint32_t M, float* gw_ws, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(save_ws && grad_ws && gw_ws && status);
    const uint32_t tiles = (M + kTile - 1) / kTile;
    const uint32_t grid = max((uint32_t)sm_count(), kJobs);
    const WgradPlan plan = make_plan(grid);
    cudaError_t e = cudaFuncSetAttribute(k_mlp_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmem);
    if (e != cudaSuccess) return (int)e;
    k_mlp_wgrad<<<grid, kWgThreads, kWgradSmem, (cudaStream_t)stream>>>(plan, (const uint8_t*)save_ws, (const uint8_t*)grad_ws, tiles, gw_ws, status);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_mlp_unpack_wgrads(const float* gw_ws, float* const* grad_weights8, float* const* grad_biases8, void* stream) {
    PVD_REQUIRE(gw_ws && grad_weights8 && grad_biases8);
    k_mlp_unpack_wgrads<<<dim3(32, 8), 256, 0, (cudaStream_t)stream>>>(gw_ws, grad_weights8, grad_biases8);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
