// field_tail.cuh -- the sigma_net / color_net tail shared by the "hash" and "mlp" fields (distill_mutual/network.py:413-437):
// packed-weight blob layout, TMEM columns of the forward, and the five-layer forward for one 128-sample tile.
#pragma once
#include "field_common.cuh"
#include "shenc.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

// byte offsets of the weight operand tiles inside the packed blob / shared memory
constexpr uint32_t kWB1 = 0;      // sigma_net.0 : 64 rows (out) x 32 cols (in, 2L zero padded)
constexpr uint32_t kWB2 = 4096;   // sigma_net.1 : 16 x 64
constexpr uint32_t kWB3 = 6144;   // color_net.0 : 64 x 32 (in = 16 SH + 15 geo + 1 pad)
constexpr uint32_t kWB4 = 10240;  // color_net.1 : 64 x 64
constexpr uint32_t kWB5 = 18432;  // color_net.2 : 16 (3 + pad) x 64
static_assert(kWB5 + 2048 == PVD_FIELD_WBLOB_BYTES, "blob size");

// TMEM columns
constexpr uint32_t kD = 0;      // [0,64)   layer output / data-gradient accumulator
constexpr uint32_t kD16 = 64;   // [64,80)  sigma_net.1 output
constexpr uint32_t kD5 = 80;    // [80,96)  color_net.2 output
constexpr uint32_t kAW5 = 96;   // [96,112)   dW5^T  [64 in ][16 out]
constexpr uint32_t kAW4 = 112;  // [112,176)  dW4    [64 out][64 in ]
constexpr uint32_t kAW3 = 176;  // [176,208)  dW3    [64 out][32 in ]
constexpr uint32_t kAW2 = 208;  // [208,224)  dW2^T  [64 in ][16 out]
constexpr uint32_t kAW1 = 224;  // [224,256)  dW1    [64 out][32 in ]


struct FieldArgs {
    const void* table;
    const int32_t* offsets;
    const uint8_t* wblob;
    uint32_t L, H;
    float S, bound, clip_min, clip_max, density_scale;
};

struct FwdRegs {  // what the backward needs from the recomputed forward of this thread's sample
    float o0_raw, o0c, rgb[3];
};

// Layers 1..5 for one tile.  X tile must already hold the encoding.  Tiles: X, H1, CIN, H3, H4 (H3/H4 may alias X-independent
// buffers in the forward-only kernel).  Returns sigma (scaled) and rgb for this thread's row.  `row` is the thread's row inside the
// tile (0..127); the cooperating threads are the team of `p` (the whole CTA, or one 128-thread warpgroup of a larger CTA).
__device__ __forceinline__ void mlp_forward(Pipe& p, const FieldArgs& a, uint8_t* smw, uint8_t* X, uint8_t* H1, uint8_t* CIN,
                                            uint8_t* H3, uint8_t* H4, const float* __restrict__ dir, uint32_t row, float& sigma,
                                            float (&o16)[16], FwdRegs& r) {
    const uint32_t tid = threadIdx.x;
    const uint32_t lane_base = ((tid >> 5) & 3u) * 32;  // TMEM lanes a warp may touch: 32 * (warp % 4)
    const bool leader = team_leader(p);
    const uint32_t trow = tc5::tmem_addr(p.tmem, lane_base, 0);
    const uint32_t sw = tc5::smem_u32(smw);
    // ---- sigma_net.0 : [128 x 32] x [64 x 32]^T
    operands_ready(p);
    if (leader) {
        tc5::fence_after_sync();
        weights_ready(p);
        issue_fwd(p.tmem + kD, tc5::smem_u32(X), 32, sw + kWB1, 64, 64);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    if (leader) PVD_T(p.trec, 6);
    relu_to_tile<4>(trow + kD, H1, row);
    // ---- sigma_net.1 : [128 x 64] x [16 x 64]^T
    operands_ready(p);
    if (leader) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kD16, tc5::smem_u32(H1), 64, sw + kWB2, 16, 16);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    if (leader) PVD_T(p.trec, 7);
    tc5::tmem_ld16(trow + kD16, o16);
#pragma unroll
    for (int i = 0; i < 16; ++i) o16[i] = __half2float(__float2half_rn(o16[i]));  // the reference's fp16 activations
    r.o0_raw = o16[0];
    r.o0c = clampf(o16[0], a.clip_min, a.clip_max);  // network.py:418-420
    o16[0] = r.o0c;
    sigma = a.density_scale * __expf(r.o0c);          // trunc_exp forward (tools/activation.py:9-12), renderer.py:440
    {
        float sh[16];
        sh_basis4(dir[0], dir[1], dir[2], sh);
        float geo[16];
#pragma unroll
        for (int i = 0; i < 15; ++i) geo[i] = o16[i + 1];
        geo[15] = 0.0f;
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 0)) = tc5::pack8(sh);
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 1)) = tc5::pack8(sh + 8);
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 2)) = tc5::pack8(geo);
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 3)) = tc5::pack8(geo + 8);
    }
    // ---- color_net.0 : [128 x 32] x [64 x 32]^T
    operands_ready(p);
    if (leader) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kD, tc5::smem_u32(CIN), 32, sw + kWB3, 64, 64);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    if (leader) PVD_T(p.trec, 8);
    relu_to_tile<4>(trow + kD, H3, row);
    // ---- color_net.1 : [128 x 64] x [64 x 64]^T
    operands_ready(p);
    if (leader) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kD, tc5::smem_u32(H3), 64, sw + kWB4, 64, 64);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    if (leader) PVD_T(p.trec, 9);
    relu_to_tile<4>(trow + kD, H4, row);
    // ---- color_net.2 : [128 x 64] x [16 x 64]^T , sigmoid
    operands_ready(p);
    if (leader) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kD5, tc5::smem_u32(H4), 64, sw + kWB5, 16, 16);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    if (leader) PVD_T(p.trec, 10);
    float c16[16];
    tc5::tmem_ld16(trow + kD5, c16);
#pragma unroll
    for (int i = 0; i < 3; ++i) r.rgb[i] = 1.0f / (1.0f + __expf(-c16[i]));
}

}  // namespace pvd
