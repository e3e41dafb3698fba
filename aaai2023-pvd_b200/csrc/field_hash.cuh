// field_hash.cuh -- pieces shared by the hash-field kernels (field_hash.cu: fp16 tensor-core path; field_hash_f32.cu: fp32 path):
// the weight-gradient workspace layout, the per-level table descriptor, corner indices / weights of one sample, the position map.
#pragma once
#include "common.cuh"
#include "gridenc.cuh"
#include "shenc.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

// layout of the weight-gradient workspace (floats), the kernel-native accumulator shapes
constexpr uint32_t kGW1 = 0;                  // [64][32]
constexpr uint32_t kGW2 = kGW1 + 64 * 32;     // [64][16]  (transposed: [in][out])
constexpr uint32_t kGW3 = kGW2 + 64 * 16;     // [64][32]
constexpr uint32_t kGW4 = kGW3 + 64 * 32;     // [64][64]
constexpr uint32_t kGW5 = kGW4 + 64 * 64;     // [64][16]  (transposed: [in][out])
constexpr uint32_t kGWTotal = kGW5 + 64 * 16; // 10240
static_assert(kGWTotal == PVD_FIELD_GW_FLOATS, "workspace size");

struct __align__(16) LevelInfo {
    float scale;
    uint32_t res1;    // resolution + 1 (dense stride)
    uint32_t offset;  // first entry
    uint32_t size;    // entries
    uint32_t mode;    // 0 dense (index < size, no modulo), 1 hashed with power-of-two size, 2 generic
    uint32_t mask;
    uint32_t pad0, pad1;
};

// explicit ld.shared of one table entry (a generic-pointer access made ptxas insert an S2R + window computation per level)
__device__ __forceinline__ LevelInfo ld_level(uint32_t lv_saddr, uint32_t l) {
    LevelInfo v;
    uint32_t a, b, c, d, e, f;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(lv_saddr + l * 32u));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(e), "=r"(f) : "r"(lv_saddr + l * 32u + 16u));
    v.scale = __uint_as_float(a); v.res1 = b; v.offset = c; v.size = d; v.mode = e; v.mask = f; v.pad0 = 0; v.pad1 = 0;
    return v;
}

__device__ __forceinline__ void level_info_init(LevelInfo* lv, const int32_t* offsets, uint32_t L, float S, uint32_t H) {
    const uint32_t l = threadIdx.x;
    if (l < L) {
        const GridLevel g = grid_level(offsets, l, S, H);
        LevelInfo v;
        v.scale = g.scale;
        v.res1 = g.resolution + 1;
        v.offset = g.offset;
        v.size = g.size;
        v.mask = g.size - 1;
        const uint64_t dense = (uint64_t)v.res1 * v.res1 * v.res1;
        // gridencoder.cu:54-72: the running stride exceeds the level size exactly when the dense grid does not fit
        const bool fits = ((uint64_t)v.res1 <= g.size) && ((uint64_t)v.res1 * v.res1 <= g.size) && (dense <= g.size);
        v.mode = fits ? 0u : (((g.size & (g.size - 1)) == 0) ? 1u : 2u);
        v.pad0 = v.pad1 = 0;
        lv[l] = v;
    }
}

// one sample's corner set at one level: 8 entry indices (in entries) and weights, reference order (bit d of idx = upper
// vertex along d, weight = ((wx)*wy)*wz)
struct Corners {
    uint32_t idx[8];
    float w[8];
};

__device__ __forceinline__ void level_corners(const LevelInfo& lv, const float (&x01)[3], Corners& c) {
    uint32_t cell[3];
    float frac[3];
    grid_locate<3>(x01, lv.scale, false, cell, frac);
    const float wx[2] = {1.0f - frac[0], frac[0]}, wy[2] = {1.0f - frac[1], frac[1]}, wz[2] = {1.0f - frac[2], frac[2]};
    if (lv.mode == 0) {
        const uint32_t s1 = lv.res1, s2 = lv.res1 * lv.res1;
        const uint32_t ix[2] = {cell[0], cell[0] + 1}, iy[2] = {cell[1] * s1, cell[1] * s1 + s1},
                       iz[2] = {cell[2] * s2, cell[2] * s2 + s2};
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) c.idx[i] = ix[i & 1] + iy[(i >> 1) & 1] + iz[(i >> 2) & 1];
    } else if (lv.mode == 1) {
        const uint32_t hx[2] = {cell[0], cell[0] + 1}, hy[2] = {cell[1] * 2654435761u, (cell[1] + 1) * 2654435761u},
                       hz[2] = {cell[2] * 805459861u, (cell[2] + 1) * 805459861u};
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) c.idx[i] = (hx[i & 1] ^ hy[(i >> 1) & 1] ^ hz[(i >> 2) & 1]) & lv.mask;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            const uint32_t v[3] = {cell[0] + (i & 1), cell[1] + ((i >> 1) & 1), cell[2] + ((i >> 2) & 1)};
            c.idx[i] = grid_index<3>(0u, false, lv.size, lv.res1 - 1, v);
        }
    }
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) c.w[i] = __fmul_rn(__fmul_rn(wx[i & 1], wy[(i >> 1) & 1]), wz[(i >> 2) & 1]);
}

// map world position to the encoder's [0,1] range: (x + bound) / (2*bound)   (gridencoder/grid.py:211)
__device__ __forceinline__ void to_unit(const float* __restrict__ p, float bound, float (&x01)[3], bool& oob) {
    const float two_b = 2.0f * bound;
#pragma unroll
    for (int d = 0; d < 3; ++d) x01[d] = __fdiv_rn(__fadd_rn(p[d], bound), two_b);
    oob = grid_oob<3>(x01);
}

// interpolate 8 consecutive features (4 levels) of one sample.  All 32 gathers of the four levels are issued before any
// of them is consumed (branch-free: a level beyond L re-reads level 0 and is multiplied by zero), because at 9-12 resident
// warps per SM the gather is latency-bound and memory-level parallelism per thread is what hides it.
template <typename T>
__device__ __forceinline__ void encode4(const T* __restrict__ table, uint32_t lv_saddr, uint32_t l0, uint32_t L,
                                        const float (&x01)[3], bool oob, float (&f)[8]) {
    float2 fv[4][8];
    float w[4][8];
#pragma unroll
    for (uint32_t q = 0; q < 4; ++q) {
        const bool on = (l0 + q < L) && !oob;
        const LevelInfo v = ld_level(lv_saddr, on ? l0 + q : 0u);
        Corners c;
        level_corners(v, x01, c);
        const T* tab = table + (size_t)v.offset * 2;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            fv[q][i] = tab_load2(tab, (size_t)c.idx[i] * 2);
            w[q][i] = on ? c.w[i] : 0.0f;
        }
    }
#pragma unroll
    for (uint32_t q = 0; q < 4; ++q) {
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            a0 = __fmaf_rn(w[q][i], fv[q][i].x, a0);
            a1 = __fmaf_rn(w[q][i], fv[q][i].y, a1);
        }
        f[2 * q] = a0;
        f[2 * q + 1] = a1;
    }
}

}  // namespace pvd
