// gridenc.cuh -- multiresolution hash / tiled grid addressing shared by the stand-alone encoder
// kernels (gridenc.cu) and the fused field kernels (field_*.cu).
//
// Restates gridencoder/src/gridencoder.cu:35-72 (hash + index), :125-138 (level scale / cell) and
// :146-167 (D-linear blend).  The float expressions are pinned with explicit FMA intrinsics in the
// same association order the reference compiles to (checked against its SASS), so that an fp32 table
// gives bit-identical features.
#pragma once
#include "common.cuh"

namespace pvd {

// per-level constants (gridencoder.cu:125-127)
struct GridLevel {
    float scale;         // exp2f(level*S)*H - 1
    uint32_t resolution; // ceil(scale) + 1
    uint32_t offset;     // first entry of the level in the table
    uint32_t size;       // entries in the level (hashmap_size)
};

__device__ __forceinline__ GridLevel grid_level(const int32_t* __restrict__ offsets, uint32_t level, float S, uint32_t H) {
    GridLevel g;
    g.offset = (uint32_t)offsets[level];
    g.size = (uint32_t)offsets[level + 1] - g.offset;
    g.scale = __fmaf_rn(exp2f(__fmul_rn((float)level, S)), (float)H, -1.0f);
    g.resolution = (uint32_t)ceilf(g.scale) + 1u;
    return g;
}

template <uint32_t D>
__device__ __forceinline__ uint32_t grid_hash(const uint32_t (&p)[D]) {  // gridencoder.cu:35-51
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

// entry index (in entries, not scalars) of a grid vertex (gridencoder.cu:54-72)
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_index(uint32_t gridtype, bool align_corners, uint32_t size, uint32_t resolution,
                                               const uint32_t (&p)[D]) {
    uint32_t stride = 1, index = 0;
    const uint32_t step = align_corners ? resolution : resolution + 1;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= size) {
            index += p[d] * stride;
            stride *= step;
        }
    }
    if (gridtype == 0 && stride > size) index = grid_hash<D>(p);
    return index % size;
}

// cell + fractional position of an input in [0,1]^D at one level (gridencoder.cu:133-138)
template <uint32_t D>
__device__ __forceinline__ void grid_locate(const float (&x)[D], float scale, bool align_corners, uint32_t (&cell)[D],
                                            float (&frac)[D]) {
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const float p = __fmaf_rn(x[d], scale, align_corners ? 0.0f : 0.5f);
        const float f = floorf(p);
        cell[d] = (uint32_t)f;
        frac[d] = p - f;
    }
}

// weight of corner `idx` (bit d set = upper vertex along d); product in dimension order (gridencoder.cu:147-159)
template <uint32_t D>
__device__ __forceinline__ float grid_corner(const uint32_t (&cell)[D], const float (&frac)[D], uint32_t idx,
                                             uint32_t (&vert)[D]) {
    float w = 1.0f;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if ((idx >> d) & 1u) {
            w = __fmul_rn(w, frac[d]);
            vert[d] = cell[d] + 1;
        } else {
            w = __fmul_rn(w, 1.0f - frac[d]);
            vert[d] = cell[d];
        }
    }
    return w;
}

template <uint32_t D>
__device__ __forceinline__ bool grid_oob(const float (&x)[D]) {  // gridencoder.cu:99-105
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) oob |= (x[d] < 0.0f) | (x[d] > 1.0f);
    return oob;
}

// table element access
__device__ __forceinline__ float tab_load(const float* t, size_t i) { return __ldg(t + i); }
__device__ __forceinline__ float tab_load(const __half* t, size_t i) { return __half2float(__ldg(t + i)); }
__device__ __forceinline__ float2 tab_load2(const float* t, size_t i) { return __ldg(reinterpret_cast<const float2*>(t + i)); }
__device__ __forceinline__ float2 tab_load2(const __half* t, size_t i) {
    return __half22float2(__ldg(reinterpret_cast<const __half2*>(t + i)));
}
__device__ __forceinline__ void tab_store(float* t, size_t i, float v) { t[i] = v; }
__device__ __forceinline__ void tab_store(__half* t, size_t i, float v) { t[i] = __float2half_rn(v); }

// gradient accumulation into a table: vectorised reductions without return value
__device__ __forceinline__ void tab_red(float* t, size_t i, float v) { atomicAdd(t + i, v); }
__device__ __forceinline__ void tab_red(__half* t, size_t i, float v) { atomicAdd(t + i, __float2half_rn(v)); }
__device__ __forceinline__ void tab_red2(float* t, size_t i, float a, float b) {
    // red.global.add.v2.f32 (sm_90+): one 8-byte reduction instead of two 4-byte ones
    // volatile keeps the reduction; no "memory" clobber, so the compiler may overlap the next level's address math with it
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(t + i), "f"(a), "f"(b));
}
// the x-neighbour pair of one (y, z) corner: entries e0, e1 with values (a0, b0), (a1, b1).  When the two entries are the two
// halves of one 16-byte slot (e0 ^ e1 == 1: always on hashed levels when x is even, on dense ones when e0 is even) a single
// red.global.add.v4.f32 replaces two v2 reductions -- the scatter is bound by reduction lane-operations per SM.
__device__ __forceinline__ void tab_red_pair(float* t, uint32_t e0, uint32_t e1, float a0, float b0, float a1, float b1) {
    if ((e0 ^ e1) == 1u) {
        const bool lo = (e0 & 1u) == 0u;  // e0 is the lower entry of the slot
        float* p = t + (size_t)(e0 & ~1u) * 2;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(lo ? a0 : a1), "f"(lo ? b0 : b1), "f"(lo ? a1 : a0),
                     "f"(lo ? b1 : b0));
    } else {
        tab_red2(t, (size_t)e0 * 2, a0, b0);
        tab_red2(t, (size_t)e1 * 2, a1, b1);
    }
}
__device__ __forceinline__ void tab_red2(__half* t, size_t i, float a, float b) {
    atomicAdd(reinterpret_cast<__half2*>(t + i), __floats2half2_rn(a, b));
}

}  // namespace pvd
