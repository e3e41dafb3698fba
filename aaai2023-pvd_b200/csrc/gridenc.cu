// gridenc.cu -- stand-alone multiresolution hash-grid encoder (forward, table-gradient scatter, input
// gradient) behind pvd_grid_encode_forward / pvd_grid_encode_backward.
//
// Replaces gridencoder/src/gridencoder.cu.  Same [L,B,C] level-major activation layout as the reference
// (so one level's table stays cache-hot while a grid row of CTAs sweeps the batch); the differences are
// vector (8-byte) feature loads and `red.global.add.v2.f32` / f16x2 reductions for the scatter, fp32
// accumulation of the interpolation for half tables, and a working half path for C == 1.
#include "gridenc.cuh"

namespace pvd {

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_fwd(const float* __restrict__ inputs, const T* __restrict__ table,
                                                 const int32_t* __restrict__ offsets, T* __restrict__ outputs, uint32_t B,
                                                 uint32_t L, float S, uint32_t H, bool calc_grad_inputs, T* __restrict__ dy_dx,
                                                 uint32_t gridtype, bool align_corners, bool sample_major) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    T* out = outputs + (sample_major ? ((size_t)b * L + level) : ((size_t)level * B + b)) * C;
    T* jac = calc_grad_inputs ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;  // [B, L, D, C]

    if (grid_oob<D>(x)) {  // gridencoder.cu:107-123
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) tab_store(out, c, 0.0f);
        if (calc_grad_inputs) {
#pragma unroll
            for (uint32_t i = 0; i < D * C; ++i) tab_store(jac, i, 0.0f);
        }
        return;
    }

    const GridLevel g = grid_level(offsets, level, S, H);
    const T* tab = table + (size_t)g.offset * C;
    uint32_t cell[D];
    float frac[D];
    grid_locate<D>(x, g.scale, align_corners, cell, frac);

    float acc[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) acc[c] = 0.0f;
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
        uint32_t v[D];
        const float w = grid_corner<D>(cell, frac, idx, v);
        const size_t e = (size_t)grid_index<D>(gridtype, align_corners, g.size, g.resolution, v) * C;
        if constexpr (C % 2 == 0) {
#pragma unroll
            for (uint32_t c = 0; c < C; c += 2) {
                const float2 f = tab_load2(tab, e + c);
                acc[c] = __fmaf_rn(w, f.x, acc[c]);
                acc[c + 1] = __fmaf_rn(w, f.y, acc[c + 1]);
            }
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) acc[c] = __fmaf_rn(w, tab_load(tab, e + c), acc[c]);
        }
    }
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) tab_store(out, c, acc[c]);

    if (calc_grad_inputs) {  // d(out)/d(x_gd): finite difference of the two faces, gridencoder.cu:180-222
#pragma unroll
        for (uint32_t gd = 0; gd < D; ++gd) {
            float ga[C];
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) ga[c] = 0.0f;
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
                float w = g.scale;
                uint32_t v[D];
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; ++nd) {
                    const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                    if ((idx >> nd) & 1u) {
                        w *= frac[d];
                        v[d] = cell[d] + 1;
                    } else {
                        w *= 1.0f - frac[d];
                        v[d] = cell[d];
                    }
                }
                v[gd] = cell[gd];
                const size_t el = (size_t)grid_index<D>(gridtype, align_corners, g.size, g.resolution, v) * C;
                v[gd] = cell[gd] + 1;
                const size_t er = (size_t)grid_index<D>(gridtype, align_corners, g.size, g.resolution, v) * C;
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) ga[c] += w * (tab_load(tab, er + c) - tab_load(tab, el + c));
            }
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) tab_store(jac, gd * C + c, ga[c]);
        }
    }
}

// scatter: one thread per (sample, level, feature pair)
template <typename T, uint32_t D, uint32_t C, uint32_t NC>
__global__ void __launch_bounds__(256) k_grid_bwd(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                 const int32_t* __restrict__ offsets, T* __restrict__ grad_table, uint32_t B,
                                                 uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners,
                                                 bool sample_major) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = tid * NC / C;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const uint32_t ch = tid * NC - b * C;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    if (grid_oob<D>(x)) return;

    const GridLevel g = grid_level(offsets, level, S, H);
    T* gt = grad_table + (size_t)g.offset * C;
    uint32_t cell[D];
    float frac[D];
    grid_locate<D>(x, g.scale, align_corners, cell, frac);

    float gv[NC];
    const T* gp = grad + (sample_major ? ((size_t)b * L + level) : ((size_t)level * B + b)) * C + ch;
#pragma unroll
    for (uint32_t c = 0; c < NC; ++c) gv[c] = tab_load(gp, c);

#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
        uint32_t v[D];
        const float w = grid_corner<D>(cell, frac, idx, v);
        const size_t e = (size_t)grid_index<D>(gridtype, align_corners, g.size, g.resolution, v) * C + ch;
        if constexpr (NC == 2) {
            tab_red2(gt, e, w * gv[0], w * gv[1]);
        } else {
            tab_red(gt, e, w * gv[0]);
        }
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (gridencoder.cu:317-343)
template <typename T, uint32_t D, uint32_t C>
__global__ void k_grid_input_bwd(const T* __restrict__ grad, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs,
                                 uint32_t B, uint32_t L, bool sample_major) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T* j = dy_dx + (size_t)b * L * D * C;
    float r = 0.0f;
    for (uint32_t l = 0; l < L; ++l) {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c)
            r += tab_load(grad, (sample_major ? ((size_t)b * L + l) : ((size_t)l * B + b)) * C + c) * tab_load(j, ((size_t)l * D + d) * C + c);
    }
    tab_store(grad_inputs, t, r);
}

// per-level (scale, resolution) exactly as every kernel of this library derives them
__global__ void k_grid_level_table(const int32_t* __restrict__ offsets, uint32_t L, float S, uint32_t H, float* __restrict__ scales,
                                   int32_t* __restrict__ resolutions) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const GridLevel g = grid_level(offsets, l, S, H);
    scales[l] = g.scale;
    resolutions[l] = (int32_t)g.resolution;
}

template <typename T, uint32_t D, uint32_t C>
int launch_fwd(const float* inputs, const void* emb, const int32_t* offsets, void* out, uint32_t B, uint32_t L, float S,
               uint32_t H, bool cgi, void* dy_dx, uint32_t gridtype, bool ac, bool sm, cudaStream_t st) {
    const dim3 grid(ceil_div(B, 256), L, 1);
    k_grid_fwd<T, D, C><<<grid, 256, 0, st>>>(inputs, (const T*)emb, offsets, (T*)out, B, L, S, H, cgi, (T*)dy_dx, gridtype, ac, sm);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

template <typename T, uint32_t D>
int dispatch_fwd(uint32_t C, const float* inputs, const void* emb, const int32_t* offsets, void* out, uint32_t B, uint32_t L,
                 float S, uint32_t H, bool cgi, void* dy_dx, uint32_t gridtype, bool ac, bool sm, cudaStream_t st) {
    switch (C) {
        case 1: return launch_fwd<T, D, 1>(inputs, emb, offsets, out, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
        case 2: return launch_fwd<T, D, 2>(inputs, emb, offsets, out, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
        case 4: return launch_fwd<T, D, 4>(inputs, emb, offsets, out, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
        case 8: return launch_fwd<T, D, 8>(inputs, emb, offsets, out, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
        default: return PVD_EUNSUPPORTED;  // "GridEncoding: C must be 1, 2, 4, or 8." gridencoder.cu:355
    }
}

template <typename T, uint32_t D, uint32_t C>
int launch_bwd(const void* grad, const float* inputs, const int32_t* offsets, void* gemb, uint32_t B, uint32_t L, float S,
               uint32_t H, bool cgi, const void* dy_dx, void* grad_inputs, uint32_t gridtype, bool ac, bool sm, cudaStream_t st) {
    constexpr uint32_t NC = C >= 2 ? 2 : 1;
    const dim3 grid(ceil_div(B * C / NC, 256), L, 1);
    k_grid_bwd<T, D, C, NC><<<grid, 256, 0, st>>>((const T*)grad, inputs, offsets, (T*)gemb, B, L, S, H, gridtype, ac, sm);
    PVD_LAUNCH_CHECK();
    if (cgi) {
        k_grid_input_bwd<T, D, C><<<ceil_div(B * D, 256), 256, 0, st>>>((const T*)grad, (const T*)dy_dx, (T*)grad_inputs, B, L, sm);
        PVD_LAUNCH_CHECK();
    }
    return PVD_OK;
}

template <typename T, uint32_t D>
int dispatch_bwd(uint32_t C, const void* grad, const float* inputs, const int32_t* offsets, void* gemb, uint32_t B, uint32_t L,
                 float S, uint32_t H, bool cgi, const void* dy_dx, void* grad_inputs, uint32_t gridtype, bool ac,
                 bool sm, cudaStream_t st) {
    switch (C) {
        case 1: return launch_bwd<T, D, 1>(grad, inputs, offsets, gemb, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
        case 2: return launch_bwd<T, D, 2>(grad, inputs, offsets, gemb, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
        case 4: return launch_bwd<T, D, 4>(grad, inputs, offsets, gemb, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
        case 8: return launch_bwd<T, D, 8>(grad, inputs, offsets, gemb, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
        default: return PVD_EUNSUPPORTED;
    }
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_grid_level_table(const int32_t* offsets, uint32_t L, float S, uint32_t H, float* scales, int32_t* resolutions,
                         void* stream) {
    if (L == 0) return PVD_OK;
    PVD_REQUIRE(offsets && scales && resolutions);
    k_grid_level_table<<<ceil_div(L, 32), 32, 0, (cudaStream_t)stream>>>(offsets, L, S, H, scales, resolutions);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs, uint32_t B,
                            uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, void* dy_dx,
                            uint32_t gridtype, int align_corners, int dtype, int sample_major, void* stream) {
    if (B == 0 || L == 0) return PVD_OK;
    PVD_REQUIRE(inputs && embeddings && offsets && outputs);
    PVD_REQUIRE(!calc_grad_inputs || dy_dx);
    PVD_REQUIRE(dtype == PVD_DTYPE_F32 || dtype == PVD_DTYPE_F16);
    cudaStream_t st = (cudaStream_t)stream;
    const bool cgi = calc_grad_inputs != 0, ac = align_corners != 0, sm = sample_major != 0;
    if (dtype == PVD_DTYPE_F32) {
        if (D == 3) return dispatch_fwd<float, 3>(C, inputs, embeddings, offsets, outputs, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
        if (D == 2) return dispatch_fwd<float, 2>(C, inputs, embeddings, offsets, outputs, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
    } else {
        if (D == 3) return dispatch_fwd<__half, 3>(C, inputs, embeddings, offsets, outputs, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
        if (D == 2) return dispatch_fwd<__half, 2>(C, inputs, embeddings, offsets, outputs, B, L, S, H, cgi, dy_dx, gridtype, ac, sm, st);
    }
    return PVD_EUNSUPPORTED;  // D must be 2 or 3, gridencoder.cu:370
}

int pvd_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                             void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             int calc_grad_inputs, const void* dy_dx, void* grad_inputs, uint32_t gridtype,
                             int align_corners, int dtype, int sample_major, void* stream) {
    (void)embeddings;
    if (B == 0 || L == 0) return PVD_OK;
    PVD_REQUIRE(grad && inputs && offsets && grad_embeddings);
    PVD_REQUIRE(!calc_grad_inputs || (dy_dx && grad_inputs));
    PVD_REQUIRE(dtype == PVD_DTYPE_F32 || dtype == PVD_DTYPE_F16);
    cudaStream_t st = (cudaStream_t)stream;
    const bool cgi = calc_grad_inputs != 0, ac = align_corners != 0, sm = sample_major != 0;
    if (dtype == PVD_DTYPE_F32) {
        if (D == 3) return dispatch_bwd<float, 3>(C, grad, inputs, offsets, grad_embeddings, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
        if (D == 2) return dispatch_bwd<float, 2>(C, grad, inputs, offsets, grad_embeddings, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
    } else {
        if (D == 3) return dispatch_bwd<__half, 3>(C, grad, inputs, offsets, grad_embeddings, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
        if (D == 2) return dispatch_bwd<__half, 2>(C, grad, inputs, offsets, grad_embeddings, B, L, S, H, cgi, dy_dx, grad_inputs, gridtype, ac, sm, st);
    }
    return PVD_EUNSUPPORTED;
}

}  // extern "C"
