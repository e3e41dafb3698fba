// field_hash_f32.cu -- the fused "hash" field query in FP32 end to end (forward and backward), for runs WITHOUT --fp16 / -O
// (the reference then evaluates NeRFNetwork.forward in fp32: fp32 table gather, gridencoder.cu with scalar_t = float, fp32 cuBLAS
// GEMMs; distill_mutual/network.py:335-437) and for north_star's 1e-4 fp32 parity bound, which the fp16 tensor-core kernels of
// field_hash.cu cannot meet by construction.
//
// kind::f16 MMAs round their operands to 11 bits and kind::tf32 to 11 as well, so the small MLPs (28->64->16, 31->64->64->3:
// 9.4 kFMA per sample) run on the CUDA cores here:
//   * one thread = one sample, 128 samples per block; activations live in shared memory as fp32 COLUMNS A[k][sample] (a thread's
//     own column: any row pitch is bank-conflict-free), weights in shared memory in BOTH orientations: Wt[in][out] for the forward
//     (for a fixed input k, 16 consecutive outputs are four broadcast 16-byte loads feeding 16 independent FMA chains) and
//     W[out][in] for the data gradients;
//   * backward: forward recomputed into shared memory, then per layer (i) the WEIGHT gradient as a block-level outer-product
//     reduction over the 128 samples -- every thread owns a fixed set of (out, in) register blocks, reads G[out][s..s+3] and
//     A[in][s..s+3] as 16-byte vectors, and keeps its accumulators in registers across ALL tiles the persistent block processes
//     (74 per thread; they leave the SM once, into the same replicated workspace layout the fp16 kernels use) -- and (ii) the DATA
//     gradient per thread, masked in place over the activation it came from;
//   * the table gradient is scattered with the fp16 path's red.global.add.v2.f32 helper (gridencoder.cu:227-314 semantics).
// Arithmetic: fp32 FMAs accumulated in input order; exp / sigmoid through expf.
#include "field_common.cuh"
#include "field_hash.cuh"

namespace pvd {

constexpr uint32_t kB = 128;            // samples per tile = threads per block
constexpr uint32_t kLD = kB + 4;        // row pitch of an activation column block (floats): 16-byte aligned rows, skewed banks
// shared-memory weights (floats).  Forward orientation Wt[k][o]; data-gradient orientation W[o][k].
constexpr uint32_t kT1 = 0;                    // Wt1 [32][64]   sigma_net.0 (in padded 28 -> 32)
constexpr uint32_t kT2 = kT1 + 32 * 64;        // Wt2 [64][16]   sigma_net.1
constexpr uint32_t kT3 = kT2 + 64 * 16;        // Wt3 [32][64]   color_net.0 (in = 16 SH + 15 geo + pad)
constexpr uint32_t kT4 = kT3 + 32 * 64;        // Wt4 [64][64]   color_net.1
constexpr uint32_t kT5 = kT4 + 64 * 64;        // Wt5 [64][4]    color_net.2 (out padded 3 -> 4)
constexpr uint32_t kTEnd = kT5 + 64 * 4;       // 9472
constexpr uint32_t kO1 = kTEnd;                // W1 [64][32]
constexpr uint32_t kO2 = kO1 + 64 * 32;        // W2 [16][64]
constexpr uint32_t kO3 = kO2 + 16 * 64;        // W3 [64][32]
constexpr uint32_t kO4 = kO3 + 64 * 32;        // W4 [64][64]
constexpr uint32_t kO5 = kO4 + 64 * 64;        // W5 [4][64]
constexpr uint32_t kOEnd = kO5 + 4 * 64;       // 18944

struct F32Args {
    const float* table;
    const int32_t* offsets;
    const float* w[5];      // fp32 row-major [out][in]: [64][in_dim] [16][64] [64][31] [64][64] [3][64]
    uint32_t L, H, in_dim;
    float S, bound, clip_min, clip_max, density_scale;
};

__device__ __forceinline__ void stage_weights_f32(float* sw, const F32Args& a, bool both) {
    const uint32_t out[5] = {64, 16, 64, 64, 3}, in[5] = {a.in_dim, 64, 31, 64, 64};
    const uint32_t pin[5] = {32, 64, 32, 64, 64}, pout[5] = {64, 16, 64, 64, 4};
    const uint32_t toff[5] = {kT1, kT2, kT3, kT4, kT5}, ooff[5] = {kO1, kO2, kO3, kO4, kO5};
    for (uint32_t l = 0; l < 5; ++l) {
        for (uint32_t e = threadIdx.x; e < pin[l] * pout[l]; e += blockDim.x) {
            const uint32_t k = e / pout[l], o = e - k * pout[l];
            const float v = (o < out[l] && k < in[l]) ? __ldg(a.w[l] + (size_t)o * in[l] + k) : 0.0f;
            sw[toff[l] + k * pout[l] + o] = v;
            if (both) sw[ooff[l] + o * pin[l] + k] = v;
        }
    }
}

// out[o] = sum_k Wt[k][o] * in[k][tid]   for o in [0, OUT), OUT a multiple of 16 (or 4): accumulation in input order
template <uint32_t K, uint32_t OUT, bool RELU, uint32_t LD = kLD>
__device__ __forceinline__ void layer_fwd(const float* __restrict__ wt, const float* __restrict__ in, float* __restrict__ outp, uint32_t tid) {
    constexpr uint32_t OB = OUT < 16 ? OUT : 16;
#pragma unroll 1
    for (uint32_t ob = 0; ob < OUT; ob += OB) {
        float acc[OB];
#pragma unroll
        for (uint32_t i = 0; i < OB; ++i) acc[i] = 0.0f;
#pragma unroll 4
        for (uint32_t k = 0; k < K; ++k) {
            const float x = in[k * LD + tid];
#pragma unroll
            for (uint32_t i = 0; i < OB; i += 4) {
                const float4 w = *reinterpret_cast<const float4*>(wt + k * OUT + ob + i);
                acc[i] = fmaf(w.x, x, acc[i]);
                acc[i + 1] = fmaf(w.y, x, acc[i + 1]);
                acc[i + 2] = fmaf(w.z, x, acc[i + 2]);
                acc[i + 3] = fmaf(w.w, x, acc[i + 3]);
            }
        }
#pragma unroll
        for (uint32_t i = 0; i < OB; ++i) outp[(ob + i) * LD + tid] = RELU ? fmaxf(acc[i], 0.0f) : acc[i];
    }
}

// din[k] = sum_o W[o][k] * g[o][tid]   for k in [0, K); optionally masked by act[k][tid] > 0 and written over it
template <uint32_t K, uint32_t OUT, bool MASK>
__device__ __forceinline__ void layer_dgrad(const float* __restrict__ w, const float* __restrict__ g, float* __restrict__ act_io, uint32_t tid) {
#pragma unroll 1
    for (uint32_t kb = 0; kb < K; kb += 16) {
        float acc[16];
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) acc[i] = 0.0f;
#pragma unroll 4
        for (uint32_t o = 0; o < OUT; ++o) {
            const float gv = g[o * kLD + tid];
#pragma unroll
            for (uint32_t i = 0; i < 16; i += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(w + o * K + kb + i);
                acc[i] = fmaf(wv.x, gv, acc[i]);
                acc[i + 1] = fmaf(wv.y, gv, acc[i + 1]);
                acc[i + 2] = fmaf(wv.z, gv, acc[i + 2]);
                acc[i + 3] = fmaf(wv.w, gv, acc[i + 3]);
            }
        }
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) {
            float* p = act_io + (kb + i) * kLD + tid;
            *p = MASK ? ((*p > 0.0f) ? acc[i] : 0.0f) : acc[i];
        }
    }
}

// dW[o][k] += sum_s G[o][s] * A[k][s] over the tile: this thread's NB register blocks of 4 (o) x 4 (k); block b covers
// o = 4 * ((b * 128 + tid) / (K/4)), k = 4 * ((b * 128 + tid) % (K/4))
template <uint32_t OUT, uint32_t K, uint32_t NB>
__device__ __forceinline__ void layer_wgrad(const float* __restrict__ g, const float* __restrict__ act, float (&acc)[NB][16], uint32_t tid) {
    static_assert(OUT * K == NB * 16 * kB, "register blocks must tile the matrix");
#pragma unroll
    for (uint32_t b = 0; b < NB; ++b) {
        const uint32_t e = b * kB + tid, o0 = 4u * (e / (K / 4u)), k0 = 4u * (e % (K / 4u));
#pragma unroll 2
        for (uint32_t s = 0; s < kB; s += 4) {
            float4 gv[4], av[4];
#pragma unroll
            for (uint32_t i = 0; i < 4; ++i) {
                gv[i] = *reinterpret_cast<const float4*>(g + (o0 + i) * kLD + s);
                av[i] = *reinterpret_cast<const float4*>(act + (k0 + i) * kLD + s);
            }
#pragma unroll
            for (uint32_t i = 0; i < 4; ++i)
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    float r = acc[b][4 * i + j];
                    r = fmaf(gv[i].x, av[j].x, r);
                    r = fmaf(gv[i].y, av[j].y, r);
                    r = fmaf(gv[i].z, av[j].z, r);
                    r = fmaf(gv[i].w, av[j].w, r);
                    acc[b][4 * i + j] = r;
                }
        }
    }
}

template <uint32_t OUT, uint32_t K, uint32_t NB>
__device__ __forceinline__ void flush_wgrad(const float (&acc)[NB][16], float* __restrict__ dst, uint32_t pitch, bool transposed, uint32_t tid) {
#pragma unroll
    for (uint32_t b = 0; b < NB; ++b) {
        const uint32_t e = b * kB + tid, o0 = 4u * (e / (K / 4u)), k0 = 4u * (e % (K / 4u));
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i)
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                const uint32_t o = o0 + i, k = k0 + j;
                atomicAdd(dst + (transposed ? k * pitch + o : o * pitch + k), acc[b][4 * i + j]);
            }
    }
}

struct F32Tile {   // shared-memory column blocks of one tile (floats)
    float *X, *H1, *CIN, *H3, *H4, *G16, *G5;
};

// forward of this thread's sample into the tile's column blocks; returns sigma (scaled), rgb, o16 (channel 0 clamped), raw o0
template <uint32_t LD = kLD>
__device__ __forceinline__ void f32_forward(const F32Args& a, const float* sw, const F32Tile& t, uint32_t lv_saddr, const float (&pos)[3],
                                            const float (&dir)[3], bool live, uint32_t tid, float& sigma, float (&rgb)[3], float (&o16)[16],
                                            float& o0_raw) {
    float x01[3] = {0.5f, 0.5f, 0.5f};
    bool oob = true;
    if (live) to_unit(pos, a.bound, x01, oob);
    for (uint32_t l0 = 0; l0 < 16; l0 += 4) {
        float f[8];
        encode4<float>(a.table, lv_saddr, l0, a.L, x01, oob || !live, f);
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) t.X[(2 * l0 + i) * LD + tid] = f[i];
    }
    layer_fwd<32, 64, true, LD>(sw + kT1, t.X, t.H1, tid);
    layer_fwd<64, 16, false, LD>(sw + kT2, t.H1, t.G16, tid);     // G16 doubles as the 16-wide sigma_net output column block
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) o16[i] = t.G16[i * LD + tid];
    o0_raw = o16[0];
    const float o0c = clampf(o16[0], a.clip_min, a.clip_max);  // network.py:418-420
    o16[0] = o0c;
    sigma = a.density_scale * expf(o0c);                       // trunc_exp forward (tools/activation.py:9-12)
    float sh[16];
    sh_basis4(dir[0], dir[1], dir[2], sh);
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) t.CIN[i * LD + tid] = sh[i];
#pragma unroll
    for (uint32_t i = 0; i < 15; ++i) t.CIN[(16 + i) * LD + tid] = o16[i + 1];
    t.CIN[31 * LD + tid] = 0.0f;
    layer_fwd<32, 64, true, LD>(sw + kT3, t.CIN, t.H3, tid);
    layer_fwd<64, 64, true, LD>(sw + kT4, t.H3, t.H4, tid);
    layer_fwd<64, 4, false, LD>(sw + kT5, t.H4, t.G5, tid);
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i) rgb[i] = 1.0f / (1.0f + expf(-t.G5[i * LD + tid]));
}

// forward: 256 samples per block (8 warps per SM; the column blocks of 256 samples + the weights fill the SM's shared memory)
constexpr uint32_t kBF = 256, kLDF = kBF + 4;
__global__ void __launch_bounds__(kBF) k_hash_field_fwd_f32(F32Args a, const float* __restrict__ xyzs, const float* __restrict__ dirs, uint32_t M,
                                                          float* __restrict__ sigmas, float* __restrict__ rgbs, float* __restrict__ feat16) {
    extern __shared__ __align__(16) float smf[];
    __shared__ LevelInfo lv[16];
    float* sw = smf;                                  // kTEnd floats (forward orientation only)
    float* col = smf + kTEnd;
    F32Tile t;
    t.X = col; t.CIN = col;                           // 32 rows, CIN over X (dead after sigma_net.0)
    t.H1 = col + 32 * kLDF; t.H3 = t.H1;               // 64 rows
    t.H4 = t.H1 + 64 * kLDF;                           // 64 rows
    t.G16 = t.H4 + 64 * kLDF;                          // 16 rows
    t.G5 = t.G16 + 16 * kLDF;                          // 4 rows
    const uint32_t tid = threadIdx.x;
    stage_weights_f32(sw, a, false);
    level_info_init(lv, a.offsets, a.L, a.S, a.H);
    __syncthreads();
    const uint32_t lv_saddr = (uint32_t)__cvta_generic_to_shared(lv);
    const uint32_t n_tiles = (M + kBF - 1) / kBF;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row = tile * kBF + tid;
        const bool live = row < M;
        float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
        if (live) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
                dir[d] = __ldg(dirs + 3 * (size_t)row + d);
            }
        }
        float sigma, rgb[3], o16[16], o0_raw;
        f32_forward<kLDF>(a, sw, t, lv_saddr, pos, dir, live, tid, sigma, rgb, o16, o0_raw);   // thread-private columns: no barrier needed
        if (live) {
            sigmas[row] = sigma;
            rgbs[3 * (size_t)row] = rgb[0]; rgbs[3 * (size_t)row + 1] = rgb[1]; rgbs[3 * (size_t)row + 2] = rgb[2];
            if (feat16) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(feat16 + 16 * (size_t)row + 4 * q) = make_float4(o16[4 * q], o16[4 * q + 1], o16[4 * q + 2], o16[4 * q + 3]);
            }
        }
    }
}

__global__ void __launch_bounds__(kB, 1) k_hash_field_bwd_f32(F32Args a, const float* __restrict__ xyzs, const float* __restrict__ dirs,
                                                             const float* __restrict__ grad_sigmas, const float* __restrict__ grad_rgbs,
                                                             const float* __restrict__ grad_feat, uint32_t M, const int32_t* __restrict__ n_valid_p,
                                                             float* __restrict__ grad_table, float* __restrict__ gw) {
    extern __shared__ __align__(16) float smf[];
    __shared__ LevelInfo lv[16];
    float* sw = smf;                                  // kOEnd floats (both orientations)
    float* col = smf + kOEnd;
    F32Tile t;
    t.X = col;                     // 32
    t.H1 = t.X + 32 * kLD;         // 64
    t.CIN = t.H1 + 64 * kLD;       // 32
    t.H3 = t.CIN + 32 * kLD;       // 64
    t.H4 = t.H3 + 64 * kLD;        // 64
    t.G16 = t.H4 + 64 * kLD;       // 16
    t.G5 = t.G16 + 16 * kLD;       // 4
    const uint32_t tid = threadIdx.x;
    stage_weights_f32(sw, a, true);
    level_info_init(lv, a.offsets, a.L, a.S, a.H);
    __syncthreads();
    const uint32_t lv_saddr = (uint32_t)__cvta_generic_to_shared(lv);
    const uint32_t m_end = n_valid_p ? min((uint32_t)max(*n_valid_p, 0), M) : M;
    const uint32_t n_tiles = (m_end + kB - 1) / kB;
    // weight-gradient accumulators, persistent over the block's tiles: dW1 64x32, dW2 16x64, dW3 64x32, dW4 64x64, dW5 4x64
    float a1[1][16], a2[1][16] /* 16x64 = 1024 = half a block: see below */, a3[1][16], a4[2][16];
    float a5[2];   // dW5 [4][64] = 256 entries: 2 per thread
#pragma unroll
    for (int i = 0; i < 16; ++i) { a1[0][i] = a2[0][i] = a3[0][i] = a4[0][i] = a4[1][i] = 0.0f; }
    a5[0] = a5[1] = 0.0f;
    bool any = false;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        any = true;
        const uint32_t row = tile * kB + tid;
        const bool live = row < m_end;
        float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
        float gsig = 0.0f, grgb[3] = {0.f, 0.f, 0.f};
        if (live) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
                dir[d] = __ldg(dirs + 3 * (size_t)row + d);
                grgb[d] = __ldg(grad_rgbs + 3 * (size_t)row + d);
            }
            gsig = __ldg(grad_sigmas + row);
        }
        float sigma, rgb[3], o16[16], o0_raw;
        f32_forward(a, sw, t, lv_saddr, pos, dir, live, tid, sigma, rgb, o16, o0_raw);
        // ---- color_net.2: G5 = grad_rgb * rgb (1 - rgb)
#pragma unroll
        for (uint32_t i = 0; i < 3; ++i) t.G5[i * kLD + tid] = live ? grgb[i] * rgb[i] * (1.0f - rgb[i]) : 0.0f;
        t.G5[3 * kLD + tid] = 0.0f;
        __syncthreads();
        {   // dW5 [4][64] += G5 H4^T : entries e = tid, tid + 128 -> (o, k) = (e / 64, e % 64)
#pragma unroll
            for (uint32_t q = 0; q < 2; ++q) {
                const uint32_t e = q * kB + tid, o = e >> 6, k = e & 63u;
                float r = a5[q];
                for (uint32_t s = 0; s < kB; s += 4) {
                    const float4 gv = *reinterpret_cast<const float4*>(t.G5 + o * kLD + s), av = *reinterpret_cast<const float4*>(t.H4 + k * kLD + s);
                    r = fmaf(gv.x, av.x, r); r = fmaf(gv.y, av.y, r); r = fmaf(gv.z, av.z, r); r = fmaf(gv.w, av.w, r);
                }
                a5[q] = r;
            }
        }
        __syncthreads();
        layer_dgrad<64, 4, true>(sw + kO5, t.G5, t.H4, tid);          // G4 over H4
        __syncthreads();
        layer_wgrad<64, 64, 2>(t.H4, t.H3, a4, tid);                  // dW4 += G4 H3^T
        __syncthreads();
        layer_dgrad<64, 64, true>(sw + kO4, t.H4, t.H3, tid);         // G3 over H3
        __syncthreads();
        layer_wgrad<64, 32, 1>(t.H3, t.CIN, a3, tid);                 // dW3 += G3 CIN^T
        __syncthreads();
        layer_dgrad<32, 64, false>(sw + kO3, t.H3, t.CIN, tid);       // dCIN over CIN (rows 16..30 = d geo)
        // ---- sigma_net.1 output gradient: channel 0 through trunc_exp + clamp, channels 1..15 = geo part of dCIN, + grad_feat
        {
            const bool inside = (o0_raw >= a.clip_min) && (o0_raw <= a.clip_max);
            float g[16];
            g[0] = gsig * a.density_scale * expf(clampf(o16[0], -12.0f, 12.0f));   // tools/activation.py:15-21
#pragma unroll
            for (uint32_t i = 0; i < 15; ++i) g[i + 1] = t.CIN[(16 + i) * kLD + tid];
            if (grad_feat && live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 gf = __ldg(reinterpret_cast<const float4*>(grad_feat + 16 * (size_t)row) + q);
                    g[4 * q] += gf.x; g[4 * q + 1] += gf.y; g[4 * q + 2] += gf.z; g[4 * q + 3] += gf.w;
                }
            }
            if (!inside) g[0] = 0.0f;
#pragma unroll
            for (uint32_t i = 0; i < 16; ++i) t.G16[i * kLD + tid] = live ? g[i] : 0.0f;
        }
        __syncthreads();
        {   // dW2 [16][64] += G2 H1^T : 1024 entries, 8 per thread as a 2 (o) x 4 (k) block: o0 = 2 * (tid / 16), k0 = 4 * (tid % 16)
            const uint32_t o0 = 2u * (tid >> 4), k0 = 4u * (tid & 15u);
            for (uint32_t s = 0; s < kB; s += 4) {
                float4 gv[2], av[4];
#pragma unroll
                for (uint32_t i = 0; i < 2; ++i) gv[i] = *reinterpret_cast<const float4*>(t.G16 + (o0 + i) * kLD + s);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const float4*>(t.H1 + (k0 + j) * kLD + s);
#pragma unroll
                for (uint32_t i = 0; i < 2; ++i)
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        float r = a2[0][4 * i + j];
                        r = fmaf(gv[i].x, av[j].x, r); r = fmaf(gv[i].y, av[j].y, r); r = fmaf(gv[i].z, av[j].z, r); r = fmaf(gv[i].w, av[j].w, r);
                        a2[0][4 * i + j] = r;
                    }
            }
        }
        __syncthreads();
        layer_dgrad<64, 16, true>(sw + kO2, t.G16, t.H1, tid);        // G1 over H1
        __syncthreads();
        layer_wgrad<64, 32, 1>(t.H1, t.X, a1, tid);                   // dW1 += G1 X^T
        __syncthreads();
        layer_dgrad<32, 64, false>(sw + kO1, t.H1, t.X, tid);         // dX over X
        // ---- scatter d(encoding) into the table gradient (gridencoder.cu:227-314, fp32)
        if (live) {
            float x01[3];
            bool oob;
            to_unit(pos, a.bound, x01, oob);
            if (!oob) {
                for (uint32_t l = 0; l < a.L; ++l) {
                    const LevelInfo v = ld_level(lv_saddr, l);
                    Corners c;
                    level_corners(v, x01, c);
                    float* gt = grad_table + (size_t)v.offset * 2;
                    const float g0 = t.X[(2 * l) * kLD + tid], g1 = t.X[(2 * l + 1) * kLD + tid];
                    if (g0 != 0.0f || g1 != 0.0f) {
#pragma unroll
                        for (uint32_t i = 0; i < 8; ++i) tab_red2(gt, (size_t)c.idx[i] * 2, c.w[i] * g0, c.w[i] * g1);
                    }
                }
            }
        }
        __syncthreads();   // the next tile's forward rewrites the column blocks other threads' weight-gradient loops read
    }
    if (any) {   // weight gradients leave the SM once, in the fp16 kernels' workspace layout (pvd_field_unpack_wgrads reads it)
        gw += (size_t)(blockIdx.x % PVD_FIELD_GW_COPIES) * PVD_FIELD_GW_FLOATS;
        flush_wgrad<64, 32, 1>(a1, gw + kGW1, 32, false, tid);
        {   // dW2 -> dW2^T [64 in][16 out]
            const uint32_t o0 = 2u * (tid >> 4), k0 = 4u * (tid & 15u);
#pragma unroll
            for (uint32_t i = 0; i < 2; ++i)
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) atomicAdd(gw + kGW2 + (k0 + j) * 16u + (o0 + i), a2[0][4 * i + j]);
        }
        flush_wgrad<64, 32, 1>(a3, gw + kGW3, 32, false, tid);
        flush_wgrad<64, 64, 2>(a4, gw + kGW4, 64, false, tid);
#pragma unroll
        for (uint32_t q = 0; q < 2; ++q) {   // dW5 -> dW5^T [64 in][16 out]
            const uint32_t e = q * kB + tid, o = e >> 6, k = e & 63u;
            if (o < 3u) atomicAdd(gw + kGW5 + k * 16u + o, a5[q]);
        }
    }
}

constexpr size_t kF32FwdSmem = (kTEnd + (32 + 64 + 64 + 16 + 4) * kLDF) * sizeof(float);             // 225 088
constexpr size_t kF32BwdSmem = (kOEnd + (32 + 64 + 32 + 64 + 64 + 16 + 4) * kLD) * sizeof(float);    // 221 504

static F32Args to_f32_args(const PvdHashField* f, const PvdFieldWeightsF32* w) {
    F32Args a;
    a.table = reinterpret_cast<const float*>(f->table);
    a.offsets = f->offsets;
    a.w[0] = w->sigma0; a.w[1] = w->sigma1; a.w[2] = w->color0; a.w[3] = w->color1; a.w[4] = w->color2;
    a.L = f->L; a.H = f->H; a.in_dim = 2 * f->L;
    a.S = f->S; a.bound = f->bound; a.clip_min = f->sigma_clip_min; a.clip_max = f->sigma_clip_max; a.density_scale = f->density_scale;
    return a;
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_hash_field_forward_f32(const PvdHashField* f, const PvdFieldWeightsF32* w, const float* xyzs, const float* dirs, uint32_t M,
                               float* sigmas, float* rgbs, float* feat16, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && w && f->table && f->offsets && w->sigma0 && w->sigma1 && w->color0 && w->color1 && w->color2 && xyzs && dirs && sigmas && rgbs);
    if (f->L == 0 || f->L > 16 || f->table_dtype != PVD_DTYPE_F32) return PVD_EUNSUPPORTED;
    const F32Args a = to_f32_args(f, w);
    cudaError_t e = cudaFuncSetAttribute(k_hash_field_fwd_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF32FwdSmem);
    if (e != cudaSuccess) return (int)e;
    const uint32_t tiles = (M + kBF - 1) / kBF, grid = min(tiles, (uint32_t)sm_count());
    k_hash_field_fwd_f32<<<grid, kBF, kF32FwdSmem, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, feat16);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_hash_field_backward_f32(const PvdHashField* f, const PvdFieldWeightsF32* w, const float* xyzs, const float* dirs,
                                const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M, const int32_t* n_valid,
                                float* grad_table, float* gw_ws, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && w && f->table && f->offsets && w->sigma0 && w->sigma1 && w->color0 && w->color1 && w->color2 && xyzs && dirs && grad_sigmas &&
                grad_rgbs && grad_table && gw_ws);
    if (f->L == 0 || f->L > 16 || f->table_dtype != PVD_DTYPE_F32) return PVD_EUNSUPPORTED;
    const F32Args a = to_f32_args(f, w);
    cudaError_t e = cudaFuncSetAttribute(k_hash_field_bwd_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF32BwdSmem);
    if (e != cudaSuccess) return (int)e;
    const uint32_t tiles = (M + kB - 1) / kB, grid = min(tiles, (uint32_t)sm_count());
    k_hash_field_bwd_f32<<<grid, kB, kF32BwdSmem, (cudaStream_t)stream>>>(a, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat16, M, n_valid, grad_table, gw_ws);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
