// field_tensors.cu -- the "tensors" (Plenoxels-style dense voxel) field and on-device ray generation.
//
// (1) NeRFNetwork.forward for model_type "tensors" (distill_mutual/network.py:184-191 init, :311-322 compute_plenoxel_fea,
//     :383-409 forward): ONE trilinear F.grid_sample (align_corners=True, zero padding) of a [1, C, D, H, W] volume with
//     C = 3 * degree^2 + 1 channels, then   sigma = trunc_exp(clamp(h0)),   rgb_k = sigmoid(sum_j h[1 + 9 k + j] * SH_j(dir))
//     -- no MLP.  The reference's channel-first volume makes every one of the 8 corners C separate cache lines (C * 8 = 224 sectors
//     per sample); here the parameter keeps its shape but lives in torch.channels_last_3d memory ([D][H][W][C]), so a corner is C
//     contiguous floats (112 bytes at degree 3) and the gather is 8 coalesced reads per sample.  The volume (235 MB at 128^3 x 28)
//     does not fit L2: this field is HBM-bound, 8 * C * 4 = 896 B/sample forward, the same again as reductions backward.
//     Work decomposition: 8 lanes per sample (4 samples per warp), lane l holds channels 4l..4l+3 as one float4.
// (2) get_rays (distill_mutual/utils.py:324-404): pixel index -> camera-space direction -> world ray, one thread per ray.
#include "common.cuh"
#include "shenc.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

struct TensorsArgs {
    const float* volume;   // [D][H][W][C] (channels_last_3d memory of a [1, C, D, H, W] parameter)
    uint32_t D, H, W, C, degree;
    float aabb[6];
    float clip_min, clip_max, density_scale;
};

constexpr uint32_t kTLanes = 8;   // lanes per sample: up to 32 channels (degree <= 3: 28)

struct TriFoot {
    uint32_t idx[8];   // voxel index (z * H + y) * W + x of the 8 corners
    float w[8];        // trilinear weights, 0 for corners outside the volume (grid_sample zero padding)
};

// grid_sample, align_corners=True: coordinate c in [-1, 1] -> (c + 1) / 2 * (size - 1); x indexes W, y indexes H, z indexes D
__device__ __forceinline__ void tri_foot(const float (&xn)[3], uint32_t D, uint32_t H, uint32_t W, TriFoot& f) {
    const float ix = (xn[0] + 1.0f) * 0.5f * (float)(W - 1), iy = (xn[1] + 1.0f) * 0.5f * (float)(H - 1), iz = (xn[2] + 1.0f) * 0.5f * (float)(D - 1);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int xx = x0 + (k & 1), yy = y0 + ((k >> 1) & 1), zz = z0 + (k >> 2);
        const bool in = xx >= 0 && xx < (int)W && yy >= 0 && yy < (int)H && zz >= 0 && zz < (int)D;
        f.idx[k] = in ? (uint32_t)((zz * (int)H + yy) * (int)W + xx) : 0u;
        f.w[k] = in ? ((k & 1) ? tx : 1.0f - tx) * (((k >> 1) & 1) ? ty : 1.0f - ty) * ((k >> 2) ? tz : 1.0f - tz) : 0.0f;
    }
}

__device__ __forceinline__ void tensors_normalise(const float* pos, const float* aabb, float (&xn)[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = __fdiv_rn(2.0f * (pos[d] - aabb[d]), aabb[3 + d] - aabb[d]) - 1.0f;   // network.py:384-389
}

// this lane's four channels of the interpolated feature vector of one sample
__device__ __forceinline__ float4 tensors_gather(const TensorsArgs& a, const TriFoot& f, uint32_t ch, bool lane_on) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        v[k] = lane_on ? __ldg(reinterpret_cast<const float4*>(a.volume + (size_t)f.idx[k] * a.C + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        h.x = __fmaf_rn(f.w[k], v[k].x, h.x);
        h.y = __fmaf_rn(f.w[k], v[k].y, h.y);
        h.z = __fmaf_rn(f.w[k], v[k].z, h.z);
        h.w = __fmaf_rn(f.w[k], v[k].w, h.w);
    }
    return h;
}

// SH coefficient multiplying channel c (c >= 1): sh[(c - 1) % n_sh]; colour it belongs to: (c - 1) / n_sh   (network.py:401-405)
template <bool BWD>
__global__ void __launch_bounds__(256) k_tensors_field(TensorsArgs a, const float* __restrict__ xyzs, const float* __restrict__ dirs, uint32_t M,
                                                      float* __restrict__ sigmas, float* __restrict__ rgbs,
                                                      const float* __restrict__ grad_sigmas, const float* __restrict__ grad_rgbs,
                                                      const int32_t* __restrict__ n_valid_p, float* __restrict__ grad_volume) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t row = gtid / kTLanes, l = gtid % kTLanes;
    const uint32_t limit = (BWD && n_valid_p) ? min((uint32_t)max(*n_valid_p, 0), M) : M;
    const bool live = row < limit;          // whole 8-lane groups are live or not: the shuffles below stay inside a group
    const uint32_t ch = 4u * l;
    const bool lane_on = live && ch < a.C;
    const uint32_t n_sh = a.degree * a.degree;
    float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f};
    if (live) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
            dir[d] = __ldg(dirs + 3 * (size_t)row + d);
        }
    }
    float xn[3];
    tensors_normalise(pos, a.aabb, xn);
    TriFoot f;
    tri_foot(xn, a.D, a.H, a.W, f);
    const float4 h = tensors_gather(a, f, ch, lane_on);
    float sh[16];
    sh_basis<float>(dir[0], dir[1], dir[2], a.degree, [&](uint32_t i, float v) { if (i < 16) sh[i] = v; });
    // this lane's contribution to the three colour pre-activations
    const float hv[4] = {h.x, h.y, h.z, h.w};
    float part[3] = {0.f, 0.f, 0.f};
    float coef[4] = {0.f, 0.f, 0.f, 0.f};
    int col[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t c = ch + q;
        if (c >= 1u && c < a.C) {
            col[q] = (int)((c - 1u) / n_sh);
            coef[q] = sh[(c - 1u) % n_sh];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (col[q] == k) part[k] = __fmaf_rn(hv[q], coef[q], part[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        part[k] += __shfl_xor_sync(0xffffffffu, part[k], 1);
        part[k] += __shfl_xor_sync(0xffffffffu, part[k], 2);
        part[k] += __shfl_xor_sync(0xffffffffu, part[k], 4);
    }
    const float h0 = __shfl_sync(0xffffffffu, h.x, (int)((threadIdx.x & 31u) & ~(kTLanes - 1u)));   // channel 0 lives in lane 0 of the group
    const float h0c = clampf(h0, a.clip_min, a.clip_max);                                          // network.py:391-397
    float rgb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[k] = 1.0f / (1.0f + __expf(-part[k]));
    if (!BWD) {
        if (live && l == 0) {
            sigmas[row] = a.density_scale * __expf(h0c);
            rgbs[3 * (size_t)row] = rgb[0];
            rgbs[3 * (size_t)row + 1] = rgb[1];
            rgbs[3 * (size_t)row + 2] = rgb[2];
        }
        return;
    }
    if (!lane_on) return;
    // d(loss)/d(h): channel 0 through trunc_exp (tools/activation.py:15-21) and the clamp mask; colour channels through the sigmoid
    const float gs = __ldg(grad_sigmas + row);
    float dpre[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dpre[k] = __ldg(grad_rgbs + 3 * (size_t)row + k) * rgb[k] * (1.0f - rgb[k]);
    float dh[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t c = ch + q;
        float g = 0.0f;
        if (c == 0u) {
            const bool inside = (h0 >= a.clip_min) && (h0 <= a.clip_max);
            g = inside ? gs * a.density_scale * __expf(clampf(h0c, -12.0f, 12.0f)) : 0.0f;
        } else if (c < a.C) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (col[q] == k) g = dpre[k] * coef[q];
        }
        dh[q] = g;
    }
    if (dh[0] == 0.0f && dh[1] == 0.0f && dh[2] == 0.0f && dh[3] == 0.0f) return;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (f.w[k] == 0.0f) continue;
        float* dst = grad_volume + (size_t)f.idx[k] * a.C + ch;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(f.w[k] * dh[0]), "f"(f.w[k] * dh[1]), "f"(f.w[k] * dh[2]),
                     "f"(f.w[k] * dh[3])
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- get_rays
// inds == nullptr: all H*W pixels of every pose (N = H*W).  Pixel p: i = p % W + 0.5, j = p / W + 0.5 (utils.py:342-347),
// direction ((i - cx) / fx, (j - cy) / fy, 1) normalised, rotated by the pose's upper-left 3x3; origin = its translation column.
__global__ void __launch_bounds__(256) k_get_rays(const float* __restrict__ poses, float fx, float fy, float cx, float cy, uint32_t W,
                                                 const int64_t* __restrict__ inds, uint32_t inds_batch_stride, uint32_t B, uint32_t N,
                                                 float* __restrict__ rays_o, float* __restrict__ rays_d) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * N) return;
    const uint32_t b = t / N, n = t - b * N;
    const int64_t p = inds ? inds[(size_t)b * inds_batch_stride + n] : (int64_t)n;
    const float i = (float)(p % (int64_t)W) + 0.5f, j = (float)(p / (int64_t)W) + 0.5f;
    const float x = __fdiv_rn(i - cx, fx), y = __fdiv_rn(j - cy, fy), z = 1.0f;
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fmaf_rn(x, x, __fmaf_rn(y, y, z * z))));
    const float dx = x * inv, dy = y * inv, dz = z * inv;
    const float* P = poses + 16 * (size_t)b;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        rays_d[3 * (size_t)t + k] = __fmaf_rn(dz, P[4 * k + 2], __fmaf_rn(dy, P[4 * k + 1], dx * P[4 * k]));
        rays_o[3 * (size_t)t + k] = P[4 * k + 3];
    }
}

static bool to_tensors_args(const PvdTensorsField* f, TensorsArgs& a) {
    if (!f || !f->volume || f->degree < 1 || f->degree > 3) return false;
    a.volume = f->volume;
    a.D = f->res[0]; a.H = f->res[1]; a.W = f->res[2];
    a.degree = f->degree;
    a.C = 3u * f->degree * f->degree + 1u;
    if (a.D < 2 || a.H < 2 || a.W < 2 || (a.C & 3u) != 0u || a.C > 4u * kTLanes) return false;   // degree 1 (C = 4), 3 (C = 28); see the header
    if ((reinterpret_cast<uintptr_t>(a.volume) & 15u) != 0) return false;
    for (int i = 0; i < 6; ++i) a.aabb[i] = f->aabb[i];
    a.clip_min = f->sigma_clip_min; a.clip_max = f->sigma_clip_max; a.density_scale = f->density_scale;
    return true;
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_tensors_field_forward(const PvdTensorsField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                              void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(xyzs && dirs && sigmas && rgbs);
    TensorsArgs a;
    if (!to_tensors_args(f, a)) return PVD_EUNSUPPORTED;
    const uint64_t threads = (uint64_t)M * kTLanes;
    k_tensors_field<false><<<(uint32_t)((threads + 255u) / 256u), 256, 0, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, nullptr, nullptr,
                                                                                              nullptr, nullptr);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_tensors_field_backward(const PvdTensorsField* f, const float* xyzs, const float* dirs, const float* grad_sigmas, const float* grad_rgbs,
                               uint32_t M, const int32_t* n_valid, float* grad_volume, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(xyzs && dirs && grad_sigmas && grad_rgbs && grad_volume);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(grad_volume) & 15u) == 0);
    TensorsArgs a;
    if (!to_tensors_args(f, a)) return PVD_EUNSUPPORTED;
    const uint64_t threads = (uint64_t)M * kTLanes;
    k_tensors_field<true><<<(uint32_t)((threads + 255u) / 256u), 256, 0, (cudaStream_t)stream>>>(a, xyzs, dirs, M, nullptr, nullptr, grad_sigmas,
                                                                                             grad_rgbs, n_valid, grad_volume);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, const int64_t* inds,
                 uint32_t inds_batch_stride, uint32_t B, uint32_t N, float* rays_o, float* rays_d, void* stream) {
    if (B == 0 || N == 0) return PVD_OK;
    PVD_REQUIRE(poses && rays_o && rays_d && H >= 1 && W >= 1 && fx != 0.0f && fy != 0.0f);
    PVD_REQUIRE(inds != nullptr || N == H * W);
    const uint64_t total = (uint64_t)B * N;
    PVD_REQUIRE(total < (1ull << 32));
    k_get_rays<<<(uint32_t)((total + 255u) / 256u), 256, 0, (cudaStream_t)stream>>>(poses, fx, fy, cx, cy, W, inds, inds_batch_stride, B, N, rays_o,
                                                                                rays_d);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
