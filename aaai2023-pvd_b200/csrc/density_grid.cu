// density_grid.cu -- occupancy-grid upkeep: NeRFRenderer.update_extra_state of the reference
// (distill_mutual/renderer.py:647-773), which every 16 training steps
//   1. builds one jittered query point per grid cell (full sweep, first 16 updates: the H^3 cells in the order of a meshgrid, written
//      through morton3D; later: H^3/4 random cells + as many cells drawn from the occupied ones) -- meshgrid, cat, morton3D,
//      float conversion, two scalings, torch.rand_like and an add, ~12 launches;
//   2. queries the field's density there;
//   3. tmp_grid[cas, indices] = sigma * density_scale; valid = (grid >= 0) & (tmp >= 0); grid[valid] = max(grid[valid] * decay, tmp[valid]);
//   4. mean_density = mean(clamp(grid, 0)).item()  (a host sync), thresh = min(mean_density, density_thresh), packbits.
// With a training step of 0.1 ms, ~1.5 ms of upkeep every 16 steps costs as much as the steps between two updates.  Here steps 1,
// 3 and 4 are three kernels and nothing is read back: the threshold is formed on the device from the accumulated sum.
//
// Arithmetic of the points is torch's, operation by operation in fp32 (explicit round-to-nearest intrinsics, no contraction), so
// that with the same noise the fused path produces bit-identical positions:
//     x = ((2 c) * fl(1 / (H - 1)) - 1) * (bound - half) + (2 u - 1) * half,    half = bound / H        (renderer.py:679-694)
#include "common.cuh"

namespace pvd {

__global__ void __launch_bounds__(256) k_density_grid_points(const int32_t* __restrict__ indices, const float* __restrict__ noise,
                                                             uint32_t n, float rcp_hm1, float scale, float half, float* __restrict__ xyzs) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t ind = indices ? (uint32_t)indices[j] : j;   // full sweep: cell j itself (Morton order covers every cell once)
    const uint32_t c[3] = {compact3(ind), compact3(ind >> 1), compact3(ind >> 2)};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // ATen's CUDA division by a host scalar multiplies by the fp32 reciprocal (BinaryDivTrueKernel.cu), not an IEEE divide
        const float base = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, (float)c[d]), rcp_hm1), -1.0f);
        const float jit = __fmul_rn(__fadd_rn(__fmul_rn(__ldg(noise + 3 * (size_t)j + d), 2.0f), -1.0f), half);
        xyzs[3 * (size_t)j + d] = __fadd_rn(__fmul_rn(base, scale), jit);
    }
}

// partial update only: tmp[indices[j]] = max over duplicates of sigma_j * density_scale (non-negative floats order like their bits);
// the reference's index_put_ keeps an arbitrary one of the duplicates (renderer.py:737)
__global__ void __launch_bounds__(256) k_density_tmp_scatter(const int32_t* __restrict__ indices, const float* __restrict__ sigmas,
                                                             uint32_t n, float density_scale, float* __restrict__ tmp) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float s = __fmul_rn(__ldg(sigmas + j), density_scale);
    // candidates are stored as bits + 1: 0 = cell not visited, and the order of non-negative floats is the order of their bits
    if (s >= 0.0f) atomicMax(reinterpret_cast<unsigned int*>(tmp) + (uint32_t)indices[j], __float_as_uint(s) + 1u);
}

// EMA update of one cascade + partial sum of clamp(grid, 0).  tmp: per-cell candidate (float bits + 1, 0 = not visited), or
// NULL for the full sweep where cell i's candidate is sigmas[i] * density_scale.
__global__ void __launch_bounds__(256) k_density_ema(float* __restrict__ grid, const float* __restrict__ tmp, const float* __restrict__ sigmas,
                                                     uint32_t n_cells, float density_scale, float decay, double* __restrict__ sum) {
    __shared__ float sh[8];
    float acc = 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        float t;
        if (tmp) {
            const uint32_t raw = __float_as_uint(tmp[i]);
            t = (raw == 0u) ? -1.0f : __uint_as_float(raw - 1u);
        } else {
            t = __fmul_rn(__ldg(sigmas + i), density_scale);
        }
        if (g >= 0.0f && t >= 0.0f) {   // renderer.py:746-749
            g = fmaxf(__fmul_rn(g, decay), t);
            grid[i] = g;
        }
        acc += fmaxf(g, 0.0f);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31u) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(sum, (double)v);
    }
}

// packbits (raymarching.cu:270-291) with thresh = min(sum / count, density_thresh) formed on the device (renderer.py:750-759)
__global__ void __launch_bounds__(256) k_packbits_mean(const float* __restrict__ grid, uint32_t N, const double* __restrict__ sum,
                                                       double inv_count, float density_thresh, uint8_t* __restrict__ bitfield,
                                                       float* __restrict__ mean_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const float mean = (float)(*sum * inv_count);
    if (n == 0 && mean_out) *mean_out = mean;
    if (n >= N) return;
    const float thresh = fminf(mean, density_thresh);
    const float4* g = reinterpret_cast<const float4*>(grid) + 2 * (size_t)n;
    const float4 a = __ldg(g), b = __ldg(g + 1);
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;
    bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;
    bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;
    bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;
    bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_density_grid_points(const int32_t* indices, const float* noise, uint32_t n, uint32_t H, float bound_cas, float* xyzs,
                            void* stream) {
    if (n == 0) return PVD_OK;
    PVD_REQUIRE(noise && xyzs && H >= 2 && H <= 1024);
    // the reference forms both scalars in Python doubles and multiplies fp32 tensors by them (the scalar is cast to fp32)
    const float scale = (float)((double)bound_cas - (double)bound_cas / (double)H);
    const float half = (float)((double)bound_cas / (double)H);
    k_density_grid_points<<<ceil_div(n, 256u), 256, 0, (cudaStream_t)stream>>>(indices, noise, n, 1.0f / (float)(H - 1), scale, half, xyzs);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_density_grid_update(float* grid, float* tmp, const int32_t* indices, const float* sigmas, uint32_t n, uint32_t n_cells,
                            float density_scale, float decay, double* sum, void* stream) {
    if (n_cells == 0) return PVD_OK;
    PVD_REQUIRE(grid && sigmas && sum);
    cudaStream_t st = (cudaStream_t)stream;
    if (indices != nullptr) {   // partial update: candidates go through tmp
        PVD_REQUIRE(tmp != nullptr);
        cudaError_t e = cudaMemsetAsync(tmp, 0, (size_t)n_cells * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
        if (n) k_density_tmp_scatter<<<ceil_div(n, 256u), 256, 0, st>>>(indices, sigmas, n, density_scale, tmp);
        PVD_LAUNCH_CHECK();
    } else {
        PVD_REQUIRE(n == n_cells);
    }
    const uint32_t grid_x = min(ceil_div(n_cells, 256u), 148u * 8u);
    k_density_ema<<<grid_x, 256, 0, st>>>(grid, indices ? tmp : nullptr, sigmas, n_cells, density_scale, decay, sum);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_packbits_mean(const float* grid, uint32_t N, const double* sum, uint32_t count, float density_thresh, uint8_t* bitfield,
                      float* mean_out, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(grid && sum && bitfield && count > 0);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15u) == 0);
    k_packbits_mean<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(grid, N, sum, 1.0 / (double)count, density_thresh, bitfield,
                                                                        mean_out);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
