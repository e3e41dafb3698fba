// optim.cu -- the fused optimizer step around the hot path (include/pvd_b200_optim.h).
//
// The reference spends, per iteration and OUTSIDE its kernels, several dense passes over the 42 MB hash table that exist only
// because optimizer, GradScaler and encoder wrapper are separate torch modules: zeros_like(embeddings) (gridencoder/grid.py:106),
// embeddings.to(half) (grid.py:52), the fp16 -> fp32 gradient cast, GradScaler.unscale_ (one pass per tensor), AdamW (~12 foreach
// kernels).  Here it is ONE pass: k_adamw_multi reads (param, exp_avg, exp_avg_sq, grad), writes the three back, writes the fp16
// shadow the field kernels gather from, and zeroes the gradient accumulator -- 34 B per element, HBM-bound (the table, its two
// moments and its gradient are 170 MB: larger than L2).
//
// Arithmetic = torch.optim.AdamW, single-tensor CUDA path (torch/optim/adam.py::_single_tensor_adam), operation by operation:
//     param.mul_(1 - lr * wd)                                   -> p = p * decay
//     exp_avg.lerp_(grad, 1 - beta1)                            -> m = fma(w1, g - m, m)          (ATen lerp, |w| < 0.5 branch)
//     exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-b2)   -> v = fma(w2, g*g, v*beta2)       (or fma(w2*g, g, .), see flags)
//     denom = (exp_avg_sq.sqrt() / bc2_sqrt).add_(eps)          -> d = sqrt(v) * fl32(1 / bc2_sqrt) + eps (ATen divides a tensor by a CPU
//                                                                  scalar by multiplying with its reciprocal, formed in double)
//     param.addcdiv_(exp_avg, denom, value=-step_size)          -> p = fma(-step_size, m / d, p)
// with explicit round-to-nearest intrinsics so that nvcc's contraction choices cannot change the result.
#include "common.cuh"
#include "../../include/pvd_b200_optim.h"

namespace pvd {

struct AdamConsts {
    float w1, w2, beta2, inv_bc2_sqrt, eps, grad_scale;
    bool left;
};

__device__ __forceinline__ void adam_one(float& p, float& m, float& v, float g, const AdamConsts& c, float decay, float neg_step, float gmul) {
    g = __fmul_rn(__fmul_rn(g, gmul), c.grad_scale);                       // GradScaler.unscale_: grad * inv_scale
    p = __fmul_rn(p, decay);
    m = __fmaf_rn(c.w1, __fsub_rn(g, m), m);
    const float vb = __fmul_rn(v, c.beta2);
    v = c.left ? __fmaf_rn(__fmul_rn(c.w2, g), g, vb) : __fmaf_rn(c.w2, __fmul_rn(g, g), vb);
    const float d = __fadd_rn(__fmul_rn(__fsqrt_rn(v), c.inv_bc2_sqrt), c.eps);
    p = __fmaf_rn(neg_step, __fdiv_rn(m, d), p);
}

__device__ __forceinline__ float4 ld_grad4(const PvdAdamSlot& s, uint64_t i4) {
    if (s.grad_f16 != nullptr) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(s.grad_f16) + i4);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return *(reinterpret_cast<const float4*>(s.grad) + i4);
}

__global__ void __launch_bounds__(256) k_adamw_multi(const PvdAdamState* __restrict__ state, const PvdAdamSlot* __restrict__ slots) {
    const PvdAdamSlot s = slots[blockIdx.y];
    const PvdAdamState st = *state;
    const bool skip = st.found_inf != 0;   // GradScaler.step: no update when a gradient was non-finite; gradients are still cleared
    const AdamConsts c{st.w1, st.w2, st.beta2_f, st.inv_bc2_sqrt, st.eps, st.grad_scale, (st.flags & PVD_ADAM_ADDCMUL_LEFT) != 0u};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(s.param) | reinterpret_cast<uintptr_t>(s.exp_avg) | reinterpret_cast<uintptr_t>(s.exp_avg_sq) |
                       reinterpret_cast<uintptr_t>(s.grad) | reinterpret_cast<uintptr_t>(s.shadow_f16)) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(s.grad_f16) & 7u) == 0;
    const uint64_t n4 = vec ? s.n / 4u : 0u;
    for (uint64_t i = t0; i < n4; i += stride) {
        if (!skip) {
            float4 p = *(reinterpret_cast<const float4*>(s.param) + i);
            float4 m = *(reinterpret_cast<const float4*>(s.exp_avg) + i);
            float4 v = *(reinterpret_cast<const float4*>(s.exp_avg_sq) + i);
            const float4 g = ld_grad4(s, i);
            adam_one(p.x, m.x, v.x, g.x, c, s.decay, s.neg_step_size, s.grad_mul);
            adam_one(p.y, m.y, v.y, g.y, c, s.decay, s.neg_step_size, s.grad_mul);
            adam_one(p.z, m.z, v.z, g.z, c, s.decay, s.neg_step_size, s.grad_mul);
            adam_one(p.w, m.w, v.w, g.w, c, s.decay, s.neg_step_size, s.grad_mul);
            *(reinterpret_cast<float4*>(s.param) + i) = p;
            *(reinterpret_cast<float4*>(s.exp_avg) + i) = m;
            *(reinterpret_cast<float4*>(s.exp_avg_sq) + i) = v;
            if (s.shadow_f16 != nullptr) {
                const __half2 lo = __floats2half2_rn(p.x, p.y), hi = __floats2half2_rn(p.z, p.w);
                *(reinterpret_cast<uint2*>(s.shadow_f16) + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
            }
        }
        if (s.zero_grad && s.grad != nullptr) *(reinterpret_cast<float4*>(s.grad) + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (uint64_t i = 4u * n4 + t0; i < s.n; i += stride) {   // unaligned tensors / the tail
        if (!skip) {
            float p = s.param[i], m = s.exp_avg[i], v = s.exp_avg_sq[i];
            const float g = s.grad_f16 != nullptr ? __half2float(reinterpret_cast<const __half*>(s.grad_f16)[i]) : s.grad[i];
            adam_one(p, m, v, g, c, s.decay, s.neg_step_size, s.grad_mul);
            s.param[i] = p; s.exp_avg[i] = m; s.exp_avg_sq[i] = v;
            if (s.shadow_f16 != nullptr) reinterpret_cast<__half*>(s.shadow_f16)[i] = __float2half_rn(p);
        }
        if (s.zero_grad && s.grad != nullptr) s.grad[i] = 0.0f;
    }
}

// One thread: the scalar bookkeeping torch does on the host in double precision (adam.py: bias_correction1 = 1 - beta1 ** step,
// step_size = lr / bias_correction1, bias_correction2_sqrt = bias_correction2 ** 0.5), kept on the device so that a captured graph
// advances the step count by itself.
__global__ void k_adamw_advance(PvdAdamState* state, PvdAdamSlot* slots, uint32_t n_slots) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    PvdAdamState st = *state;
    if (st.found_inf != 0) {
        st.skipped += 1;
    } else {
        st.step += 1;
    }
    const double step = (double)max(st.step, 1);
    const double bc1 = 1.0 - pow(st.beta1, step);
    const double bc2 = 1.0 - pow(st.beta2, step);
    // ATen divides a tensor by a CPU scalar by multiplying with the scalar's reciprocal, formed in DOUBLE and then rounded to fp32
    // (scripts/micro/adam_probe.py: 1.0f / (float)b differs from it by an ulp at most steps, and with it 6 % of the updated parameters)
    st.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    st.w1 = (float)(1.0 - st.beta1);
    st.w2 = (float)(1.0 - st.beta2);
    st.beta2_f = (float)st.beta2;
    *state = st;
    for (uint32_t i = 0; i < n_slots; ++i) {
        const double lr = slots[i].lr;
        slots[i].neg_step_size = (float)(-(lr / bc1));
        slots[i].decay = (float)(1.0 - lr * slots[i].weight_decay);
    }
}

// after the step: found_inf is consumed
__global__ void k_adamw_finish(PvdAdamState* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state->found_inf = 0;
}

__global__ void __launch_bounds__(256) k_grad_nonfinite(const float4* __restrict__ g32, const uint2* __restrict__ g16, uint64_t n4,
                                                        const float* __restrict__ tail32, const __half* __restrict__ tail16, uint32_t n_tail,
                                                        PvdAdamState* state) {
    bool bad = false;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (g16 != nullptr) {
            const uint2 u = __ldg(g16 + i);
            // fp16 non-finite: exponent bits all ones
            bad |= ((u.x & 0x7c00u) == 0x7c00u) || ((u.x & 0x7c000000u) == 0x7c000000u) || ((u.y & 0x7c00u) == 0x7c00u) ||
                   ((u.y & 0x7c000000u) == 0x7c000000u);
        } else {
            const float4 v = __ldg(g32 + i);
            bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < n_tail)
        bad |= g16 != nullptr ? !isfinite(__half2float(tail16[threadIdx.x])) : !isfinite(tail32[threadIdx.x]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(&state->found_inf, 1);
}

// the same check over every slot of an optimizer in one launch (blockIdx.y = slot), on whichever gradient the step will read
__global__ void __launch_bounds__(256) k_grad_nonfinite_multi(PvdAdamState* state, const PvdAdamSlot* __restrict__ slots) {
    const PvdAdamSlot s = slots[blockIdx.y];
    bool bad = false;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s.grad_f16 != nullptr) {
        const __half* g = reinterpret_cast<const __half*>(s.grad_f16);
        const bool vec = (reinterpret_cast<uintptr_t>(g) & 7u) == 0;
        const uint64_t n4 = vec ? s.n / 4u : 0u;
        for (uint64_t i = t0; i < n4; i += stride) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(g) + i);
            bad |= ((u.x & 0x7c00u) == 0x7c00u) || ((u.x & 0x7c000000u) == 0x7c000000u) || ((u.y & 0x7c00u) == 0x7c00u) ||
                   ((u.y & 0x7c000000u) == 0x7c000000u);
        }
        for (uint64_t i = 4u * n4 + t0; i < s.n; i += stride) bad |= !isfinite(__half2float(g[i]));
    } else if (s.grad != nullptr) {
        const bool vec = (reinterpret_cast<uintptr_t>(s.grad) & 15u) == 0;
        const uint64_t n4 = vec ? s.n / 4u : 0u;
        for (uint64_t i = t0; i < n4; i += stride) {
            const float4 v = *(reinterpret_cast<const float4*>(s.grad) + i);
            bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
        }
        for (uint64_t i = 4u * n4 + t0; i < s.n; i += stride) bad |= !isfinite(s.grad[i]);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(&state->found_inf, 1);
}

__global__ void __launch_bounds__(256) k_f16_to_f32(const uint2* __restrict__ src, float4* __restrict__ dst, uint64_t n4, float scale) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 u = __ldg(src + i);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        dst[i] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
    }
}

__global__ void __launch_bounds__(256) k_f32_to_f16_scaled(const float4* __restrict__ src, uint2* __restrict__ dst, uint64_t n4, float scale) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        auto sat = [scale](float x) { return fminf(65504.0f, fmaxf(-65504.0f, x * scale)); };   // NaN propagates, finite never becomes inf
        const __half2 lo = __floats2half2_rn(sat(v.x), sat(v.y)), hi = __floats2half2_rn(sat(v.z), sat(v.w));
        dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}

// several fp32 -> fp16 casts in one launch (blockIdx.y = tensor): the vm model's 12 plane / line tensors -> their fp16 shadows
__global__ void __launch_bounds__(256) k_f32_to_f16_multi(const PvdCastDesc* __restrict__ descs) {
    const PvdCastDesc d = descs[blockIdx.y];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(d.src) & 15u) | (reinterpret_cast<uintptr_t>(d.dst) & 7u)) == 0;
    const uint64_t n4 = vec ? d.n / 4u : 0u;
    for (uint64_t i = t0; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(d.src) + i);
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        reinterpret_cast<uint2*>(d.dst)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    for (uint64_t i = 4u * n4 + t0; i < d.n; i += stride) reinterpret_cast<__half*>(d.dst)[i] = __float2half_rn(d.src[i]);
}

// Warm the L2 with a buffer the next kernels will gather from at random: one cp.async.bulk.prefetch.L2 (TMA prefetch, no destination)
// per 32 KB chunk.  A 42 MB hash table streamed this way costs ~7 us of HBM time on a side stream; gathered cold, the same bytes
// arrive as 1.3 M scattered 32-byte sector misses in the middle of the forward kernel.
__global__ void __launch_bounds__(128) k_l2_prefetch(const uint8_t* __restrict__ p, uint64_t bytes) {
    constexpr uint64_t kChunk = 32768;
    const uint64_t n_chunks = (bytes + kChunk - 1) / kChunk;
    if ((threadIdx.x & 31u) != 0u) return;   // the bulk prefetch takes warp-uniform operands: one lane per warp issues
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t c = warp; c < n_chunks; c += n_warps) {
        const uint64_t off = c * kChunk;
        const uint32_t sz = (uint32_t)min((unsigned long long)kChunk, (unsigned long long)(bytes - off)) & ~15u;
        if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(sz) : "memory");
    }
}

static inline uint32_t stream_grid(uint64_t n_vec, uint32_t per_sm) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (n_vec + 255u) / 256u;
    return (uint32_t)max(1ull, min((unsigned long long)want, (unsigned long long)sms * per_sm));
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_grad_nonfinite(const float* grad, const void* grad_f16, uint64_t n, PvdAdamState* state, void* stream) {
    PVD_REQUIRE(state && (grad || grad_f16));
    if (n == 0) return PVD_OK;
    const bool h = grad_f16 != nullptr;
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(h ? grad_f16 : (const void*)grad) & (h ? 7u : 15u)) == 0);
    const uint64_t n4 = n / 4u;
    const uint32_t tail = (uint32_t)(n - 4u * n4);
    k_grad_nonfinite<<<stream_grid(n4, 8), 256, 0, (cudaStream_t)stream>>>(
        h ? nullptr : reinterpret_cast<const float4*>(grad), h ? reinterpret_cast<const uint2*>(grad_f16) : nullptr, n4,
        h ? nullptr : grad + 4u * n4, h ? reinterpret_cast<const __half*>(grad_f16) + 4u * n4 : nullptr, tail, state);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_grad_nonfinite_slots(PvdAdamState* state, const PvdAdamSlot* slots, uint32_t n_slots, uint64_t max_n, void* stream) {
    PVD_REQUIRE(state && slots);
    if (n_slots == 0) return PVD_OK;
    PVD_REQUIRE(n_slots <= 65535u);
    const dim3 grid(stream_grid((max_n + 3u) / 4u, 8), n_slots, 1);
    k_grad_nonfinite_multi<<<grid, 256, 0, (cudaStream_t)stream>>>(state, slots);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_adamw_advance(PvdAdamState* state, PvdAdamSlot* slots, uint32_t n_slots, void* stream) {
    PVD_REQUIRE(state && (slots || n_slots == 0));
    k_adamw_advance<<<1, 32, 0, (cudaStream_t)stream>>>(state, slots, n_slots);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_adamw_step(const PvdAdamState* state, const PvdAdamSlot* slots, uint32_t n_slots, uint64_t max_n, void* stream) {
    PVD_REQUIRE(state && slots);
    if (n_slots == 0) return PVD_OK;
    PVD_REQUIRE(n_slots <= 65535u);
    const dim3 grid(stream_grid((max_n + 3u) / 4u, 8), n_slots, 1);
    k_adamw_multi<<<grid, 256, 0, (cudaStream_t)stream>>>(state, slots);
    PVD_LAUNCH_CHECK();
    k_adamw_finish<<<1, 32, 0, (cudaStream_t)stream>>>(const_cast<PvdAdamState*>(state));
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_cast_f32_to_f16_multi(const PvdCastDesc* descs_dev, uint32_t n_descs, uint64_t max_n, void* stream) {
    PVD_REQUIRE(descs_dev != nullptr || n_descs == 0);
    if (n_descs == 0) return PVD_OK;
    PVD_REQUIRE(n_descs <= 65535u);
    const dim3 grid(stream_grid((max_n + 3u) / 4u, 8), n_descs, 1);
    k_f32_to_f16_multi<<<grid, 256, 0, (cudaStream_t)stream>>>(descs_dev);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_l2_prefetch(const void* ptr, uint64_t bytes, void* stream) {
    if (bytes == 0) return PVD_OK;
    PVD_REQUIRE(ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0);
    const uint64_t chunks = (bytes + 32767u) / 32768u;
    k_l2_prefetch<<<(uint32_t)min((unsigned long long)((chunks + 3u) / 4u), 148ull), 128, 0, (cudaStream_t)stream>>>((const uint8_t*)ptr, bytes);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_cast_f16_to_f32(const void* src, float* dst, uint64_t elem_count, float scale, void* stream) {
    if (elem_count == 0) return PVD_OK;
    PVD_REQUIRE(src && dst && (elem_count % 4u) == 0);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(src) & 7u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
    const uint64_t n4 = elem_count / 4u;
    k_f16_to_f32<<<stream_grid(n4, 16), 256, 0, (cudaStream_t)stream>>>((const uint2*)src, (float4*)dst, n4, scale);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_cast_f32_to_f16_scaled(const float* src, void* dst, uint64_t elem_count, float scale, void* stream) {
    if (elem_count == 0) return PVD_OK;
    PVD_REQUIRE(src && dst && (elem_count % 4u) == 0);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0);
    const uint64_t n4 = elem_count / 4u;
    k_f32_to_f16_scaled<<<stream_grid(n4, 16), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint2*)dst, n4, scale);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
