// shenc.cu -- stand-alone spherical-harmonics direction encoder behind pvd_sh_encode_forward/backward.
// Replaces shencoder/src/shencoder.cu.  The Jacobian needed for calc_grad_inputs comes from evaluating
// the same basis template on forward-mode dual numbers (shenc.cuh).
#include "shenc.cuh"

namespace pvd {

__global__ void __launch_bounds__(256) k_sh_fwd(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B,
                                               uint32_t D, uint32_t degree, bool calc_grad_inputs,
                                               float* __restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t C2 = degree * degree;
    const float* in = inputs + (size_t)b * D;
    const float x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
    float* out = outputs + (size_t)b * C2;
    if (!calc_grad_inputs) {
        sh_basis<float>(x, y, z, degree, [&](int i, float v) { out[i] = v; });
    } else {
        float* jx = dy_dx + (size_t)b * D * C2;  // [B, 3, C2]  (shencoder.cu:127-129)
        float* jy = jx + C2;
        float* jz = jy + C2;
        const Dual3 X{x, 1.f, 0.f, 0.f}, Y{y, 0.f, 1.f, 0.f}, Z{z, 0.f, 0.f, 1.f};
        sh_basis<Dual3>(X, Y, Z, degree, [&](int i, Dual3 v) {
            out[i] = v.v;
            jx[i] = v.dx;
            jy[i] = v.dy;
            jz[i] = v.dz;
        });
    }
}

// grad_inputs[b,d] += sum_c grad[b,c] * dy_dx[b,d,c]   (shencoder.cu:359-383)
__global__ void __launch_bounds__(256) k_sh_bwd(const float* __restrict__ grad, uint32_t B, uint32_t D, uint32_t degree,
                                               const float* __restrict__ dy_dx, float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = t / D;
    if (b >= B) return;
    const uint32_t d = t - b * D;
    const uint32_t C2 = degree * degree;
    const float* g = grad + (size_t)b * C2;
    const float* j = dy_dx + ((size_t)b * D + d) * C2;
    float acc = 0.0f;
    for (uint32_t c = 0; c < C2; ++c) acc = __fmaf_rn(__ldg(g + c), __ldg(j + c), acc);
    grad_inputs[t] += acc;
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C, int calc_grad_inputs,
                          float* dy_dx, void* stream) {
    if (B == 0) return PVD_OK;
    PVD_REQUIRE(inputs && outputs);
    PVD_REQUIRE(!calc_grad_inputs || dy_dx);
    if (D != 3 || C < 1 || C > 8) return PVD_EUNSUPPORTED;  // sphere_harmonics.py:75-78
    k_sh_fwd<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(inputs, outputs, B, D, C, calc_grad_inputs != 0, dy_dx);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t C, const float* dy_dx,
                           float* grad_inputs, void* stream) {
    (void)inputs;
    if (B == 0) return PVD_OK;
    PVD_REQUIRE(grad && dy_dx && grad_inputs);
    if (D != 3 || C < 1 || C > 8) return PVD_EUNSUPPORTED;
    k_sh_bwd<<<ceil_div(B * D, 256), 256, 0, (cudaStream_t)stream>>>(grad, B, D, C, dy_dx, grad_inputs);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
