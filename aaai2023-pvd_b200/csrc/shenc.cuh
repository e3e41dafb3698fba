// shenc.cuh -- real spherical-harmonics basis (degree 1..8 => 1..64 coefficients), written once as a
// template over the scalar type: instantiated with float it is the forward encoder; instantiated with
// the forward-mode dual number `Dual3` it yields d(basis)/d(x,y,z) without 192 hand-written derivative
// polynomials (the reference spells all of them out, shencoder/src/shencoder.cu:131-351).
//
// Basis convention and constants follow shencoder/src/shencoder.cu:50-121 (the tiny-cuda-nn ordering).
#pragma once
#include "common.cuh"

namespace pvd {

struct Dual3 {
    float v, dx, dy, dz;
};
__host__ __device__ __forceinline__ Dual3 operator+(Dual3 a, Dual3 b) { return {a.v + b.v, a.dx + b.dx, a.dy + b.dy, a.dz + b.dz}; }
__host__ __device__ __forceinline__ Dual3 operator-(Dual3 a, Dual3 b) { return {a.v - b.v, a.dx - b.dx, a.dy - b.dy, a.dz - b.dz}; }
__host__ __device__ __forceinline__ Dual3 operator-(Dual3 a) { return {-a.v, -a.dx, -a.dy, -a.dz}; }
__host__ __device__ __forceinline__ Dual3 operator*(Dual3 a, Dual3 b) {
    return {a.v * b.v, a.dx * b.v + a.v * b.dx, a.dy * b.v + a.v * b.dy, a.dz * b.v + a.v * b.dz};
}
__host__ __device__ __forceinline__ Dual3 operator*(float s, Dual3 a) { return {s * a.v, s * a.dx, s * a.dy, s * a.dz}; }
__host__ __device__ __forceinline__ Dual3 operator*(Dual3 a, float s) { return s * a; }
__host__ __device__ __forceinline__ Dual3 operator+(Dual3 a, float s) { return {a.v + s, a.dx, a.dy, a.dz}; }
__host__ __device__ __forceinline__ Dual3 operator+(float s, Dual3 a) { return a + s; }
__host__ __device__ __forceinline__ Dual3 operator-(Dual3 a, float s) { return {a.v - s, a.dx, a.dy, a.dz}; }
__host__ __device__ __forceinline__ Dual3 operator-(float s, Dual3 a) { return {s - a.v, -a.dx, -a.dy, -a.dz}; }

__host__ __device__ __forceinline__ float sh_const(float, float c) { return c; }
__host__ __device__ __forceinline__ Dual3 sh_const(Dual3, float c) { return {c, 0.f, 0.f, 0.f}; }

// out[0 .. degree*degree) = Y_lm(x, y, z).  `Out` is any callable out(index, value).
template <typename T, typename Out>
__host__ __device__ __forceinline__ void sh_basis(T x, T y, T z, uint32_t degree, Out&& out) {
    out(0, sh_const(x, 0.28209479177387814f));
    if (degree <= 1) return;
    // l = 1
    out(1, -0.48860251190291987f * y);
    out(2, 0.48860251190291987f * z);
    out(3, -0.48860251190291987f * x);
    if (degree <= 2) return;
    const T xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    // l = 2
    out(4, 1.0925484305920792f * xy);
    out(5, -1.0925484305920792f * yz);
    out(6, 0.94617469575755997f * z2 - 0.31539156525251999f);
    out(7, -1.0925484305920792f * xz);
    out(8, 0.54627421529603959f * x2 - 0.54627421529603959f * y2);
    if (degree <= 3) return;
    // l = 3
    out(9, 0.59004358992664352f * y * (-3.0f * x2 + y2));
    out(10, 2.8906114426405538f * xy * z);
    out(11, 0.45704579946446572f * y * (1.0f - 5.0f * z2));
    out(12, 0.3731763325901154f * z * (5.0f * z2 - 3.0f));
    out(13, 0.45704579946446572f * x * (1.0f - 5.0f * z2));
    out(14, 1.4453057213202769f * z * (x2 - y2));
    out(15, 0.59004358992664352f * x * (-x2 + 3.0f * y2));
    if (degree <= 4) return;
    const T x4 = x2 * x2, y4 = y2 * y2, z4 = z2 * z2;
    // l = 4
    out(16, 2.5033429417967046f * xy * (x2 - y2));
    out(17, 1.7701307697799304f * yz * (-3.0f * x2 + y2));
    out(18, 0.94617469575756008f * xy * (7.0f * z2 - 1.0f));
    out(19, 0.66904654355728921f * yz * (3.0f - 7.0f * z2));
    out(20, -3.1735664074561294f * z2 + 3.7024941420321507f * z4 + 0.31735664074561293f);
    out(21, 0.66904654355728921f * xz * (3.0f - 7.0f * z2));
    out(22, 0.47308734787878004f * (x2 - y2) * (7.0f * z2 - 1.0f));
    out(23, 1.7701307697799304f * xz * (-x2 + 3.0f * y2));
    out(24, -3.7550144126950569f * x2 * y2 + 0.62583573544917614f * x4 + 0.62583573544917614f * y4);
    if (degree <= 5) return;
    // l = 5
    out(25, 0.65638205684017015f * y * (10.0f * x2 * y2 - 5.0f * x4 - y4));
    out(26, 8.3026492595241645f * xy * z * (x2 - y2));
    out(27, -0.48923829943525038f * y * (3.0f * x2 - y2) * (9.0f * z2 - 1.0f));
    out(28, 4.7935367849733241f * xy * z * (3.0f * z2 - 1.0f));
    out(29, 0.45294665119569694f * y * (14.0f * z2 - 21.0f * z4 - 1.0f));
    out(30, 0.1169503224534236f * z * (-70.0f * z2 + 63.0f * z4 + 15.0f));
    out(31, 0.45294665119569694f * x * (14.0f * z2 - 21.0f * z4 - 1.0f));
    out(32, 2.3967683924866621f * z * (x2 - y2) * (3.0f * z2 - 1.0f));
    out(33, -0.48923829943525038f * x * (x2 - 3.0f * y2) * (9.0f * z2 - 1.0f));
    out(34, 2.0756623148810411f * z * (-6.0f * x2 * y2 + x4 + y4));
    out(35, 0.65638205684017015f * x * (10.0f * x2 * y2 - x4 - 5.0f * y4));
    if (degree <= 6) return;
    const T x6 = x4 * x2, y6 = y4 * y2, z6 = z4 * z2;
    // l = 6
    out(36, 1.3663682103838286f * xy * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4));
    out(37, 2.3666191622317521f * yz * (10.0f * x2 * y2 - 5.0f * x4 - y4));
    out(38, 2.0182596029148963f * xy * (x2 - y2) * (11.0f * z2 - 1.0f));
    out(39, -0.92120525951492349f * yz * (3.0f * x2 - y2) * (11.0f * z2 - 3.0f));
    out(40, 0.92120525951492349f * xy * (-18.0f * z2 + 33.0f * z4 + 1.0f));
    out(41, 0.58262136251873131f * yz * (30.0f * z2 - 33.0f * z4 - 5.0f));
    out(42, 6.6747662381009842f * z2 - 20.024298714302954f * z4 + 14.684485723822165f * z6 - 0.31784601133814211f);
    out(43, 0.58262136251873131f * xz * (30.0f * z2 - 33.0f * z4 - 5.0f));
    out(44, 0.46060262975746175f * (x2 - y2) * (11.0f * z2 * (3.0f * z2 - 1.0f) - 7.0f * z2 + 1.0f));
    out(45, -0.92120525951492349f * xz * (x2 - 3.0f * y2) * (11.0f * z2 - 3.0f));
    out(46, 0.50456490072872406f * (11.0f * z2 - 1.0f) * (-6.0f * x2 * y2 + x4 + y4));
    out(47, 2.3666191622317521f * xz * (10.0f * x2 * y2 - x4 - 5.0f * y4));
    out(48, 10.247761577878714f * x2 * y4 - 10.247761577878714f * x4 * y2 + 0.6831841051919143f * x6 - 0.6831841051919143f * y6);
    if (degree <= 7) return;
    // l = 7
    out(49, 0.70716273252459627f * y * (-21.0f * x2 * y4 + 35.0f * x4 * y2 - 7.0f * x6 + y6));
    out(50, 5.2919213236038001f * xy * z * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4));
    out(51, -0.51891557872026028f * y * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + 5.0f * x4 + y4));
    out(52, 4.1513246297620823f * xy * z * (x2 - y2) * (13.0f * z2 - 3.0f));
    out(53, -0.15645893386229404f * y * (3.0f * x2 - y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f));
    out(54, 0.44253269244498261f * xy * z * (-110.0f * z2 + 143.0f * z4 + 15.0f));
    out(55, 0.090331607582517306f * y * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f));
    out(56, 0.068284276912004949f * z * (315.0f * z2 - 693.0f * z4 + 429.0f * z6 - 35.0f));
    out(57, 0.090331607582517306f * x * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f));
    out(58, 0.07375544874083044f * z * (x2 - y2) * (143.0f * z2 * (3.0f * z2 - 1.0f) - 187.0f * z2 + 45.0f));
    out(59, -0.15645893386229404f * x * (x2 - 3.0f * y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f));
    out(60, 1.0378311574405206f * z * (13.0f * z2 - 3.0f) * (-6.0f * x2 * y2 + x4 + y4));
    out(61, -0.51891557872026028f * x * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + x4 + 5.0f * y4));
    out(62, 2.6459606618019f * z * (15.0f * x2 * y4 - 15.0f * x4 * y2 + x6 - y6));
    out(63, 0.70716273252459627f * x * (-35.0f * x2 * y4 + 21.0f * x4 * y2 - x6 + 7.0f * y6));
}

// degree-4 specialisation used by the fused field kernels: 16 coefficients into a register array
__device__ __forceinline__ void sh_basis4(float x, float y, float z, float (&o)[16]) {
    sh_basis<float>(x, y, z, 4u, [&](int i, float v) { o[i] = v; });
}

}  // namespace pvd
