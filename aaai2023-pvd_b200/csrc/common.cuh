// common.cuh -- shared device helpers for the PVD B200 hot path (sm_100a).
//
// Arithmetic that decides INTEGER results (cell indices, sample counts) is written with explicit
// round-to-nearest intrinsics (__fmaf_rn, __fmul_rn, __fadd_rn, __fdiv_rn) so that the result does
// not depend on nvcc's FMA-contraction choices; the CPU oracle (oracle/pvd_oracle.c) mirrors each
// of them with fmaf()/plain IEEE ops.  The comments cite the reference expression being restated.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <float.h>

#include "../../include/pvd_b200.h"

#define PVD_LAUNCH_CHECK()                         \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define PVD_REQUIRE(cond)                 \
    do {                                  \
        if (!(cond)) return PVD_EINVAL;   \
    } while (0)

// ---- diagnostic timeline (tools/trace build only: -DPVD_TRACE; the shipped library compiles these to nothing) --------------
#ifdef PVD_TRACE
#define PVD_TRACE_TU(setter)                                                                   \
    static __device__ unsigned long long* g_trace_buf;                                         \
    static __device__ unsigned int g_trace_cap;                                                \
    extern "C" int setter(unsigned long long* buf, unsigned int records) {                     \
        cudaError_t e = cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf));                    \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_trace_cap, &records, sizeof(records));  \
        return (int)e;                                                                         \
    }
#define PVD_T(rec, slot)                                                                       \
    do {                                                                                       \
        if (g_trace_buf != nullptr && (unsigned int)(rec) < g_trace_cap) g_trace_buf[(size_t)(rec) * 16 + (slot)] = clock64(); \
    } while (0)
#define PVD_TV(rec, slot, val)                                                                 \
    do {                                                                                       \
        if (g_trace_buf != nullptr && (unsigned int)(rec) < g_trace_cap) g_trace_buf[(size_t)(rec) * 16 + (slot)] = (unsigned long long)(val); \
    } while (0)
#else
#define PVD_TRACE_TU(setter)
#define PVD_T(rec, slot) do { } while (0)
#define PVD_TV(rec, slot, val) do { } while (0)
#endif

namespace pvd {

// ---- programmatic dependent launch (griddepcontrol): a kernel launched with launch_pdl() may start -- prologue, anything that does not
// read its predecessor's output -- as soon as every CTA of the predecessor has called pdl_launch_dependents() (or exited); it must call
// pdl_wait() before touching what the predecessor writes (that returns once the predecessor grid has completed and its memory is
// visible).  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
// PVD_PDL=0 turns the attribute off (every launch a full dependency): the A/B switch of DESIGN.md's measurement
inline bool pdl_enabled() {
    static const bool on = []() { const char* v = getenv("PVD_PDL"); return !(v != nullptr && v[0] == '0'); }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

constexpr float kSqrt3 = 1.7320508075688772f;  // raymarching.cu:21
constexpr float kRPi = 0.3183098861837907f;    // raymarching.cu:24

__host__ __device__ __forceinline__ uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float clampf(float x, float lo, float hi) {  // raymarching.cu:36-38
    return fminf(hi, fmaxf(lo, x));
}

// ---- Morton codes, 10 bits per axis (raymarching.cu:58-83) ---------------------------------
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xC30C30C3u;
    x = (x | (x >> 4)) & 0x0F00F00Fu;
    x = (x | (x >> 8)) & 0xFF0000FFu;
    x = (x | (x >> 16)) & 0x0000FFFFu;
    return x;
}

// ---- frexpf exponent without the libm call ---------------------------------------------------
// frexpf(m, &e) returns e with m = f * 2^e, f in [0.5,1); e = 0 for m == 0 (raymarching.cu:44-56).
__device__ __forceinline__ int frexp_exponent(float m) {
    if (m == 0.0f) return 0;
    uint32_t u = __float_as_uint(m);
    int be = (int)((u >> 23) & 0xFFu);
    if (be == 0) {  // subnormal: normalise
        int lz = __clz(u << 9);
        return -126 - lz;
    }
    if (be == 255) return 0;  // inf/nan: unspecified by the standard; pick 0
    return be - 126;
}

// mip level of a point / of a step size (raymarching.cu:44-56). C = cascade count.
__device__ __forceinline__ int mip_from_pos(float x, float y, float z, uint32_t C) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const int e = frexp_exponent(mx);
    return min((int)C - 1, max(0, e));
}
__device__ __forceinline__ int mip_from_dt(float dt, float Hf, uint32_t C) {
    // reference: dt * H * 0.5 (the 0.5 is a double literal, the product is rounded back to float;
    // a multiplication by 0.5 is exact, so float arithmetic gives the same value).
    const float mx = __fmul_rn(__fmul_rn(dt, Hf), 0.5f);
    const int e = frexp_exponent(mx);
    return min((int)C - 1, max(0, e));
}

// ---- PCG32 (same generator the reference vendors: raymarching/src/pcg32.h:57-72,107-116,149-170)
constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dULL;

struct Pcg32 {
    uint64_t state;
    uint64_t inc;
};

// seed(initstate, initseq=1): state=0; inc=(seq<<1)|1; step; state+=initstate; step.
__host__ __device__ __forceinline__ Pcg32 pcg32_seeded(uint64_t initstate, uint64_t initseq = 1u) {
    Pcg32 r;
    r.inc = (initseq << 1u) | 1u;
    r.state = 0u;
    r.state = r.state * kPcgMult + r.inc;
    r.state += initstate;
    r.state = r.state * kPcgMult + r.inc;
    return r;
}

// jump ahead by `delta` steps: O(log delta) LCG exponentiation (pcg32.h:149-170).
__host__ __device__ __forceinline__ void pcg32_advance(Pcg32& r, uint64_t delta) {
    uint64_t cur_mult = kPcgMult, cur_plus = r.inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    r.state = acc_mult * r.state + acc_plus;
}

__host__ __device__ __forceinline__ uint32_t pcg32_next_uint(Pcg32& r) {
    const uint64_t old = r.state;
    r.state = old * kPcgMult + r.inc;
    const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}

// uniform in [0,1): mantissa trick (pcg32.h:107-116)
__device__ __forceinline__ float pcg32_next_float(Pcg32& r) {
    const uint32_t u = (pcg32_next_uint(r) >> 9) | 0x3f800000u;
    return __uint_as_float(u) - 1.0f;
}

// ---- occupancy-grid marching step, shared by the training and inference marchers ---------------
// One evaluation of the loop body of raymarching.cu:362-403 at parameter t.
struct MarchCtx {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, rbound, dt_gamma, dt_min, dt_max, rH, Hf, Hm1f;
    float halfH, rH2;      // 0.5*H and 2/H: exact power-of-two rescalings of H and 1/H
    int incx, incy, incz;  // 1 where the ray moves towards +axis (sign bit of d clear), else 0
    uint32_t C, H, H3;
};

__device__ __forceinline__ void march_ctx_init(MarchCtx& c, const float* o, const float* d, float bound,
                                               float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    c.ox = o[0]; c.oy = o[1]; c.oz = o[2];
    c.dx = d[0]; c.dy = d[1]; c.dz = d[2];
    c.rdx = __fdiv_rn(1.0f, c.dx); c.rdy = __fdiv_rn(1.0f, c.dy); c.rdz = __fdiv_rn(1.0f, c.dz);
    c.bound = bound; c.rbound = __fdiv_rn(1.0f, bound); c.dt_gamma = dt_gamma;
    c.Hf = (float)H; c.Hm1f = (float)(H - 1); c.rH = __fdiv_rn(1.0f, c.Hf);
    // dt_min = 2*SQRT3/max_steps ; dt_max = 2*SQRT3*(1<<(C-1))/H   (raymarching.cu:346-347)
    const float two_s3 = __fmul_rn(2.0f, kSqrt3);
    c.dt_min = __fdiv_rn(two_s3, (float)max_steps);
    c.dt_max = __fdiv_rn(__fmul_rn(two_s3, (float)(1u << (C - 1))), c.Hf);
    c.halfH = 0.5f * c.Hf; c.rH2 = 2.0f * c.rH;
    c.incx = (int)((__float_as_uint(c.dx) >> 31) ^ 1u);
    c.incy = (int)((__float_as_uint(c.dy) >> 31) ^ 1u);
    c.incz = (int)((__float_as_uint(c.dz) >> 31) ^ 1u);
    c.C = C; c.H = H; c.H3 = H * H * H;
}

__device__ __forceinline__ float march_dt(const MarchCtx& c, float t) {
    return clampf(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max);
}

// position on the ray, clamped to the scene cube (raymarching.cu:364-366; nvcc contracts o + t*d to FMA)
__device__ __forceinline__ void march_pos(const MarchCtx& c, float t, float& x, float& y, float& z) {
    x = clampf(__fmaf_rn(t, c.dx, c.ox), -c.bound, c.bound);
    y = clampf(__fmaf_rn(t, c.dy, c.oy), -c.bound, c.bound);
    z = clampf(__fmaf_rn(t, c.dz, c.oz), -c.bound, c.bound);
}

// cell coordinate along one axis: (int)clamp(0.5*(x*rb+1)*H, 0, H-1)   (raymarching.cu:377-379).
// The reference evaluates 0.5*(..)*H in double and rounds to float when calling clamp(); with the
// inner FMA done in float, the double product of a float by 0.5 and by an integer-valued H is exact,
// so a single correctly-rounded float multiply gives the identical float.
// 0.5*a is exact, so fl(fl(0.5*a)*H) = fl(a*(0.5*H)): one multiply by the precomputed 0.5*H.
__device__ __forceinline__ int march_cell(float x, float mip_rbound, float halfH, float Hm1f) {
    const float a = __fmaf_rn(x, mip_rbound, 1.0f);
    const float p = __fmul_rn(a, halfH);
    return (int)clampf(p, 0.0f, Hm1f);
}

// distance to the exit of the current (empty) cell along one axis (raymarching.cu:393-395)
// n + 0.5 + 0.5*sign(d) is the exact integer n + inc (inc = 1 for d >= +0, 0 for d <= -0); fl(a*rH)*2 = fl(a*(2*rH)).
__device__ __forceinline__ float march_exit(int n, int inc, float rd, float x, float rH2, float mip_bound) {
    const float a = (float)(n + inc);
    const float q = __fadd_rn(__fmul_rn(a, rH2), -1.0f);        // a*rH*2 - 1
    return __fmul_rn(__fmaf_rn(q, mip_bound, -x), rd);          // (q*mb - x) * rd
}

// The occupancy cell of a sample, split from the load of its bit so that callers can put several loads in flight.
struct MarchCell {
    int nx, ny, nz;
    float mip_bound;
    uint32_t index;  // bit index into the density bitfield (level * H^3 + morton)
};

__device__ __forceinline__ MarchCell march_locate(const MarchCtx& c, float dt, float x, float y, float z) {
    MarchCell m;
    // level in [0, C-1] (raymarching.cu:371); a single cascade needs no exponent arithmetic at all
    const int level = (c.C > 1) ? max(mip_from_pos(x, y, z, c.C), mip_from_dt(dt, c.Hf, c.C)) : 0;
    // mip_bound = min(2^level, bound), mip_rbound = 1 / mip_bound (:373-374).  The IEEE quotient 1 / 2^level is exactly
    // 2^-level, so the division is only ever needed for 1 / bound, which is hoisted into the context.
    const float p2 = (float)(1u << level);
    const bool pow2 = p2 <= c.bound;
    m.mip_bound = pow2 ? p2 : c.bound;
    const float mip_rbound = pow2 ? __uint_as_float((uint32_t)(127 - level) << 23) : c.rbound;
    m.nx = march_cell(x, mip_rbound, c.halfH, c.Hm1f);
    m.ny = march_cell(y, mip_rbound, c.halfH, c.Hm1f);
    m.nz = march_cell(z, mip_rbound, c.halfH, c.Hm1f);
    m.index = (uint32_t)level * c.H3 + morton3((uint32_t)m.nx, (uint32_t)m.ny, (uint32_t)m.nz);
    return m;
}

// parameter at which the ray leaves the cell (tt of raymarching.cu:397)
__device__ __forceinline__ float march_leave(const MarchCtx& c, const MarchCell& m, float t, float x, float y, float z) {
    const float tx = march_exit(m.nx, c.incx, c.rdx, x, c.rH2, m.mip_bound);
    const float ty = march_exit(m.ny, c.incy, c.rdy, y, c.rH2, m.mip_bound);
    const float tz = march_exit(m.nz, c.incz, c.rdz, z, c.rH2, m.mip_bound);
    return __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
}

// Returns true if the sample at t is in an occupied cell. When it is not, `t_skip` receives the
// parameter at which the ray leaves the cell.
__device__ __forceinline__ bool march_probe(const MarchCtx& c, const uint8_t* __restrict__ grid, float t,
                                            float dt, float x, float y, float z, float& t_skip) {
    const MarchCell m = march_locate(c, dt, x, y, z);
    const bool occ = (__ldg(grid + (m.index >> 3)) >> (m.index & 7u)) & 1u;
    if (!occ) t_skip = march_leave(c, m, t, x, y, z);
    return occ;
}

// warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inclusive scans over the warp
__device__ __forceinline__ float warp_scan_mul(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v *= u;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v += u;
    }
    return v;
}

}  // namespace pvd
