// field_vm.cu -- fused "vm" (TensoRF vector-matrix) field query for sm_100a, forward and backward.
//
// Replaces, for model_type "vm", NeRFNetwork.forward of the reference (distill_mutual/network.py:216-309,344-382):
// 12 F.grid_sample launches (3 planes + 3 lines, for 16 sigma and 48 colour components each) with ~20 elementwise kernels,
// basis_mat (144 -> 15), clamps, trunc_exp, SH, color_net and sigmoid -- and the mirror image of all that in autograd.
//
// Layout.  The planes/lines are read IN PLACE as channels-last fp32: a parameter of logical shape [1, R, H, W] stored with
// torch.channels_last strides is [H][W][R] in memory, so one bilinear tap is R contiguous floats (64 B sigma + 192 B colour)
// instead of R loads H*W*4 bytes apart as in the reference's channel-first storage ("VM layout").  The shapes in
// the state_dict do not change.
//
// Work decomposition.
//   gather  : HALF A WARP PER SAMPLE, lanes = components (lanes 0-3 the 16 sigma components, lanes 4-15 the 48 colour ones,
//             four per lane) -> every tap is one coalesced 256-byte request of 128-bit loads; plane x line products go straight
//             into the fp16 operand tile APP[128 x 144] (colour) and a group-reduced scalar (sigma).  4608 B/sample algorithmic.
//   MLP     : thread per sample around tcgen05 GEMMs, exactly as in field_hash.cu: basis_mat (K = 144, 9 MMAs), clamps,
//             exp, SH concat, color_net (3 layers), sigmoid.
//   backward: forward recomputed; weight gradients accumulate in TMEM (basis as three M=64 pieces); d(APP) comes back in
//             three 48-column data-gradient GEMMs; then half a warp per sample re-gathers plane/line values and scatters
//             d(plane) = d(prod) * line, d(line) = d(prod) * plane with red.global.add.v4.f32 into channels-last gradients
//             (coalesced 256-byte reductions).
#include <stdlib.h>
#include "field_common.cuh"
#include "shenc.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

// weight blob (bytes)
constexpr uint32_t kVB = 0;        // basis_mat : 16 rows (15 out + pad) x 144 cols
constexpr uint32_t kVB3 = 4608;    // color_net.0 : 64 x 32
constexpr uint32_t kVB4 = 8704;    // color_net.1 : 64 x 64
constexpr uint32_t kVB5 = 16896;   // color_net.2 : 16 x 64
static_assert(kVB5 + 2048 == PVD_VM_WBLOB_BYTES, "vm blob size");
// TMEM columns
constexpr uint32_t kVD = 0;        // [0,64)
constexpr uint32_t kVD16 = 64;     // [64,80)  basis_mat output
constexpr uint32_t kVD5 = 80;      // [80,96)
constexpr uint32_t kVAW5 = 96;     // [96,112)   dW5^T [64][16]
constexpr uint32_t kVAW4 = 112;    // [112,176)  dW4   [64][64]
constexpr uint32_t kVAW3 = 176;    // [176,208)  dW3   [64][32]
constexpr uint32_t kVAB = 208;     // [208,256)  dBasis^T as three [64][16] pieces (app channels 0-63, 64-127, 128-143)
// gradient workspace (floats)
constexpr uint32_t kVGB = 0;       // [192][16]
constexpr uint32_t kVG3 = 3072;    // [64][32]
constexpr uint32_t kVG4 = 5120;    // [64][64]
constexpr uint32_t kVG5 = 9216;    // [64][16]
static_assert(kVG5 + 1024 == PVD_FIELD_GW_FLOATS, "vm workspace size");

struct VmArgs {
    const void* smat[3];   // planes / lines, channels-last: fp32 (the parameters themselves) or fp16 (a shadow copy: half the
    const void* svec[3];   // gather bytes; the kernels are instantiated for either, template parameter PF16)
    const void* cmat[3];
    const void* cvec[3];
    const uint8_t* wblob;
    uint32_t res[3];
    float aabb[6];
    float clip_min, clip_max, density_scale;
};
struct VmGradPtrs {
    float* smat[3];
    float* svec[3];
    float* cmat[3];
    float* cvec[3];
};

// bilinear footprint of one sample on plane i and line i (grid_sample, align_corners=True, zeros padding)
struct Foot {
    uint32_t pidx[4];  // texel index (y*W + x) of the 4 plane taps
    float pw[4];       // weights (0 for out-of-range taps)
    uint32_t lidx[2];
    float lw[2];
    int32_t pkey, lkey;  // identity of the plane cell (x0, y0) and of the line cell l0 the sample falls in
};

__device__ __forceinline__ void vm_normalise(const float* pos, const float* aabb, float (&xn)[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d)  // 2 * (x - lo) / (hi - lo) - 1   (network.py:345-350)
        xn[d] = __fdiv_rn(2.0f * (pos[d] - aabb[d]), aabb[3 + d] - aabb[d]) - 1.0f;
}

__device__ __forceinline__ void vm_foot(const float (&xn)[3], const uint32_t (&res)[3], int i, Foot& f) {
    // mat_ids = [[0,1],[0,2],[1,2]], vec_ids = [2,1,0]  (network.py:76-77)
    const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2, av = 2 - i;
    const uint32_t W = res[a0], Hh = res[a1], D = res[av];
    const float ix = (xn[a0] + 1.0f) * 0.5f * (float)(W - 1);
    const float iy = (xn[a1] + 1.0f) * 0.5f * (float)(Hh - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    const int x0 = (int)fx, y0 = (int)fy;
    f.pkey = ((y0 + 2) << 16) | ((x0 + 2) & 0xffff);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        const bool in = (xx >= 0) && (xx < (int)W) && (yy >= 0) && (yy < (int)Hh);
        f.pidx[t] = in ? (uint32_t)(yy * (int)W + xx) : 0u;
        f.pw[t] = in ? ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0) : 0.0f;
    }
    const float il = (xn[av] + 1.0f) * 0.5f * (float)(D - 1);
    const float fl = floorf(il);
    const float wl1 = il - fl;
    const int l0 = (int)fl;
    f.lkey = l0 + 2;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int ll = l0 + t;
        const bool in = (ll >= 0) && (ll < (int)D);
        f.lidx[t] = in ? (uint32_t)ll : 0u;
        f.lw[t] = in ? (t ? wl1 : 1.0f - wl1) : 0.0f;
    }
}

// Lane mapping of the gather / scatter: HALF a warp per sample, four consecutive components per lane.  Lane L works on sample
// (L >> 4) of the current pair; its 16-lane group covers the 64 components of a tap: lanes 0-3 the 16 sigma components, lanes 4-15
// the 48 colour ones.  A tap of one sample is then 16 lanes x 16 B = the same contiguous 256 bytes as before, but every load is a
// 128-bit load and every reduction a red.global.add.v4.f32: half the instructions and -- what bounds the backward -- half the
// reduction lane-operations (the SM retires roughly one reduction lane-operation per 1.3 cycles whatever its width).
struct LaneMap {
    bool sig;        // this lane holds sigma components
    uint32_t R;      // components per texel of its tensors (16 / 48)
    uint32_t ch;     // first of its four components
    uint32_t half;   // which sample of the pair
    __device__ __forceinline__ LaneMap() {
        const uint32_t lane = threadIdx.x & 31u, l16 = lane & 15u;
        sig = l16 < 4u;
        R = sig ? 16u : 48u;
        ch = sig ? 4u * l16 : 4u * (l16 - 4u);
        half = lane >> 4;
    }
};

// All 18 taps (3 planes x 4 + 3 lines x 2) of one sample for this lane's four components, and their bilinear weights.  The loads
// are ISSUED here and consumed later: the gather is latency-bound (ncu: 62% of the stall samples of a one-sample-at-a-time loop
// were long-scoreboard waits on these loads), so a warp keeps the taps of two samples (9 KB) in flight.
// 128-bit read-only load that stays where it is written (volatile): all 18 loads of a sample are issued back to back before the
// first one is consumed; left to itself the compiler sinks each plane's loads next to their use (three serial latencies per sample)
__device__ __forceinline__ float4 ldg_v4_issue(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// A tap of four consecutive components starting at element `e` of a plane / line: 16 bytes of fp32, or 8 bytes of the fp16 shadow
// (kept packed in two registers until it is consumed: `Tap` is what stays in flight).
template <bool PF16> struct TapT { using type = float4; };
template <> struct TapT<true> { using type = uint2; };
template <bool PF16>
__device__ __forceinline__ typename TapT<PF16>::type tap_issue(const void* base, size_t e) {
    if constexpr (PF16) {
        uint2 v;
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(reinterpret_cast<const __half*>(base) + e));
        return v;
    } else {
        return ldg_v4_issue(reinterpret_cast<const float*>(base) + e);
    }
}
__device__ __forceinline__ float4 tap_value(const float4& v) { return v; }
__device__ __forceinline__ float4 tap_value(const uint2& v) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

template <bool PF16>
struct SampleTaps {
    typename TapT<PF16>::type t[3][4], u[3][2];
    float pw[3][4], lw[3][2];
};

template <bool PF16>
__device__ __forceinline__ void vm_issue(const VmArgs& a, const float (&pos)[3], const LaneMap& m, SampleTaps<PF16>& T) {
    float xn[3];
    vm_normalise(pos, a.aabb, xn);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Foot f;
        vm_foot(xn, a.res, i, f);
        const void* __restrict__ mat = m.sig ? a.smat[i] : a.cmat[i];
        const void* __restrict__ vec = m.sig ? a.svec[i] : a.cvec[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            T.t[i][k] = tap_issue<PF16>(mat, (size_t)f.pidx[k] * m.R + m.ch);
            T.pw[i][k] = f.pw[k];
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            T.u[i][k] = tap_issue<PF16>(vec, (size_t)f.lidx[k] * m.R + m.ch);
            T.lw[i][k] = f.lw[k];
        }
    }
}

// interpolated plane and line value of pair i (grid_sample's bilinear sum, tap order 00, 01, 10, 11)
template <bool PF16>
__device__ __forceinline__ void vm_interp(const SampleTaps<PF16>& T, int i, float4& pv, float4& lv) {
    pv = make_float4(0.f, 0.f, 0.f, 0.f);
    lv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 t = tap_value(T.t[i][k]);
        pv.x = __fmaf_rn(T.pw[i][k], t.x, pv.x);
        pv.y = __fmaf_rn(T.pw[i][k], t.y, pv.y);
        pv.z = __fmaf_rn(T.pw[i][k], t.z, pv.z);
        pv.w = __fmaf_rn(T.pw[i][k], t.w, pv.w);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 u = tap_value(T.u[i][k]);
        lv.x = __fmaf_rn(T.lw[i][k], u.x, lv.x);
        lv.y = __fmaf_rn(T.lw[i][k], u.y, lv.y);
        lv.z = __fmaf_rn(T.lw[i][k], u.z, lv.z);
        lv.w = __fmaf_rn(T.lw[i][k], u.w, lv.w);
    }
}

// products -> APP tile row r (colour, fp16) and sfeat[r] (sigma, summed over the four sigma lanes of the half-warp)
template <bool PF16>
__device__ __forceinline__ void vm_finish(const SampleTaps<PF16>& T, const LaneMap& m, uint32_t r, uint8_t* APP, float* sfeat) {
    float sacc = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float4 pv, lv;
        vm_interp(T, i, pv, lv);
        const float p0 = pv.x * lv.x, p1 = pv.y * lv.y, p2 = pv.z * lv.z, p3 = pv.w * lv.w;
        if (m.sig) {
            sacc += (p0 + p1) + (p2 + p3);
        } else {
            const uint32_t col = (uint32_t)i * 48u + m.ch;
            const __half2 lo = __floats2half2_rn(p0, p1), hi = __floats2half2_rn(p2, p3);
            *reinterpret_cast<uint2*>(APP + tc5::chunk_off(kTile, r, col >> 3) + (col & 7u) * 2u) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
    sacc = m.sig ? sacc : 0.0f;
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
    if ((threadIdx.x & 15u) == 0u) sfeat[r] = sacc;
}

// Gather phase for the 32 samples of this warp, two at a time: APP tile (colour products, fp16) and sfeat[row] (sigma feature,
// fp32).  Lane l first fetches the position of the warp's sample l; a sample's position is then a shuffle away.
template <bool PF16>
__device__ __forceinline__ void vm_gather(const VmArgs& a, const float* __restrict__ xyzs, uint32_t tile_row0, uint32_t M,
                                          uint8_t* APP, float* sfeat) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const LaneMap m;
    float mine[3] = {0.f, 0.f, 0.f};
    {
        const uint32_t row = tile_row0 + warp * 32 + lane;
        if (row < M) {
#pragma unroll
            for (int d = 0; d < 3; ++d) mine[d] = __ldg(xyzs + 3 * (size_t)row + d);
        }
    }
#pragma unroll 1
    for (uint32_t s = 0; s < 32; s += 2) {
        float pos[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) pos[d] = __shfl_sync(0xffffffffu, mine[d], (int)(s + m.half));
        SampleTaps<PF16> T;
        vm_issue<PF16>(a, pos, m, T);
        vm_finish<PF16>(T, m, warp * 32 + s + m.half, APP, sfeat);
    }
}

struct VmRegs {
    float sf_raw, sfc, rgb[3];
    float cf_raw[15];
};

// basis_mat + color_net for one tile; APP and sfeat are ready.  CIN may alias APP (forward only).
__device__ __forceinline__ void vm_mlp_forward(Pipe& p, const VmArgs& a, uint8_t* smw, uint8_t* APP, uint8_t* CIN, uint8_t* H3,
                                               uint8_t* H4, const float* sfeat, const float* dir, uint32_t row, float& sigma,
                                               float (&feat)[16], VmRegs& r) {
    const uint32_t tid = threadIdx.x;
    const uint32_t trow = tc5::tmem_addr(p.tmem, (tid >> 5) * 32, 0);
    const uint32_t sw = tc5::smem_u32(smw);
    operands_ready();
    if (tid == 0) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kVD16, tc5::smem_u32(APP), 144, sw + kVB, 16, 16);  // colour features: [128 x 144] x [16 x 144]^T
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    float cf[16];
    tc5::tmem_ld16(trow + kVD16, cf);
    r.sf_raw = sfeat[row];
    r.sfc = clampf(r.sf_raw, a.clip_min, a.clip_max);                         // network.py:353-358
    feat[0] = r.sfc;
#pragma unroll
    for (int i = 0; i < 15; ++i) {
        r.cf_raw[i] = __half2float(__float2half_rn(cf[i]));                   // basis_mat runs in fp16 under autocast
        feat[i + 1] = clampf(r.cf_raw[i], a.clip_min, a.clip_max);            // network.py:359-361
    }
    sigma = a.density_scale * __expf(r.sfc);
    {
        float sh[16], geo[16];
        sh_basis4(dir[0], dir[1], dir[2], sh);
#pragma unroll
        for (int i = 0; i < 15; ++i) geo[i] = feat[i + 1];
        geo[15] = 0.0f;
        // CIN may alias APP: every thread has finished reading TMEM, and the basis MMA has completed (waited above)
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 0)) = tc5::pack8(sh);
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 1)) = tc5::pack8(sh + 8);
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 2)) = tc5::pack8(geo);
        *reinterpret_cast<uint4*>(CIN + tc5::chunk_off(kTile, row, 3)) = tc5::pack8(geo + 8);
    }
    operands_ready();
    if (tid == 0) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kVD, tc5::smem_u32(CIN), 32, sw + kVB3, 64, 64);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    relu_to_tile<4>(trow + kVD, H3, row);
    operands_ready();
    if (tid == 0) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kVD, tc5::smem_u32(H3), 64, sw + kVB4, 64, 64);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    relu_to_tile<4>(trow + kVD, H4, row);
    operands_ready();
    if (tid == 0) {
        tc5::fence_after_sync();
        issue_fwd(p.tmem + kVD5, tc5::smem_u32(H4), 64, sw + kVB5, 16, 16);
        tc5::mma_commit(p.bar);
    }
    mma_wait(p);
    float c16[16];
    tc5::tmem_ld16(trow + kVD5, c16);
#pragma unroll
    for (int i = 0; i < 3; ++i) r.rgb[i] = 1.0f / (1.0f + __expf(-c16[i]));
}

// =============================================================================================== forward
template <bool PF16>
__global__ void __launch_bounds__(128, 4) k_vm_field_fwd(VmArgs a, const float* __restrict__ xyzs, const float* __restrict__ dirs,
                                                      uint32_t M, float* __restrict__ sigmas, float* __restrict__ rgbs,
                                                      float* __restrict__ feat16, int32_t* status) {
    extern __shared__ __align__(128) uint8_t smem[];   // no-swizzle operand tiles need 16 B; 1024 would pad the static part by 1.5 KB and cost the 4th CTA
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float sfeat[kTile];
    uint8_t* smw = smem;                          // 18944
    // APP is dead once the basis GEMM has completed: CIN aliases its first 8 KB and the H3 / H4 activation tile the next 16 KB, so the
    // CTA needs 54.5 KB and FOUR of them fit an SM (592 slots for the 576 tiles of a 4096-ray batch: one wave instead of 1.3)
    uint8_t* APP = smem + PVD_VM_WBLOB_BYTES;     // 36864 ; CIN aliases its first 8 KB
    uint8_t* H = APP + 8192;                      // 16384 ; H3 then H4, inside the dead APP tile
    const uint32_t tid = threadIdx.x;
    // TMEM first: the SM does not launch the next CTA of a tcgen05 kernel until the previous one has relinquished its allocation
    // permit (measured: scripts/micro/cta_launch.cu), so anything placed before the alloc delays every later CTA of the SM.
    if (tid < 32) tc5::tmem_alloc(&tmem_base_s, 128);
    stage_blob(smw, a.wblob, PVD_VM_WBLOB_BYTES);
    if (tid == 0) {
        tc5::mbar_init(&bar, 1);
        tc5::mbar_fence_init();
    }
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    Pipe p{&bar, 0u, tmem_base_s, status};
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();  // previous tile's sfeat / APP consumers are done
        vm_gather<PF16>(a, xyzs, tile * kTile, M, APP, sfeat);
        const uint32_t row = tile * kTile + tid;
        const bool live = row < M;
        float dir[3] = {0.f, 0.f, 0.f};
        if (live) {
#pragma unroll
            for (int d = 0; d < 3; ++d) dir[d] = __ldg(dirs + 3 * (size_t)row + d);
        }
        float sigma, feat[16];
        VmRegs r;
        vm_mlp_forward(p, a, smw, APP, APP, H, H, sfeat, dir, tid, sigma, feat, r);
        if (live) {
            sigmas[row] = sigma;
            rgbs[3 * (size_t)row] = r.rgb[0];
            rgbs[3 * (size_t)row + 1] = r.rgb[1];
            rgbs[3 * (size_t)row + 2] = r.rgb[2];
            if (feat16) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(feat16 + 16 * (size_t)row + 4 * q) =
                        make_float4(feat[4 * q], feat[4 * q + 1], feat[4 * q + 2], feat[4 * q + 3]);
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc5::tmem_dealloc(p.tmem, 128);
}

// =============================================================================================== backward
template <bool PF16>
__global__ void __launch_bounds__(128, 2) k_vm_field_bwd(VmArgs a, VmGradPtrs g, const float* __restrict__ xyzs,
                                                      const float* __restrict__ dirs, const float* __restrict__ grad_sigmas,
                                                      const float* __restrict__ grad_rgbs, const float* __restrict__ grad_feat,
                                                      uint32_t M, const int32_t* __restrict__ n_valid_p, float* __restrict__ gw,
                                                      int32_t* status, uint32_t diag_skip, uint8_t* __restrict__ dapp_ws,
                                                      float* __restrict__ dsf_ws) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float sfeat[kTile];
    __shared__ float dsf[kTile];
    uint8_t* smw = smem;                          // 18944
    uint8_t* APP = smem + PVD_VM_WBLOB_BYTES;     // 36864  (later d(APP))
    uint8_t* CIN = APP + 36864;                   // 8192
    uint8_t* H3 = CIN + 8192;                     // 16384
    uint8_t* H4 = H3 + 16384;                     // 16384
    uint8_t* G16 = H4 + 16384;                    // 4096
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    // TMEM first: the SM does not launch the next CTA of a tcgen05 kernel until the previous one has relinquished its allocation
    // permit (measured: scripts/micro/cta_launch.cu), so anything placed before the alloc delays every later CTA of the SM.
    if (tid < 32) tc5::tmem_alloc(&tmem_base_s, 256);
    stage_blob(smw, a.wblob, PVD_VM_WBLOB_BYTES);
    if (tid == 0) {
        tc5::mbar_init(&bar, 1);
        tc5::mbar_fence_init();
    }
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    Pipe p{&bar, 0u, tmem_base_s, status};
    const uint32_t trow = tc5::tmem_addr(p.tmem, warp * 32, 0);
    const uint32_t sw = tc5::smem_u32(smw);
    const uint32_t n_valid = n_valid_p ? min((uint32_t)max(*n_valid_p, 0), M) : M;
    const uint32_t n_tiles = (n_valid + kTile - 1) / kTile;
    bool first = true;
    bool pdl_pending = true;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        vm_gather<PF16>(a, xyzs, tile * kTile, M, APP, sfeat);
        const uint32_t row = tile * kTile + tid;
        const bool live = row < n_valid;
        float dir[3] = {0.f, 0.f, 0.f}, gsig = 0.0f, grgb[3] = {0.f, 0.f, 0.f};
        auto load_grads = [&]() {   // what the preceding loss kernel wrote
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; ++d) grgb[d] = __ldg(grad_rgbs + 3 * (size_t)row + d);
                gsig = __ldg(grad_sigmas + row);
            }
        };
        if (live) {
#pragma unroll
            for (int d = 0; d < 3; ++d) dir[d] = __ldg(dirs + 3 * (size_t)row + d);
        }
        // programmatic dependent launch (see k_hash_field_bwd): the first tile's gather + forward recomputation run under the loss
        // kernel; the upstream gradients are read after pdl_wait()
        if (!pdl_pending) load_grads();
        float sigma, feat[16];
        VmRegs r;
        vm_mlp_forward(p, a, smw, APP, CIN, H3, H4, sfeat, dir, tid, sigma, feat, r);
        if (pdl_pending) {
            pdl_wait();
            pdl_pending = false;
            load_grads();
        }
        // ---- colour net backward (same sequence as the hash field)
        {
            float gq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 3; ++i) gq[i] = grgb[i] * r.rgb[i] * (1.0f - r.rgb[i]);
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 0)) = tc5::pack8(gq);
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 1)) = make_uint4(0, 0, 0, 0);
        }
        operands_ready();
        if (tid == 0) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kVAW5, tc5::smem_u32(H4), tc5::smem_u32(G16), 16, first);
            issue_dgrad(p.tmem + kVD, tc5::smem_u32(G16), 16, sw + kVB5, 16, 64);
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        mask_grad_in_place(trow + kVD, H4, tid);
        operands_ready();
        if (tid == 0) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kVAW4, tc5::smem_u32(H4), tc5::smem_u32(H3), 64, first);
            issue_dgrad(p.tmem + kVD, tc5::smem_u32(H4), 64, sw + kVB4, 64, 64);
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        mask_grad_in_place(trow + kVD, H3, tid);
        operands_ready();
        if (tid == 0) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kVAW3, tc5::smem_u32(H3), tc5::smem_u32(CIN), 32, first);
            issue_dgrad(p.tmem + kVD, tc5::smem_u32(H3), 64, sw + kVB3, 64, 32);
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        // ---- d(colour features) through the clamp, d(sigma feature) through trunc_exp + clamp
        {
            float dc[16], gq[16], gf[16];
            tc5::tmem_ld16(trow + kVD + 16, dc);
#pragma unroll
            for (int i = 0; i < 16; ++i) gf[i] = 0.0f;
            if (grad_feat && live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(grad_feat + 16 * (size_t)row) + q);
                    gf[4 * q] = v.x; gf[4 * q + 1] = v.y; gf[4 * q + 2] = v.z; gf[4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 15; ++i) {
                const bool in = (r.cf_raw[i] >= a.clip_min) && (r.cf_raw[i] <= a.clip_max);
                gq[i] = in ? dc[i] + gf[i + 1] : 0.0f;
            }
            gq[15] = 0.0f;
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 0)) = tc5::pack8(gq);
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 1)) = tc5::pack8(gq + 8);
            const bool sin = (r.sf_raw >= a.clip_min) && (r.sf_raw <= a.clip_max);
            dsf[tid] = (sin && live) ? gsig * a.density_scale * __expf(clampf(r.sfc, -12.0f, 12.0f)) + gf[0] : 0.0f;
        }
        operands_ready();
        if (tid == 0) {
            tc5::fence_after_sync();
            // dBasis^T [144 x 16] += APP^T G as three M=64 pieces (the last covers channels 128..191; rows >= 144 are never flushed)
            const uint32_t idesc = tc5::instr_desc_f16(64, 16, 1, 1);
            for (uint32_t piece = 0; piece < 3; ++piece)
                for (uint32_t s0 = 0; s0 < kTile; s0 += 16)
                    tc5::mma_f16_ss(p.tmem + kVAB + 16 * piece, tc5::desc_mnmajor(tc5::smem_u32(APP), kTile, s0, 64 * piece),
                                    tc5::desc_mnmajor(tc5::smem_u32(G16), kTile, s0, 0), idesc, !(first && s0 == 0));
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        first = false;
        // d(APP) = G [128 x 16] x basis [16 x 144], 48 columns at a time, written (fp16) over APP
        for (uint32_t piece = 0; piece < 3; ++piece) {
            operands_ready();
            if (tid == 0) {
                tc5::fence_after_sync();
                const uint32_t idesc = tc5::instr_desc_f16(128, 48, 0, 1);
                tc5::mma_f16_ss(p.tmem + kVD, tc5::desc_kmajor(tc5::smem_u32(G16), kTile, 0),
                                tc5::desc_mnmajor(sw + kVB, 16, 0, 48 * piece), idesc, 0u);
                tc5::mma_commit(p.bar);
            }
            mma_wait(p);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[16];
                tc5::tmem_ld16(trow + kVD + 16 * c, v);
                *reinterpret_cast<uint4*>(APP + tc5::chunk_off(kTile, tid, 6 * piece + 2 * c)) = tc5::pack8(v);
                *reinterpret_cast<uint4*>(APP + tc5::chunk_off(kTile, tid, 6 * piece + 2 * c + 1)) = tc5::pack8(v + 8);
            }
        }
        if (dapp_ws != nullptr) {
            // two-kernel backward: d(APP) (fp16, the tile's chunk layout, 36 KB per tile) and d(sigma feature) go to a workspace;
            // k_vm_scatter reduces them into the plane / line gradients at four times this kernel's occupancy
            uint8_t* dst = dapp_ws + (size_t)tile * 36864u;
#pragma unroll
            for (uint32_t j = 0; j < 18; ++j)
                *reinterpret_cast<uint4*>(dst + tc5::chunk_off(kTile, tid, j)) = *reinterpret_cast<const uint4*>(APP + tc5::chunk_off(kTile, tid, j));
            dsf_ws[(size_t)tile * kTile + tid] = dsf[tid];
            continue;   // the loop head's __syncthreads() orders the next gather after these reads
        }
        __syncthreads();  // d(APP) and dsf complete in shared memory
        // ---- scatter: half a warp per RUN of 16 consecutive samples, four components per lane (LaneMap).  Consecutive samples of a
        //      ray are dt = 2 sqrt(3) / 1024 apart, half a texel of a 300^2 plane: they fall into the same plane cell / line cell about
        //      every other step, so their contributions are summed in registers while the cell does not change and leave the SM as
        //      ONE red.global.add.v4.f32 per tap and run -- the backward is bound by reduction lane-operations (~1.3 cycles each per SM).
        {
            const LaneMap m;
            const uint32_t base = tile * kTile + warp * 32 + 16u * m.half;   // first sample of this half-warp's run
            const uint32_t n_s = ((diag_skip & 1u) == 0u && base < n_valid) ? min(16u, n_valid - base) : 0u;
            const uint32_t n_it = __reduce_max_sync(0xffffffffu, n_s);       // warp-uniform trip count (loads use shuffles)
            const uint32_t l16 = lane & 15u;
            float mine[3] = {0.f, 0.f, 0.f};
            if (l16 < n_s) {
#pragma unroll
                for (int d = 0; d < 3; ++d) mine[d] = __ldg(xyzs + 3 * (size_t)(base + l16) + d);
            }
            float4 accp[3][4], accl[3][2];
            int32_t keyp[3] = {-1, -1, -1}, keyl[3] = {-1, -1, -1};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) accp[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 2; ++k) accl[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            auto red4 = [](float* dst, const float4& v) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            };
            auto flush_plane = [&](int i) {
                if (keyp[i] < 0) return;
                const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2;
                const int W = (int)a.res[a0], Hh = (int)a.res[a1];
                const int x0 = (keyp[i] & 0xffff) - 2, y0 = (keyp[i] >> 16) - 2;
                float* gm = m.sig ? g.smat[i] : g.cmat[i];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
                    if (xx >= 0 && xx < W && yy >= 0 && yy < Hh) red4(gm + (size_t)(yy * W + xx) * m.R + m.ch, accp[i][k]);
                    accp[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto flush_line = [&](int i) {
                if (keyl[i] < 0) return;
                const int D = (int)a.res[2 - i];
                const int l0 = keyl[i] - 2;
                float* gv = m.sig ? g.svec[i] : g.cvec[i];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int ll = l0 + k;
                    if (ll >= 0 && ll < D) red4(gv + (size_t)ll * m.R + m.ch, accl[i][k]);
                    accl[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
#pragma unroll 1
            for (uint32_t s = 0; s < n_it; ++s) {
                float pos[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) pos[d] = __shfl_sync(0xffffffffu, mine[d], (int)(16u * m.half + (s & 15u)));
                SampleTaps<PF16> T;
                vm_issue<PF16>(a, pos, m, T);
                if (s >= n_s) continue;   // this half-warp's run is shorter (tail of the valid rows)
                const uint32_t rr = warp * 32 + 16u * m.half + s;
                float xn[3];
                vm_normalise(pos, a.aabb, xn);
                const float ds = dsf[rr];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Foot f;
                    vm_foot(xn, a.res, i, f);   // weights and cell identity again (cheap ALU; T carries only the tap values)
                    float4 pv, lv;
                    vm_interp<PF16>(T, i, pv, lv);
                    float4 dp;
                    if (m.sig) {
                        dp = make_float4(ds, ds, ds, ds);
                    } else {
                        const uint32_t col = (uint32_t)i * 48u + m.ch;
                        const uint2 raw = *reinterpret_cast<const uint2*>(APP + tc5::chunk_off(kTile, rr, col >> 3) + (col & 7u) * 2u);
                        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                        dp = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                    const float4 dpv = make_float4(dp.x * lv.x, dp.y * lv.y, dp.z * lv.z, dp.w * lv.w);
                    const float4 dlv = make_float4(dp.x * pv.x, dp.y * pv.y, dp.z * pv.z, dp.w * pv.w);
                    if (f.pkey != keyp[i]) {
                        flush_plane(i);
                        keyp[i] = f.pkey;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        accp[i][k].x = __fmaf_rn(f.pw[k], dpv.x, accp[i][k].x);
                        accp[i][k].y = __fmaf_rn(f.pw[k], dpv.y, accp[i][k].y);
                        accp[i][k].z = __fmaf_rn(f.pw[k], dpv.z, accp[i][k].z);
                        accp[i][k].w = __fmaf_rn(f.pw[k], dpv.w, accp[i][k].w);
                    }
                    if (f.lkey != keyl[i]) {
                        flush_line(i);
                        keyl[i] = f.lkey;
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        accl[i][k].x = __fmaf_rn(f.lw[k], dlv.x, accl[i][k].x);
                        accl[i][k].y = __fmaf_rn(f.lw[k], dlv.y, accl[i][k].y);
                        accl[i][k].z = __fmaf_rn(f.lw[k], dlv.z, accl[i][k].z);
                        accl[i][k].w = __fmaf_rn(f.lw[k], dlv.w, accl[i][k].w);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                flush_plane(i);
                flush_line(i);
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    gw += (size_t)(blockIdx.x % PVD_FIELD_GW_COPIES) * PVD_FIELD_GW_FLOATS;
    if (!first) {
        flush_acc(p.tmem, kVAW3, 32, gw + kVG3);
        flush_acc(p.tmem, kVAW4, 64, gw + kVG4);
        flush_acc(p.tmem, kVAW5, 16, gw + kVG5);
        flush_acc(p.tmem, kVAB, 16, gw + kVGB);
        flush_acc(p.tmem, kVAB + 16, 16, gw + kVGB + 64 * 16);
        flush_acc(p.tmem, kVAB + 32, 16, gw + kVGB + 128 * 16, 16);
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc5::tmem_dealloc(p.tmem, 256);
}

// =============================================================================================== stand-alone gradient scatter
// Second kernel of the two-kernel backward: reads d(APP) / d(sigma feature) from the workspace, re-gathers the plane / line values
// and reduces into the channels-last gradients.  Half a warp walks a run of 16 consecutive samples (LaneMap), the (plane, line)
// pair is the OUTER loop, so only 6 taps and 6 accumulators are live at a time: ~90 registers, 256-thread CTAs, 16-24 warps per
// SM instead of the 8 the MLP kernel can hold (212 registers, 100 KB of shared memory) -- this phase is bound by the latency of
// its own loads and reductions, not by a chip-level unit (its 6 M sector reductions are fewer than k_hash_scatter's 8 M).
template <bool PF16>
__global__ void __launch_bounds__(256, 2) k_vm_scatter(VmArgs a, VmGradPtrs g, const float* __restrict__ xyzs,
                                                       const uint8_t* __restrict__ dapp_ws, const float* __restrict__ dsf_ws, uint32_t M,
                                                       const int32_t* __restrict__ n_valid_p) {
    const uint32_t n_valid = n_valid_p ? min((uint32_t)max(*n_valid_p, 0), M) : M;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, l16 = lane & 15u;
    const LaneMap m;
    const uint32_t base = (blockIdx.x * 8u + warp) * 32u + 16u * m.half;   // first sample of this half-warp's run
    const uint32_t n_s = (base < n_valid) ? min(16u, n_valid - base) : 0u;
    const uint32_t n_it = __reduce_max_sync(0xffffffffu, n_s);
    if (n_it == 0) return;
    float mine[3] = {0.f, 0.f, 0.f};
    float ds_mine = 0.0f;
    if (l16 < n_s) {
#pragma unroll
        for (int d = 0; d < 3; ++d) mine[d] = __ldg(xyzs + 3 * (size_t)(base + l16) + d);
        ds_mine = __ldg(dsf_ws + base + l16);
    }
    auto red4 = [](float* dst, const float4& v) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    };
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // unrolled: a runtime index into the kernel-parameter arrays would move them to local memory
        const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2;
        const int W = (int)a.res[a0], Hh = (int)a.res[a1], D = (int)a.res[2 - i];
        const void* __restrict__ mat = m.sig ? a.smat[i] : a.cmat[i];
        const void* __restrict__ vec = m.sig ? a.svec[i] : a.cvec[i];
        float* gm = m.sig ? g.smat[i] : g.cmat[i];
        float* gv = m.sig ? g.svec[i] : g.cvec[i];
        float4 accp[4], accl[2];
        int32_t keyp = -1, keyl = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) accp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) accl[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        auto flush_plane = [&]() {
            if (keyp < 0) return;
            const int x0 = (keyp & 0xffff) - 2, y0 = (keyp >> 16) - 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
                if (xx >= 0 && xx < W && yy >= 0 && yy < Hh) red4(gm + (size_t)(yy * W + xx) * m.R + m.ch, accp[k]);
                accp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto flush_line = [&]() {
            if (keyl < 0) return;
            const int l0 = keyl - 2;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ll = l0 + k;
                if (ll >= 0 && ll < D) red4(gv + (size_t)ll * m.R + m.ch, accl[k]);
                accl[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // software pipeline over the run: the taps and the d(APP) slice of sample s+1 are in flight under the arithmetic of sample s
        typename TapT<PF16>::type t[2][4], u[2][2];
        Foot f[2];
        uint2 draw[2];
        auto issue = [&](uint32_t s2, int b) {
            float pos[3], xn[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) pos[d] = __shfl_sync(0xffffffffu, mine[d], (int)(16u * m.half + (s2 & 15u)));
            vm_normalise(pos, a.aabb, xn);
            vm_foot(xn, a.res, i, f[b]);
#pragma unroll
            for (int k = 0; k < 4; ++k) t[b][k] = tap_issue<PF16>(mat, (size_t)f[b].pidx[k] * m.R + m.ch);
#pragma unroll
            for (int k = 0; k < 2; ++k) u[b][k] = tap_issue<PF16>(vec, (size_t)f[b].lidx[k] * m.R + m.ch);
            draw[b] = make_uint2(0u, 0u);
            if (!m.sig && s2 < n_s) {
                const uint32_t row = base + s2, col = (uint32_t)i * 48u + m.ch;
                draw[b] = __ldg(reinterpret_cast<const uint2*>(dapp_ws + (size_t)(row / kTile) * 36864u +
                                                               tc5::chunk_off(kTile, row % kTile, col >> 3) + (col & 7u) * 2u));
            }
        };
        issue(0, 0);
#pragma unroll 1
        for (uint32_t s = 0; s < n_it; s += 2) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const uint32_t sc = s + (uint32_t)b;
                if (sc >= n_it) break;                       // warp-uniform
                if (sc + 1 < n_it) issue(sc + 1, b ^ 1);
                const float ds = __shfl_sync(0xffffffffu, ds_mine, (int)(16u * m.half + (sc & 15u)));
                if (sc >= n_s) continue;                     // this half-warp's run is shorter
                float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), lv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 tv = tap_value(t[b][k]);
                    pv.x = __fmaf_rn(f[b].pw[k], tv.x, pv.x); pv.y = __fmaf_rn(f[b].pw[k], tv.y, pv.y);
                    pv.z = __fmaf_rn(f[b].pw[k], tv.z, pv.z); pv.w = __fmaf_rn(f[b].pw[k], tv.w, pv.w);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float4 uv = tap_value(u[b][k]);
                    lv.x = __fmaf_rn(f[b].lw[k], uv.x, lv.x); lv.y = __fmaf_rn(f[b].lw[k], uv.y, lv.y);
                    lv.z = __fmaf_rn(f[b].lw[k], uv.z, lv.z); lv.w = __fmaf_rn(f[b].lw[k], uv.w, lv.w);
                }
                float4 dp;
                if (m.sig) {
                    dp = make_float4(ds, ds, ds, ds);
                } else {
                    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&draw[b].x));
                    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&draw[b].y));
                    dp = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
                const float4 dpv = make_float4(dp.x * lv.x, dp.y * lv.y, dp.z * lv.z, dp.w * lv.w);
                const float4 dlv = make_float4(dp.x * pv.x, dp.y * pv.y, dp.z * pv.z, dp.w * pv.w);
                if (f[b].pkey != keyp) {
                    flush_plane();
                    keyp = f[b].pkey;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    accp[k].x = __fmaf_rn(f[b].pw[k], dpv.x, accp[k].x); accp[k].y = __fmaf_rn(f[b].pw[k], dpv.y, accp[k].y);
                    accp[k].z = __fmaf_rn(f[b].pw[k], dpv.z, accp[k].z); accp[k].w = __fmaf_rn(f[b].pw[k], dpv.w, accp[k].w);
                }
                if (f[b].lkey != keyl) {
                    flush_line();
                    keyl = f[b].lkey;
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    accl[k].x = __fmaf_rn(f[b].lw[k], dlv.x, accl[k].x); accl[k].y = __fmaf_rn(f[b].lw[k], dlv.y, accl[k].y);
                    accl[k].z = __fmaf_rn(f[b].lw[k], dlv.z, accl[k].z); accl[k].w = __fmaf_rn(f[b].lw[k], dlv.w, accl[k].w);
                }
            }
        }
        flush_plane();
        flush_line();
    }
}

__global__ void k_vm_pack_weights(const float* __restrict__ basis, const float* __restrict__ wc0, const float* __restrict__ wc1,
                                  const float* __restrict__ wc2, uint8_t* __restrict__ blob) {
    pack_matrix(basis, 15, 144, blob + kVB, 16, 144);
    pack_matrix(wc0, 64, 31, blob + kVB3, 64, 32);
    pack_matrix(wc1, 64, 64, blob + kVB4, 64, 64);
    pack_matrix(wc2, 3, 64, blob + kVB5, 16, 64);
}

__global__ void k_vm_unpack_wgrads(const float* __restrict__ gw, float* __restrict__ gb, float* __restrict__ g0,
                                   float* __restrict__ g1, float* __restrict__ g2) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    auto sum = [&](uint32_t i) {
        float acc = 0.0f;
#pragma unroll
        for (uint32_t c = 0; c < PVD_FIELD_GW_COPIES; ++c) acc += gw[(size_t)c * PVD_FIELD_GW_FLOATS + i];
        return acc;
    };
    if (t < 15 * 144) { const uint32_t o = t / 144, i = t - o * 144; gb[t] += sum(kVGB + i * 16 + o); }
    if (t < 64 * 31) { const uint32_t o = t / 31, i = t - o * 31; g0[t] += sum(kVG3 + o * 32 + i); }
    if (t < 64 * 64) { g1[t] += sum(kVG4 + t); }
    if (t < 3 * 64) { const uint32_t o = t / 64, i = t - o * 64; g2[t] += sum(kVG5 + i * 16 + o); }
}

static bool vm_planes_f16(const PvdVmField* f) { return f->plane_dtype == PVD_DTYPE_F16; }

static bool to_vm_args(const PvdVmField* f, VmArgs& a) {
    if (f->plane_dtype != PVD_DTYPE_F16 && f->plane_dtype != PVD_DTYPE_F32) return false;
    for (int i = 0; i < 3; ++i) {
        a.smat[i] = f->sigma_mat[i]; a.svec[i] = f->sigma_vec[i]; a.cmat[i] = f->color_mat[i]; a.cvec[i] = f->color_vec[i];
        a.res[i] = f->res[i];
        if (!a.smat[i] || !a.svec[i] || !a.cmat[i] || !a.cvec[i] || a.res[i] < 2) return false;
    }
    for (int i = 0; i < 6; ++i) a.aabb[i] = f->aabb[i];
    a.wblob = reinterpret_cast<const uint8_t*>(f->wblob);
    a.clip_min = f->sigma_clip_min; a.clip_max = f->sigma_clip_max; a.density_scale = f->density_scale;
    return a.wblob != nullptr;
}

constexpr size_t kVmFwdSmem = PVD_VM_WBLOB_BYTES + 36864;                               // 55808
constexpr size_t kVmBwdSmem = PVD_VM_WBLOB_BYTES + 36864 + 8192 + 16384 + 16384 + 4096; // 100864

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_vm_pack_weights(const float* basis_mat, const float* w_color0, const float* w_color1, const float* w_color2, void* wblob,
                        void* stream) {
    PVD_REQUIRE(basis_mat && w_color0 && w_color1 && w_color2 && wblob);
    k_vm_pack_weights<<<16, 256, 0, (cudaStream_t)stream>>>(basis_mat, w_color0, w_color1, w_color2, (uint8_t*)wblob);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_vm_field_forward(const PvdVmField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                         float* feat16, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && xyzs && dirs && sigmas && rgbs && status);
    VmArgs a;
    if (!to_vm_args(f, a)) return PVD_EINVAL;
    const uint32_t tiles = (M + kTile - 1) / kTile;
    const uint32_t grid = min(tiles, (uint32_t)(4 * sm_count()));
    if (vm_planes_f16(f)) {
        cudaError_t e = cudaFuncSetAttribute(k_vm_field_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVmFwdSmem);
        if (e != cudaSuccess) return (int)e;
        k_vm_field_fwd<true><<<grid, 128, kVmFwdSmem, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, feat16, status);
    } else {
        cudaError_t e = cudaFuncSetAttribute(k_vm_field_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVmFwdSmem);
        if (e != cudaSuccess) return (int)e;
        k_vm_field_fwd<false><<<grid, 128, kVmFwdSmem, (cudaStream_t)stream>>>(a, xyzs, dirs, M, sigmas, rgbs, feat16, status);
    }
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

static int vm_backward(const PvdVmField* f, const PvdVmGrads* grads, const float* xyzs, const float* dirs, const float* grad_sigmas,
                       const float* grad_rgbs, const float* grad_feat16, uint32_t M, const int32_t* n_valid, float* gw_ws,
                       void* scatter_ws, int32_t* status, void* stream);

int pvd_vm_field_backward(const PvdVmField* f, const PvdVmGrads* grads, const float* xyzs, const float* dirs,
                          const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                          const int32_t* n_valid, float* gw_ws, int32_t* status, void* stream) {
    return vm_backward(f, grads, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat16, M, n_valid, gw_ws, nullptr, status, stream);
}

uint64_t pvd_vm_backward_workspace_bytes(uint32_t M) {
    const uint64_t tiles = ((uint64_t)M + kTile - 1) / kTile;
    return tiles * 36864u + tiles * kTile * sizeof(float);
}

int pvd_vm_field_backward_ws(const PvdVmField* f, const PvdVmGrads* grads, const float* xyzs, const float* dirs,
                             const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                             const int32_t* n_valid, float* gw_ws, void* scatter_ws, int32_t* status, void* stream) {
    PVD_REQUIRE(scatter_ws != nullptr && (reinterpret_cast<uintptr_t>(scatter_ws) & 15u) == 0);
    return vm_backward(f, grads, xyzs, dirs, grad_sigmas, grad_rgbs, grad_feat16, M, n_valid, gw_ws, scatter_ws, status, stream);
}

static int vm_backward(const PvdVmField* f, const PvdVmGrads* grads, const float* xyzs, const float* dirs, const float* grad_sigmas,
                       const float* grad_rgbs, const float* grad_feat16, uint32_t M, const int32_t* n_valid, float* gw_ws,
                       void* scatter_ws, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && grads && xyzs && dirs && grad_sigmas && grad_rgbs && gw_ws && status);
    VmArgs a;
    if (!to_vm_args(f, a)) return PVD_EINVAL;
    VmGradPtrs g;
    for (int i = 0; i < 3; ++i) {
        g.smat[i] = grads->sigma_mat[i]; g.svec[i] = grads->sigma_vec[i]; g.cmat[i] = grads->color_mat[i]; g.cvec[i] = grads->color_vec[i];
        PVD_REQUIRE(g.smat[i] && g.svec[i] && g.cmat[i] && g.cvec[i]);
    }
    const uint32_t tiles = (M + kTile - 1) / kTile;
    const uint32_t grid = min(tiles, (uint32_t)(2 * sm_count()));
    const bool pf16 = vm_planes_f16(f);
    cudaError_t e = pf16 ? cudaFuncSetAttribute(k_vm_field_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVmBwdSmem)
                         : cudaFuncSetAttribute(k_vm_field_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVmBwdSmem);
    if (e != cudaSuccess) return (int)e;
    // PVD_VM_DIAG_SKIP=1 (timing diagnostics only, wrong gradients): leave out the plane / line gradient scatter
    static const uint32_t diag_skip = []() { const char* v = getenv("PVD_VM_DIAG_SKIP"); return v ? (uint32_t)atoi(v) : 0u; }();
    uint8_t* dapp_ws = reinterpret_cast<uint8_t*>(scatter_ws);
    float* dsf_ws = scatter_ws ? reinterpret_cast<float*>(dapp_ws + (size_t)tiles * 36864u) : nullptr;
    e = pf16 ? launch_pdl(k_vm_field_bwd<true>, dim3(grid), dim3(128), kVmBwdSmem, (cudaStream_t)stream, a, g, xyzs, dirs, grad_sigmas, grad_rgbs,
                          grad_feat16, M, n_valid, gw_ws, status, diag_skip, dapp_ws, dsf_ws)
             : launch_pdl(k_vm_field_bwd<false>, dim3(grid), dim3(128), kVmBwdSmem, (cudaStream_t)stream, a, g, xyzs, dirs, grad_sigmas, grad_rgbs,
                          grad_feat16, M, n_valid, gw_ws, status, diag_skip, dapp_ws, dsf_ws);
    if (e != cudaSuccess) return (int)e;
    PVD_LAUNCH_CHECK();
    if (scatter_ws != nullptr && !(diag_skip & 1u)) {
        if (pf16)
            k_vm_scatter<true><<<ceil_div(M, 256u), 256, 0, (cudaStream_t)stream>>>(a, g, xyzs, dapp_ws, dsf_ws, M, n_valid);
        else
            k_vm_scatter<false><<<ceil_div(M, 256u), 256, 0, (cudaStream_t)stream>>>(a, g, xyzs, dapp_ws, dsf_ws, M, n_valid);
        PVD_LAUNCH_CHECK();
    }
    return PVD_OK;
}

int pvd_vm_unpack_wgrads(const float* gw_ws, float* g_basis, float* gw_color0, float* gw_color1, float* gw_color2, void* stream) {
    PVD_REQUIRE(gw_ws && g_basis && gw_color0 && gw_color1 && gw_color2);
    k_vm_unpack_wgrads<<<16, 256, 0, (cudaStream_t)stream>>>(gw_ws, g_basis, gw_color0, gw_color1, gw_color2);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
