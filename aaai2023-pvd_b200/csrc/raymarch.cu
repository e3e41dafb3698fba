// raymarch.cu -- ray/AABB intersection, occupancy-grid ray marching and alpha compositing for sm_100a.
//
// Replaces the reference's raymarching extension (raymarching/src/raymarching.cu) behind the C ABI in
// include/pvd_b200.h.  Design differences (B200-first, not a translation):
//   * march_rays_train is ONE occupancy march per ray (the reference marches every ray twice), a
//     stash of the accepted (t, dt) pairs, a block-scan that hands out DETERMINISTIC sample offsets in
//     ray-id order (the reference uses atomicAdd arrival order), and a warp-per-ray expansion that
//     writes xyzs/dirs/deltas with lane-contiguous stores.
//   * composite_rays_train fwd/bwd run one WARP per ray: the transmittance product and the running
//     colour / depth sums are warp shuffle scans over 32 samples at a time instead of a serial loop
//     per thread, so sample reads are coalesced across the warp.
//   * every launch goes to the caller's stream and reports launch errors.
#include <type_traits>

#include "common.cuh"
#include "../../include/pvd_b200_fused.h"
PVD_TRACE_TU(pvd_debug_trace_raymarch)

namespace pvd {

// =============================================================================================
// utilities
// =============================================================================================

// ray / axis-aligned box slab test (raymarching.cu:94-147)
__device__ __forceinline__ void near_far_one(const float* __restrict__ o, const float* __restrict__ d,
                                             const float* __restrict__ aabb, float min_near, float& near_out,
                                             float& far_out) {
    const float ox = o[0], oy = o[1], oz = o[2];
    const float rdx = __fdiv_rn(1.0f, d[0]), rdy = __fdiv_rn(1.0f, d[1]), rdz = __fdiv_rn(1.0f, d[2]);
    float near = __fmul_rn(aabb[0] - ox, rdx), far = __fmul_rn(aabb[3] - ox, rdx);
    if (near > far) { float s = near; near = far; far = s; }
    float near_y = __fmul_rn(aabb[1] - oy, rdy), far_y = __fmul_rn(aabb[4] - oy, rdy);
    if (near_y > far_y) { float s = near_y; near_y = far_y; far_y = s; }
    if (near > far_y || near_y > far) { near_out = far_out = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(aabb[2] - oz, rdz), far_z = __fmul_rn(aabb[5] - oz, rdz);
    if (near_z > far_z) { float s = near_z; near_z = far_z; far_z = s; }
    if (near > far_z || near_z > far) { near_out = far_out = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_out = near;
    far_out = far;
}

__global__ void k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                           const float* __restrict__ aabb, uint32_t N, float min_near, float* __restrict__ nears,
                           float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float a, b;
    near_far_one(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, aabb, min_near, a, b);
    nears[n] = a;
    fars[n] = b;
}

// background-sphere polar coordinates (raymarching.cu:165-200)
__global__ void k_polar(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                        uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* o = rays_o + 3 * (size_t)n;
    const float* d = rays_d + 3 * (size_t)n;
    const float ox = o[0], oy = o[1], oz = o[2], dx = d[0], dy = d[1], dz = d[2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float Cq = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * Cq)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[2 * (size_t)n] = 2 * theta * kRPi - 1;
    coords[2 * (size_t)n + 1] = phi * kRPi;
}

__global__ void k_morton3D(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t* c = coords + 3 * (size_t)n;
    indices[n] = (int32_t)morton3((uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2]);
}

__global__ void k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t ind = indices[n];  // signed shifts, as raymarching.cu:251-255
    int32_t* c = coords + 3 * (size_t)n;
    c[0] = (int32_t)compact3((uint32_t)(ind >> 0));
    c[1] = (int32_t)compact3((uint32_t)(ind >> 1));
    c[2] = (int32_t)compact3((uint32_t)(ind >> 2));
}

// 8 density cells -> one byte; two float4 loads per thread (raymarching.cu:270-291)
__global__ void k_packbits(const float* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
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

// =============================================================================================
// training march: count+stash -> scan -> expand
// =============================================================================================

// Coarse rejection mask (single cascade, H = 128 -- the reference's only grid size, renderer.py:36).  The H^3 occupancy
// bitfield is Morton ordered, so an aligned block of 2x2x2 cells is exactly one BYTE of it; any[b] = "some cell of block b is
// occupied" on the 64^3 block lattice, and the mask bit of a block is the OR of any[] over the block and its 26 neighbours
// (dilation by one block).  A lattice point of a ray lying within 0.9 block (Euclidean) of a probe point whose block is NOT
// masked sits in a cell of that probe's 3x3x3 block neighbourhood, i.e. in an empty cell (see k_march_count).
// Both arrays: 64^3 bits indexed x + 64 y + 4096 z, 8192 words.
constexpr uint32_t kCoarseB = 64;                 // blocks per axis
constexpr uint32_t kCoarseMaskWords = 64 * 64 * 64 / 32;

__global__ void __launch_bounds__(256) k_coarse_any(const uint8_t* __restrict__ grid, uint32_t* __restrict__ any) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;  // output word: 32 consecutive x at fixed (y, z)
    if (w >= kCoarseMaskWords) return;
    const uint32_t x0 = (w & 1u) * 32u, y = (w >> 1) & 63u, z = w >> 7;
    const uint32_t myz = (spread3(y) << 1) | (spread3(z) << 2);
    uint32_t word = 0;
#pragma unroll 8
    for (uint32_t i = 0; i < 32; ++i) word |= (__ldg(grid + (spread3(x0 + i) | myz)) ? 1u : 0u) << i;
    any[w] = word;
}

__global__ void __launch_bounds__(256) k_coarse_dilate(const uint32_t* __restrict__ any, uint32_t* __restrict__ mask) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= kCoarseMaskWords) return;
    const int half = (int)(w & 1u), y = (int)((w >> 1) & 63u), z = (int)(w >> 7);
    uint64_t acc = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy, zz = z + dz;
            if (yy < 0 || yy > 63 || zz < 0 || zz > 63) continue;
            const uint32_t r = (uint32_t)(yy * 2 + zz * 128);
            const uint64_t row = (uint64_t)__ldg(any + r) | ((uint64_t)__ldg(any + r + 1) << 32);
            acc |= row | (row << 1) | (row >> 1);
        }
    mask[w] = half ? (uint32_t)(acc >> 32) : (uint32_t)acc;
}

// Pass 1 -- ONE WARP PER RAY.  The reference walks each ray with one thread (raymarching.cu:357-403): a chain of several
// hundred dependent iterations, each with a scattered byte load, at one warp per SM.  Here the 32 lanes of a warp test 32
// CONSECUTIVE lattice points of the same ray at once and the serial skipping logic is replayed on ballots, which gives
// 4096 resident warps instead of 128 and 32 independent occupancy loads in flight per ray.
//
// Why this is exact.  The marcher's parameter only ever advances by `t += clamp(t*dt_gamma, dt_min, dt_max)`, in the
// occupied branch (:389) and in the skip loop (:400) alike, so the visited parameters are a subset of one fixed lattice
// t_0 = t0, t_{k+1} = fl(t_k + dtf(t_k)) that does not depend on the grid.  A lattice point is "landed" if the serial loop
// evaluates it: point 0 is; after an occupied landed point k, k+1 is; after an empty landed point k with cell exit tt_k,
// the first j > k with t_j >= tt_k is.  Each lane evaluates its lattice point exactly as the serial loop would
// (same position, cell, level, exit expressions), and the landed set is then resolved by pointer doubling over the successor
// map (see `close_chains` / `resolve` below).  One warp per CTA makes every branch provably warp-uniform for the compiler
// (no WARPSYNC / re-convergence code around the shuffles).  Lattice values are
// produced either by the 32-step serial recurrence (every lane runs it redundantly and keeps its own element) or, when
// dt is constant and the window stays inside one binade, from the observation that fl(t + dt) advances the bit pattern of
// t by a constant number of ulps c (round-to-nearest of dt/ulp; ties are detected and sent to the serial path), so
// bits(t_{k0+i}) = bits(t_{k0}) + i*c.
// The accepted (t, dt) pairs go to the stash in march order; num_steps[n] and t0[n] to the workspace.
__global__ void __launch_bounds__(32) k_march_count(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                    const uint8_t* __restrict__ grid, float bound, float dt_gamma,
                                                    uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                                    const float* __restrict__ nears, const float* __restrict__ fars,
                                                    uint32_t perturb, Pcg32 rng, int32_t* __restrict__ num_steps_out,
                                                    float* __restrict__ t0_out, float2* __restrict__ stash,
                                                    const uint32_t* __restrict__ coarse, const float* __restrict__ aabb,
                                                    float min_near, float* __restrict__ nears_out, float* __restrict__ fars_out) {
    const uint32_t n = blockIdx.x;  // one warp per CTA: the ray index, and with it all control flow below, is warp-uniform
    const uint32_t lane = threadIdx.x;
    if (n >= N) return;
#ifdef PVD_TRACE
    unsigned int tr_groups = 0, tr_general = 0, tr_iters = 0;
    long long tr_c0 = 0, tr_probe = 0, tr_prep = 0, tr_res = 0, tr_gen = 0, tr_head = 0;
    if (lane == 0) {
        unsigned long long gt; unsigned int smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        PVD_T(n, 0); PVD_TV(n, 8, gt); PVD_TV(n, 7, smid);
    }
#endif
    MarchCtx c;
    march_ctx_init(c, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H);
    float far, t0;
    if (aabb != nullptr) {  // fused near_far_from_aabb (raymarching.cu:94-147): saves a launch and the nears/fars round trip
        near_far_one(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, aabb, min_near, t0, far);
        if (lane == 0) {
            nears_out[n] = t0;
            fars_out[n] = far;
        }
    } else {
        far = fars[n];
        t0 = nears[n];
    }
    if (perturb) {  // raymarching.cu:351-354: the same jitter every call (seed 42, advance(n))
        pcg32_advance(rng, (uint64_t)n);
        t0 = __fmaf_rn(c.dt_min, pcg32_next_float(rng), t0);
    }
    // ---- exact pruning with the dilated 64^3 block mask: probe [t0, far] every 0.9 block.  Every parameter t the marcher can
    // evaluate is within half a probe spacing of a probe, and a point within 0.9 block of a probe has its cell in the probe's
    // 3x3x3 block neighbourhood (clamping to the cube moves both the same way), so around an UNMASKED probe every cell is empty.
    //   * no probe masked  -> no evaluated cell is occupied: zero samples (most rays of a batch);
    //   * beyond the last masked probe (+ 3/4 spacing) nothing can be emitted any more, so the march may stop there: `far` is
    //     pulled in.  (The part BEFORE the first masked probe cannot be skipped: which lattice points are evaluated later
    //     depends on the chain of cell exits from t0 on.)
    if (coarse != nullptr && t0 > 0.0f && t0 < far) {
        const float dlen = sqrtf(c.dx * c.dx + c.dy * c.dy + c.dz * c.dz);
        const float step = 0.9f * (2.0f * bound / (float)kCoarseB) / dlen;
        const float span = far - t0;
        if (dlen > 0.0f && step > 0.0f && span < 2048.0f * step) {  // finite, sane ray; otherwise fall through to the marcher
            const uint32_t probes = (uint32_t)(span / step) + 2u;
            uint32_t t_hit = 0;  // bit pattern of the largest masked probe parameter (t > 0, so bit order = float order)
            const float bscale = 0.5f * (float)kCoarseB;
            for (uint32_t k = lane; k < probes; k += 32) {
                const float t = fminf(t0 + (float)k * step, far);
                float x, y, z;
                march_pos(c, t, x, y, z);
                const int cx = min((int)kCoarseB - 1, max(0, (int)((x * c.rbound + 1.0f) * bscale)));
                const int cy = min((int)kCoarseB - 1, max(0, (int)((y * c.rbound + 1.0f) * bscale)));
                const int cz = min((int)kCoarseB - 1, max(0, (int)((z * c.rbound + 1.0f) * bscale)));
                const uint32_t ci = (uint32_t)(cx + 64 * cy + 4096 * cz);
                if ((__ldg(coarse + (ci >> 5)) >> (ci & 31u)) & 1u) t_hit = max(t_hit, __float_as_uint(t));  // t >= t0 > 0
            }
            t_hit = __reduce_max_sync(0xffffffffu, t_hit);
            if (t_hit == 0u) {
                if (lane == 0) {
                    num_steps_out[n] = 0;
                    t0_out[n] = t0;
                    PVD_T(n, 1); PVD_T(n, 2);
                }
                return;
            }
            far = fminf(far, __uint_as_float(t_hit) + 0.75f * step);
        }
    }
    if (lane == 0) PVD_T(n, 1);
    float2* st = stash + (size_t)n * max_steps;
    const bool const_dt = (dt_gamma == 0.0f);
    const float dt_c = march_dt(c, 0.0f);
    const uint32_t lt_mask = (1u << lane) - 1u;

    float t_base = t0;
    float carry_tt = 0.0f;
    bool have_carry = false;
    uint32_t count = 0;
    bool done = false;

    // Replays the serial control flow of raymarching.cu:362-403 over one window of 32 evaluated lattice points WITHOUT a serial
    // loop.  Every point i has a successor nxt(i): i+1 after an occupied point (:389), the first j > i with t_j >= tt_i after an
    // empty one (the do-while of :398-402), 32 when that lies beyond the window.  The points the serial loop evaluates are the
    // chain start -> nxt(start) -> ...; nxt is found for all 32 points at once (integer arithmetic on the lattice, or a
    // 6-shuffle lower bound over the lanes' ascending t) and the chains are closed by pointer doubling (5 rounds: V =
    // points on the chain from i, as a bit mask; `close_chains`), all of it independent of where the chain enters the window -- so the four
    // windows of a group are prepared with full instruction-level parallelism and only `resolve` (a dozen instructions) is
    // serial from window to window.
    struct Prep {
        uint32_t V, valid_mask, occ_mask;
    };
    // first window index j with t_j >= tt, by a 6-shuffle binary search over the lanes' ascending t (any lattice)
    auto lower_bound_shfl = [&](float t_i, float tt) -> uint32_t {
        uint32_t pos = 0;
#pragma unroll
        for (uint32_t sft = 16; sft >= 1; sft >>= 1) {
            const float v = __shfl_sync(0xffffffffu, t_i, (int)(pos + sft - 1u));
            if (v < tt) pos += sft;
        }
        const float v31 = __shfl_sync(0xffffffffu, t_i, 31);
        if (pos == 31u && v31 < tt) pos = 32u;
        return pos;
    };
    // the same on the uniform lattice bits(t_j) = bw + j*cstep (positive floats: bit order = value order), no shuffles:
    // j = ceil((bits(tt) - bw) / cstep) clamped to [0, 32]; the quotient is estimated in float and corrected exactly.
    auto lower_bound_lattice = [&](uint32_t bw, uint32_t cstep, float rcstep, float tt) -> uint32_t {
        const uint32_t ttb = __float_as_uint(tt);
        const uint32_t diff = ttb - bw;                         // meaningful when ttb > bw
        uint32_t q = (uint32_t)((float)min(diff, 32u * cstep) * rcstep);  // cstep < 2^16 (group condition): exact float, |error| <= 1
        if (q * cstep > diff) --q;
        if ((q + 1u) * cstep <= diff) ++q;
        q += (q * cstep < diff) ? 1u : 0u;
        if (diff > 31u * cstep) q = 32u;                        // beyond t_31
        if (ttb <= bw) q = 0u;
        return q;
    };
    // successor of every point, then the chains closed by pointer doubling; W windows side by side so that their shuffle
    // chains interleave (the rounds are the outer loop).
    auto successors = [&](uint32_t pos, bool valid, bool occ) -> uint32_t {
        uint32_t P = (valid && occ) ? lane + 1u : max(pos, lane + 1u);
        if (!valid) P = 32u;  // landing on a point with t >= far ends the loop (:362)
        return P;
    };
    auto close_chains = [&](auto& pr, auto& P, auto W) {
        constexpr int kW = decltype(W)::value;
        uint32_t any_occ = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) any_occ |= pr[w].occ_mask;
        if (any_occ == 0u) {
            // nothing to emit in these windows (the common case): only each chain's last point matters, for the pending exit.
            // Q = successor with a self-loop at the end of the chain; five squarings reach the end from every start.
            uint32_t Q[kW];
#pragma unroll
            for (int w = 0; w < kW; ++w) Q[w] = (P[w] < 32u) ? P[w] : lane;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
#pragma unroll
                for (int w = 0; w < kW; ++w) Q[w] = __shfl_sync(0xffffffffu, Q[w], (int)Q[w]);
            }
#pragma unroll
            for (int w = 0; w < kW; ++w) pr[w].V = (1u << lane) | (1u << Q[w]);
        } else {
            uint32_t V[kW];
#pragma unroll
            for (int w = 0; w < kW; ++w) V[w] = 1u << lane;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
#pragma unroll
                for (int w = 0; w < kW; ++w) {
                    const uint32_t Pv = __shfl_sync(0xffffffffu, P[w], (int)(P[w] & 31u));
                    const uint32_t Vv = __shfl_sync(0xffffffffu, V[w], (int)(P[w] & 31u));
                    const bool in = P[w] < 32u;
                    V[w] |= in ? Vv : 0u;
                    P[w] = in ? Pv : P[w];
                }
            }
#pragma unroll
            for (int w = 0; w < kW; ++w) pr[w].V = V[w];
        }
    };
    // Straight-line on purpose: every shuffle / ballot is executed by the whole warp on every call and the (warp-uniform) state
    // is updated under predicates, so the compiler needs no re-convergence points between the windows of a group.
    auto resolve = [&](const Prep& r, float t_i, float dt_i, float tt) {
        const uint32_t ge = __ballot_sync(0xffffffffu, t_i >= carry_tt);
        const uint32_t cur = (have_carry && ge != 0u) ? (uint32_t)__ffs((int)ge) - 1u : 0u;
        uint32_t L = __shfl_sync(0xffffffffu, r.V, (int)cur);  // the evaluated points of this window, in chain order = bit order
        bool act = !done;
        if (act && r.valid_mask == 0u) {  // t only grows: nothing further can satisfy t < far
            done = true;
            act = false;
        }
        if (have_carry && ge == 0u) act = false;  // the pending skip passes over the whole window; it stays pending
        if (!act) L = 0u;
#ifdef PVD_TRACE
        if (act) ++tr_iters;
#endif
        const bool hit_far = (L & ~r.valid_mask) != 0u;  // the chain reaches a point with t >= far (its last point)
        L &= r.valid_mask;
        const uint32_t want = L & r.occ_mask;
        const uint32_t room = max_steps - count;
        const uint32_t rank = __popc(want & lt_mask);
        const bool full = (uint32_t)__popc(want) >= room;  // num_steps reaches max_steps inside this window (:362)
        const uint32_t emit = __ballot_sync(0xffffffffu, ((want >> lane) & 1u) && rank < room);
        if ((emit >> lane) & 1u) st[count + rank] = make_float2(t_i, dt_i);
        count += (uint32_t)__popc(emit);
        const uint32_t last = L ? 31u - (uint32_t)__clz((int)L) : 0u;
        const float tt_last = __shfl_sync(0xffffffffu, tt, (int)last);
        if (act) {
            if (hit_far || full) done = true;
            // the chain left the window through an empty cell: its exit is still pending
            have_carry = (L != 0u) && !((r.occ_mask >> last) & 1u);
            if (have_carry) carry_tt = tt_last;
        }
    };

#ifdef PVD_TRACE
    tr_c0 = clock64();
#endif
    // A valid ray evaluates at most (far - near) / dt_min <= max_steps * bound lattice points, 32 per window: the cap is twice that.
    // It only ever triggers on garbage input (non-finite rays, a zero direction), where the reference's loop would not terminate.
    const uint32_t max_windows = max_steps * (uint32_t)ceilf(fmaxf(bound, 1.0f)) / 16u + 64u;
    uint32_t windows = 0;
    while (!done && windows++ < max_windows) {
        // ---- fast path: constant dt and a group of four windows (128 lattice points) inside one binade.  All four occupancy
        // probes of a lane are issued before any is consumed, which hides the L2 latency of the bitfield loads, and the four
        // chain preparations interleave.
        bool group = false;
        uint32_t b = 0, cstep = 0;
        if (const_dt && t_base > 0.0f) {
            b = __float_as_uint(t_base);
            const uint32_t e = b >> 23;
            const uint32_t b1 = __float_as_uint(__fadd_rn(t_base, dt_c));
            if (e > 22 && e < 254 && (b1 >> 23) == e) {
                cstep = b1 - b;
                // tie check: dt / ulp(t) with fractional part exactly one half rounds by the parity of t -> serial path
                const float r = __fmul_rn(dt_c, __uint_as_float((uint32_t)(277 - (int)e) << 23));  // dt * 2^(150 - e)
                const bool tie = (r - floorf(r)) == 0.5f;
                group = !tie && cstep > 0 && ((b + 128u * cstep) >> 23) == e;
            }
        }
#ifdef PVD_TRACE
        if (group) ++tr_groups; else ++tr_general;
        tr_head += clock64() - tr_c0;
        tr_c0 = clock64();
#endif
        if (group) {
            // branch-free: a clamped position always has a cell, so points beyond `far` are located and loaded too (and
            // ignored) -- no divergent region, the four bitfield loads are in flight together while the exits are computed.
            float t_i[4], tt[4];
            bool valid[4], occ[4];
            uint32_t bits[4], sh[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                t_i[w] = __uint_as_float(b + ((uint32_t)w * 32u + lane) * cstep);
                valid[w] = t_i[w] < far;
                float x, y, z;
                march_pos(c, t_i[w], x, y, z);
                const MarchCell m = march_locate(c, dt_c, x, y, z);
                bits[w] = __ldg(grid + (m.index >> 3));
                sh[w] = m.index & 7u;
                tt[w] = march_leave(c, m, t_i[w], x, y, z);
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) occ[w] = (bits[w] >> sh[w]) & 1u;
#ifdef PVD_TRACE
            if (__ballot_sync(0xffffffffu, occ[0] | occ[1] | occ[2] | occ[3]) == 0xdeadbeefu) ++tr_iters;  // consume the loads
            tr_probe += clock64() - tr_c0;
            tr_c0 = clock64();
#endif
            Prep pr[4];
            uint32_t P[4];
            const float rcstep = 1.0f / (float)cstep;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                pr[w].valid_mask = __ballot_sync(0xffffffffu, valid[w]);
                pr[w].occ_mask = __ballot_sync(0xffffffffu, valid[w] && occ[w]);
                P[w] = successors(lower_bound_lattice(b + (uint32_t)w * 32u * cstep, cstep, rcstep, tt[w]), valid[w], occ[w]);
            }
            close_chains(pr, P, std::integral_constant<int, 4>{});
#ifdef PVD_TRACE
            if ((pr[0].V ^ pr[1].V ^ pr[2].V ^ pr[3].V) == 0xdeadbeefu) ++tr_iters;
            tr_prep += clock64() - tr_c0;
            tr_c0 = clock64();
#endif
#pragma unroll
            for (int w = 0; w < 4; ++w) resolve(pr[w], t_i[w], dt_c, tt[w]);
#ifdef PVD_TRACE
            tr_res += clock64() - tr_c0;
            tr_c0 = clock64();
#endif
            t_base = __uint_as_float(b + 128u * cstep);
        } else {
            // ---- general path: one window, lattice by the serial recurrence (every lane runs it and keeps its own element)
            float t = t_base, t_i = t_base;
#pragma unroll 4
            for (uint32_t j = 0; j < 32; ++j) {
                if (j == lane) t_i = t;
                t = __fadd_rn(t, const_dt ? dt_c : march_dt(c, t));
            }
            const bool valid = t_i < far;
            float x, y, z, tt = 0.0f;
            march_pos(c, t_i, x, y, z);
            const float dt_i = const_dt ? dt_c : march_dt(c, t_i);
            bool occ = false;
            if (valid) occ = march_probe(c, grid, t_i, dt_i, x, y, z, tt);
            Prep pr[1];
            uint32_t P[1];
            pr[0].valid_mask = __ballot_sync(0xffffffffu, valid);
            pr[0].occ_mask = __ballot_sync(0xffffffffu, valid && occ);
            P[0] = successors(lower_bound_shfl(t_i, tt), valid, occ);
            close_chains(pr, P, std::integral_constant<int, 1>{});
            resolve(pr[0], t_i, dt_i, tt);
            t_base = t;
#ifdef PVD_TRACE
            tr_gen += clock64() - tr_c0;
            tr_c0 = clock64();
#endif
        }
    }
    if (lane == 0) {
        num_steps_out[n] = (int32_t)count;
        t0_out[n] = t0;
#ifdef PVD_TRACE
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        PVD_T(n, 2); PVD_TV(n, 3, tr_groups); PVD_TV(n, 4, tr_general); PVD_TV(n, 5, tr_iters); PVD_TV(n, 6, count); PVD_TV(n, 9, gt);
        PVD_TV(n, 10, tr_probe); PVD_TV(n, 11, tr_prep); PVD_TV(n, 12, tr_res); PVD_TV(n, 13, tr_gen); PVD_TV(n, 14, tr_head);
#endif
    }
}

// Pass 2 (one CTA): exclusive prefix sum of num_steps in ray-id order; rays[n] = (n, offset, count);
// counter[0] += total, counter[1] += N  (what the reference's two atomics accumulate, raymarching.cu:408-409).
__global__ void __launch_bounds__(1024) k_march_scan(const int32_t* __restrict__ num_steps, uint32_t N,
                                                    int32_t* __restrict__ rays, int32_t* __restrict__ counter) {
    // 8 consecutive rays per thread (8192 per pass): serial prefix inside the thread, shuffle scan of the thread totals
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < N; base += 8u * blockDim.x) {
        const uint32_t n0 = base + 8u * threadIdx.x;
        int32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (n0 + i < N) ? __ldg(num_steps + n0 + i) : 0;
        int32_t tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += v[i];
        int32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += u;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if ((int)lane >= o) w += u;
            }
            warp_tot[lane] = w;  // inclusive
        }
        __syncthreads();
        const int32_t carry = carry_s;
        int32_t off = carry + (wid ? warp_tot[wid - 1] : 0) + inc - tot;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t n = n0 + i;
            if (n < N) {
                rays[3 * (size_t)n] = (int32_t)n;
                rays[3 * (size_t)n + 1] = off;
                rays[3 * (size_t)n + 2] = v[i];
            }
            off += v[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counter[0] += carry_s;
        counter[1] += (int32_t)N;
    }
}

// Pass 3 (one warp per ray): expand the stash into xyzs / dirs / deltas (raymarching.cu:454-469).
__global__ void __launch_bounds__(128) k_march_expand(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                     float bound, uint32_t max_steps, uint32_t N, uint32_t M,
                                                     const int32_t* __restrict__ rays, const float* __restrict__ t0s,
                                                     const float2* __restrict__ stash, float* __restrict__ xyzs,
                                                     float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)rays[3 * (size_t)n + 1];
    const uint32_t cnt = (uint32_t)rays[3 * (size_t)n + 2];
    if (cnt == 0 || offset + cnt >= M) return;  // dropped exactly as raymarching.cu:418-419
    const float ox = rays_o[3 * (size_t)n], oy = rays_o[3 * (size_t)n + 1], oz = rays_o[3 * (size_t)n + 2];
    const float dx = rays_d[3 * (size_t)n], dy = rays_d[3 * (size_t)n + 1], dz = rays_d[3 * (size_t)n + 2];
    const float t0 = t0s[n];
    const float2* st = stash + (size_t)n * max_steps;
    for (uint32_t i = lane; i < cnt; i += 32) {
        const float2 s = st[i];
        float last_t = t0;
        if (i > 0) {
            const float2 p = st[i - 1];
            last_t = __fadd_rn(p.x, p.y);
        }
        const float t_next = __fadd_rn(s.x, s.y);
        const size_t row = (size_t)offset + i;
        xyzs[3 * row + 0] = clampf(__fmaf_rn(s.x, dx, ox), -bound, bound);
        xyzs[3 * row + 1] = clampf(__fmaf_rn(s.x, dy, oy), -bound, bound);
        xyzs[3 * row + 2] = clampf(__fmaf_rn(s.x, dz, oz), -bound, bound);
        dirs[3 * row + 0] = dx;
        dirs[3 * row + 1] = dy;
        dirs[3 * row + 2] = dz;
        *reinterpret_cast<float2*>(deltas + 2 * row) = make_float2(s.y, __fadd_rn(t_next, -last_t));
    }
}

// =============================================================================================
// training composite: one warp per ray, shuffle scans
// =============================================================================================

struct Chunk {
    float w, T_after, r, g, b, tsum;
};

__global__ void __launch_bounds__(32) k_composite_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                      const float* __restrict__ deltas, const int32_t* __restrict__ rays,
                                                      uint32_t M, uint32_t N, float* __restrict__ weights_sum,
                                                      float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x;  // one warp per CTA: n, and every branch below, is provably warp-uniform (no WARPSYNC around the scans)
    const uint32_t lane = threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n];
    const uint32_t offset = (uint32_t)rays[3 * (size_t)n + 1];
    const uint32_t cnt = (uint32_t)rays[3 * (size_t)n + 2];
    if (cnt == 0 || offset + cnt >= M) {  // raymarching.cu:525-532
        if (lane == 0) {
            weights_sum[index] = 0;
            depth[index] = 0;
            image[3 * (size_t)index] = 0;
            image[3 * (size_t)index + 1] = 0;
            image[3 * (size_t)index + 2] = 0;
        }
        return;
    }
    float T = 1.0f, tcarry = 0.0f;           // carried across 32-sample chunks
    float r = 0, g = 0, b = 0, ws = 0, d = 0;  // per-lane partial sums
    // blocks of 4 chunks (128 samples): the loads of all four chunks are issued before the first one goes through its two shuffle
    // scans, so a ray costs one memory latency per 128 samples instead of one per 32 (most rays: one latency in all)
    for (uint32_t base0 = 0; base0 < cnt; base0 += 128) {
        float sg[4], c0[4], c1[4], c2[4];
        float2 dl[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = base0 + 32u * q + lane;
            const bool ok = i < cnt;
            const size_t row = (size_t)offset + (ok ? i : 0);
            sg[q] = ok ? __ldg(sigmas + row) : 0.0f;
            dl[q] = ok ? __ldg(reinterpret_cast<const float2*>(deltas + 2 * row)) : make_float2(0.f, 0.f);
            c0[q] = ok ? __ldg(rgbs + 3 * row) : 0.f;
            c1[q] = ok ? __ldg(rgbs + 3 * row + 1) : 0.f;
            c2[q] = ok ? __ldg(rgbs + 3 * row + 2) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t base = base0 + 32u * q;
            if (base >= cnt) break;  // warp-uniform
            const bool ok = base + lane < cnt;
            const float alpha = ok ? 1.0f - __expf(-sg[q] * dl[q].x) : 0.0f;  // raymarching.cu:546
            const float om = 1.0f - alpha;
            const float incl = warp_scan_mul(om, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float Ti = T * excl;
            const float w = alpha * Ti;
            const float tin = tcarry + warp_scan_add(dl[q].y, lane);
            if (ok) {
                r += w * c0[q];
                g += w * c1[q];
                b += w * c2[q];
                d += w * tin;
                ws += w;
            }
            T *= __shfl_sync(0xffffffffu, incl, 31);
            tcarry = __shfl_sync(0xffffffffu, tin, 31);
        }
    }
    r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[3 * (size_t)index] = r;
        image[3 * (size_t)index + 1] = g;
        image[3 * (size_t)index + 2] = b;
    }
}

// FUSED_MSE: instead of reading upstream gradients, derive them from the photometric loss of renderer.py:445 + the
// trainer's criterion (just_train_tea/utils.py:841-846):  pred = image + (1 - ws) * bg ; loss = mean((pred - gt)^2).
// grad_ws then points to gt [N,3], grad_img to bg [3]; `loss_scale` multiplies the gradients (GradScaler) and
// loss_out[0] accumulates the UNSCALED loss, loss_out[1] the number of rays that contributed.
template <bool FUSED_MSE>
__global__ void __launch_bounds__(32) k_composite_bwd(const float* __restrict__ grad_ws, const float* __restrict__ grad_img,
                                                      const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                      const float* __restrict__ deltas, const int32_t* __restrict__ rays,
                                                      const float* __restrict__ weights_sum, const float* __restrict__ image,
                                                      uint32_t M, uint32_t N, float* __restrict__ grad_sigmas,
                                                      float* __restrict__ grad_rgbs, float loss_scale, float* __restrict__ loss_out) {
    const uint32_t n = blockIdx.x;  // one warp per CTA: n, and every branch below, is provably warp-uniform (no WARPSYNC around the scans)
    const uint32_t lane = threadIdx.x;
    if (n >= N) return;
#ifdef PVD_TRACE
    if (lane == 0) {
        unsigned long long gt_; unsigned int smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        PVD_T(4096u + n, 0); PVD_TV(4096u + n, 8, gt_); PVD_TV(4096u + n, 7, smid);
    }
#endif
    const uint32_t index = (uint32_t)rays[3 * (size_t)n];
    const uint32_t offset = (uint32_t)rays[3 * (size_t)n + 1];
    const uint32_t cnt = (uint32_t)rays[3 * (size_t)n + 2];
    const bool skip = (cnt == 0 || offset + cnt >= M);  // raymarching.cu:629
    const float r_final = image[3 * (size_t)index], g_final = image[3 * (size_t)index + 1],
                b_final = image[3 * (size_t)index + 2], ws_final = weights_sum[index];
    float gws, gr, gg, gb;
    if constexpr (FUSED_MSE) {
        const float* gt = grad_ws + 3 * (size_t)index;
        const float bgr = grad_img[0], bgg = grad_img[1], bgb = grad_img[2];
        const float om = 1.0f - ws_final;
        const float dr = r_final + om * bgr - gt[0], dg = g_final + om * bgg - gt[1], db = b_final + om * bgb - gt[2];
        const float k = 2.0f / (3.0f * (float)N);
        if (lane == 0) {
            float* slot = loss_out + 2u * (n % PVD_LOSS_SLOTS);
            atomicAdd(slot, (dr * dr + dg * dg + db * db) / (3.0f * (float)N));
            if (!skip) atomicAdd(slot + 1, 1.0f);
        }
        gr = k * dr * loss_scale; gg = k * dg * loss_scale; gb = k * db * loss_scale;
        gws = -(gr * bgr + gg * bgg + gb * bgb);
    } else {
        gws = grad_ws[index];
        gr = grad_img[3 * (size_t)index]; gg = grad_img[3 * (size_t)index + 1]; gb = grad_img[3 * (size_t)index + 2];
    }
    if (lane == 0) { PVD_T(4096u + n, 1); PVD_TV(4096u + n, 6, cnt); }
    if (skip) {
        // a ray that does not fit the sample buffer contributes nothing (raymarching.cu:629), but its rows below M are still read by
        // the field backward: zero them here, so that the caller need not clear the whole gradient buffers every step
        for (uint32_t i = offset + lane; i < min(offset + cnt, M); i += 32) {
            grad_sigmas[i] = 0.0f;
            grad_rgbs[3 * (size_t)i] = 0.0f;
            grad_rgbs[3 * (size_t)i + 1] = 0.0f;
            grad_rgbs[3 * (size_t)i + 2] = 0.0f;
        }
        return;
    }
    float T = 1.0f, rc = 0, gc = 0, bc = 0, wc = 0;  // carries (running sums up to the previous chunk)
    for (uint32_t base0 = 0; base0 < cnt; base0 += 128) {  // 4 chunks of loads in flight at once, as in the forward
        float sgv[4], d0v[4], c0v[4], c1v[4], c2v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = base0 + 32u * q + lane;
            const bool ok = i < cnt;
            const size_t row = (size_t)offset + (ok ? i : 0);
            sgv[q] = ok ? __ldg(sigmas + row) : 0.0f;
            d0v[q] = ok ? __ldg(deltas + 2 * row) : 0.0f;
            c0v[q] = ok ? __ldg(rgbs + 3 * row) : 0.f;
            c1v[q] = ok ? __ldg(rgbs + 3 * row + 1) : 0.f;
            c2v[q] = ok ? __ldg(rgbs + 3 * row + 2) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
        const uint32_t base = base0 + 32u * q;
        if (base >= cnt) break;  // warp-uniform
        const uint32_t i = base + lane;
        const bool ok = i < cnt;
        const size_t row = (size_t)offset + (ok ? i : 0);
        const float sigma = sgv[q], d0 = d0v[q], cr = c0v[q], cg = c1v[q], cb = c2v[q];
        const float alpha = ok ? 1.0f - __expf(-sigma * d0) : 0.0f;
        const float om = 1.0f - alpha;
        const float incl = warp_scan_mul(om, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = alpha * (T * excl);
        const float T_after = T * incl;  // T after `T *= 1 - alpha` (raymarching.cu:660)
        const float r_run = rc + warp_scan_add(w * cr, lane);
        const float g_run = gc + warp_scan_add(w * cg, lane);
        const float b_run = bc + warp_scan_add(w * cb, lane);
        const float w_run = wc + warp_scan_add(w, lane);
        if (ok) {
            grad_rgbs[3 * row] = gr * w;
            grad_rgbs[3 * row + 1] = gg * w;
            grad_rgbs[3 * row + 2] = gb * w;
            grad_sigmas[row] = d0 * (gr * (T_after * cr - (r_final - r_run)) + gg * (T_after * cg - (g_final - g_run)) +
                                     gb * (T_after * cb - (b_final - b_run)) + gws * (T_after - (ws_final - w_run)));
        }
        T *= __shfl_sync(0xffffffffu, incl, 31);
        rc = __shfl_sync(0xffffffffu, r_run, 31);
        gc = __shfl_sync(0xffffffffu, g_run, 31);
        bc = __shfl_sync(0xffffffffu, b_run, 31);
        wc = __shfl_sync(0xffffffffu, w_run, 31);
        }
    }
#ifdef PVD_TRACE
    if (lane == 0) {
        unsigned long long gt_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));
        PVD_T(4096u + n, 2); PVD_TV(4096u + n, 9, gt_);
    }
#endif
}

// Training composite, forward AND backward in one launch, for the MSE criterion (the engine's path): the warp that owns a ray runs
// the forward sweep (raymarching.cu:505-579), derives d(loss)/d(image, weights_sum) from the finished pixel exactly as
// k_composite_bwd<true> does, and sweeps the ray again for the sample gradients (raymarching.cu:597-697) -- the second sweep hits
// L1/L2, and one launch, one set of header loads and one kernel tail disappear from the step's critical path.
__global__ void __launch_bounds__(32) k_composite_train_mse(const float* __restrict__ gt_rgb, const float* __restrict__ bg,
                                                           float loss_scale, const float* __restrict__ sigmas,
                                                           const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                                           const int32_t* __restrict__ rays, uint32_t M, uint32_t N,
                                                           float* __restrict__ weights_sum, float* __restrict__ depth,
                                                           float* __restrict__ image, float* __restrict__ grad_sigmas,
                                                           float* __restrict__ grad_rgbs, float* __restrict__ loss_out) {
    pdl_wait();                // (launched with launch_pdl behind the field forward: resident early, starts the moment that grid is done)
    pdl_launch_dependents();   // the field backward (launched with launch_pdl) may start its prologue + forward recomputation now
    const uint32_t n = blockIdx.x;
    const uint32_t lane = threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n];
    const uint32_t offset = (uint32_t)rays[3 * (size_t)n + 1];
    const uint32_t cnt = (uint32_t)rays[3 * (size_t)n + 2];
    const bool skip = (cnt == 0 || offset + cnt >= M);  // raymarching.cu:525-532, :629
    const float gt0 = gt_rgb[3 * (size_t)index], gt1 = gt_rgb[3 * (size_t)index + 1], gt2 = gt_rgb[3 * (size_t)index + 2];
    const float bgr = bg[0], bgg = bg[1], bgb = bg[2];
    float r = 0, g = 0, b = 0, ws = 0, d = 0;
    if (!skip) {
        float T = 1.0f, tcarry = 0.0f;
        for (uint32_t base0 = 0; base0 < cnt; base0 += 128) {
            float sg[4], c0[4], c1[4], c2[4];
            float2 dl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = base0 + 32u * q + lane;
                const bool ok = i < cnt;
                const size_t row = (size_t)offset + (ok ? i : 0);
                sg[q] = ok ? __ldg(sigmas + row) : 0.0f;
                dl[q] = ok ? __ldg(reinterpret_cast<const float2*>(deltas + 2 * row)) : make_float2(0.f, 0.f);
                c0[q] = ok ? __ldg(rgbs + 3 * row) : 0.f;
                c1[q] = ok ? __ldg(rgbs + 3 * row + 1) : 0.f;
                c2[q] = ok ? __ldg(rgbs + 3 * row + 2) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t base = base0 + 32u * q;
                if (base >= cnt) break;  // warp-uniform
                const bool ok = base + lane < cnt;
                const float alpha = ok ? 1.0f - __expf(-sg[q] * dl[q].x) : 0.0f;
                const float om = 1.0f - alpha;
                const float incl = warp_scan_mul(om, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.0f;
                const float w = alpha * (T * excl);
                const float tin = tcarry + warp_scan_add(dl[q].y, lane);
                if (ok) {
                    r += w * c0[q];
                    g += w * c1[q];
                    b += w * c2[q];
                    d += w * tin;
                    ws += w;
                }
                T *= __shfl_sync(0xffffffffu, incl, 31);
                tcarry = __shfl_sync(0xffffffffu, tin, 31);
            }
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    }
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[3 * (size_t)index] = r;
        image[3 * (size_t)index + 1] = g;
        image[3 * (size_t)index + 2] = b;
    }
    // ---- loss and its gradient w.r.t. the pixel (renderer.py:445 background mix, MSELoss over N x 3 values)
    const float r_final = r, g_final = g, b_final = b, ws_final = ws;
    const float om_f = 1.0f - ws_final;
    const float dr = r_final + om_f * bgr - gt0, dg = g_final + om_f * bgg - gt1, db = b_final + om_f * bgb - gt2;
    const float k = 2.0f / (3.0f * (float)N);
    if (lane == 0) {
        float* slot = loss_out + 2u * (n % PVD_LOSS_SLOTS);
        atomicAdd(slot, (dr * dr + dg * dg + db * db) / (3.0f * (float)N));
        if (!skip) atomicAdd(slot + 1, 1.0f);
    }
    const float gr = k * dr * loss_scale, gg = k * dg * loss_scale, gb = k * db * loss_scale;
    const float gws = -(gr * bgr + gg * bgg + gb * bgb);
    if (skip) {  // rows below M of a ray that does not fit are still read by the field backward: clear them
        for (uint32_t i = offset + lane; i < min(offset + cnt, M); i += 32) {
            grad_sigmas[i] = 0.0f;
            grad_rgbs[3 * (size_t)i] = 0.0f;
            grad_rgbs[3 * (size_t)i + 1] = 0.0f;
            grad_rgbs[3 * (size_t)i + 2] = 0.0f;
        }
        return;
    }
    float T = 1.0f, rc = 0, gc = 0, bc = 0, wc = 0;
    for (uint32_t base0 = 0; base0 < cnt; base0 += 128) {
        float sgv[4], d0v[4], c0v[4], c1v[4], c2v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = base0 + 32u * q + lane;
            const bool ok = i < cnt;
            const size_t row = (size_t)offset + (ok ? i : 0);
            sgv[q] = ok ? __ldg(sigmas + row) : 0.0f;
            d0v[q] = ok ? __ldg(deltas + 2 * row) : 0.0f;
            c0v[q] = ok ? __ldg(rgbs + 3 * row) : 0.f;
            c1v[q] = ok ? __ldg(rgbs + 3 * row + 1) : 0.f;
            c2v[q] = ok ? __ldg(rgbs + 3 * row + 2) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t base = base0 + 32u * q;
            if (base >= cnt) break;  // warp-uniform
            const uint32_t i = base + lane;
            const bool ok = i < cnt;
            const size_t row = (size_t)offset + (ok ? i : 0);
            const float sigma = sgv[q], d0 = d0v[q], cr = c0v[q], cg = c1v[q], cb = c2v[q];
            const float alpha = ok ? 1.0f - __expf(-sigma * d0) : 0.0f;
            const float om = 1.0f - alpha;
            const float incl = warp_scan_mul(om, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float w = alpha * (T * excl);
            const float T_after = T * incl;
            const float r_run = rc + warp_scan_add(w * cr, lane);
            const float g_run = gc + warp_scan_add(w * cg, lane);
            const float b_run = bc + warp_scan_add(w * cb, lane);
            const float w_run = wc + warp_scan_add(w, lane);
            if (ok) {
                grad_rgbs[3 * row] = gr * w;
                grad_rgbs[3 * row + 1] = gg * w;
                grad_rgbs[3 * row + 2] = gb * w;
                grad_sigmas[row] = d0 * (gr * (T_after * cr - (r_final - r_run)) + gg * (T_after * cg - (g_final - g_run)) +
                                         gb * (T_after * cb - (b_final - b_run)) + gws * (T_after - (ws_final - w_run)));
            }
            T *= __shfl_sync(0xffffffffu, incl, 31);
            rc = __shfl_sync(0xffffffffu, r_run, 31);
            gc = __shfl_sync(0xffffffffu, g_run, 31);
            bc = __shfl_sync(0xffffffffu, b_run, 31);
            wc = __shfl_sync(0xffffffffu, w_run, 31);
        }
    }
}

// =============================================================================================
// inference kernels: one thread per alive ray, n_step <= 8 samples per call
// =============================================================================================

__global__ void k_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                             const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                             const float* __restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
                             uint32_t C, uint32_t H, const uint8_t* __restrict__ grid, const float* __restrict__ nears,
                             const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                             float* __restrict__ deltas, uint32_t perturb, Pcg32 rng) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    float t = rays_t[n];
    MarchCtx c;
    march_ctx_init(c, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H);
    const float far = fars[index];
    float* px = xyzs + (size_t)n * n_step * 3;
    float* pd = dirs + (size_t)n * n_step * 3;
    float* pl = deltas + (size_t)n * n_step * 2;
    if (perturb) {  // raymarching.cu:749-752 (advance by the slot n, not the ray id)
        pcg32_advance(rng, (uint64_t)n);
        t = __fmaf_rn(c.dt_min, pcg32_next_float(rng), t);
    }
    float last_t = t;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        float x, y, z, tt;
        march_pos(c, t, x, y, z);
        const float dt = march_dt(c, t);
        if (march_probe(c, grid, t, dt, x, y, z, tt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
            t = __fadd_rn(t, dt);
            pl[0] = dt;
            pl[1] = __fadd_rn(t, -last_t);
            last_t = t;
            px += 3; pd += 3; pl += 2;
            ++step;
        } else {
            do { t = __fadd_rn(t, march_dt(c, t)); } while (t < tt);
        }
    }
}

__global__ void k_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                                 float* __restrict__ rays_t, const float* __restrict__ sigmas,
                                 const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                 float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    float t = rays_t[n];
    const float* ps = sigmas + (size_t)n * n_step;
    const float* pc = rgbs + (size_t)n * n_step * 3;
    const float* pl = deltas + (size_t)n * n_step * 2;
    float ws = weights_sum[index], d = depth[index];
    float r = image[3 * (size_t)index], g = image[3 * (size_t)index + 1], b = image[3 * (size_t)index + 2];
    uint32_t step = 0;
    while (step < n_step) {
        if (pl[0] == 0) break;  // ray ran out of samples (raymarching.cu:862)
        const float alpha = 1.0f - __expf(-ps[0] * pl[0]);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t += pl[1];
        d += w * t;
        r += w * pc[0];
        g += w * pc[1];
        b += w * pc[2];
        if (T < 1e-4f) break;  // early termination (raymarching.cu:886)
        ps += 1; pc += 3; pl += 2;
        ++step;
    }
    rays_t[n] = (step < n_step) ? -1.0f : t;
    weights_sum[index] = ws;
    depth[index] = d;
    image[3 * (size_t)index] = r;
    image[3 * (size_t)index + 1] = g;
    image[3 * (size_t)index + 2] = b;
}

// stable compaction of the alive list by one CTA (ballot + popc scan)
__global__ void __launch_bounds__(1024) k_compact_rays(uint32_t n_alive, int32_t* __restrict__ rays_alive,
                                                      const int32_t* __restrict__ rays_alive_old,
                                                      float* __restrict__ rays_t, const float* __restrict__ rays_t_old,
                                                      int32_t* __restrict__ alive_counter) {
    __shared__ int32_t warp_cnt[32];
    __shared__ int32_t carry_s;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = alive_counter[0];
    __syncthreads();
    for (uint32_t base = 0; base < n_alive; base += blockDim.x) {
        const uint32_t n = base + threadIdx.x;
        const float told = (n < n_alive) ? rays_t_old[n] : -1.0f;
        const bool keep = (n < n_alive) && (told >= 0);  // raymarching.cu:934
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(mask);
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_cnt[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if ((int)lane >= o) w += u;
            }
            warp_cnt[lane] = w;
        }
        __syncthreads();
        const int32_t carry = carry_s;
        if (keep) {
            const int32_t pos = carry + (wid ? warp_cnt[wid - 1] : 0) + __popc(mask & ((1u << lane) - 1u));
            rays_alive[pos] = rays_alive_old[n];
            rays_t[pos] = told;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_cnt[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) alive_counter[0] = carry_s;
}

}  // namespace pvd

// =============================================================================================
// C ABI
// =============================================================================================
using namespace pvd;

extern "C" {

int pvd_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                           float* nears, float* fars, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(rays_o && rays_d && aabb && nears && fars);
    k_near_far<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_polar_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                       void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(rays_o && rays_d && coords);
    k_polar<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(coords && indices);
    k_morton3D<<<ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(coords && indices);
    k_morton3D_invert<<<ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(grid && bitfield);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15u) == 0);
    k_packbits<<<ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

constexpr uint64_t kCoarseWords = 2ull * kCoarseMaskWords;  // dilated 64^3 block mask, then the undilated one

// The block mask is built from cascade 0 of a 128^3 grid and probed with cells mapped by 1 / bound; cascade 0 spans
// [-min(1, bound), min(1, bound)] (march_locate), so the two agree only for bound <= 1 (then C == 1 anyway, renderer.py:81).
static inline bool coarse_mask_applies(uint32_t C, uint32_t H, float bound) { return (C == 1) && (H == 2 * kCoarseB) && (bound <= 1.0f); }

int pvd_march_coarse_mask(const uint8_t* grid, uint32_t C, uint32_t H, float bound, int32_t* ws_i32, void* stream) {
    PVD_REQUIRE(grid && ws_i32);
    if (!coarse_mask_applies(C, H, bound)) return PVD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* coarse = reinterpret_cast<uint32_t*>(ws_i32);
    uint32_t* any = coarse + kCoarseMaskWords;
    k_coarse_any<<<kCoarseMaskWords / 256, 256, 0, st>>>(grid, any);
    PVD_LAUNCH_CHECK();
    k_coarse_dilate<<<kCoarseMaskWords / 256, 256, 0, st>>>(any, coarse);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

uint64_t pvd_march_rays_train_workspace_words(uint32_t N, uint32_t max_steps) {
    // coarse mask[8192] + any[8192] | num_steps[N] | t0[N] | stash[N*max_steps] float2   (stash kept 8-byte aligned)
    const uint64_t head = ((2ull * N + 1ull) / 2ull) * 2ull;
    return kCoarseWords + head + 2ull * (uint64_t)N * max_steps;
}

int pvd_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                               uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars,
                               int32_t* rays, int32_t* counter, uint32_t perturb, int32_t* ws_i32, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && ws_i32);
    PVD_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(ws_i32) & 7u) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* coarse = reinterpret_cast<uint32_t*>(ws_i32);
    int32_t* num_steps = ws_i32 + kCoarseWords;
    float* t0 = reinterpret_cast<float*>(ws_i32 + kCoarseWords + N);
    const uint64_t head = ((2ull * N + 1ull) / 2ull) * 2ull;
    float2* stash = reinterpret_cast<float2*>(ws_i32 + kCoarseWords + head);
    const Pcg32 rng = pcg32_seeded(42u);  // hard-coded seed, raymarching.cu:488
    // coarse pruning: one cascade, the reference's 128^3 grid (a 2x2x2 block of cells = one byte of the Morton bitfield)
    const bool use_coarse = coarse_mask_applies(C, H, bound);
    if (use_coarse) {
        uint32_t* any = reinterpret_cast<uint32_t*>(ws_i32) + kCoarseMaskWords;
        k_coarse_any<<<kCoarseMaskWords / 256, 256, 0, st>>>(grid, any);
        PVD_LAUNCH_CHECK();
        k_coarse_dilate<<<kCoarseMaskWords / 256, 256, 0, st>>>(any, coarse);
        PVD_LAUNCH_CHECK();
    }
    k_march_count<<<N, 32, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars,
                                                  perturb, rng, num_steps, t0, stash, use_coarse ? coarse : nullptr, nullptr, 0.0f,
                                                  nullptr, nullptr);
    PVD_LAUNCH_CHECK();
    k_march_scan<<<1, 1024, 0, st>>>(num_steps, N, rays, counter);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_march_rays_train_count_aabb(const float* rays_o, const float* rays_d, const uint8_t* grid, const float* aabb, float min_near,
                                    float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                    float* nears, float* fars, int32_t* rays, int32_t* counter, uint32_t perturb,
                                    uint32_t reuse_coarse, int32_t* ws_i32, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(rays_o && rays_d && grid && aabb && nears && fars && rays && counter && ws_i32);
    PVD_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(ws_i32) & 7u) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* coarse = reinterpret_cast<uint32_t*>(ws_i32);
    int32_t* num_steps = ws_i32 + kCoarseWords;
    float* t0 = reinterpret_cast<float*>(ws_i32 + kCoarseWords + N);
    const uint64_t head = ((2ull * N + 1ull) / 2ull) * 2ull;
    float2* stash = reinterpret_cast<float2*>(ws_i32 + kCoarseWords + head);
    const Pcg32 rng = pcg32_seeded(42u);
    const bool use_coarse = coarse_mask_applies(C, H, bound);
    if (use_coarse && !reuse_coarse) {
        uint32_t* any = reinterpret_cast<uint32_t*>(ws_i32) + kCoarseMaskWords;
        k_coarse_any<<<kCoarseMaskWords / 256, 256, 0, st>>>(grid, any);
        PVD_LAUNCH_CHECK();
        k_coarse_dilate<<<kCoarseMaskWords / 256, 256, 0, st>>>(any, coarse);
        PVD_LAUNCH_CHECK();
    }
    k_march_count<<<N, 32, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nullptr, nullptr, perturb,
                                                  rng, num_steps, t0, stash, use_coarse ? coarse : nullptr, aabb, min_near, nears, fars);
    PVD_LAUNCH_CHECK();
    k_march_scan<<<1, 1024, 0, st>>>(num_steps, N, rays, counter);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_march_rays_train_write(const float* rays_o, const float* rays_d, float bound, uint32_t max_steps, uint32_t N,
                               uint32_t M, const int32_t* rays, const int32_t* ws_i32, float* xyzs, float* dirs,
                               float* deltas, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(rays_o && rays_d && rays && ws_i32 && xyzs && dirs && deltas);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(ws_i32) & 7u) == 0 && (reinterpret_cast<uintptr_t>(deltas) & 7u) == 0);
    const float* t0 = reinterpret_cast<const float*>(ws_i32 + kCoarseWords + N);
    const uint64_t head = ((2ull * N + 1ull) / 2ull) * 2ull;
    const float2* stash = reinterpret_cast<const float2*>(ws_i32 + kCoarseWords + head);
    k_march_expand<<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, bound, max_steps, N, M, rays, t0, stash,
                                                                     xyzs, dirs, deltas);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                         const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                         uint32_t perturb, int32_t* ws_i32, void* stream) {
    int rc = pvd_march_rays_train_count(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, rays, counter,
                                        perturb, ws_i32, stream);
    if (rc != PVD_OK) return rc;
    return pvd_march_rays_train_write(rays_o, rays_d, bound, max_steps, N, M, rays, ws_i32, xyzs, dirs, deltas, stream);
}

int pvd_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                     uint32_t M, uint32_t N, float* weights_sum, float* depth, float* image,
                                     void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && depth && image);
    k_composite_fwd<<<N, 32, 0, (cudaStream_t)stream>>>(sigmas, rgbs, deltas, rays, M, N, weights_sum,
                                                                      depth, image);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                      const float* rgbs, const float* deltas, const int32_t* rays,
                                      const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                      float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(grad_weights_sum && grad_image && sigmas && rgbs && deltas && rays && weights_sum && image &&
                grad_sigmas && grad_rgbs);
    k_composite_bwd<false><<<N, 32, 0, (cudaStream_t)stream>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas,
                                                                             rays, weights_sum, image, M, N, grad_sigmas,
                                                                             grad_rgbs, 1.0f, nullptr);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_composite_rays_train_backward_mse(const float* gt_rgb, const float* bg_color, float loss_scale, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int32_t* rays,
                                          const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                          float* grad_sigmas, float* grad_rgbs, float* loss_out, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(gt_rgb && bg_color && sigmas && rgbs && deltas && rays && weights_sum && image && grad_sigmas && grad_rgbs &&
                loss_out);
    k_composite_bwd<true><<<N, 32, 0, (cudaStream_t)stream>>>(gt_rgb, bg_color, sigmas, rgbs, deltas, rays,
                                                                            weights_sum, image, M, N, grad_sigmas, grad_rgbs,
                                                                            loss_scale, loss_out);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_composite_rays_train_mse(const float* gt_rgb, const float* bg_color, float loss_scale, const float* sigmas,
                                 const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                 float* weights_sum, float* depth, float* image, float* grad_sigmas, float* grad_rgbs,
                                 float* loss_out, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(gt_rgb && bg_color && sigmas && rgbs && deltas && rays && weights_sum && depth && image && grad_sigmas &&
                grad_rgbs && loss_out);
    cudaError_t e = launch_pdl(k_composite_train_mse, dim3(N), dim3(32), 0, (cudaStream_t)stream, gt_rgb, bg_color, loss_scale, sigmas, rgbs, deltas, rays,
                               M, N, weights_sum, depth, image, grad_sigmas, grad_rgbs, loss_out);
    if (e != cudaSuccess) return (int)e;
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                   const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                   uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                   float* dirs, float* deltas, uint32_t perturb, void* stream) {
    if (n_alive == 0) return PVD_OK;
    PVD_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas);
    PVD_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1 && n_step >= 1);
    const Pcg32 rng = pcg32_seeded((uint64_t)perturb);  // raymarching.cu:816
    k_march_rays<<<ceil_div(n_alive, 64), 64, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o,
                                                                         rays_d, bound, dt_gamma, max_steps, C, H, grid,
                                                                         nears, fars, xyzs, dirs, deltas, perturb, rng);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t, const float* sigmas,
                       const float* rgbs, const float* deltas, float* weights_sum, float* depth, float* image,
                       void* stream) {
    if (n_alive == 0) return PVD_OK;
    PVD_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image);
    k_composite_rays<<<ceil_div(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, sigmas,
                                                                               rgbs, deltas, weights_sum, depth, image);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                     const float* rays_t_old, int32_t* alive_counter, void* stream) {
    if (n_alive == 0) return PVD_OK;
    PVD_REQUIRE(rays_alive && rays_alive_old && rays_t && rays_t_old && alive_counter);
    k_compact_rays<<<1, 1024, 0, (cudaStream_t)stream>>>(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old,
                                                         alive_counter);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
