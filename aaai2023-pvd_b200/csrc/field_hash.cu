// field_hash.cu -- fused "hash" field query for sm_100a: multiresolution hash-grid encode + degree-4 SH +
// sigma_net (2L->64->16) + color_net (31->64->64->3), forward and backward, one kernel per direction.
//
// Replaces, for model_type "hash", the chain of ~15 launches per direction that the reference issues from
// NeRFNetwork.forward (distill_mutual/network.py:335-437): GridEncoder kernel + permute, 5 cuBLAS GEMMs with
// ReLU/clamp/exp/sigmoid/cat kernels between them, the SH kernel, and their autograd mirrors.
//
// Work decomposition: one CTA = 128 threads = one tile of 128 consecutive samples; thread t owns sample row t.
//   * gather phase: each thread interpolates its sample's 2L features from the (L2-resident) table with
//     8-byte/4-byte vector loads, 32 independent gathers in flight per thread;
//   * MLP phase: activations live in shared memory as fp16 "chunk" tiles (tc5.cuh) that are simultaneously valid
//     K-major and MN-major tcgen05 operands; every layer is 2-4 tcgen05.mma instructions (M = 128 samples) issued by
//     one thread, accumulating in TMEM; each thread then pulls ITS OWN row back with tcgen05.ld (32x32b: TMEM lane ==
//     sample row), applies ReLU / clamp / exp / sigmoid in registers and writes the next operand tile;
//   * backward: forward recomputed from the saved fp16 encoding, then per layer one weight-gradient GEMM
//     (reduction over the 128 samples, both operands MN-major views of the SAME tiles, fp32 accumulators persistent
//     in TMEM across all tiles a CTA processes) and one data-gradient GEMM (weights as MN-major operand, so no
//     transposed weight copy exists); the encoding gradient is scattered with red.global.add.v2/v4.f32 by four extra
//     scatter warps of the same CTA (default) or by a second kernel (k_hash_scatter).
// CTAs are persistent (grid = min(#tiles, 2 x #SMs)) so weights are staged and TMEM is allocated once per CTA and
// weight gradients leave the SM once.
#include <stdlib.h>
#include "common.cuh"
PVD_TRACE_TU(pvd_debug_trace_field_hash)
#include "gridenc.cuh"
#include "shenc.cuh"
#include "tc5.cuh"
#include "field_common.cuh"
#include "field_tail.cuh"
#include "field_hash.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

__device__ __forceinline__ void stage_weights(uint8_t* smw, const uint8_t* __restrict__ blob) { stage_blob(smw, blob, PVD_FIELD_WBLOB_BYTES); }

// =============================================================================================== forward kernel
// TMA_IN: the tile's sample block -- 128 consecutive rows of xyzs and of dirs, 1536 contiguous bytes each -- is staged into shared
// memory by two TMA bulk copies (cp.async.bulk, mbarrier complete_tx) issued by one thread: the first tile's while the CTA is still in
// its prologue (TMEM allocation, level table, barrier), a further tile's as soon as the previous tile's last layer has completed.
// The staging area is the first 3 KB of the H1/H3/H4 activation tile, which is dead between a tile's last MMA and the next tile's
// first epilogue -- no shared memory is added (a first version with its own double buffer cost 8 KB per CTA, the fourth resident
// CTA per SM with it, and ran 48.8 us instead of 32.0: profiles/README.md).  A tile whose byte count is not a multiple of 16 (only
// the last, when M % 4 != 0) is loaded per thread.
template <typename T, bool TMA_IN>
__global__ void __launch_bounds__(128) k_hash_field_fwd(FieldArgs a, const float* __restrict__ xyzs, const float* __restrict__ dirs,
                                                        uint32_t M, float* __restrict__ sigmas, float* __restrict__ rgbs,
                                                        __half* __restrict__ enc, float* __restrict__ feat16, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, wbar, sbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ LevelInfo lv[16];
    // Forward-only tiles alias: the encoding X is dead once layer 1's MMAs completed, so the colour-net input CIN reuses its
    // buffer; H1, H3, H4 are each dead when the next one is written (its consumer MMA has been waited for).  44 KB per CTA
    // instead of 68 KB -> 4 resident CTAs per SM (the TMEM limit) instead of 3, i.e. one tile per CTA at 4096 rays.
    uint8_t* smw = smem;                               // 20480
    uint8_t* X = smem + PVD_FIELD_WBLOB_BYTES;         // 8192  (X, then CIN)
    uint8_t* CIN = X;
    uint8_t* HA = X + 8192;                            // 16384 (H1, then H3, then H4)
    uint8_t* HB = HA;
    float* const sxyz = reinterpret_cast<float*>(HA);            // TMA_IN staging: 2 x 1536 B at the head of the activation tile
    float* const sdir = sxyz + kTile * 3;
    const uint32_t tid = threadIdx.x;

    if (tid == 0) {
        PVD_T(blockIdx.x, 0);
#ifdef PVD_TRACE
        unsigned long long gt; unsigned int smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        PVD_TV(blockIdx.x, 12, gt);
        PVD_TV(blockIdx.x, 13, smid);
#endif
    }
    // TMEM first: the SM does not launch the next CTA of a tcgen05 kernel until the previous one has relinquished its allocation
    // permit (measured: scripts/micro/cta_launch.cu), so anything placed before the alloc delays every later CTA of the SM.
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    auto tile_rows = [&](uint32_t tile) { return min(kTile, M - tile * kTile); };
    auto tma_tile = [&](uint32_t tile) { return TMA_IN && (tile_rows(tile) & 3u) == 0u; };   // 12 B rows: a multiple of 16 bytes
    auto stage_samples = [&](uint32_t tile) {   // thread 0
        const uint32_t bytes = tile_rows(tile) * 12u;
        tc5::mbar_expect_tx(&sbar, 2u * bytes);
        tc5::bulk_g2s(tc5::smem_u32(sxyz), xyzs + 3 * (size_t)tile * kTile, bytes, &sbar);
        tc5::bulk_g2s(tc5::smem_u32(sdir), dirs + 3 * (size_t)tile * kTile, bytes, &sbar);
    };
    if (tid < 32) tc5::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) {
        tc5::mbar_init(&bar, 1);
        tc5::mbar_init(&wbar, 1);
        if (TMA_IN) tc5::mbar_init(&sbar, 1);
        tc5::mbar_fence_init();
        if (TMA_IN && blockIdx.x < n_tiles && tma_tile(blockIdx.x)) stage_samples(blockIdx.x);  // first: the gather waits on it
        // programmatic dependent launch: this kernel may have started under the weight-pack kernel (a step that re-stages its
        // parameters); samples and table are older than that, only the packed tiles must wait for it
        pdl_wait();
        stage_blob_async(smw, a.wblob, PVD_FIELD_WBLOB_BYTES, &wbar);  // TMA: lands while the first tile is gathered
    }
    pdl_launch_dependents();   // the loss kernel may be made resident now; it waits for this grid before it reads sigmas / rgbs
    level_info_init(lv, a.offsets, a.L, a.S, a.H);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    Pipe p{&bar, 0u, tmem_base_s, status};
    p.wbar = &wbar;
    p.trec = blockIdx.x;
    if (tid == 0) PVD_T(blockIdx.x, 1);
    const T* table = reinterpret_cast<const T*>(a.table);
    const uint32_t lv_saddr = tc5::smem_u32(lv);

    uint32_t n_staged = 0;   // TMA-staged tiles so far: the parity of `sbar` to wait for
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row = tile * kTile + tid;
        const bool live = row < M;
        float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
        if (tma_tile(tile)) {
            // a tile after the first: the previous tile's last MMA (the only reader of the activation tile) has been waited for by
            // every thread, so the staging area is free; the first tile's copy was issued in the prologue
            if (tile != blockIdx.x && tid == 0) stage_samples(tile);
            if (!tc5::mbar_wait(&sbar, n_staged & 1u)) atomicExch(status, 4);
            ++n_staged;
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    pos[d] = sxyz[3 * tid + d];
                    dir[d] = sdir[3 * tid + d];
                }
            }
        } else if (live) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
                dir[d] = __ldg(dirs + 3 * (size_t)row + d);
            }
        }
        float x01[3];
        bool oob;
        to_unit(pos, a.bound, x01, oob);
        if (oob) x01[0] = x01[1] = x01[2] = 0.0f;  // keeps the (ignored) gathers of an out-of-range sample inside the table
#pragma unroll 1
        for (uint32_t j = 0; j < 4; ++j) {
            float f[8];
            encode4<T>(table, lv_saddr, 4 * j, a.L, x01, oob, f);
            const uint4 u = tc5::pack8(f);
            *reinterpret_cast<uint4*>(X + tc5::chunk_off(kTile, tid, j)) = u;
            if (enc && live) *reinterpret_cast<uint4*>(enc + (size_t)row * PVD_FIELD_ENC_STRIDE + 8 * j) = u;
            if (tid == 0) PVD_T(blockIdx.x, 2 + j);
        }
        float sigma, o16[16];
        FwdRegs r;
        mlp_forward(p, a, smw, X, HA, CIN, HB, HA, dir, tid, sigma, o16, r);
        if (live) {
            sigmas[row] = sigma;
            rgbs[3 * (size_t)row] = r.rgb[0];
            rgbs[3 * (size_t)row + 1] = r.rgb[1];
            rgbs[3 * (size_t)row + 2] = r.rgb[2];
            if (feat16) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(feat16 + 16 * (size_t)row + 4 * q) =
                        make_float4(o16[4 * q], o16[4 * q + 1], o16[4 * q + 2], o16[4 * q + 3]);
            }
        }
    }
    if (tid == 0) PVD_T(blockIdx.x, 11);
    if (tid == 0) weights_ready(p);  // a CTA without tiles must not exit under its own in-flight bulk copy
    tc5::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc5::tmem_dealloc(p.tmem, 128);
#ifdef PVD_TRACE
    if (tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        PVD_TV(blockIdx.x, 14, gt);
    }
#endif
}

// =============================================================================================== persistent inference kernel
// The evaluation branch of NeRFRenderer.run_cuda (distill_mutual/renderer.py:450-543) is a HOST loop: march_rays (n_step <= 8 samples
// per alive ray) -> NeRFNetwork.forward -> composite_rays -> compact_rays, one D2H read of the alive count per iteration, every
// intermediate ([n_alive * n_step] xyzs / dirs / deltas / sigmas / rgbs) through HBM.  Here the whole loop is ONE persistent kernel
// for a hash field: a CTA holds 16 ray slots x 8 steps = the 128 samples of one tensor-core tile and iterates
//     refill dead slots from a global ray queue (atomicAdd) -> march each live ray by up to 8 samples (the reference's loop,
//     raymarching.cu:705-793) -> gather + tcgen05 MLP for the 128 samples -> composite each ray's 8 samples in order
//     (raymarching.cu:826-909, early termination at T < 1e-4) -> retire finished rays (write weights_sum / depth / image)
// until the queue is empty and every slot has retired.  Nothing but the final per-ray outputs leaves the SM.  A ray's result does not
// depend on how its samples are chunked (the march continues from the composited t, the termination test is per sample), so with
// perturb off -- the reference's evaluation setting -- the outputs equal the host loop's bit for bit.
constexpr uint32_t kRenderRays = 16, kRenderSteps = 8;
static_assert(kRenderRays * kRenderSteps == kTile, "one MLP tile per iteration");

struct RenderArgs {
    const float* rays_o;
    const float* rays_d;
    const uint8_t* grid;
    const float* nears;
    const float* fars;
    float bound, dt_gamma;
    uint32_t max_steps, C, H, N;
};

template <typename T>
__global__ void __launch_bounds__(128) k_hash_render_persistent(FieldArgs a, RenderArgs ra, int32_t* __restrict__ queue,
                                                                float* __restrict__ weights_sum, float* __restrict__ depth,
                                                                float* __restrict__ image, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, wbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ LevelInfo lv[16];
    __shared__ float s_pos[kTile * 3], s_dir[kRenderRays * 3], s_delta[kTile * 2], s_sig[kTile], s_rgb[kTile * 3];
    __shared__ int32_t s_ray[kRenderRays], s_cnt[kRenderRays];
    __shared__ float s_t[kRenderRays], s_ws[kRenderRays], s_d[kRenderRays], s_img[kRenderRays * 3];
    uint8_t* smw = smem;
    uint8_t* X = smem + PVD_FIELD_WBLOB_BYTES;
    uint8_t* CIN = X;
    uint8_t* HA = X + 8192;
    uint8_t* HB = HA;
    const uint32_t tid = threadIdx.x;
    if (tid < 32) tc5::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) {
        tc5::mbar_init(&bar, 1);
        tc5::mbar_init(&wbar, 1);
        tc5::mbar_fence_init();
        stage_blob_async(smw, a.wblob, PVD_FIELD_WBLOB_BYTES, &wbar);
    }
    level_info_init(lv, a.offsets, a.L, a.S, a.H);
    if (tid < kRenderRays) s_ray[tid] = -1;
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    Pipe p{&bar, 0u, tmem_base_s, status};
    p.wbar = &wbar;
    const T* table = reinterpret_cast<const T*>(a.table);
    const uint32_t lv_saddr = tc5::smem_u32(lv);

    for (;;) {
        // ---- refill + march: thread r < 16 owns ray slot r
        int have_samples = 0;
        if (tid < kRenderRays) {
            if (s_ray[tid] == -1) {
                const int32_t idx = atomicAdd(queue, 1);
                if ((uint32_t)idx < ra.N) {
                    s_ray[tid] = idx;
                    s_t[tid] = __ldg(ra.nears + idx);
                    s_ws[tid] = 0.0f; s_d[tid] = 0.0f;
                    s_img[3 * tid] = 0.0f; s_img[3 * tid + 1] = 0.0f; s_img[3 * tid + 2] = 0.0f;
                    s_cnt[tid] = 0;
                } else {
                    s_ray[tid] = -2;   // queue exhausted: this slot stays empty
                }
            }
            uint32_t step = 0;
            float* px = s_pos + 3 * kRenderSteps * tid;
            float* pl = s_delta + 2 * kRenderSteps * tid;
            if (s_ray[tid] >= 0) {
                const int32_t index = s_ray[tid];
                MarchCtx c;
                march_ctx_init(c, ra.rays_o + 3 * (size_t)index, ra.rays_d + 3 * (size_t)index, ra.bound, ra.dt_gamma, ra.max_steps, ra.C, ra.H);
                s_dir[3 * tid] = c.dx; s_dir[3 * tid + 1] = c.dy; s_dir[3 * tid + 2] = c.dz;
                const float far = __ldg(ra.fars + index);
                float t = s_t[tid];
                float last_t = t;
                while (t < far && step < kRenderSteps) {   // raymarching.cu:757-790
                    float x, y, z, tt;
                    march_pos(c, t, x, y, z);
                    const float dt = march_dt(c, t);
                    if (march_probe(c, ra.grid, t, dt, x, y, z, tt)) {
                        px[3 * step] = x; px[3 * step + 1] = y; px[3 * step + 2] = z;
                        t = __fadd_rn(t, dt);
                        pl[2 * step] = dt;
                        pl[2 * step + 1] = __fadd_rn(t, -last_t);
                        last_t = t;
                        ++step;
                    } else {
                        do { t = __fadd_rn(t, march_dt(c, t)); } while (t < tt);
                    }
                }
            }
            have_samples = step > 0;
            for (uint32_t sidx = step; sidx < kRenderSteps; ++sidx) {   // unused slots: zeros, like the reference's fresh buffers
                px[3 * sidx] = 0.0f; px[3 * sidx + 1] = 0.0f; px[3 * sidx + 2] = 0.0f;
                pl[2 * sidx] = 0.0f; pl[2 * sidx + 1] = 0.0f;
            }
            if (s_ray[tid] < 0) { s_dir[3 * tid] = 0.0f; s_dir[3 * tid + 1] = 0.0f; s_dir[3 * tid + 2] = 0.0f; }
        }
        // ---- field query for the 128 samples (skipped when no ray of the CTA produced a sample)
        if (__syncthreads_or(have_samples)) {
            float pos[3], dir[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pos[d] = s_pos[3 * tid + d];
                dir[d] = s_dir[3 * (tid / kRenderSteps) + d];
            }
            float x01[3];
            bool oob;
            to_unit(pos, a.bound, x01, oob);
            if (oob) x01[0] = x01[1] = x01[2] = 0.0f;
#pragma unroll 1
            for (uint32_t j = 0; j < 4; ++j) {
                float f[8];
                encode4<T>(table, lv_saddr, 4 * j, a.L, x01, oob, f);
                *reinterpret_cast<uint4*>(X + tc5::chunk_off(kTile, tid, j)) = tc5::pack8(f);
            }
            float sigma, o16[16];
            FwdRegs r;
            mlp_forward(p, a, smw, X, HA, CIN, HB, HA, dir, tid, sigma, o16, r);
            s_sig[tid] = sigma;
            s_rgb[3 * tid] = r.rgb[0]; s_rgb[3 * tid + 1] = r.rgb[1]; s_rgb[3 * tid + 2] = r.rgb[2];
        }
        __syncthreads();
        // ---- composite each live ray's samples in order, retire finished rays (raymarching.cu:826-909)
        int idle = 1;
        if (tid < kRenderRays) {
            if (s_ray[tid] >= 0) {
                const int32_t index = s_ray[tid];
                const float* ps = s_sig + kRenderSteps * tid;
                const float* pc = s_rgb + 3 * kRenderSteps * tid;
                const float* pl = s_delta + 2 * kRenderSteps * tid;
                float t = s_t[tid], ws = s_ws[tid], d = s_d[tid];
                float r = s_img[3 * tid], g = s_img[3 * tid + 1], b = s_img[3 * tid + 2];
                uint32_t step = 0;
                while (step < kRenderSteps) {
                    if (pl[2 * step] == 0.0f) break;
                    const float alpha = 1.0f - __expf(-ps[step] * pl[2 * step]);
                    const float Tr = 1.0f - ws;
                    const float w = alpha * Tr;
                    ws += w;
                    t += pl[2 * step + 1];
                    d += w * t;
                    r += w * pc[3 * step];
                    g += w * pc[3 * step + 1];
                    b += w * pc[3 * step + 2];
                    if (Tr < 1e-4f) break;
                    ++step;
                }
                s_cnt[tid] += (int32_t)kRenderSteps;
                const bool dead = (step < kRenderSteps) || ((uint32_t)s_cnt[tid] >= ra.max_steps);
                if (dead) {
                    weights_sum[index] = ws;
                    depth[index] = d;
                    image[3 * (size_t)index] = r;
                    image[3 * (size_t)index + 1] = g;
                    image[3 * (size_t)index + 2] = b;
                    s_ray[tid] = -1;
                } else {
                    s_t[tid] = t; s_ws[tid] = ws; s_d[tid] = d;
                    s_img[3 * tid] = r; s_img[3 * tid + 1] = g; s_img[3 * tid + 2] = b;
                }
            }
            idle = (s_ray[tid] == -2);
        }
        if (__syncthreads_and(idle)) break;   // queue exhausted and every slot retired
    }
    if (tid == 0) weights_ready(p);
    tc5::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc5::tmem_dealloc(p.tmem, 128);
}

// Stand-alone scatter of d(encoding) [M,32] fp16 into the fp32 table gradient: one thread per (sample, level), 256-thread CTAs,
// ~40 registers -> full occupancy, so the reductions' issue latency is hidden by other warps instead of stalling a 128-thread
// MLP CTA that also holds 88 KB of shared memory and 256 TMEM columns (gridencoder.cu:227-314 semantics, fp32 accumulation).
constexpr uint32_t kAggMaxRes1 = 700;  // aggregate runs on levels whose resolution is below this (cell edge > ~0.85 dt at 1024 steps)

// One warp = 32 consecutive samples at ONE level: reductions of w * g into the level's gradient slice (gridencoder.cu:227-314
// semantics, fp32 accumulation).  `active` lanes carry a sample with a non-zero gradient `g` inside the unit cube.
__device__ __forceinline__ void scatter_level(const LevelInfo& v, const float (&x01)[3], bool active, float2 g, uint32_t lane,
                                              float* __restrict__ grad_table) {
    Corners c;
    level_corners(v, x01, c);
    float* gt = grad_table + (size_t)v.offset * 2;
    // Coarse levels: consecutive samples of a ray share a cell for many steps (cell edge / dt = 37 samples at level 0), so
    // most of the 8 x 32 reductions of a warp would hit the same 8 addresses and serialise in the L2 atomic unit.  Lanes of a run
    // of equal cells are summed with a segmented warp scan first and only the head lane of each run issues reductions.
    if (v.res1 > kAggMaxRes1) {
        if (active) {
#pragma unroll
            for (uint32_t i = 0; i < 8; i += 2)
                tab_red_pair(gt, c.idx[i], c.idx[i + 1], c.w[i] * g.x, c.w[i] * g.y, c.w[i + 1] * g.x, c.w[i + 1] * g.y);
        }
        return;
    }
    // the lower corner identifies the cell exactly (indices of a dense level are unique; for a hashed level compare all 8)
    uint32_t key0 = active ? c.idx[0] : 0xffffffffu, key7 = active ? c.idx[7] : 0xffffffffu;
    const uint32_t p0 = __shfl_up_sync(0xffffffffu, key0, 1), p7 = __shfl_up_sync(0xffffffffu, key7, 1);
    const uint32_t p3 = __shfl_up_sync(0xffffffffu, c.idx[3], 1), p5 = __shfl_up_sync(0xffffffffu, c.idx[5], 1);
    const bool head = (lane == 0) || !active || (p0 != key0) || (p7 != key7) || (p3 != c.idx[3]) || (p5 != c.idx[5]);
    const uint32_t head_mask = __ballot_sync(0xffffffffu, head);
    float val[16];
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        val[2 * i] = active ? c.w[i] * g.x : 0.0f;
        val[2 * i + 1] = active ? c.w[i] * g.y : 0.0f;
    }
    const uint32_t after = (lane == 31u) ? 0u : (head_mask >> (lane + 1u));  // head flags of the lanes above this one
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const bool same = (lane + d < 32u) && ((after & ((1u << d) - 1u)) == 0u);  // lanes lane+1 .. lane+d start no new run
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) {
            const float o = __shfl_down_sync(0xffffffffu, val[i], d);
            if (same) val[i] += o;
        }
    }
    if (head && active) {
#pragma unroll
        for (uint32_t i = 0; i < 8; i += 2)
            tab_red_pair(gt, c.idx[i], c.idx[i + 1], val[2 * i], val[2 * i + 1], val[2 * i + 2], val[2 * i + 3]);
    }
}


// =============================================================================================== backward kernel

// FUSE = true: 128 + 32 * kScatterWarps threads.  Threads 0-127 are the MLP group (everything below, on named barrier 1); the rest
// are SCATTER warps: warp w takes the 32 samples [32 (w % 4), +32) of every tile through the levels l = w / 4 (mod kScatterWarps/4)
// (scatter_level) while the MLP group is already in the tensor-core chain of the CTA's next tile.
// Measured on B200 (hash workload, 73 k samples): two launches 31 + 31 us; fused with 4 scatter warps 54.9 us; with 12 scatter
// warps and the register file re-partitioned by setmaxnreg (launch: 64 registers per thread; scatter warpgroups 40, MLP warpgroup
// 128 -- ptxas honours it, USETMAXREG in the SASS) 54.2 us.  The reductions are bound by a chip-level resource (the L2 atomic
// units: 8.2 M sector reductions per step), not by the number of warps issuing them, so the simple variant is the default; the
// gain over two launches is the MLP chain of tile i+1 running under the reductions of tile i.  d(encoding) changes hands in shared memory (two 8 KB buffers,
// dx_full / dx_empty mbarriers): the scatter is bound by the SM's reduction throughput, the MLP chain by tcgen05 latency, so the
// two overlap almost perfectly and neither d(encoding) nor a second launch touches memory.
constexpr uint32_t kScatterWarps = 4;    // FUSE: scatter warps beside the MLP warpgroup (4, 8 or 12; see the comment below)
template <typename T, bool FUSE>
__global__ void __launch_bounds__(FUSE ? 128 + 32 * kScatterWarps : 128, 2) k_hash_field_bwd(FieldArgs a, const float* __restrict__ xyzs, const float* __restrict__ dirs,
                                                        const __half* __restrict__ enc, const float* __restrict__ grad_sigmas,
                                                        const float* __restrict__ grad_rgbs, const float* __restrict__ grad_feat,
                                                        uint32_t row0, uint32_t M, const int32_t* __restrict__ n_valid_p,
                                                        __half* __restrict__ dx_out, float* __restrict__ grad_table,
                                                        float* __restrict__ gw, int32_t* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, wbar, dx_full[2], dx_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ LevelInfo lv[16];
    uint8_t* smw = smem;                        // 20480
    uint8_t* X = smem + PVD_FIELD_WBLOB_BYTES;  // 8192
    uint8_t* CIN = X + 8192;                    // 8192
    uint8_t* H1 = CIN + 8192;                   // 16384
    uint8_t* H3 = H1 + 16384;                   // 16384
    uint8_t* H4 = H3 + 16384;                   // 16384
    uint8_t* G16 = H4 + 16384;                  // 4096
    uint8_t* DXB = G16 + 4096;                  // FUSE: 2 x 8192, d(encoding) tiles handed to the scatter warps
    const uint32_t tid = threadIdx.x;
    const uint32_t lane_base = ((tid >> 5) & 3u) * 32;

    // TMEM first: the SM does not launch the next CTA of a tcgen05 kernel until the previous one has relinquished its allocation
    // permit (measured: scripts/micro/cta_launch.cu), so anything placed before the alloc delays every later CTA of the SM.
    if (tid < 32) tc5::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) PVD_T(2048u + blockIdx.x, 0);
    if (tid == 0) {
        tc5::mbar_init(&bar, 1);
        tc5::mbar_init(&wbar, 1);
        if (FUSE) {
            for (int q = 0; q < 2; ++q) {
                tc5::mbar_init(&dx_full[q], 128);
                tc5::mbar_init(&dx_empty[q], 32 * kScatterWarps);
            }
        }
        tc5::mbar_fence_init();
        stage_blob_async(smw, a.wblob, PVD_FIELD_WBLOB_BYTES, &wbar);  // TMA: lands while the first tile's rows are loaded
    }
    level_info_init(lv, a.offsets, a.L, a.S, a.H);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    Pipe p{&bar, 0u, tmem_base_s, status};
    p.wbar = &wbar;
    p.trec = 2048u + blockIdx.x;  // PVD_TRACE timeline record of this CTA (the last tile it processes wins)
    if (FUSE) p.team = 1u;
    const bool leader = tid == 0;
    if (tid == 0) PVD_T(p.trec, 1);
    const uint32_t trow = tc5::tmem_addr(p.tmem, lane_base, 0);
    const uint32_t sw = tc5::smem_u32(smw);
    // this launch covers rows [row0, row0 + M) of the sample buffers, cut at n_valid (rows of padding contribute nothing)
    const uint32_t row_end = n_valid_p ? min((uint32_t)max(*n_valid_p, 0), row0 + M) : row0 + M;
    const uint32_t n_valid = row_end;
    const uint32_t lv_saddr = tc5::smem_u32(lv);

    const uint32_t n_tiles = (row_end > row0) ? (row_end - row0 + kTile - 1) / kTile : 0u;
    bool first = true;
    if (FUSE && tid >= 128) {
        // ------------------------------------------------------------------ scatter warps
        if (kScatterWarps > 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const uint32_t sw = (tid - 128u) >> 5, lane = tid & 31u;
        const uint32_t sgroup = sw & 3u, lsel = sw >> 2;      // sample group of the tile, level residue class
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t r = sgroup * 32u + lane;           // row inside the tile
            const uint32_t srow = row0 + tile * kTile + r;
            bool in = srow < n_valid;
            float x01[3] = {0.5f, 0.5f, 0.5f};
            if (in) {
                float pos[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) pos[d] = __ldg(xyzs + 3 * (size_t)srow + d);
                bool oob;
                to_unit(pos, a.bound, x01, oob);
                in = !oob;
                if (oob) x01[0] = x01[1] = x01[2] = 0.5f;
            }
            const uint32_t bsel = it & 1u;
            if (!tc5::mbar_wait(&dx_full[bsel], (it >> 1) & 1u)) atomicExch(status, 3);
            const uint8_t* dxt = DXB + bsel * 8192u;
#pragma unroll 1
            for (uint32_t level = lsel; level < a.L; level += kScatterWarps / 4u) {
                const __half2 gh = *reinterpret_cast<const __half2*>(dxt + tc5::chunk_off(kTile, r, level >> 2) + (level & 3u) * 4u);
                const float2 g = __half22float2(gh);
                const bool active = in && !(g.x == 0.0f && g.y == 0.0f);
                scatter_level(ld_level(lv_saddr, level), x01, active, g, lane, grad_table);
            }
            tc5::mbar_arrive(&dx_empty[bsel]);
        }
    } else {
    if (FUSE && kScatterWarps > 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    bool pdl_pending = true;
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t row = row0 + tile * kTile + tid;
        const bool live = row < n_valid;
        float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
        float gsig = 0.0f, grgb[3] = {0.f, 0.f, 0.f};
        auto load_grads = [&]() {   // what the preceding loss kernel wrote
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; ++d) grgb[d] = __ldg(grad_rgbs + 3 * (size_t)row + d);
                gsig = __ldg(grad_sigmas + row);
            }
        };
        if (live) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pos[d] = __ldg(xyzs + 3 * (size_t)row + d);
                dir[d] = __ldg(dirs + 3 * (size_t)row + d);
            }
        }
        // Programmatic dependent launch: this kernel may have started while the loss kernel (composite / pair combine) was still
        // running.  Its first tile's forward recomputation needs only the samples and the saved encoding (older kernels); the
        // upstream gradients are read after pdl_wait().  Later tiles issue the loads early, under the recomputation.
        if (!pdl_pending) load_grads();
        // saved encoding -> X tile
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (live) u = __ldg(reinterpret_cast<const uint4*>(enc + (size_t)row * PVD_FIELD_ENC_STRIDE + 8 * j));
            *reinterpret_cast<uint4*>(X + tc5::chunk_off(kTile, tid, j)) = u;
        }
        if (tid == 0) PVD_T(p.trec, 2);
        float sigma, o16[16];
        FwdRegs r;
        mlp_forward(p, a, smw, X, H1, CIN, H3, H4, dir, tid, sigma, o16, r);
        if (pdl_pending) {
            pdl_wait();
            pdl_pending = false;
            load_grads();
        }

        // ---- d(color_net.2 pre-activation) = grad_rgb * rgb * (1 - rgb)
        {
            float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 3; ++i) g[i] = grgb[i] * r.rgb[i] * (1.0f - r.rgb[i]);
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 0)) = tc5::pack8(g);
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 1)) = make_uint4(0, 0, 0, 0);
        }
        operands_ready(p);
        if (leader) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kAW5, tc5::smem_u32(H4), tc5::smem_u32(G16), 16, first);   // dW5^T += H4^T G5
            issue_dgrad(p.tmem + kD, tc5::smem_u32(G16), 16, sw + kWB5, 16, 64);             // dH4 = G5 W5
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        if (tid == 0) PVD_T(p.trec, 3);
        mask_grad_in_place(trow + kD, H4, tid);                                              // G4 (over H4)
        operands_ready(p);
        if (leader) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kAW4, tc5::smem_u32(H4), tc5::smem_u32(H3), 64, first);    // dW4 += G4^T H3
            issue_dgrad(p.tmem + kD, tc5::smem_u32(H4), 64, sw + kWB4, 64, 64);              // dH3 = G4 W4
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        if (tid == 0) PVD_T(p.trec, 4);
        mask_grad_in_place(trow + kD, H3, tid);                                              // G3 (over H3)
        operands_ready(p);
        if (leader) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kAW3, tc5::smem_u32(H3), tc5::smem_u32(CIN), 32, first);   // dW3 += G3^T CIN
            issue_dgrad(p.tmem + kD, tc5::smem_u32(H3), 64, sw + kWB3, 64, 32);              // dCIN = G3 W3
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        if (tid == 0) PVD_T(p.trec, 5);
        // ---- d(sigma_net.1 output): channel 0 through trunc_exp + clamp, channels 1..15 = geo part of dCIN
        {
            float dc[16];
            tc5::tmem_ld16(trow + kD + 16, dc);  // columns 16..31 = d(geo 0..14), pad
            float g[16];
            const bool inside = (r.o0_raw >= a.clip_min) && (r.o0_raw <= a.clip_max);  // clamp backward
            // trunc_exp backward: g * exp(clamp(x, -12, 12)) (tools/activation.py:15-21)
            g[0] = gsig * a.density_scale * __expf(clampf(r.o0c, -12.0f, 12.0f));
#pragma unroll
            for (int i = 0; i < 15; ++i) g[i + 1] = dc[i];
            if (grad_feat && live) {  // d(loss)/d(feature_sigma_color): the distillation feature / sigma_l losses
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 gf = __ldg(reinterpret_cast<const float4*>(grad_feat + 16 * (size_t)row) + q);
                    g[4 * q] += gf.x; g[4 * q + 1] += gf.y; g[4 * q + 2] += gf.z; g[4 * q + 3] += gf.w;
                }
            }
            if (!inside) g[0] = 0.0f;  // clamp backward (network.py:418-420)
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 0)) = tc5::pack8(g);
            *reinterpret_cast<uint4*>(G16 + tc5::chunk_off(kTile, tid, 1)) = tc5::pack8(g + 8);
        }
        operands_ready(p);
        if (leader) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kAW2, tc5::smem_u32(H1), tc5::smem_u32(G16), 16, first);   // dW2^T += H1^T G2
            issue_dgrad(p.tmem + kD, tc5::smem_u32(G16), 16, sw + kWB2, 16, 64);             // dH1 = G2 W2
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        if (tid == 0) PVD_T(p.trec, 11);
        mask_grad_in_place(trow + kD, H1, tid);                                              // G1 (over H1)
        operands_ready(p);
        if (leader) {
            tc5::fence_after_sync();
            issue_wgrad(p.tmem + kAW1, tc5::smem_u32(H1), tc5::smem_u32(X), 32, first);     // dW1 += G1^T X
            issue_dgrad(p.tmem + kD, tc5::smem_u32(H1), 64, sw + kWB1, 64, 32);              // dX = G1 W1
            tc5::mma_commit(p.bar);
        }
        mma_wait(p);
        if (tid == 0) PVD_T(p.trec, 12);
        first = false;
        // ---- scatter d(encoding) into the table gradient (gridencoder.cu:227-314 semantics, fp32 accumulation)
        float dx[32];
        tc5::tmem_ld16(trow + kD, *reinterpret_cast<float(*)[16]>(&dx[0]));
        tc5::tmem_ld16(trow + kD + 16, *reinterpret_cast<float(*)[16]>(&dx[16]));
        if (FUSE) {  // hand d(encoding) to this CTA's scatter warps through shared memory
            const uint32_t bsel = it & 1u;
            if (it >= 2u && !tc5::mbar_wait(&dx_empty[bsel], ((it >> 1) - 1u) & 1u)) atomicExch(status, 3);
            uint8_t* dxt = DXB + bsel * 8192u;
            float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(dxt + tc5::chunk_off(kTile, tid, j)) = tc5::pack8(live ? dx + 8 * j : z);
            tc5::mbar_arrive(&dx_full[bsel]);
        } else if (dx_out != nullptr) {  // split mode: hand d(encoding) to the stand-alone, high-occupancy scatter kernel
            if (live) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(dx_out + (size_t)row * PVD_FIELD_ENC_STRIDE + 8 * j) = tc5::pack8(dx + 8 * j);
            }
        } else if (live) {
            float x01[3];
            bool oob;
            to_unit(pos, a.bound, x01, oob);
            if (!oob) {
#pragma unroll
                for (uint32_t l = 0; l < 16; ++l) {
                    if (l < a.L) {
                        const LevelInfo v = ld_level(lv_saddr, l);
                        Corners c;
                        level_corners(v, x01, c);
                        float* gt = grad_table + (size_t)v.offset * 2;
                        const float g0 = dx[2 * l], g1 = dx[2 * l + 1];
#pragma unroll
                        for (uint32_t i = 0; i < 8; ++i) tab_red2(gt, (size_t)c.idx[i] * 2, c.w[i] * g0, c.w[i] * g1);
                    }
                }
            }
        }
    }
    if (tid == 0) PVD_T(p.trec, 13);
    if (tid == 0) weights_ready(p);  // a CTA without tiles must not exit under its own in-flight bulk copy
    // ---- weight gradients leave the SM once per CTA
    operands_ready(p);
    tc5::fence_after_sync();
    gw += (size_t)(blockIdx.x % PVD_FIELD_GW_COPIES) * PVD_FIELD_GW_FLOATS;
    if (!first) {
        flush_acc(p.tmem, kAW1, 32, gw + kGW1);
        flush_acc(p.tmem, kAW2, 16, gw + kGW2);
        flush_acc(p.tmem, kAW3, 32, gw + kGW3);
        flush_acc(p.tmem, kAW4, 64, gw + kGW4);
        flush_acc(p.tmem, kAW5, 16, gw + kGW5);
    }
    if (tid == 0) PVD_T(p.trec, 14);
    }  // MLP group
    tc5::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc5::tmem_dealloc(p.tmem, 256);
}

__global__ void __launch_bounds__(256) k_hash_scatter(FieldArgs a, const float* __restrict__ xyzs, const __half* __restrict__ dx,
                                                      uint32_t row0, uint32_t M, const int32_t* __restrict__ n_valid_p,
                                                      float* __restrict__ grad_table) {
    __shared__ LevelInfo lvs;
    const uint32_t level = blockIdx.y;
    if (threadIdx.x == 0) {
        const GridLevel g = grid_level(a.offsets, level, a.S, a.H);
        LevelInfo v;
        v.scale = g.scale; v.res1 = g.resolution + 1; v.offset = g.offset; v.size = g.size; v.mask = g.size - 1;
        const uint64_t dense = (uint64_t)v.res1 * v.res1 * v.res1;
        const bool fits = ((uint64_t)v.res1 <= g.size) && ((uint64_t)v.res1 * v.res1 <= g.size) && (dense <= g.size);
        v.mode = fits ? 0u : (((g.size & (g.size - 1)) == 0) ? 1u : 2u);
        v.pad0 = v.pad1 = 0;
        lvs = v;
    }
    __syncthreads();
    const uint32_t n_valid = n_valid_p ? min((uint32_t)max(*n_valid_p, 0), row0 + M) : row0 + M;
    const uint32_t b = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const LevelInfo v = lvs;
    bool active = b < n_valid;
    float2 g = make_float2(0.f, 0.f);
    float x01[3] = {0.5f, 0.5f, 0.5f};
    if (active) {
        g = __half22float2(__ldg(reinterpret_cast<const __half2*>(dx + (size_t)b * PVD_FIELD_ENC_STRIDE + 2 * level)));
        float pos[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) pos[d] = __ldg(xyzs + 3 * (size_t)b + d);
        bool oob;
        to_unit(pos, a.bound, x01, oob);
        active = !oob && !(g.x == 0.0f && g.y == 0.0f);
    }
    scatter_level(v, x01, active, g, lane, grad_table);
}

__global__ void k_pack_weights(const float* __restrict__ ws0, const float* __restrict__ ws1, const float* __restrict__ wc0,
                               const float* __restrict__ wc1, const float* __restrict__ wc2, uint32_t in_dim,
                               uint8_t* __restrict__ blob) {
    pdl_launch_dependents();   // the forward behind it (launch_pdl) gathers its first tile meanwhile and waits before it stages the tiles
    pack_matrix(ws0, 64, in_dim, blob + kWB1, 64, 32);
    pack_matrix(ws1, 16, 64, blob + kWB2, 16, 64);
    pack_matrix(wc0, 64, 31, blob + kWB3, 64, 32);
    pack_matrix(wc1, 64, 64, blob + kWB4, 64, 64);
    pack_matrix(wc2, 3, 64, blob + kWB5, 16, 64);
}

// un-pad / transpose the kernel-native weight-gradient workspace into the parameter shapes (accumulating)
__global__ void k_unpack_wgrads(const float* __restrict__ gw, uint32_t in_dim, float* __restrict__ g0, float* __restrict__ g1,
                                float* __restrict__ g2, float* __restrict__ g3, float* __restrict__ g4) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    auto sum = [&](uint32_t i) {
        float acc = 0.0f;
#pragma unroll
        for (uint32_t c = 0; c < PVD_FIELD_GW_COPIES; ++c) acc += gw[(size_t)c * PVD_FIELD_GW_FLOATS + i];
        return acc;
    };
    if (t < 64 * in_dim) { const uint32_t o = t / in_dim, i = t - o * in_dim; g0[t] += sum(kGW1 + o * 32 + i); }
    if (t < 16 * 64) { const uint32_t o = t / 64, i = t - o * 64; g1[t] += sum(kGW2 + i * 16 + o); }
    if (t < 64 * 31) { const uint32_t o = t / 31, i = t - o * 31; g2[t] += sum(kGW3 + o * 32 + i); }
    if (t < 64 * 64) { g3[t] += sum(kGW4 + t); }
    if (t < 3 * 64) { const uint32_t o = t / 64, i = t - o * 64; g4[t] += sum(kGW5 + i * 16 + o); }
}

static FieldArgs to_args(const PvdHashField* f) {
    FieldArgs a;
    a.table = f->table;
    a.offsets = f->offsets;
    a.wblob = reinterpret_cast<const uint8_t*>(f->wblob);
    a.L = f->L;
    a.H = f->H;
    a.S = f->S;
    a.bound = f->bound;
    a.clip_min = f->sigma_clip_min;
    a.clip_max = f->sigma_clip_max;
    a.density_scale = f->density_scale;
    return a;
}



constexpr size_t kFwdSmem = PVD_FIELD_WBLOB_BYTES + 8192 + 16384;                          // 45056
constexpr size_t kBwdSmem = PVD_FIELD_WBLOB_BYTES + 8192 + 8192 + 16384 * 3 + 4096;         // 90112

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_field_pack_weights(const float* w_sigma0, const float* w_sigma1, const float* w_color0, const float* w_color1,
                           const float* w_color2, uint32_t in_dim, void* wblob, void* stream) {
    PVD_REQUIRE(w_sigma0 && w_sigma1 && w_color0 && w_color1 && w_color2 && wblob);
    if (in_dim == 0 || in_dim > 32 || (in_dim & 1u)) return PVD_EUNSUPPORTED;
    k_pack_weights<<<16, 256, 0, (cudaStream_t)stream>>>(w_sigma0, w_sigma1, w_color0, w_color1, w_color2, in_dim, (uint8_t*)wblob);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_hash_field_forward(const PvdHashField* f, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                           void* enc, float* feat16, int32_t* status, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(f && f->table && f->offsets && f->wblob && xyzs && dirs && sigmas && rgbs && status);
    if (f->L == 0 || f->L > 16) return PVD_EUNSUPPORTED;
    const FieldArgs a = to_args(f);
    const uint32_t tiles = (M + kTile - 1) / kTile;
    const uint32_t grid = min(tiles, (uint32_t)(4 * sm_count()));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    // sample blocks through TMA bulk copies (default) or per-thread loads (PVD_FWD_TMA=0, and whenever the buffers are not
    // 16-byte aligned)
    static const bool want_tma = []() { const char* v = getenv("PVD_FWD_TMA"); return !(v != nullptr && v[0] == '0'); }();
    const bool tma = want_tma && ((reinterpret_cast<uintptr_t>(xyzs) | reinterpret_cast<uintptr_t>(dirs)) & 15u) == 0;
    auto launch = [&](auto kern) -> int {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
        if (e != cudaSuccess) return (int)e;
        e = launch_pdl(kern, dim3(grid), dim3(128), kFwdSmem, st, a, xyzs, dirs, M, sigmas, rgbs, (__half*)enc, feat16, status);
        return e == cudaSuccess ? PVD_OK : (int)e;
    };
    int rc;
    if (f->table_dtype == PVD_DTYPE_F16) {
        rc = tma ? launch(k_hash_field_fwd<__half, true>) : launch(k_hash_field_fwd<__half, false>);
    } else if (f->table_dtype == PVD_DTYPE_F32) {
        rc = tma ? launch(k_hash_field_fwd<float, true>) : launch(k_hash_field_fwd<float, false>);
    } else {
        return PVD_EUNSUPPORTED;
    }
    if (rc != PVD_OK) return rc;
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

static int hash_backward_rows(const PvdHashField* f, const float* xyzs, const float* dirs, const void* enc, const float* grad_sigmas,
                              const float* grad_rgbs, const float* grad_feat16, uint32_t row0, uint32_t rows, const int32_t* n_valid,
                              float* grad_table, float* gw_ws, void* dx_ws, int32_t* status, uint32_t phases, void* stream) {
    if (rows == 0) return PVD_OK;
    PVD_REQUIRE(f && f->offsets && f->wblob && xyzs && dirs && enc && grad_sigmas && grad_rgbs && grad_table && gw_ws && status);
    if (f->L == 0 || f->L > 16) return PVD_EUNSUPPORTED;
    const FieldArgs a = to_args(f);
    cudaStream_t st = (cudaStream_t)stream;
    if (phases & PVD_BWD_MLP) {
        const uint32_t tiles = (rows + kTile - 1) / kTile;
        const uint32_t grid = min(tiles, (uint32_t)(2 * sm_count()));
        // the table is not read in the backward (the encoding was saved); one instantiation serves both table dtypes
        if (dx_ws == nullptr) {  // one launch: the CTA's scatter warps reduce tile i while its MLP warps are in tile i+1
            cudaError_t e = cudaFuncSetAttribute(k_hash_field_bwd<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(kBwdSmem + 16384));
            if (e != cudaSuccess) return (int)e;
            e = launch_pdl(k_hash_field_bwd<float, true>, dim3(grid), dim3(128 + 32 * kScatterWarps), kBwdSmem + 16384, st, a, xyzs, dirs,
                           (const __half*)enc, grad_sigmas, grad_rgbs, grad_feat16, row0, rows, n_valid, (__half*)nullptr, grad_table, gw_ws, status);
            if (e != cudaSuccess) return (int)e;
        } else {
            cudaError_t e = cudaFuncSetAttribute(k_hash_field_bwd<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
            if (e != cudaSuccess) return (int)e;
            e = launch_pdl(k_hash_field_bwd<float, false>, dim3(grid), dim3(128), kBwdSmem, st, a, xyzs, dirs, (const __half*)enc, grad_sigmas,
                           grad_rgbs, grad_feat16, row0, rows, n_valid, (__half*)dx_ws, grad_table, gw_ws, status);
            if (e != cudaSuccess) return (int)e;
        }
        PVD_LAUNCH_CHECK();
    }
    if ((phases & PVD_BWD_SCATTER) && dx_ws != nullptr) {
        const dim3 sgrid((rows + 255) / 256, f->L, 1);
        k_hash_scatter<<<sgrid, 256, 0, st>>>(a, xyzs, (const __half*)dx_ws, row0, rows, n_valid, grad_table);
        PVD_LAUNCH_CHECK();
    }
    return PVD_OK;
}

int pvd_hash_field_backward(const PvdHashField* f, const float* xyzs, const float* dirs, const void* enc,
                            const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                            const int32_t* n_valid, float* grad_table, float* gw_ws, void* dx_ws, int32_t* status, void* stream) {
    return hash_backward_rows(f, xyzs, dirs, enc, grad_sigmas, grad_rgbs, grad_feat16, 0u, M, n_valid, grad_table, gw_ws, dx_ws, status,
                              PVD_BWD_MLP | PVD_BWD_SCATTER, stream);
}

int pvd_hash_field_backward_rows(const PvdHashField* f, const float* xyzs, const float* dirs, const void* enc,
                                 const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t row0,
                                 uint32_t rows, const int32_t* n_valid, float* grad_table, float* gw_ws, void* dx_ws, int32_t* status,
                                 uint32_t phases, void* stream) {
    PVD_REQUIRE(dx_ws != nullptr && (phases & (PVD_BWD_MLP | PVD_BWD_SCATTER)) != 0);
    return hash_backward_rows(f, xyzs, dirs, enc, grad_sigmas, grad_rgbs, grad_feat16, row0, rows, n_valid, grad_table, gw_ws, dx_ws,
                              status, phases, stream);
}

int pvd_hash_render_persistent(const PvdHashField* f, const float* rays_o, const float* rays_d, const uint8_t* grid, const float* nears,
                               const float* fars, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, uint32_t N,
                               int32_t* queue, float* weights_sum, float* depth, float* image, int32_t* status, void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(f && f->table && f->offsets && f->wblob && rays_o && rays_d && grid && nears && fars && queue && weights_sum && depth && image && status);
    PVD_REQUIRE(C >= 1 && C <= 16 && H >= 1 && H <= 1024 && max_steps >= 1);
    if (f->L == 0 || f->L > 16) return PVD_EUNSUPPORTED;
    const FieldArgs a = to_args(f);
    RenderArgs ra{rays_o, rays_d, grid, nears, fars, bound, dt_gamma, max_steps, C, H, N};
    const uint32_t groups = (N + kRenderRays - 1) / kRenderRays;
    const uint32_t blocks = min(groups, (uint32_t)(4 * sm_count()));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (f->table_dtype == PVD_DTYPE_F16) {
        e = cudaFuncSetAttribute(k_hash_render_persistent<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
        if (e != cudaSuccess) return (int)e;
        k_hash_render_persistent<__half><<<blocks, 128, kFwdSmem, st>>>(a, ra, queue, weights_sum, depth, image, status);
    } else if (f->table_dtype == PVD_DTYPE_F32) {
        e = cudaFuncSetAttribute(k_hash_render_persistent<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
        if (e != cudaSuccess) return (int)e;
        k_hash_render_persistent<float><<<blocks, 128, kFwdSmem, st>>>(a, ra, queue, weights_sum, depth, image, status);
    } else {
        return PVD_EUNSUPPORTED;
    }
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_field_unpack_wgrads(const float* gw_ws, uint32_t in_dim, float* gw_sigma0, float* gw_sigma1, float* gw_color0,
                            float* gw_color1, float* gw_color2, void* stream) {
    PVD_REQUIRE(gw_ws && gw_sigma0 && gw_sigma1 && gw_color0 && gw_color1 && gw_color2);
    if (in_dim == 0 || in_dim > 32) return PVD_EUNSUPPORTED;
    k_unpack_wgrads<<<16, 256, 0, (cudaStream_t)stream>>>(gw_ws, in_dim, gw_sigma0, gw_sigma1, gw_color0, gw_color1, gw_color2);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
