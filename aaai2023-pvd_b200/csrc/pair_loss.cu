// pair_loss.cu -- the loss side of the (teacher, student) distillation step at SHARED samples.
//
// Reference: Trainer.train_step of distill_mutual/utils.py:954-1189.  Student and teacher are rendered on the same samples
// (renderer.py:374-394); with the default loss_type "normL2" (main_distill_mutual.py, get_loss :940-951) the stage-3 loss is
//     loss = rate_rgb   * || pred_tea - pred_stu ||_2          over [N,3]   (:1110-1111; pred = image + (1 - ws) * bg, renderer.py:445)
//          + rate_fea   * || feat_stu - feat_tea ||_2          over [M,16]  (:1137-1149, `feature_sigma_color`)
//          + rate_color * || color_l_stu - color_l_tea ||_2    over [M,3]   (:1158-1165, `color_l` = per-sample rgb)
//          + rate_sigma * || sigma_l_stu - sigma_l_tea ||_2    over [M]     (:1166-1173, `sigma_l` = feat[..., 0])
// (stage 1 keeps only the feature term, stage 2 drops the rgb term: :1046-1108).  The reference evaluates it with ~40 elementwise /
// reduction launches over the [M,16] / [M,3] / [M] activations of BOTH networks plus two composite launches and their autograd
// mirrors.  Every term is a global L2 norm, d||x|| / dx = x / ||x||, so the step needs the four sums of squares before any gradient
// can be final.  Here:
//   k_pair_sample_sq   one pass over the per-sample outputs of both fields -> sums of squares (feature, colour, sigma), 64 slots;
//   k_pair_composite   one warp per ray: composites teacher AND student in the same sweep (deltas read once), accumulates
//                      || pred_s - pred_t ||^2, then sweeps the ray again for the student's sample gradients with the UN-normalised
//                      upstream (pred_s - pred_t) -- the 1/norm factor is uniform over the batch and is applied afterwards;
//   k_pair_combine     one pass over the samples: reads the sums, forms the four 1/norm coefficients, and writes the final
//                      grad_sigmas / grad_rgbs / grad_feat16 the student's fused backward consumes (x loss_scale, GradScaler).
// Padding rows (rows >= the march's sample count, up to M) are evaluated by both networks at xyz = dir = 0 and DO enter the
// per-sample terms, as in the reference (raymarching.py:240-242 zero-fills them and network.forward sees all M rows);
// k_zero_sample_tail restores those zeros, because the engine's persistent sample buffers would otherwise keep stale rows.
#include "common.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

constexpr uint32_t kSumRgb = 0, kSumFea = 1, kSumColor = 2, kSumSigma = 3;  // float index inside a PVD_PAIR_SUM_STRIDE slot

__device__ __forceinline__ float block_sum_256(float v, float* sh /* [8] */) {
    v = warp_sum(v);
    const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31u;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.0f;
    if (w == 0) t = warp_sum(t);
    return t;  // valid in warp 0
}

__global__ void __launch_bounds__(256) k_pair_sample_sq(const float* __restrict__ feat_t, const float* __restrict__ feat_s,
                                                        const float* __restrict__ rgb_t, const float* __restrict__ rgb_s, uint32_t M,
                                                        float* __restrict__ sums) {
    __shared__ float sh[8];
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    float fea = 0.0f, col = 0.0f, sig = 0.0f;
    if (row < M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(feat_s + 16 * (size_t)row) + q);
            const float4 b = __ldg(reinterpret_cast<const float4*>(feat_t + 16 * (size_t)row) + q);
            const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
            if (q == 0) sig = dx * dx;
            fea += dx * dx + dy * dy + dz * dz + dw * dw;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = __ldg(rgb_s + 3 * (size_t)row + c) - __ldg(rgb_t + 3 * (size_t)row + c);
            col += d * d;
        }
    }
    fea = block_sum_256(fea, sh);
    col = block_sum_256(col, sh);
    sig = block_sum_256(sig, sh);
    if (threadIdx.x == 0) {
        float* slot = sums + PVD_PAIR_SUM_STRIDE * (blockIdx.x % PVD_LOSS_SLOTS);
        atomicAdd(slot + kSumFea, fea);
        atomicAdd(slot + kSumColor, col);
        atomicAdd(slot + kSumSigma, sig);
    }
}

// one warp per ray (= per CTA: every branch is warp-uniform, as in k_composite_train_mse)
__global__ void __launch_bounds__(32) k_pair_composite(const float* __restrict__ bg, const float* __restrict__ sig_t,
                                                      const float* __restrict__ rgb_t, const float* __restrict__ sig_s,
                                                      const float* __restrict__ rgb_s, const float* __restrict__ deltas,
                                                      const int32_t* __restrict__ rays, uint32_t M, uint32_t N,
                                                      float* __restrict__ pred_t, float* __restrict__ weights_sum,
                                                      float* __restrict__ depth, float* __restrict__ image,
                                                      float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs,
                                                      float* __restrict__ sums) {
    const uint32_t n = blockIdx.x;
    const uint32_t lane = threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n];
    const uint32_t offset = (uint32_t)rays[3 * (size_t)n + 1];
    const uint32_t cnt = (uint32_t)rays[3 * (size_t)n + 2];
    const bool skip = (cnt == 0 || offset + cnt >= M);  // raymarching.cu:525-532, :629
    const float bgr = bg[0], bgg = bg[1], bgb = bg[2];
    float r = 0, g = 0, b = 0, ws = 0, d = 0;   // student
    float rt = 0, gt = 0, bt = 0, wt = 0;       // teacher
    if (!skip) {
        float T = 1.0f, Tt = 1.0f, tcarry = 0.0f;
        for (uint32_t base0 = 0; base0 < cnt; base0 += 64) {  // 2 chunks x 2 networks of loads in flight
            float sg[2], c0[2], c1[2], c2[2], sgt[2], t0[2], t1[2], t2[2];
            float2 dl[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t i = base0 + 32u * q + lane;
                const bool ok = i < cnt;
                const size_t row = (size_t)offset + (ok ? i : 0);
                dl[q] = ok ? __ldg(reinterpret_cast<const float2*>(deltas + 2 * row)) : make_float2(0.f, 0.f);
                sg[q] = ok ? __ldg(sig_s + row) : 0.0f;
                c0[q] = ok ? __ldg(rgb_s + 3 * row) : 0.f;
                c1[q] = ok ? __ldg(rgb_s + 3 * row + 1) : 0.f;
                c2[q] = ok ? __ldg(rgb_s + 3 * row + 2) : 0.f;
                sgt[q] = ok ? __ldg(sig_t + row) : 0.0f;
                t0[q] = ok ? __ldg(rgb_t + 3 * row) : 0.f;
                t1[q] = ok ? __ldg(rgb_t + 3 * row + 1) : 0.f;
                t2[q] = ok ? __ldg(rgb_t + 3 * row + 2) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t base = base0 + 32u * q;
                if (base >= cnt) break;  // warp-uniform
                const bool ok = base + lane < cnt;
                // student (raymarching.cu:546-560)
                const float alpha = ok ? 1.0f - __expf(-sg[q] * dl[q].x) : 0.0f;
                const float incl = warp_scan_mul(1.0f - alpha, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.0f;
                const float w = alpha * (T * excl);
                const float tin = tcarry + warp_scan_add(dl[q].y, lane);
                // teacher, same samples
                const float alpha_t = ok ? 1.0f - __expf(-sgt[q] * dl[q].x) : 0.0f;
                const float incl_t = warp_scan_mul(1.0f - alpha_t, lane);
                float excl_t = __shfl_up_sync(0xffffffffu, incl_t, 1);
                if (lane == 0) excl_t = 1.0f;
                const float w_t = alpha_t * (Tt * excl_t);
                if (ok) {
                    r += w * c0[q]; g += w * c1[q]; b += w * c2[q]; d += w * tin; ws += w;
                    rt += w_t * t0[q]; gt += w_t * t1[q]; bt += w_t * t2[q]; wt += w_t;
                }
                T *= __shfl_sync(0xffffffffu, incl, 31);
                Tt *= __shfl_sync(0xffffffffu, incl_t, 31);
                tcarry = __shfl_sync(0xffffffffu, tin, 31);
            }
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
        rt = warp_sum(rt); gt = warp_sum(gt); bt = warp_sum(bt); wt = warp_sum(wt);
    }
    // pixels with the background mixed in (renderer.py:445), for both networks
    const float om_s = 1.0f - ws, om_t = 1.0f - wt;
    const float pr_t = rt + om_t * bgr, pg_t = gt + om_t * bgg, pb_t = bt + om_t * bgb;
    const float dr = (r + om_s * bgr) - pr_t, dg = (g + om_s * bgg) - pg_t, db = (b + om_s * bgb) - pb_t;
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[3 * (size_t)index] = r;
        image[3 * (size_t)index + 1] = g;
        image[3 * (size_t)index + 2] = b;
        pred_t[3 * (size_t)index] = pr_t;
        pred_t[3 * (size_t)index + 1] = pg_t;
        pred_t[3 * (size_t)index + 2] = pb_t;
        atomicAdd(sums + PVD_PAIR_SUM_STRIDE * (n % PVD_LOSS_SLOTS) + kSumRgb, dr * dr + dg * dg + db * db);
    }
    if (skip) {  // rows below M of a ray that does not fit still reach k_pair_combine: no composite gradient for them
        for (uint32_t i = offset + lane; i < min(offset + cnt, M); i += 32) {
            grad_sigmas[i] = 0.0f;
            grad_rgbs[3 * (size_t)i] = 0.0f;
            grad_rgbs[3 * (size_t)i + 1] = 0.0f;
            grad_rgbs[3 * (size_t)i + 2] = 0.0f;
        }
        return;
    }
    // ---- student backward sweep (raymarching.cu:597-697) with the un-normalised upstream d||pred_t - pred_s|| ~ (pred_s - pred_t)
    const float gr = dr, gg = dg, gb = db;
    const float gws = -(gr * bgr + gg * bgg + gb * bgb);
    const float r_final = r, g_final = g, b_final = b, ws_final = ws;
    float T = 1.0f, rc = 0, gc = 0, bc = 0, wc = 0;
    for (uint32_t base0 = 0; base0 < cnt; base0 += 128) {
        float sgv[4], d0v[4], c0v[4], c1v[4], c2v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = base0 + 32u * q + lane;
            const bool ok = i < cnt;
            const size_t row = (size_t)offset + (ok ? i : 0);
            sgv[q] = ok ? __ldg(sig_s + row) : 0.0f;
            d0v[q] = ok ? __ldg(deltas + 2 * row) : 0.0f;
            c0v[q] = ok ? __ldg(rgb_s + 3 * row) : 0.f;
            c1v[q] = ok ? __ldg(rgb_s + 3 * row + 1) : 0.f;
            c2v[q] = ok ? __ldg(rgb_s + 3 * row + 2) : 0.f;
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
            const float incl = warp_scan_mul(1.0f - alpha, lane);
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

__global__ void __launch_bounds__(256) k_pair_combine(const float* __restrict__ feat_t, const float* __restrict__ feat_s,
                                                      const float* __restrict__ rgb_t, const float* __restrict__ rgb_s,
                                                      const float* __restrict__ sums, PvdPairRates rates, float loss_scale,
                                                      uint32_t M, const int32_t* __restrict__ n_composite_p,
                                                      float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs,
                                                      float* __restrict__ grad_feat, float* __restrict__ loss_out) {
    pdl_launch_dependents();   // the student field backward may start its prologue + forward recomputation now
    __shared__ float coef[4];
    if (threadIdx.x < 32) {  // warp 0: the four sums over the slots -> norms -> gradient coefficients rate / ||.||
        float s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            s[k] = warp_sum(__ldg(sums + PVD_PAIR_SUM_STRIDE * threadIdx.x + k) + __ldg(sums + PVD_PAIR_SUM_STRIDE * (threadIdx.x + 32u) + k));
        if (threadIdx.x == 0) {
            const float rate[4] = {rates.rgb, rates.fea, rates.color, rates.sigma};
            float total = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float nrm = sqrtf(s[k]);
                coef[k] = (nrm > 0.0f && rate[k] != 0.0f) ? loss_scale * rate[k] / nrm : 0.0f;  // torch.norm's backward at 0 is 0
                total += rate[k] * nrm;
                if (blockIdx.x == 0) loss_out[1 + k] = nrm;   // the four un-weighted terms (what the trainer logs, :1179-1187)
            }
            if (blockIdx.x == 0) loss_out[0] = total;
        }
    }
    static_assert(PVD_LOSS_SLOTS == 64, "k_pair_combine reads two slots per lane");
    __syncthreads();
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    // rows that carry a composite gradient: those below the march's sample count (the rest is padding no ray owns)
    const uint32_t n_comp = n_composite_p ? min((uint32_t)max(*n_composite_p, 0), M) : 0u;
    const float c_rgb = coef[kSumRgb], c_fea = coef[kSumFea], c_col = coef[kSumColor], c_sig = coef[kSumSigma];
    float gs = 0.0f, gc[3] = {0.f, 0.f, 0.f};
    if (row < n_comp) {
        gs = grad_sigmas[row];
#pragma unroll
        for (int c = 0; c < 3; ++c) gc[c] = grad_rgbs[3 * (size_t)row + c];
    }
    grad_sigmas[row] = c_rgb * gs;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float d = __ldg(rgb_s + 3 * (size_t)row + c) - __ldg(rgb_t + 3 * (size_t)row + c);
        grad_rgbs[3 * (size_t)row + c] = c_rgb * gc[c] + c_col * d;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(feat_s + 16 * (size_t)row) + q);
        const float4 b = __ldg(reinterpret_cast<const float4*>(feat_t + 16 * (size_t)row) + q);
        float4 o = make_float4(c_fea * (a.x - b.x), c_fea * (a.y - b.y), c_fea * (a.z - b.z), c_fea * (a.w - b.w));
        if (q == 0) o.x += c_sig * (a.x - b.x);  // sigma_l = feat[..., 0] (network.py:424)
        *(reinterpret_cast<float4*>(grad_feat + 16 * (size_t)row) + q) = o;
    }
}

__global__ void __launch_bounds__(256) k_zero_sample_tail(const int32_t* __restrict__ rays, const int32_t* __restrict__ counter,
                                                          uint32_t N, uint32_t M, float* __restrict__ xyzs, float* __restrict__ dirs,
                                                          float* __restrict__ deltas) {
    const uint32_t T = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    auto zero_row = [&](uint32_t row) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            xyzs[3 * (size_t)row + c] = 0.0f;
            dirs[3 * (size_t)row + c] = 0.0f;
        }
        deltas[2 * (size_t)row] = 0.0f;
        deltas[2 * (size_t)row + 1] = 0.0f;
    };
    // rows of rays that were dropped because they do not fit (raymarching.cu:419): the reference leaves them zero
    for (uint32_t n = tid; n < N; n += T) {
        const uint32_t offset = (uint32_t)rays[3 * (size_t)n + 1], cnt = (uint32_t)rays[3 * (size_t)n + 2];
        if (cnt != 0 && offset + cnt >= M && offset < M)
            for (uint32_t row = offset; row < M && row < offset + cnt; ++row) zero_row(row);
    }
    // rows past the last sample
    const uint32_t total = min((uint32_t)max(counter[0], 0), M);
    for (uint32_t row = total + tid; row < M; row += T) zero_row(row);
}


// d/dp of  weight * mean|p|  (NeRFNetwork.density_loss, distill_mutual/network.py:549-557: the L1 penalty on the vm sigma planes and
// lines that both trainers add for model_type "vm", utils.py:1135-1136 / just_train_tea/utils.py:843-844), accumulated onto grad;
// the term's value goes to the step's loss slots.
__global__ void __launch_bounds__(256) k_l1_mean_reg(const float* __restrict__ p, uint64_t n, float k_loss, float k_grad,
                                                     float* __restrict__ grad, float* __restrict__ loss_slots) {
    __shared__ float sh[8];
    float acc = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(p + i);
        acc += fabsf(v);
        grad[i] += (v > 0.0f) ? k_grad : ((v < 0.0f) ? -k_grad : 0.0f);  // torch.sign
    }
    acc = block_sum_256(acc, sh);
    if (threadIdx.x == 0) atomicAdd(loss_slots + 2u * (blockIdx.x % PVD_LOSS_SLOTS), k_loss * acc);
}

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_pair_sample_sq(const float* feat_tea, const float* feat_stu, const float* rgbs_tea, const float* rgbs_stu, uint32_t M,
                       float* sums, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(feat_tea && feat_stu && rgbs_tea && rgbs_stu && sums);
    k_pair_sample_sq<<<ceil_div(M, 256u), 256, 0, (cudaStream_t)stream>>>(feat_tea, feat_stu, rgbs_tea, rgbs_stu, M, sums);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_pair_composite(const float* bg_color, const float* sigmas_tea, const float* rgbs_tea, const float* sigmas_stu,
                       const float* rgbs_stu, const float* deltas, const int32_t* rays, uint32_t M, uint32_t N, float* pred_tea,
                       float* weights_sum, float* depth, float* image, float* grad_sigmas, float* grad_rgbs, float* sums,
                       void* stream) {
    if (N == 0) return PVD_OK;
    PVD_REQUIRE(bg_color && sigmas_tea && rgbs_tea && sigmas_stu && rgbs_stu && deltas && rays && pred_tea && weights_sum && depth &&
                image && grad_sigmas && grad_rgbs && sums);
    k_pair_composite<<<N, 32, 0, (cudaStream_t)stream>>>(bg_color, sigmas_tea, rgbs_tea, sigmas_stu, rgbs_stu, deltas, rays, M, N,
                                                         pred_tea, weights_sum, depth, image, grad_sigmas, grad_rgbs, sums);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_pair_combine(const float* feat_tea, const float* feat_stu, const float* rgbs_tea, const float* rgbs_stu, const float* sums,
                     const PvdPairRates* rates, float loss_scale, uint32_t M, const int32_t* n_composite, float* grad_sigmas,
                     float* grad_rgbs, float* grad_feat16, float* loss_out, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(feat_tea && feat_stu && rgbs_tea && rgbs_stu && sums && rates && grad_sigmas && grad_rgbs && grad_feat16 && loss_out);
    k_pair_combine<<<ceil_div(M, 256u), 256, 0, (cudaStream_t)stream>>>(feat_tea, feat_stu, rgbs_tea, rgbs_stu, sums, *rates,
                                                                       loss_scale, M, n_composite, grad_sigmas, grad_rgbs,
                                                                       grad_feat16, loss_out);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_zero_sample_tail(const int32_t* rays, const int32_t* counter, uint32_t N, uint32_t M, float* xyzs, float* dirs,
                         float* deltas, void* stream) {
    if (M == 0) return PVD_OK;
    PVD_REQUIRE(rays && counter && xyzs && dirs && deltas);
    k_zero_sample_tail<<<32, 256, 0, (cudaStream_t)stream>>>(rays, counter, N, M, xyzs, dirs, deltas);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_l1_mean_reg(const float* param, uint64_t n, float weight, float loss_scale, float* grad, float* loss_slots, void* stream) {
    if (n == 0 || weight == 0.0f) return PVD_OK;
    PVD_REQUIRE(param && grad && loss_slots);
    const uint32_t grid = (uint32_t)min((unsigned long long)((n + 1023u) / 1024u), 148ull * 8ull);
    k_l1_mean_reg<<<grid, 256, 0, (cudaStream_t)stream>>>(param, n, weight / (float)n, loss_scale * weight / (float)n, grad, loss_slots);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
