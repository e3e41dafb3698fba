/*
 * pvd_b200_fused.h -- C ABI of the FUSED field query: hash-grid encode + SH + sigma/color MLPs in one kernel
 * per direction, MLP GEMMs on tcgen05 tensor cores with TMEM accumulators (sm_100a only).
 *
 * These entry points are additive: they sit behind NeRFNetwork.forward / NeRFRenderer.run_cuda of the reference
 * (distill_mutual/network.py:335-437, distill_mutual/renderer.py:359-448) and replace, for model_type "hash",
 *   GridEncoder.forward (gridencoder/grid.py:207-232) -> sigma_net (network.py:413-417) -> clamp/trunc_exp (:418-425)
 *   -> SHEncoder (shencoder/sphere_harmonics.py:83-95) -> color_net + sigmoid (network.py:428-437)
 * and its autograd backward (5 cuBLAS GEMM pairs + the gridencoder scatter).
 *
 * Same conventions as pvd_b200.h: caller-owned device pointers, caller-allocated outputs, no hidden state,
 * stream passed as void*, int return (0 ok / cudaError_t / negative PVD_E*).
 */
#ifndef PVD_B200_FUSED_H
#define PVD_B200_FUSED_H

#include <stdint.h>
#include "pvd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PVD_FIELD_WBLOB_BYTES 20480u /* packed fp16 operand tiles of the five weight matrices */
#define PVD_FIELD_ENC_STRIDE 32u     /* halves per sample in the saved encoding (2L <= 32, zero padded) */

typedef struct PvdHashField {
    const void* table;      /* [offsets[L], 2] hash-grid features, fp16 (PVD_DTYPE_F16) or fp32 */
    const int32_t* offsets; /* [L+1] level offsets (gridencoder/grid.py:177-190) */
    const void* wblob;      /* PVD_FIELD_WBLOB_BYTES from pvd_field_pack_weights */
    int32_t table_dtype;
    uint32_t L;             /* levels, 2L <= 32 */
    uint32_t H;             /* base resolution */
    float S;                /* log2(per_level_scale) */
    float bound;            /* scene bound: xyz in [-bound, bound] is mapped to [0,1] (grid.py:211) */
    float sigma_clip_min;   /* clamp of sigma_net channel 0 (network.py:418-420; defaults -2 / 7) */
    float sigma_clip_max;
    float density_scale;    /* renderer.py:440 */
} PvdHashField;

/* Convert the five nn.Linear weights (fp32, row-major [out, in]) of
 *   sigma_net.0 [64, 2L], sigma_net.1 [16, 64], color_net.0 [64, 31], color_net.1 [64, 64], color_net.2 [3, 64]
 * (network.py:103-152, bias=False) into the fp16 shared-memory operand tiles the kernels load. Run after every
 * optimizer step. */
int pvd_field_pack_weights(const float* w_sigma0, const float* w_sigma1, const float* w_color0, const float* w_color1,
                           const float* w_color2, uint32_t in_dim, void* wblob, void* stream);

/* Forward for M samples: xyzs [M,3], dirs [M,3] -> sigmas [M] (already multiplied by density_scale), rgbs [M,3].
 * Optional outputs (NULL to skip): enc [M, 32] fp16 (hash features, needed by the backward),
 * feat16 [M,16] fp32 = sigma_net output with channel 0 clamped (`feature_sigma_color`, network.py:421).
 * `status` (device int, caller zeroes it) is set non-zero if a tensor-core wait timed out. */
int pvd_hash_field_forward(const PvdHashField* field, const float* xyzs, const float* dirs, uint32_t M, float* sigmas,
                           float* rgbs, void* enc, float* feat16, int32_t* status, void* stream);

/* Backward: d(loss)/d(sigmas) [M], d(loss)/d(rgbs) [M,3] -> ACCUMULATES into
 *   grad_table [offsets[L], 2] fp32 and gw_ws, a PVD_FIELD_GW_FLOATS workspace holding the five weight gradients in the
 *   kernel-native accumulator shapes  dW1[64][32] | dW2^T[64][16] | dW3[64][32] | dW4[64][64] | dW5^T[64][16]
 *   (zero padded; pvd_field_unpack_wgrads adds them onto parameter-shaped [out, in] buffers).
 * grad_feat16 [M,16] (or NULL) is d(loss)/d(feat16): the gradient of the distillation losses that read
 * `feature_sigma_color` / `sigma_l` directly (distill_mutual/utils.py:1046-1108); channel 0 passes the clamp mask.
 * dx_ws: optional [M, 32] fp16 scratch.  NULL (the engines' default): ONE launch -- four scatter warps inside every MLP CTA reduce
 * tile i's d(encoding), handed over in shared memory, into grad_table while the MLP warps run the tensor-core chain of tile i+1.
 * When given: TWO launches -- the MLP kernel writes d(encoding) there and a second, full-occupancy kernel (one thread per
 * sample x level) scatters it into grad_table.
 * `enc` is the tensor the forward saved.  Rows >= *n_valid (device pointer, e.g. the march counter; NULL = all M)
 * are padding and contribute nothing. */
#define PVD_FIELD_GW_FLOATS 10240u
/* The weight-gradient workspace holds PVD_FIELD_GW_COPIES replicas of that layout (CTA c accumulates into replica c mod COPIES,
 * so that at most grid/COPIES reductions hit one address: L2 serialises same-address atomics); the unpack functions sum them.
 * gw_ws must therefore be PVD_FIELD_GW_COPIES * PVD_FIELD_GW_FLOATS floats, zeroed by the caller before the backward. */
#define PVD_FIELD_GW_COPIES 16u
int pvd_hash_field_backward(const PvdHashField* field, const float* xyzs, const float* dirs, const void* enc,
                            const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                            const int32_t* n_valid, float* grad_table, float* gw_ws, void* dx_ws, int32_t* status,
                            void* stream);

/* composite_rays_train_backward (pvd_b200.h) with the upstream gradients derived in the kernel from the photometric loss:
 *   pred = image + (1 - weights_sum) * bg_color   (distill_mutual/renderer.py:445)
 *   loss = mean((pred - gt_rgb)^2) over the N x 3 values (just_train_tea/utils.py:841-846, MSELoss)
 * gradients are multiplied by loss_scale (GradScaler).  loss_out is PVD_LOSS_SLOTS pairs of floats: ray n adds into pair
 * n % PVD_LOSS_SLOTS (4096 same-address atomics would serialise in L2 and stall the SMs' memory pipes), and the step's values are the
 * sums over the slots: sum loss_out[2s] = unscaled loss, sum loss_out[2s+1] = rays that carried
 * samples.  gt_rgb [N,3], bg_color [3]. */
int pvd_composite_rays_train_backward_mse(const float* gt_rgb, const float* bg_color, float loss_scale, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int32_t* rays,
                                          const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                          float* grad_sigmas, float* grad_rgbs, float* loss_out, void* stream);

/* composite_rays_train_forward + _backward_mse in ONE launch (same arithmetic, same outputs): the warp that owns a ray sweeps it
 * forward, forms the pixel's loss gradient, and sweeps it again for the sample gradients. */
int pvd_composite_rays_train_mse(const float* gt_rgb, const float* bg_color, float loss_scale, const float* sigmas,
                                 const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                 float* weights_sum, float* depth, float* image, float* grad_sigmas, float* grad_rgbs,
                                 float* loss_out, void* stream);

/* The same backward restricted to rows [row0, row0 + rows) of the sample buffers (all pointers are the buffers' bases), and split
 * in its two kernels so that a caller can run them on different streams: PVD_BWD_MLP = tcgen05 MLP backward, leaves d(encoding) in
 * dx_ws (required here); PVD_BWD_SCATTER = reductions of dx_ws into grad_table.  Halving the rows and issuing
 * MLP(A) ; MLP(B) on one stream and SCATTER(A) ; SCATTER(B) on another overlaps the atomic-bound scatter of one half with the
 * latency-bound MLP chain of the other.  n_valid counts rows from the start of the buffers, as above. */
#define PVD_LOSS_SLOTS 64u
#define PVD_BWD_MLP 1u
#define PVD_BWD_SCATTER 2u
int pvd_hash_field_backward_rows(const PvdHashField* field, const float* xyzs, const float* dirs, const void* enc,
                                 const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t row0,
                                 uint32_t rows, const int32_t* n_valid, float* grad_table, float* gw_ws, void* dx_ws, int32_t* status,
                                 uint32_t phases, void* stream);

/* ------------------------------------------------------------------------------------------
 * The same field in FP32 end to end (csrc/field_hash_f32.cu): what NeRFNetwork.forward computes when the reference runs WITHOUT
 * --fp16 (fp32 table gather, gridencoder.cu with scalar_t = float, fp32 GEMMs; network.py:335-437), and the path north_star's
 * 1e-4 fp32 parity bound is measured on.  field->table must be the fp32 table (table_dtype PVD_DTYPE_F32); field->wblob is not
 * used: the five weight matrices are read in place as fp32 row-major [out, in].  No saved encoding: the backward recomputes the
 * forward.  Gradients go to the same grad_table / gw_ws (PVD_FIELD_GW_COPIES x PVD_FIELD_GW_FLOATS, pvd_field_unpack_wgrads)
 * as the fp16 path.
 * ---------------------------------------------------------------------------------------- */
typedef struct PvdFieldWeightsF32 {
    const float* sigma0; /* [64, 2L] */
    const float* sigma1; /* [16, 64] */
    const float* color0; /* [64, 31] */
    const float* color1; /* [64, 64] */
    const float* color2; /* [3, 64]  */
} PvdFieldWeightsF32;
int pvd_hash_field_forward_f32(const PvdHashField* field, const PvdFieldWeightsF32* weights, const float* xyzs, const float* dirs,
                               uint32_t M, float* sigmas, float* rgbs, float* feat16, void* stream);
int pvd_hash_field_backward_f32(const PvdHashField* field, const PvdFieldWeightsF32* weights, const float* xyzs, const float* dirs,
                                const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                                const int32_t* n_valid, float* grad_table, float* gw_ws, void* stream);

/* gw_* += un-padded / transposed views of gw_ws; shapes [64,in_dim], [16,64], [64,31], [64,64], [3,64]. */
int pvd_field_unpack_wgrads(const float* gw_ws, uint32_t in_dim, float* gw_sigma0, float* gw_sigma1, float* gw_color0,
                            float* gw_color1, float* gw_color2, void* stream);

/* ------------------------------------------------------------------------------------------
 * "vm" (TensoRF vector-matrix) field: network.py:72-90 (parameters), :216-309 (plane/line features), :344-382 (forward).
 * Planes [1,R,H,W] and lines [1,R,D,1] are read in place as CHANNELS-LAST fp32 (torch.channels_last strides: memory order
 * [H][W][R] / [D][R]); R = 16 for sigma, 48 for colour (network.py:73-74).  res[d] = grid resolution along axis d;
 * plane i spans axes mat_ids[i] = {0,1},{0,2},{1,2}, line i axis vec_ids[i] = 2,1,0 (network.py:76-77).
 * ---------------------------------------------------------------------------------------- */
#define PVD_VM_WBLOB_BYTES 18944u

typedef struct PvdVmField {
    const void* sigma_mat[3]; /* planes [1,16,H,W] / lines [1,16,D,1] / [1,48,..] in channels-last memory ([H][W][R]): the fp32 */
    const void* sigma_vec[3]; /* parameters themselves (plane_dtype PVD_DTYPE_F32) or an fp16 shadow copy of them (PVD_DTYPE_F16: */
    const void* color_mat[3]; /* half the gather bytes; gradients are fp32 either way)                                          */
    const void* color_vec[3];
    const void* wblob;      /* PVD_VM_WBLOB_BYTES from pvd_vm_pack_weights */
    uint32_t res[3];
    float aabb[6];          /* aabb_train: positions are mapped to [-1,1] (network.py:345-350) */
    float sigma_clip_min;   /* clamp of both sigma_feat and color_feat (network.py:353-361) */
    float sigma_clip_max;
    float density_scale;
    int32_t plane_dtype;    /* PVD_DTYPE_F32 (0) | PVD_DTYPE_F16 (1) */
} PvdVmField;

typedef struct PvdVmGrads { /* channels-last fp32 gradient buffers, same shapes as the parameters; accumulated into */
    float* sigma_mat[3];
    float* sigma_vec[3];
    float* color_mat[3];
    float* color_vec[3];
} PvdVmGrads;

/* basis_mat.weight [15,144], color_net.{0,1,2}.weight [64,31] [64,64] [3,64] (fp32) -> fp16 operand tiles */
int pvd_vm_pack_weights(const float* basis_mat, const float* w_color0, const float* w_color1, const float* w_color2, void* wblob,
                        void* stream);

/* Forward: sigmas [M] (x density_scale), rgbs [M,3], optional feat16 [M,16] = [clamped sigma_feat, clamped color_feat(15)]
 * (`feature_sigma_color`, network.py:362-364). */
int pvd_vm_field_forward(const PvdVmField* field, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                         float* feat16, int32_t* status, void* stream);

/* Backward: accumulates into the channels-last plane/line gradients of `grads` and into gw_ws (PVD_FIELD_GW_FLOATS floats:
 * dBasis^T[192][16] | dW3[64][32] | dW4[64][64] | dW5^T[64][16]; pvd_vm_unpack_wgrads adds them onto [out,in] buffers). */
int pvd_vm_field_backward(const PvdVmField* field, const PvdVmGrads* grads, const float* xyzs, const float* dirs,
                          const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                          const int32_t* n_valid, float* gw_ws, int32_t* status, void* stream);

/* The same backward as TWO kernels: the MLP kernel leaves d(appearance) (fp16) and d(sigma feature) in `scatter_ws`
 * (pvd_vm_backward_workspace_bytes(M) bytes, 16-byte aligned) and a second, high-occupancy kernel re-gathers the taps and reduces
 * into the plane / line gradients.  Same results up to the order of the fp32 reductions; 141 -> ~100 us at 73 k samples. */
uint64_t pvd_vm_backward_workspace_bytes(uint32_t M);
int pvd_vm_field_backward_ws(const PvdVmField* field, const PvdVmGrads* grads, const float* xyzs, const float* dirs,
                             const float* grad_sigmas, const float* grad_rgbs, const float* grad_feat16, uint32_t M,
                             const int32_t* n_valid, float* gw_ws, void* scatter_ws, int32_t* status, void* stream);

int pvd_vm_unpack_wgrads(const float* gw_ws, float* g_basis, float* gw_color0, float* gw_color1, float* gw_color2, void* stream);

/* ------------------------------------------------------------------------------------------
 * "mlp" (NeRF) field -- the frozen teacher of mlp -> hash distillation (BASELINE config 5), and a trainable model:
 * FreqEncoder PE(10) 3 -> 63 (tools/encoding.py:6-49), nerf_mlp = 8 Linear layers with bias, 256 wide, skip concat of the
 * encoding after the 4th (network.py:56-70,324-333), output 28, then the same sigma_net / color_net tail as the hash model.
 * ---------------------------------------------------------------------------------------- */
#define PVD_MLP_WBLOB_BYTES 876544u

typedef struct PvdMlpField {
    const void* wblob;       /* `replicas` x PVD_MLP_WBLOB_BYTES: pvd_mlp_pack_weights fills the first, pvd_mlp_replicate_weights the rest */
    const void* tail_wblob;  /* PVD_FIELD_WBLOB_BYTES from pvd_field_pack_weights (in_dim = 28) */
    float sigma_clip_min;
    float sigma_clip_max;
    float density_scale;
    uint32_t replicas;       /* 0 / 1: one copy.  R > 1: CTA b streams replica b mod R -- every CTA walks the SAME 876 KB in the same order,
                                so one copy concentrates all 148 TMA streams on the few L2 slices that hold the current piece */
} PvdMlpField;

/* weights8 / biases8: DEVICE arrays of 8 device pointers to nerf_mlp.{0..7}.weight ([256,63] [256,256]x3 [256,319] [256,256]x2
 * [28,256], fp32 row-major) and .bias. */
int pvd_mlp_pack_weights(const float* const* weights8, const float* const* biases8, void* wblob, void* stream);
/* copy replica 0 of a packed stream (`bytes` each) into replicas 1 .. replicas-1 behind it */
int pvd_mlp_replicate_weights(void* wblob, uint64_t bytes, uint32_t replicas, void* stream);

int pvd_mlp_field_forward(const PvdMlpField* field, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                          float* feat16, int32_t* status, void* stream);

/* TRAINING an mlp model (main_just_train_tea.py --model_type mlp; network.py:324-333 under autograd in the reference).
 * Forward: the same kernel, which additionally saves, per 128-sample tile, every fp16 operand tile the tensor core consumed
 * (PE tile + the ReLU outputs of layers 0..6, PVD_MLP_SAVE_TILE_BYTES per tile; save_ws = ceil(M / 128) tiles) and the 28-wide trunk
 * output `enc` [M, 32] fp16 (the tail's input, same role as the hash model's saved encoding).
 * Backward, three launches:
 *   1. the sigma/colour tail: pvd_hash_field_backward_rows(tail_field, ..., enc, ..., dx_ws = d_x28 [M,32] fp16, PVD_BWD_MLP) -- the
 *      hash model's tcgen05 tail backward, which leaves d(loss)/d(x28) in dx_ws and the five tail weight gradients in its gw_ws;
 *   2. pvd_mlp_trunk_backward: data gradients through layers 7..1 (transposed weight stream, masks from the saved activations);
 *      writes G_7, G_0..G_6 = d(loss)/d(pre-activation) per tile into grad_ws (PVD_MLP_GRAD_TILE_BYTES per tile);
 *   3. pvd_mlp_weight_grads: dW_l = G_l^T act_l, db_l = column sums of G_l, reduction over all samples on the tensor core
 *      (accumulators in TMEM, one (layer, sample-split) job per CTA); ACCUMULATES into gw_ws, PVD_MLP_GW_FLOATS floats
 *      (caller zeroes it): layers 0..6 as [256][320] (layer 4: columns 0..62 = in_pts part, 64..319 = hidden part; others from
 *      column 0), layer 7 as [256 in][32 out], then 8 x 256 bias gradients.
 *   pvd_mlp_unpack_wgrads adds them onto parameter-shaped buffers. */
#define PVD_MLP_SAVE_TILE_BYTES 475136u
#define PVD_MLP_GRAD_TILE_BYTES 466944u
#define PVD_MLP_GW_FLOATS (7u * 256u * 320u + 256u * 32u + 8u * 256u)
int pvd_mlp_field_forward_train(const PvdMlpField* field, const float* xyzs, const float* dirs, uint32_t M, float* sigmas,
                                float* rgbs, float* feat16, void* save_ws, void* enc, int32_t* status, void* stream);
/* wblob_t: PVD_MLP_WBLOB_T_BYTES from pvd_mlp_pack_weights_t (the transposed weight stream of the backward). */
#define PVD_MLP_WBLOB_T_BYTES (49u * 16384u)
int pvd_mlp_pack_weights_t(const float* const* weights8, void* wblob_t, void* stream);
/* n_valid (device pointer or NULL = all M): rows >= *n_valid are padding and receive zero gradients. */
int pvd_mlp_trunk_backward(const void* wblob_t, uint32_t replicas, const void* save_ws, const void* d_x28, uint32_t M,
                           const int32_t* n_valid, void* grad_ws, int32_t* status, void* stream);
int pvd_mlp_weight_grads(const void* save_ws, const void* grad_ws, uint32_t M, float* gw_ws, int32_t* status, void* stream);
int pvd_mlp_unpack_wgrads(const float* gw_ws, float* const* grad_weights8, float* const* grad_biases8, void* stream);

/* ------------------------------------------------------------------------------------------
 * (teacher, student) distillation at SHARED samples: the loss side of Trainer.train_step, distill_mutual/utils.py:954-1189,
 * loss_type "normL2" (get_loss :940-951):
 *   loss = rgb * ||pred_tea - pred_stu|| + fea * ||feat_stu - feat_tea|| + color * ||rgbs_stu - rgbs_tea|| + sigma * ||feat_stu[:,0] - feat_tea[:,0]||
 * with pred = image + (1 - weights_sum) * bg_color (renderer.py:445), feat = `feature_sigma_color` [M,16], rgbs = `color_l` [M,3].
 * Both fields are first evaluated on the same xyzs/dirs (pvd_*_field_forward with feat16); then, on one stream:
 *   pvd_pair_sample_sq -> pvd_pair_composite -> pvd_pair_combine -> the student's pvd_*_field_backward(grad_sigmas, grad_rgbs, grad_feat16).
 * `sums` is PVD_LOSS_SLOTS x PVD_PAIR_SUM_STRIDE floats, zeroed by the caller before pvd_pair_sample_sq:
 *   slot[0] sum (pred_stu - pred_tea)^2, slot[1] feature, slot[2] colour, slot[3] sigma; the step's values are the sums over slots.
 * Stages 1 and 2 of the reference (no compositing, :1046-1108) skip pvd_pair_composite and pass n_composite = NULL.
 * ---------------------------------------------------------------------------------------- */
#define PVD_PAIR_SUM_STRIDE 4u
typedef struct PvdPairRates { /* loss_rate_rgb / loss_rate_fea_sc / loss_rate_color / loss_rate_sigma (main_distill_mutual.py) */
    float rgb, fea, color, sigma;
} PvdPairRates;

/* per-sample sums of squares over ALL M rows (padding rows included, as in the reference: raymarching.py:240-242) */
int pvd_pair_sample_sq(const float* feat_tea, const float* feat_stu, const float* rgbs_tea, const float* rgbs_stu, uint32_t M,
                       float* sums, void* stream);
/* composite_rays_train_forward of teacher AND student on the same rays (raymarching.cu:505-582) + sum of squared pixel differences
 * + composite_rays_train_backward of the student (:607-686) with the UN-normalised upstream (pred_stu - pred_tea).
 * Outputs: pred_tea [N,3] (background mixed in), the student's weights_sum [N], depth [N], image [N,3] (raw, as the op returns them),
 * grad_sigmas [M], grad_rgbs [M,3] (un-normalised; rows of rays that own samples), sums slot[0]. */
int pvd_pair_composite(const float* bg_color, const float* sigmas_tea, const float* rgbs_tea, const float* sigmas_stu,
                       const float* rgbs_stu, const float* deltas, const int32_t* rays, uint32_t M, uint32_t N, float* pred_tea,
                       float* weights_sum, float* depth, float* image, float* grad_sigmas, float* grad_rgbs, float* sums,
                       void* stream);
/* final gradients, in place over grad_sigmas / grad_rgbs (rows >= *n_composite carry no composite gradient), and
 * grad_feat16 [M,16]; all multiplied by loss_scale.  loss_out [5] = total loss, then the four un-weighted norms (rgb, fea, color, sigma). */
int pvd_pair_combine(const float* feat_tea, const float* feat_stu, const float* rgbs_tea, const float* rgbs_stu, const float* sums,
                     const PvdPairRates* rates, float loss_scale, uint32_t M, const int32_t* n_composite, float* grad_sigmas,
                     float* grad_rgbs, float* grad_feat16, float* loss_out, void* stream);
/* zero the rows of xyzs / dirs / deltas [M,*] that no surviving ray owns (past counter[0], or owned by a ray that was dropped because
 * it does not fit): what torch.zeros gives the reference every call (raymarching.py:240-242) for persistent buffers. */
int pvd_zero_sample_tail(const int32_t* rays, const int32_t* counter, uint32_t N, uint32_t M, float* xyzs, float* dirs,
                         float* deltas, void* stream);

/* grad[i] += loss_scale * weight / n * sign(param[i]);  loss_slots[2*s] += weight / n * sum|param|  -- the gradient and value of
 * weight * mean|param|, one term of NeRFNetwork.density_loss (network.py:549-557; vm models, l1_reg_weight, utils.py:1135-1136). */
int pvd_l1_mean_reg(const float* param, uint64_t n, float weight, float loss_scale, float* grad, float* loss_slots, void* stream);

/* ---- multi-GPU: the one exchange step (rays shard, parameters replicate) ------------------------------------------
 * Sum over all ranks of elements [elem_offset, elem_offset + elem_count) of a symmetric fp16 buffer, done in the NVSwitch:
 * `multicast_ptr` is the multicast mapping of the buffer (torch symmetric memory: handle.multicast_ptr); the range is read with
 * multimem.ld_reduce (fp32 accumulation) and written back to every rank with multimem.st.  Rank r calls it on its own 1/W of the
 * buffer, between two all-rank barriers supplied by the caller.  Offsets and counts in elements, multiples of 8. */
int pvd_multimem_allreduce_f16(void* multicast_ptr, uint64_t elem_offset, uint64_t elem_count, void* stream);
/* fp32 gradient -> fp16 exchange payload (elem_count multiple of 4, 16-byte aligned source) */
/* The same reduction with the two cross-rank barriers INSIDE the kernel (one launch instead of barrier + kernel + barrier):
 * `signal_pad_ptrs_dev` = device array of the ranks' symmetric-memory signal pads (uint32 slots, all zero between calls; slots
 * [8W, 10W) are used), `local_state` = 4 zero-initialised uint32 of this rank ([2] != 0 afterwards: a barrier timed out).
 * blocks (0 = default) and unroll (2 | 4 | 8) are tuning knobs; every rank must pass the same values. */
int pvd_multimem_allreduce_f16_fused(void* multicast_ptr, uint64_t elem_offset, uint64_t elem_count, const void* signal_pad_ptrs_dev,
                                     uint32_t rank, uint32_t world, uint32_t* local_state, uint32_t blocks, uint32_t unroll, void* stream);
/* Two-shot all-reduce over peer pointers (plain NVLink loads / stores, fp32 accumulation, no switch reduction), barriers inside the
 * kernel: `buffer_ptrs_dev` = device array of the ranks' symmetric payload buffers; world 2, 4 or 8; the other arguments as above
 * (signal-pad slots [12W, 13W): monotonic epoch flags; `local_state[3]` carries the epoch between launches); tuning knobs blocks (0 = 148 CTAs of 512 threads), unroll (1 | 2 | 4 vectors per thread and
 * buffer in flight), weak (plain instead of relaxed.sys accesses; the barriers' system-scope fences order them). */
int pvd_p2p_allreduce_f16(const void* buffer_ptrs_dev, uint64_t elem_offset, uint64_t elem_count, const void* signal_pad_ptrs_dev,
                          uint32_t rank, uint32_t world, uint32_t* local_state, uint32_t blocks, uint32_t unroll, uint32_t weak,
                          void* stream);
/* NVLink rate probe (scripts/micro/exchange_probe.py): copy elem_count halves between this rank's and the next rank's symmetric
 * buffer -- mode 0 pull, 1 push, 2 both directions on alternating vectors.  No synchronisation: timing only. */
int pvd_p2p_copy_probe(const void* buffer_ptrs_dev, uint64_t elem_count, uint32_t rank, uint32_t world, uint32_t mode, uint32_t blocks, void* stream);
int pvd_cast_f32_to_f16(const float* src, void* dst, uint64_t elem_count, void* stream);

/* ------------------------------------------------------------------------------------------
 * Persistent inference: the evaluation branch of NeRFRenderer.run_cuda (distill_mutual/renderer.py:450-543: a host loop of
 * march_rays -> forward -> composite_rays -> compact_rays with one D2H read per iteration) as ONE kernel for a hash field: CTAs pull
 * rays from a global queue, march / query / composite them 16 rays x 8 steps at a time and write only the final per-ray
 * weights_sum [N], depth [N], image [N,3] (raw accumulators: the caller mixes the background and normalises the depth,
 * renderer.py:540-541).  perturb is off (the reference's evaluation setting); `queue` is one int32 of scratch.
 * ---------------------------------------------------------------------------------------- */
int pvd_hash_render_persistent(const PvdHashField* field, const float* rays_o, const float* rays_d, const uint8_t* grid, const float* nears,
                               const float* fars, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, uint32_t N,
                               int32_t* queue, float* weights_sum, float* depth, float* image, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * "tensors" (Plenoxels-style) field: NeRFNetwork.forward for model_type "tensors" (distill_mutual/network.py:184-191,311-322,383-409):
 * one trilinear grid_sample (align_corners=True, zero padding) of the [1, C, D, H, W] volume, C = 3 * degree^2 + 1, then
 * sigma = trunc_exp(clamp(h[0])), rgb_k = sigmoid(sum_j h[1 + k * degree^2 + j] * SH_j(dir)).  `volume` is the parameter in
 * torch.channels_last_3d memory ([D][H][W][C]).  Supported: degree 1 and 3 (C = 4, 28: the kernels move four channels per lane;
 * the reference's default is 3, main_distill_mutual.py:218).
 * ---------------------------------------------------------------------------------------- */
typedef struct PvdTensorsField {
    const float* volume;
    uint32_t res[3];        /* D, H, W */
    uint32_t degree;        /* plenoxel_degree */
    float aabb[6];          /* aabb_train */
    float sigma_clip_min;
    float sigma_clip_max;
    float density_scale;
} PvdTensorsField;
int pvd_tensors_field_forward(const PvdTensorsField* field, const float* xyzs, const float* dirs, uint32_t M, float* sigmas, float* rgbs,
                              void* stream);
/* accumulates into grad_volume (same layout as volume, fp32); rows at or above *n_valid (if given) contribute nothing */
int pvd_tensors_field_backward(const PvdTensorsField* field, const float* xyzs, const float* dirs, const float* grad_sigmas,
                               const float* grad_rgbs, uint32_t M, const int32_t* n_valid, float* grad_volume, void* stream);

/* get_rays (distill_mutual/utils.py:324-404) on the device: poses [B, 4, 4] row-major cam2world, intrinsics (fx, fy, cx, cy), pixel
 * indices inds [B, N] int64 (element (b, n) at inds[b * inds_batch_stride + n]; stride 0 = one index row shared by all poses, as the
 * reference's `inds.expand([B, N])`; NULL = all H*W pixels in order) -> rays_o, rays_d [B, N, 3]. */
int pvd_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, const int64_t* inds,
                 uint32_t inds_batch_stride, uint32_t B, uint32_t N, float* rays_o, float* rays_d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVD_B200_FUSED_H */
