/*
 * pvd_b200_optim.h -- C ABI of the fused optimizer step that closes the training iteration around the hot path.
 *
 * Replaces, per iteration of the reference's trainers (distill_mutual/utils.py:802-819: optimizer.zero_grad(),
 * scaler.scale(loss).backward(), scaler.step(optimizer), scaler.update(); optimizer = torch.optim.AdamW(betas=(0.9, 0.99),
 * eps=1e-15), main_distill_mutual.py:327-339) and of the encoder wrapper (gridencoder/grid.py:52: embeddings.to(half) of the
 * whole table every forward; grid.py:106: zeros_like(embeddings) every backward):
 *
 *   GradScaler.unscale_ + found_inf check | AdamW over every parameter | fp32 -> fp16 table shadow | gradient zeroing
 *
 * as ONE multi-tensor kernel (plus an optional non-finite pre-pass and a one-thread bookkeeping kernel), all CUDA-graph
 * capturable: the step counter and the bias corrections live in device memory.
 *
 * Arithmetic follows torch.optim.AdamW's single-tensor CUDA path operation by operation (lerp / mul / addcmul / sqrt /
 * div-by-scalar as multiplication by the fp32 reciprocal / add eps / addcdiv), so that with the same gradients the fp32 master
 * parameters stay bit-identical to torch's (tests/test_gpu_optim.py).
 *
 * Conventions as in pvd_b200.h: caller-owned device memory, no hidden state, stream as void*, int return.
 */
#ifndef PVD_B200_OPTIM_H
#define PVD_B200_OPTIM_H

#include <stdint.h>
#include "pvd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One parameter tensor (dense in memory, any shape: the step is elementwise). 88 bytes. */
typedef struct PvdAdamSlot {
    float* param;          /* fp32 master, n elements */
    float* exp_avg;        /* fp32 first moment */
    float* exp_avg_sq;     /* fp32 second moment */
    float* grad;           /* fp32 gradient accumulator (read unless grad_f16 is set; zeroed when zero_grad != 0), may be NULL */
    const void* grad_f16;  /* optional fp16 gradient (the reduced payload of the multi-GPU exchange); read instead of `grad` */
    void* shadow_f16;      /* optional fp16 copy of the updated parameter (the hash table the field kernels gather from) */
    uint64_t n;
    double lr;             /* per group (network.py:646-683 builds groups with lr / lr2); double: torch forms lr / bias_correction1 and */
    double weight_decay;   /* 1 - lr * weight_decay in Python doubles before rounding to fp32.  Decoupled decay (AdamW), torch default 0.01 */
    float neg_step_size;   /* derived by pvd_adamw_advance: -(lr / (1 - beta1^step)) */
    float decay;           /* derived: 1 - lr * weight_decay */
    uint32_t zero_grad;
    float grad_mul;        /* per-tensor factor applied to the gradient before grad_scale (rank-count factors of the multi-GPU exchange; 1 = exact) */
} PvdAdamSlot;

/* Shared state of one optimizer (device memory; pvd_adamw_advance updates `step` and the derived fields). */
typedef struct PvdAdamState {
    double beta1, beta2;
    float eps;
    float grad_scale;      /* gradients are multiplied by this first: 1 / loss_scale (GradScaler.unscale_), times any exchange factor */
    int32_t step;          /* completed steps */
    int32_t found_inf;     /* set by pvd_grad_nonfinite; a step with found_inf != 0 is skipped (GradScaler.step) */
    int32_t skipped;       /* number of skipped steps so far */
    uint32_t flags;        /* PVD_ADAM_* */
    float w1, w2;          /* derived: 1 - beta1, 1 - beta2 as fp32 */
    float beta2_f;         /* derived */
    float inv_bc2_sqrt;    /* derived: fp32 reciprocal of fp32(sqrt(1 - beta2^step)) */
} PvdAdamState;

#define PVD_ADAM_ADDCMUL_LEFT 1u /* exp_avg_sq += (w2 * g) * g instead of w2 * (g * g) (ATen's association differs by version) */

/* found_inf |= any non-finite value in grad[0..n) (fp32) or grad_f16[0..n): GradScaler's check, one read pass. */
int pvd_grad_nonfinite(const float* grad, const void* grad_f16, uint64_t n, PvdAdamState* state, void* stream);

/* The same check over the gradients of every slot (whichever of grad / grad_f16 the step will read), one launch. */
int pvd_grad_nonfinite_slots(PvdAdamState* state, const PvdAdamSlot* slots, uint32_t n_slots, uint64_t max_n, void* stream);

/* step += 1 (unless found_inf: then skipped += 1), bias corrections and the per-slot derived fields. One thread. */
int pvd_adamw_advance(PvdAdamState* state, PvdAdamSlot* slots, uint32_t n_slots, void* stream);

/* The multi-tensor step over `n_slots` slots (device array). Skipped entirely (gradients still zeroed) when found_inf != 0.
 * found_inf is consumed (cleared) afterwards. */
int pvd_adamw_step(const PvdAdamState* state, const PvdAdamSlot* slots, uint32_t n_slots, uint64_t max_n, void* stream);

/* Several fp32 -> fp16 casts in one launch (device array of descriptors): the fp16 shadows of the vm planes / lines. */
typedef struct PvdCastDesc {
    const float* src;
    void* dst;
    uint64_t n;
} PvdCastDesc;
int pvd_cast_f32_to_f16_multi(const PvdCastDesc* descs_dev, uint32_t n_descs, uint64_t max_n, void* stream);

/* Stream `bytes` at `ptr` (16-byte aligned) into L2 with TMA prefetches: run on a side stream ahead of a kernel that gathers from the
 * buffer at random (the hash table, the vm planes) when L2 is cold. */
int pvd_l2_prefetch(const void* ptr, uint64_t bytes, void* stream);

/* fp16 -> fp32 (the write-back of an exchanged gradient payload), scaled. */
int pvd_cast_f16_to_f32(const void* src, float* dst, uint64_t elem_count, float scale, void* stream);
/* fp32 -> fp16 with a scale and saturation to +-65504 (the payload of the multi-GPU exchange; never produces inf from finite input). */
int pvd_cast_f32_to_f16_scaled(const float* src, void* dst, uint64_t elem_count, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
