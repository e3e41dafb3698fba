/*
 * pvd_b200.h -- C ABI of the B200-native PVD volume-rendering hot path.
 *
 * One shared library (libpvd_b200.so, built for sm_100a by __graft_entry__.build())
 * exports everything below with C linkage: plain device pointers, sizes and a CUDA
 * stream handle -- no torch / ATen types anywhere in a signature.
 *
 * Conventions (they mirror the reference's native layer):
 *   - every pointer is CALLER-OWNED DEVICE memory unless the name ends in _host;
 *   - the caller allocates every output; the library never allocates and keeps no state,
 *     so every entry point is re-entrant per stream;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is
 *     what the reference always uses, e.g. raymarching/src/raymarching.cu:492);
 *   - return value: 0 on success, otherwise a cudaError_t value (launch/configuration
 *     errors are reported, unlike the reference, which checks nothing) or a negative
 *     PVD_E* code for argument errors (the reference's TORCH_CHECK / std::runtime_error).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef PVD_B200_H
#define PVD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVD_OK 0
#define PVD_EINVAL (-1)      /* bad argument value (NULL pointer, zero size where not allowed) */
#define PVD_EUNSUPPORTED (-2) /* e.g. level_dim not in {1,2,4,8}: gridencoder.cu:355,370 throws */

#define PVD_DTYPE_F32 0
#define PVD_DTYPE_F16 1

/* library identification / diagnostics */
int pvd_abi_version(void);                    /* bumped on any signature change */
const char* pvd_error_string(int code);       /* cudaGetErrorString for >0, own text for <0 */
int pvd_device_sm_count(int* out_sms);        /* SM count of the current device */

/* ------------------------------------------------------------------------------------------
 * raymarching -- utilities            (reference: raymarching/src/raymarching.h:7-11)
 * ---------------------------------------------------------------------------------------- */

/* raymarching.h:7  near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
 * kernel raymarching.cu:94-147. rays_o/rays_d [N,3] f32, aabb [6] f32 -> nears/fars [N] f32.
 * Miss => both FLT_MAX. */
int pvd_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                           uint32_t N, float min_near, float* nears, float* fars, void* stream);

/* raymarching.h:8  polar_from_ray(rays_o, rays_d, radius, N, coords); kernel raymarching.cu:165-200.
 * coords [N,2] f32 in [-1,1]. */
int pvd_polar_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                       float* coords, void* stream);

/* raymarching.h:9  morton3D(coords, N, indices); kernel raymarching.cu:216-228. coords [N,3] i32. */
int pvd_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);

/* raymarching.h:10 morton3D_invert(indices, N, coords); kernel raymarching.cu:239-256. */
int pvd_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);

/* raymarching.h:11 packbits(grid, N, density_thresh, bitfield); kernel raymarching.cu:270-291.
 * grid [8*N] f32, bitfield [N] u8, bit i of byte n = grid[8n+i] > thresh. */
int pvd_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream);

/* ------------------------------------------------------------------------------------------
 * raymarching -- training             (reference: raymarching/src/raymarching.h:13-15)
 * ---------------------------------------------------------------------------------------- */

/* raymarching.h:13 march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
 *                                   nears, fars, xyzs, dirs, deltas, rays, counter, perturb)
 * kernel raymarching.cu:314-483.
 *
 * Same outputs, but point offsets are DETERMINISTIC: offset(ray n) = exclusive prefix sum of
 * num_steps in ray-id order and rays[n] = (n, offset, num_steps) (the reference hands them out with
 * atomicAdd in arrival order, raymarching.cu:408-416). counter[0] += sum(num_steps), counter[1] += N
 * exactly as the reference's atomics leave them. Rays with offset + num_steps >= M write no samples
 * (raymarching.cu:419).  Rows of xyzs/dirs/deltas that no ray writes are left untouched (the Python
 * wrapper zero-fills like raymarching.py:240-242).
 *
 * `ws_i32` is caller scratch of pvd_march_rays_train_workspace_words(N, max_steps) int32 words.
 * C = cascade count, H = grid size, M = rows available in xyzs/dirs/deltas. */
uint64_t pvd_march_rays_train_workspace_words(uint32_t N, uint32_t max_steps);
int pvd_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                         float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                         uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                         float* deltas, int32_t* rays, int32_t* counter, uint32_t perturb,
                         int32_t* ws_i32, void* stream);

/* The same operation in two phases, so that a caller that does not know M yet (the reference's warm-up path,
 * raymarching.py:231,276-284, allocates N*max_steps rows and zero-fills 134 MB) can read counter[0] after the
 * count phase and allocate exactly the rows it needs.  count: occupancy march + offsets -> rays, counter, ws;
 * write: expands ws into xyzs/dirs/deltas under the drop rule for M. */
int pvd_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               const float* nears, const float* fars, int32_t* rays, int32_t* counter,
                               uint32_t perturb, int32_t* ws_i32, void* stream);
/* count phase with near_far_from_aabb fused in (writes nears/fars) and, when reuse_coarse != 0, re-using the coarse
 * rejection mask a previous call left in ws_i32 for the SAME grid (a caller whose bitfield is static between density updates). */
int pvd_march_rays_train_count_aabb(const float* rays_o, const float* rays_d, const uint8_t* grid, const float* aabb,
                                    float min_near, float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                                    uint32_t H, float* nears, float* fars, int32_t* rays, int32_t* counter, uint32_t perturb,
                                    uint32_t reuse_coarse, int32_t* ws_i32, void* stream);
/* (Re)build the coarse rejection mask in ws_i32 from `grid` alone -- what a caller that replays a captured step with
 * reuse_coarse != 0 runs after every density-grid update (renderer.py:647-773 rewrites density_bitfield every 16 steps).
 * No-op when the mask does not apply (C != 1, H != 128 or bound > 1). */
int pvd_march_coarse_mask(const uint8_t* grid, uint32_t C, uint32_t H, float bound, int32_t* ws_i32, void* stream);
int pvd_march_rays_train_write(const float* rays_o, const float* rays_d, float bound, uint32_t max_steps,
                               uint32_t N, uint32_t M, const int32_t* rays, const int32_t* ws_i32, float* xyzs,
                               float* dirs, float* deltas, void* stream);

/* raymarching.h:14 composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, weights_sum, depth, image)
 * kernel raymarching.cu:505-582. One warp per ray, shuffle scan of the transmittance. */
int pvd_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                     const int32_t* rays, uint32_t M, uint32_t N, float* weights_sum,
                                     float* depth, float* image, void* stream);

/* raymarching.h:15 composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays,
 *                                                weights_sum, image, M, N, grad_sigmas, grad_rgbs)
 * kernel raymarching.cu:607-686. grad_sigmas/grad_rgbs rows of dropped / empty rays are not written. */
int pvd_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                      const float* sigmas, const float* rgbs, const float* deltas,
                                      const int32_t* rays, const float* weights_sum, const float* image,
                                      uint32_t M, uint32_t N, float* grad_sigmas, float* grad_rgbs,
                                      void* stream);

/* ------------------------------------------------------------------------------------------
 * raymarching -- inference            (reference: raymarching/src/raymarching.h:17-19)
 * ---------------------------------------------------------------------------------------- */

/* raymarching.h:17 march_rays(...); kernel raymarching.cu:705-811. perturb is also the pcg32 seed. */
int pvd_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                   const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                   uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                   const float* fars, float* xyzs, float* dirs, float* deltas, uint32_t perturb,
                   void* stream);

/* raymarching.h:18 composite_rays(...); kernel raymarching.cu:826-909. In-place on weights_sum/depth/image. */
int pvd_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t,
                       const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                       float* depth, float* image, void* stream);

/* raymarching.h:19 compact_rays(...); kernel raymarching.cu:922-939. Deterministic (stable) compaction:
 * survivors keep their relative order (the reference's order is atomic-arrival order). */
int pvd_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                     const float* rays_t_old, int32_t* alive_counter, void* stream);

/* ------------------------------------------------------------------------------------------
 * gridencoder                          (reference: gridencoder/src/gridencoder.h:12-13)
 * ---------------------------------------------------------------------------------------- */

/* Diagnostic: per-level scale = exp2f(level*S)*H - 1 and resolution = ceil(scale)+1 exactly as the device computes
 * them (gridencoder.cu:126-127; exp2f is the CUDA approximation, so a CPU restatement can differ by an ulp). */
int pvd_grid_level_table(const int32_t* offsets, uint32_t L, float S, uint32_t H, float* scales,
                         int32_t* resolutions, void* stream);

/* gridencoder.h:12 grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H,
 *                                      calc_grad_inputs, dy_dx, gridtype, align_corners)
 * kernel gridencoder.cu:75-224.
 * inputs [B,D] f32 in [0,1]; embeddings [offsets[L], C] f32 or f16 (dtype); offsets [L+1] i32;
 * outputs [L,B,C] in the table dtype (the reference's level-major layout, grid.py:55) when sample_major == 0,
 * or [B,L,C] (== the [B, L*C] the reference produces with an extra permute copy, grid.py:84) when sample_major != 0;
 * dy_dx [B, L*D*C] table dtype or NULL when !calc_grad_inputs.
 * D in {2,3}; C in {1,2,4,8} else PVD_EUNSUPPORTED. S = log2(per_level_scale), H = base resolution. */
int pvd_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                            void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                            uint32_t H, int calc_grad_inputs, void* dy_dx, uint32_t gridtype,
                            int align_corners, int dtype, int sample_major, void* stream);

/* gridencoder.h:13 grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
 *                                       calc_grad_inputs, dy_dx, grad_inputs, gridtype, align_corners)
 * kernels gridencoder.cu:227-343. grad [L,B,C] (or [B,L,C] when sample_major); grad_embeddings is ACCUMULATED into (caller zeroes it,
 * grid.py:106). For f16 tables with C==1 the reference silently drops the gradient (its at::Half
 * atomicAdd stub has no body, gridencoder.cu:22-26); here it is accumulated correctly. */
int pvd_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                             const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                             const void* dy_dx, void* grad_inputs, uint32_t gridtype, int align_corners,
                             int dtype, int sample_major, void* stream);

/* ------------------------------------------------------------------------------------------
 * shencoder                            (reference: shencoder/src/shencoder.h:10,13)
 * ---------------------------------------------------------------------------------------- */

/* shencoder.h:10 sh_encode_forward(inputs, outputs, B, D, C, calc_grad_inputs, dy_dx); kernel shencoder.cu:27-356.
 * inputs [B,3] f32, outputs [B,C*C] f32, dy_dx [B,3*C*C] f32 or NULL. C = degree in [1,8]. */
int pvd_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C,
                          int calc_grad_inputs, float* dy_dx, void* stream);

/* shencoder.h:13 sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs); kernel shencoder.cu:359-383.
 * grad_inputs [B,3] is accumulated into (+=), as in the reference. */
int pvd_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t C,
                           const float* dy_dx, float* grad_inputs, void* stream);

/* ------------------------------------------------------------------------------------------
 * Density-grid upkeep: NeRFRenderer.update_extra_state (distill_mutual/renderer.py:647-773) without host round trips.
 *   pvd_density_grid_points   one jittered query point per cell: cell index indices[j] (Morton code; NULL = cell j, the full sweep
 *                             of renderer.py:657-699), noise [n,3] uniform in [0,1) (what torch.rand_like supplies, :690-693),
 *                             bound_cas = min(2^cas, bound); xyzs [n,3], bit-identical to the reference's torch arithmetic.
 *   pvd_density_grid_update   tmp[indices] = sigmas * density_scale (NULL indices: cell i <- sigmas[i]); valid = grid >= 0 & tmp >= 0;
 *                             grid[valid] = max(grid[valid] * decay, tmp[valid]) (:746-749); *sum += sum(clamp(grid, 0)) over the
 *                             n_cells cells of this cascade (double, caller zeroes it before the first cascade).  Duplicate
 *                             indices keep the LARGEST candidate (the reference's index_put_ keeps an arbitrary one).
 *                             tmp: [n_cells] scratch, only read/written when indices != NULL.
 *   pvd_packbits_mean         mean = *sum / count (:750-752), thresh = min(mean, density_thresh), packbits (raymarching.cu:270-291);
 *                             mean_out (device float, optional) receives the mean -- nothing is read back to the host.
 * ---------------------------------------------------------------------------------------- */
int pvd_density_grid_points(const int32_t* indices, const float* noise, uint32_t n, uint32_t H, float bound_cas, float* xyzs,
                            void* stream);
int pvd_density_grid_update(float* grid, float* tmp, const int32_t* indices, const float* sigmas, uint32_t n, uint32_t n_cells,
                            float density_scale, float decay, double* sum, void* stream);
int pvd_packbits_mean(const float* grid, uint32_t N, const double* sum, uint32_t count, float density_thresh, uint8_t* bitfield,
                      float* mean_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVD_B200_H */
