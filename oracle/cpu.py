"""numpy front-end to the C oracle (oracle/pvd_oracle.c).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pvd_oracle.c")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(BUILD, "libpvd_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off: only the explicit fmaf() calls fuse, like the reference's SASS."""
    os.makedirs(BUILD, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-fvisibility=hidden",
               "-mfma", "-o", LIB, SRC, "-lm"]
        subprocess.check_call(cmd)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pcg32_jitter(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().oracle_pcg32_jitter(C.c_uint64(seed), C.c_uint32(n), _p(out, C.c_float))
    return out


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o, C.c_float), _p(rays_d, C.c_float), _p(aabb, C.c_float), C.c_uint32(N),
                                    C.c_float(min_near), _p(nears, C.c_float), _p(fars, C.c_float))
    return nears, fars


def morton3D(coords):
    coords = _i32(coords).reshape(-1, 3)
    out = np.empty(coords.shape[0], np.int32)
    lib().oracle_morton3D(_p(coords, C.c_int32), C.c_uint32(coords.shape[0]), _p(out, C.c_int32))
    return out


def morton3D_invert(indices):
    indices = _i32(indices).reshape(-1)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().oracle_morton3D_invert(_p(indices, C.c_int32), C.c_uint32(indices.shape[0]), _p(out, C.c_int32))
    return out


def packbits(grid, thresh):
    grid = _f32(grid).reshape(-1)
    N = grid.shape[0] // 8
    out = np.empty(N, np.uint8)
    lib().oracle_packbits(_p(grid, C.c_float), C.c_uint32(N), C.c_float(thresh), _p(out, C.c_uint8))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, cascade, H, nears, fars, M=None, perturb=False, dt_gamma=0.0,
                     max_steps=1024, counter=None):
    """Canonical-order restatement of raymarching.march_rays_train (raymarching.py:176-289 + raymarching.cu:314-483).

    Returns (xyzs[M,3], dirs[M,3], deltas[M,2], rays[N,3], counter[2]); rows no ray wrote are zero (raymarching.py:240-242).
    """
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    nears, fars = _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    if M is None:
        M = N * max_steps
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    if counter is None:
        counter = np.zeros(2, np.int32)
    lib().oracle_march_rays_train(_p(rays_o, C.c_float), _p(rays_d, C.c_float), _p(bitfield, C.c_uint8), C.c_float(bound),
                                  C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(N), C.c_uint32(cascade),
                                  C.c_uint32(H), C.c_uint32(M), _p(nears, C.c_float), _p(fars, C.c_float),
                                  _p(xyzs, C.c_float), _p(dirs, C.c_float), _p(deltas, C.c_float), _p(rays, C.c_int32),
                                  _p(counter, C.c_int32), C.c_uint32(int(perturb)))
    return xyzs, dirs, deltas, rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, cascade, H, nears, fars,
               perturb=0, dt_gamma=0.0, max_steps=1024):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive, rays_t, nears, fars = _i32(rays_alive), _f32(rays_t), _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(rays_alive, C.c_int32), _p(rays_t, C.c_float),
                            _p(rays_o, C.c_float), _p(rays_d, C.c_float), C.c_float(bound), C.c_float(dt_gamma),
                            C.c_uint32(max_steps), C.c_uint32(cascade), C.c_uint32(H), _p(bitfield, C.c_uint8),
                            _p(nears, C.c_float), _p(fars, C.c_float), _p(xyzs, C.c_float), _p(dirs, C.c_float),
                            _p(deltas, C.c_float), C.c_uint32(int(perturb)))
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, deltas, rays):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas, C.c_float), _p(rgbs, C.c_float), _p(deltas, C.c_float),
                                              _p(rays, C.c_int32), C.c_uint32(M), C.c_uint32(N), _p(ws, C.c_float),
                                              _p(depth, C.c_float), _p(image, C.c_float))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image):
    grad_ws, grad_image = _f32(grad_ws), _f32(grad_image)
    sigmas, rgbs, deltas, rays, ws, image = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays), _f32(ws), _f32(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)  # zeros_like, raymarching.py:339-340
    lib().oracle_composite_rays_train_backward(_p(grad_ws, C.c_float), _p(grad_image, C.c_float), _p(sigmas, C.c_float),
                                               _p(rgbs, C.c_float), _p(deltas, C.c_float), _p(rays, C.c_int32),
                                               _p(ws, C.c_float), _p(image, C.c_float), C.c_uint32(M), C.c_uint32(N),
                                               _p(gs, C.c_float), _p(gc, C.c_float))
    return gs, gc


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    """In place on rays_t / weights_sum / depth / image (float32 contiguous numpy arrays)."""
    rays_alive = _i32(rays_alive)
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    lib().oracle_composite_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(rays_alive, C.c_int32), _p(rays_t, C.c_float),
                                _p(sigmas, C.c_float), _p(rgbs, C.c_float), _p(deltas, C.c_float), _p(weights_sum, C.c_float),
                                _p(depth, C.c_float), _p(image, C.c_float))


def compact_rays(n_alive, rays_alive_old, rays_t_old):
    rays_alive_old, rays_t_old = _i32(rays_alive_old), _f32(rays_t_old)
    ra, rt = np.zeros_like(rays_alive_old), np.zeros_like(rays_t_old)
    cnt = np.zeros(1, np.int32)
    lib().oracle_compact_rays(C.c_uint32(n_alive), _p(ra, C.c_int32), _p(rays_alive_old, C.c_int32), _p(rt, C.c_float),
                              _p(rays_t_old, C.c_float), _p(cnt, C.c_int32))
    return ra, rt, int(cnt[0])


def grid_offsets(input_dim=3, num_levels=14, base_resolution=16, log2_hashmap_size=19, per_level_scale=None,
                 desired_resolution=None, align_corners=False):
    """Level table of GridEncoder.__init__ (gridencoder/grid.py:157-190). Returns (offsets int32[L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


_scale_keepalive = None


def set_level_scales(scales=None):
    """Pin the per-level scales to the values a CUDA device computed (None restores libm exp2f)."""
    global _scale_keepalive
    if scales is None:
        _scale_keepalive = None
        lib().oracle_set_level_scales(None, C.c_uint32(0))
    else:
        _scale_keepalive = _f32(scales).copy()
        lib().oracle_set_level_scales(_p(_scale_keepalive, C.c_float), C.c_uint32(_scale_keepalive.shape[0]))


def grid_level_info(offsets, S, H):
    offsets = _i32(offsets)
    L = offsets.shape[0] - 1
    scales, res = np.empty(L, np.float32), np.empty(L, np.int32)
    lib().oracle_grid_level_info(_p(offsets, C.c_int32), C.c_uint32(L), C.c_float(S), C.c_uint32(H), _p(scales, C.c_float),
                                 _p(res, C.c_int32))
    return scales, res


def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                        align_corners=False):
    """Returns outputs [B, L*C] (after the permute of grid.py:84) and dy_dx [B, L*D*C] or None. fp32 table."""
    inputs, embeddings, offsets = _f32(inputs), _f32(embeddings), _i32(offsets)
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, embeddings.shape[1]
    S = np.float32(np.log2(per_level_scale))
    out = np.empty((L, B, Cc), np.float32)
    dy_dx = np.empty((B, L * D * Cc), np.float32) if calc_grad_inputs else None
    lib().oracle_grid_encode_forward(_p(inputs, C.c_float), _p(embeddings, C.c_float), _p(offsets, C.c_int32),
                                     _p(out, C.c_float), C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc), C.c_uint32(L),
                                     C.c_float(S), C.c_uint32(base_resolution), C.c_int(int(calc_grad_inputs)),
                                     _p(dy_dx, C.c_float) if calc_grad_inputs else None, C.c_uint32(gridtype),
                                     C.c_int(int(align_corners)))
    return np.ascontiguousarray(out.transpose(1, 0, 2).reshape(B, L * Cc)), dy_dx


def grid_encode_backward(grad, inputs, embeddings_shape, offsets, per_level_scale, base_resolution, dy_dx=None, gridtype=0,
                         align_corners=False):
    """grad [B, L*C] -> (grad_embeddings [sO, C], grad_inputs [B, D] or None); grid.py:93-136."""
    inputs, offsets = _f32(inputs), _i32(offsets)
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, embeddings_shape[1]
    g = np.ascontiguousarray(_f32(grad).reshape(B, L, Cc).transpose(1, 0, 2))
    S = np.float32(np.log2(per_level_scale))
    ge = np.zeros(embeddings_shape, np.float32)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    lib().oracle_grid_encode_backward(_p(g, C.c_float), _p(inputs, C.c_float), _p(offsets, C.c_int32), _p(ge, C.c_float),
                                      C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc), C.c_uint32(L), C.c_float(S),
                                      C.c_uint32(base_resolution), C.c_int(int(dy_dx is not None)),
                                      _p(_f32(dy_dx), C.c_float) if dy_dx is not None else None,
                                      _p(gi, C.c_float) if gi is not None else None, C.c_uint32(gridtype),
                                      C.c_int(int(align_corners)))
    return ge, gi


def sh_encode_forward(inputs, degree):
    inputs = _f32(inputs).reshape(-1, 3)
    B = inputs.shape[0]
    out = np.empty((B, degree * degree), np.float32)
    rc = lib().oracle_sh_encode_forward(_p(inputs, C.c_float), _p(out, C.c_float), C.c_uint32(B), C.c_uint32(degree))
    if rc != 0:
        raise ValueError("C oracle covers SH degree 1..4; use oracle.sh_reference for higher degrees")
    return out


# ------------------------------------------------------------------------------------------------ density-grid upkeep
def _morton_invert_np(ind):
    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xC30C30C3
        x = (x | (x >> 4)) & 0x0F00F00F
        x = (x | (x >> 8)) & 0xFF0000FF
        x = (x | (x >> 16)) & 0x0000FFFF
        return x
    ind = ind.astype(np.uint32)
    return np.stack([compact(ind), compact(ind >> 1), compact(ind >> 2)], axis=-1)


def density_grid_points(indices, noise, H, bound_cas):
    """Query points of NeRFRenderer.update_extra_state (distill_mutual/renderer.py:679-694) for the cells `indices` (Morton codes;
    None = all H^3 cells in Morton order) with the uniform noise `noise` [n,3]: fp32, operation by operation as torch evaluates
        xyzs = 2 * coords.float() / (H - 1) - 1 ; cas = xyzs * (bound - half) ; cas += (rand * 2 - 1) * half
    on the GPU (ATen's CUDA division by a host scalar multiplies by the fp32 reciprocal; Python scalars are cast to fp32)."""
    f = np.float32
    n = noise.shape[0]
    ind = np.arange(n, dtype=np.uint32) if indices is None else np.asarray(indices)
    c = _morton_invert_np(ind).astype(f)
    rcp = f(1.0) / f(H - 1)
    base = (f(2.0) * c) * rcp - f(1.0)
    scale = f(float(bound_cas) - float(bound_cas) / H)
    half = f(float(bound_cas) / H)
    return (base * scale + (noise.astype(f) * f(2.0) - f(1.0)) * half).astype(f)


def density_grid_update(grid, sigmas, indices, density_scale, decay):
    """tmp[indices] = sigmas * density_scale; valid = (grid >= 0) & (tmp >= 0); grid[valid] = max(grid[valid] * decay, tmp[valid])
    (renderer.py:737,746-749).  Duplicate indices keep the largest candidate (the reference keeps an arbitrary one).  Returns the new
    grid (fp32) and sum(clamp(grid, 0)) in double."""
    f = np.float32
    g = grid.astype(f).copy()
    tmp = np.full(g.shape, -1.0, f)
    s = (sigmas.astype(f) * f(density_scale)).astype(f)
    if indices is None:
        tmp[:] = s
    else:
        np.maximum.at(tmp, np.asarray(indices, dtype=np.int64), s)
    valid = (g >= 0) & (tmp >= 0)
    g[valid] = np.maximum(g[valid] * f(decay), tmp[valid])
    return g, float(np.clip(g, 0, None).astype(np.float64).sum())


def packbits_mean(grid, total_sum, density_thresh):
    """mean = sum / count; thresh = min(mean, density_thresh); packbits (renderer.py:750-759)."""
    mean = np.float32(total_sum / grid.size)
    return packbits(grid.reshape(-1), float(min(mean, np.float32(density_thresh)))), float(mean)
