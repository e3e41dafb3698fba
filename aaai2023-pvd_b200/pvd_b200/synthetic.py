"""Deterministic synthetic workload for the hot path: a Lego-shaped occupancy grid and NeRF-synthetic style rays.

No dataset exists offline, so the benchmark and the parity tests use the scene SURVEY.md section 8d specifies:
  * occupancy = union of axis-aligned boxes (base slab, tracks, body, cab, arm, bucket) inside [-0.8, 0.8]^3,
    rasterised at cell centres into the Morton-ordered density grid the reference keeps
    (`density_grid [cascade, H^3]`, distill_mutual/renderer.py:99-106) and packed with the packbits rule
    (bit i of byte n = grid[8n+i] > thresh, raymarching.cu:283-290);
  * cameras = 800x800, focal 1111.11 (camera_angle_x 0.6911 rad), poses pose_spherical(theta, phi, 4.0) mapped by
    nerf_matrix_to_ngp(scale 0.8) (distill_mutual/utils.py:53-97), one pose per step, N random pixels
    (torch.randint, utils.py:354), ray directions as get_rays computes them (utils.py:391-399).
Everything is a pure function of the seed.  numpy/torch on the CPU; callers move tensors to the GPU.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

# (xmin, ymin, zmin, xmax, ymax, zmax) in world units; y is up in the NGP frame used by the reference
LEGO_BOXES = np.array([
    [-0.62, -0.40, -0.36, 0.62, -0.30, 0.36],   # base slab
    [-0.66, -0.52, -0.44, 0.66, -0.38, -0.28],  # left track
    [-0.66, -0.52, 0.28, 0.66, -0.38, 0.44],    # right track
    [-0.40, -0.30, -0.28, 0.30, 0.00, 0.28],    # body
    [-0.34, 0.00, -0.20, 0.02, 0.26, 0.20],     # cab
    [0.10, -0.10, -0.06, 0.58, 0.02, 0.06],     # arm
    [0.50, -0.22, -0.24, 0.74, 0.10, 0.24],     # bucket
], dtype=np.float64)


def _part1by2(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3d_np(x, y, z):
    return _part1by2(x) | (_part1by2(y) << np.uint32(1)) | (_part1by2(z) << np.uint32(2))


def lego_density_grid(grid_size: int = 128, bound: float = 1.0, cascade: int = 1, boxes: np.ndarray = LEGO_BOXES) -> np.ndarray:
    """density_grid [cascade, H^3] float32 in Morton order: 1 inside the box union, 0 outside."""
    H = grid_size
    idx = np.arange(H)
    grid = np.zeros((cascade, H ** 3), np.float32)
    for cas in range(cascade):
        mip_bound = min(2.0 ** cas, bound)
        centre = (-1.0 + (2.0 * idx + 1.0) / H) * mip_bound  # centre of march cell i (raymarching.cu:377)
        X, Y, Z = np.meshgrid(centre, centre, centre, indexing="ij")
        occ = np.zeros((H, H, H), bool)
        for b in boxes:
            occ |= (X >= b[0]) & (X <= b[3]) & (Y >= b[1]) & (Y <= b[4]) & (Z >= b[2]) & (Z <= b[5])
        ix, iy, iz = np.meshgrid(idx, idx, idx, indexing="ij")
        m = morton3d_np(ix.ravel(), iy.ravel(), iz.ravel())
        grid[cas, m] = occ.ravel().astype(np.float32)
    return grid


def pack_bitfield(density_grid: np.ndarray, thresh: float = 0.01) -> np.ndarray:
    bits = (density_grid.reshape(-1, 8) > thresh).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def lego_bitfield(grid_size: int = 128, bound: float = 1.0, cascade: int = 1, thresh: float = 0.01):
    g = lego_density_grid(grid_size, bound, cascade)
    bf = pack_bitfield(g, thresh)
    return g, bf, hashlib.sha256(bf.tobytes()).hexdigest()


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """Camera-to-world of a camera on a sphere looking at the origin (NeRF-synthetic convention)."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    trans = np.eye(4, dtype=np.float32)
    trans[2, 3] = radius
    rphi = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], np.float32)
    rth = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32)
    return flip @ (rth @ (rphi @ trans))


def nerf_to_ngp(pose: np.ndarray, scale: float = 0.8) -> np.ndarray:
    return np.array([
        [pose[1, 0], -pose[1, 1], -pose[1, 2], pose[1, 3] * scale],
        [pose[2, 0], -pose[2, 1], -pose[2, 2], pose[2, 3] * scale],
        [pose[0, 0], -pose[0, 1], -pose[0, 2], pose[0, 3] * scale],
        [0, 0, 0, 1]], dtype=np.float32)


def random_pose(rng: np.random.Generator, radius: float = 4.0) -> np.ndarray:
    theta = rng.uniform(-180.0, 180.0)
    phi = rng.uniform(-80.0, 0.0)
    return nerf_to_ngp(pose_spherical(theta, phi, radius))


def rays_for_pose(pose: np.ndarray, n_rays: int, gen: torch.Generator, H: int = 800, W: int = 800,
                  focal: float = 1111.1111):
    """rays_o, rays_d [n_rays, 3] float32 (CPU) for random pixels of one camera."""
    inds = torch.randint(0, H * W, size=[n_rays], generator=gen)
    i = (inds % W).float() + 0.5
    j = (inds // W).float() + 0.5
    cx, cy = W / 2, H / 2
    dirs = torch.stack(((i - cx) / focal, (j - cy) / focal, torch.ones_like(i)), dim=-1)
    dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    p = torch.from_numpy(pose)
    rays_d = dirs @ p[:3, :3].T
    rays_o = p[:3, 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


def make_ray_batches(n_batches: int, n_rays: int, seed: int = 0):
    """A list of (rays_o, rays_d) CPU tensors, one random pose each; deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    return [rays_for_pose(random_pose(rng), n_rays, gen) for _ in range(n_batches)]
