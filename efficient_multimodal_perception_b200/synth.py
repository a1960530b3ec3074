"""Seeded synthetic nuScenes-shaped inputs (SURVEY.md §8d, S1–S5). CPU tensors, fp32.

There is no dataset in this environment; the shapes follow the reference's loaders:
points are rows of 11 fp32 (x, y, z, intensity, ring + 6 SAM label columns —
sam/create_sam_masks.py:116-167, configs/nuscenes_surf_sam.py:37-43), a sweep is 32 rings
(tools/create_range_images.py:10-13), the occupancy grid is 200x200x16 (loading.py:103).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

# Geometry A — configs/point_triplane.py:8-24 (config-exact)
GEOM_A = dict(pc_range=[-25.0, -25.0, -5.0, 25.0, 25.0, 3.0], voxel_size=(0.4, 0.4, 0.1),
              grid_size=[128, 128, 80], split=[25, 25, 20], channels=128)
# Geometry B — BASELINE.json's "200x200x16 grid" (not a reference config)
GEOM_B = dict(pc_range=[-50.0, -50.0, -5.0, 50.0, 50.0, 3.0], voxel_size=(0.5, 0.5, 0.5),
              grid_size=[200, 200, 16], split=[25, 25, 16], channels=128)
# configs/triplane_occ.py:13-17
OCC = dict(voxel_size=(0.5, 0.5, 0.5), triplane_voxel_size=(0.4, 0.4, 0.1),
           triplane_range=[-25.0, -25.0, -5.0, 25.0, 25.0, 3.0], occ_range=[-25.0, -25.0, -5.0, 25.0, 25.0, 3.0],
           channels=32, plane=128)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def lidar_sweep(n: int = 34720, seed: int = 1001, jitter: float = 0.0) -> torch.Tensor:
    """One sweep: [n, 11] fp32. 32 rings at -30..+10 deg elevation, azimuth U(0, 2pi), range a
    ground-plane / obstacle mixture clipped to 1..70 m (about 70 % lands inside +-25 m)."""
    g = _gen(seed)
    ring = torch.randint(0, 32, (n,), generator=g)
    elev = torch.deg2rad(-30.0 + ring.float() * (40.0 / 31.0))
    azim = torch.rand(n, generator=g) * (2 * math.pi)
    sensor_h = 1.84
    ground = sensor_h / torch.clamp(-torch.sin(elev), min=1e-3)  # hits the ground plane
    obstacle = 2.0 + torch.rand(n, generator=g).pow(2) * 60.0
    is_obst = (torch.rand(n, generator=g) < 0.45) | (elev >= 0)
    rng = torch.where(is_obst, torch.minimum(obstacle, torch.where(elev < 0, ground, obstacle)), ground)
    rng = (rng * (1.0 + 0.02 * torch.randn(n, generator=g))).clamp(1.0, 70.0)
    x = rng * torch.cos(elev) * torch.cos(azim)
    y = rng * torch.cos(elev) * torch.sin(azim)
    z = rng * torch.sin(elev)  # sensor frame: ground at about -1.84 m
    if jitter > 0:
        off = (torch.rand(3, generator=g) * 2 - 1) * jitter
        x, y = x + off[0], y + off[1]
    pts = torch.zeros(n, 11)
    pts[:, 0], pts[:, 1], pts[:, 2] = x, y, z
    pts[:, 3] = torch.rand(n, generator=g) * 255.0
    pts[:, 4] = ring.float()
    lab = torch.randint(0, 61, (n, 6), generator=g).float()
    lab[torch.rand(n, 6, generator=g) < 0.4] = 0.0
    pts[:, 5:] = lab
    return pts


def multi_sweep(sweeps: int = 10, n_per_sweep: int = 35000, seed: int = 1005) -> torch.Tensor:
    """S5: `sweeps` accumulated sweeps with +-2 m ego-motion jitter -> [sweeps*n, 11]."""
    return torch.cat([lidar_sweep(n_per_sweep, seed * 131 + s, jitter=2.0) for s in range(sweeps)], 0)


def point_features(n: int, channels: int, seed: int) -> torch.Tensor:
    return torch.randn(n, channels, generator=_gen(seed))


def triplane_stacked(batch: int = 1, channels: int = 32, size: int = 128, seed: int = 1002) -> torch.Tensor:
    """MixVisionTransformer neck output viewed as [B, 3, C, H, W] (triplane_occ.py:178-179)."""
    return torch.randn(batch, 3, channels, size, size, generator=_gen(seed))


def triplane_list(batch: int, channels: int, grid_size: Sequence[int], seed: int) -> List[torch.Tensor]:
    """PointTriplane planes [B,C,X,Y], [B,C,Y,Z], [B,C,X,Z] (point_triplane_projector.py:113-117)."""
    g = _gen(seed)
    X, Y, Z = grid_size
    return [torch.randn(batch, channels, X, Y, generator=g), torch.randn(batch, channels, Y, Z, generator=g),
            torch.randn(batch, channels, X, Z, generator=g)]


def roi_lattice(occ_range=OCC["occ_range"], voxel_size=OCC["voxel_size"]) -> torch.Tensor:
    """TriplaneOcc.roi() voxel centres, [99, 99, 16, 3] at the config values (triplane_occ.py:291-318)."""
    min_x = int((abs(-50 - occ_range[0]) + 0.5) / voxel_size[0])
    min_y = int((abs(-50 - occ_range[1]) + 0.5) / voxel_size[1])
    max_x = int((abs(50 - occ_range[0]) - 0.5) / voxel_size[0])
    max_y = int((abs(50 - occ_range[1]) - 0.5) / voxel_size[1])
    X, Y = max_x - min_x + 1, max_y - min_y + 1
    Z = int((occ_range[5] - occ_range[2]) / voxel_size[2])
    return lattice((X, Y, Z), voxel_size, occ_range[:3])


def lattice(dims: Tuple[int, int, int], voxel_size, origin) -> torch.Tensor:
    """Voxel centres (i + 0.5) * vs + lo on a dims grid -> [X, Y, Z, 3], z fastest."""
    X, Y, Z = dims
    xs = torch.arange(X, dtype=torch.float32).view(X, 1, 1).expand(X, Y, Z)
    ys = torch.arange(Y, dtype=torch.float32).view(1, Y, 1).expand(X, Y, Z)
    zs = torch.arange(Z, dtype=torch.float32).view(1, 1, Z).expand(X, Y, Z)
    ref = torch.stack((xs, ys, zs), -1).clone()
    for a in range(3):
        ref[..., a] = (ref[..., a] + 0.5) * voxel_size[a] + origin[a]
    return ref


def occ_gt_lattice() -> torch.Tensor:
    """BASELINE.json's 640k queries: the full 200x200x16 occupancy-GT grid (loading.py:103) at 0.5 m
    over +-50 m x [-5, 3] m -> [200, 200, 16, 3]; 3/4 of it lies outside the +-25 m planes."""
    return lattice((200, 200, 16), (0.5, 0.5, 0.5), (-50.0, -50.0, -5.0))


def uniform_queries(n: int, rng=OCC["triplane_range"], seed: int = 1002) -> torch.Tensor:
    """n uniform-random in-range queries (worst-case locality) -> [n, 3]."""
    u = torch.rand(n, 3, generator=_gen(seed + 7))
    lo = torch.tensor(rng[:3])
    hi = torch.tensor(rng[3:])
    return lo + u * (hi - lo)


def range_image_points(batch: int = 8, seed: int = 1003) -> torch.Tensor:
    """S3: range_points [B, 32, 1024, 3] with ~30 % empty pixels = (0,0,0)."""
    out = torch.zeros(batch, 32, 1024, 3)
    for b in range(batch):
        g = _gen(seed * 17 + b)
        ring = torch.arange(32).view(32, 1).expand(32, 1024)
        elev = torch.deg2rad(-30.0 + ring.float() * (40.0 / 31.0))
        azim = (torch.arange(1024).float().view(1, 1024).expand(32, 1024) + 0.5) * (2 * math.pi / 1024)
        rng = (2.0 + torch.rand(32, 1024, generator=g).pow(2) * 60.0)
        ground = 1.84 / torch.clamp(-torch.sin(elev), min=1e-3)
        rng = torch.where(elev < 0, torch.minimum(rng, ground), rng).clamp(1.0, 70.0)
        p = torch.stack((rng * torch.cos(elev) * torch.cos(azim), rng * torch.cos(elev) * torch.sin(azim),
                         rng * torch.sin(elev)), -1)
        p[torch.rand(32, 1024, generator=g) < 0.3] = 0.0
        out[b] = p
    return out


@dataclass
class CameraRig:
    lidar2image: torch.Tensor  # [6, 4, 4]
    imgs_aug: list            # per camera dict(resize, crop, flip)
    img_shape: Tuple[int, int]  # PIL size (W, H) after augmentation = (512, 256), as transforms_3d.py:169 stores it


def camera_rig(seed: int = 1004) -> CameraRig:
    """Six pinhole cameras at 60-degree yaw steps; 1600x900 images resized by 0.525 and cropped to
    256x512 with (164, 216) (transforms_3d.py:69-77), no flip. img_shape is PIL's (W, H) = (512, 256)
    (`data["img_shape"] = new_imgs[0].size`, transforms_3d.py:169), so the reference's
    `resize_dims = img_shape[::-1]` is (H, W) = (256, 512)."""
    g = _gen(seed)
    fx = fy = 1266.0
    cx, cy = 816.0, 491.0
    K = torch.tensor([[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    mats = []
    for c in range(6):
        yaw = math.radians(60.0 * c) + float(torch.randn((), generator=g)) * 0.01
        # lidar (x fwd, y left, z up) -> camera (x right, y down, z fwd), rotated by yaw about z
        cz, sz = math.cos(yaw), math.sin(yaw)
        R_yaw = torch.tensor([[cz, sz, 0], [-sz, cz, 0], [0, 0, 1.0]])
        R_axes = torch.tensor([[0, -1.0, 0], [0, 0, -1.0], [1.0, 0, 0]])
        R = R_axes @ R_yaw
        t = torch.tensor([0.0, 0.3, -0.1])
        T = torch.eye(4)
        T[:3, :3] = R
        T[:3, 3] = t
        mats.append(K @ T)
    augs = [dict(resize=0.525, crop=(164, 216), flip=False) for _ in range(6)]
    return CameraRig(torch.stack(mats), augs, (512, 256))


def batch_offsets(sizes: Sequence[int]) -> torch.Tensor:
    off = torch.zeros(len(sizes) + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor(list(sizes), dtype=torch.int64), 0)
    return off
