"""CPU oracle for the triplane hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module. The product path (``efficient_multimodal_perception_b200``) never does
and fails loudly when ``libtriplane.so`` is missing.

What it is: a torch-CPU restatement, in the reference's own op order, of the functions SURVEY.md
§8(a) lists. Where the reference calls torch ops (``F.grid_sample``, ``torch.unique``, division,
boolean-mask indexing) the oracle calls the SAME torch ops, so on those rows it *is* the reference
arithmetic. Two third-party ops are absent from this image and from /root/reference and are restated
from their published semantics:

* ``torch_scatter.scatter_max`` (pulled in by "Pytorch-geometric == 2.1.0", reference README.md:48;
  call site point_triplane_projector.py:104): per-index max, empty -> 0. Restated with
  ``Tensor.scatter_reduce_('amax', include_self=False)``.
* ``spconv.pytorch.SparseMaxPool3d`` + ``SparseConvTensor.dense()`` (spconv == 2.1.21, reference
  README.md:49; call sites point_triplane_projector.py:53-58,111-115): kernel == stride, padding 0
  -> out index = idx // k, positions with idx // k >= (n - k) // k + 1 dropped, max over the active
  inputs of a window, inactive outputs dense() to 0. Whether spconv's native path clamps at 0
  (zero-initialised output buffer) cannot be verified here: ``clamp_zero`` selects it.

Pinning status (see tests/golden/make_golden.py and DESIGN.md):
* a1 voxelize_points, a4 the five sample_points_triplane variants, a5 roi(), a2 point_to_cam:
  PINNED — golden vectors are produced by executing the reference's own function bodies (extracted
  from /root/reference with ``ast``; mmcv/mmdet are not importable) on torch-CPU, and this oracle
  must reproduce them bit for bit.
* a3 PointTriplaneProjector.forward: the reference class itself is executed with stub modules for
  torch_scatter / spconv built from the restatements above, so everything except those two
  third-party ops is pinned; the two ops themselves are "parity unpinned".
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# a1  voxelize_points — mmdet3d/models/detectors/point_triplane.py:133-161
# ------------------------------------------------------------------------------------------------
def voxelize_points(points: Sequence[torch.Tensor], pc_range, voxel_size):
    cropped_points, grid_ind = [], []
    for pts in points:
        crop_mask = (
            (pts[..., 0] > pc_range[0]) & (pts[..., 0] < pc_range[3])
            & (pts[..., 1] > pc_range[1]) & (pts[..., 1] < pc_range[4])
            & (pts[..., 2] > pc_range[2]) & (pts[..., 2] < pc_range[5])
        )  # :148-150
        cropped = pts[crop_mask]  # :152
        voxel_ind = torch.zeros((cropped.shape[0], 3), dtype=pts.dtype)
        voxel_ind[..., 0] = (cropped[..., 0] - pc_range[0]) / voxel_size[0]  # :154
        voxel_ind[..., 1] = (cropped[..., 1] - pc_range[1]) / voxel_size[1]
        voxel_ind[..., 2] = (cropped[..., 2] - pc_range[2]) / voxel_size[2]
        cropped_points.append(cropped)
        grid_ind.append(voxel_ind.type(torch.int))  # :159
    return cropped_points, grid_ind


def voxel_index_mul_rcp(points: torch.Tensor, pc_range, voxel_size) -> torch.Tensor:
    """torch-CUDA's evaluation of ``(p - lo) / python_float``: multiply by the fp32 reciprocal
    (ATen BinaryDivTrueKernel.cu scalar fast path). Emulated on CPU for the MUL_RCP arith mode."""
    out = torch.zeros((points.shape[0], 3), dtype=torch.float32)
    for a in range(3):
        rcp = (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(voxel_size[a], dtype=torch.float32))
        out[:, a] = (points[:, a] - pc_range[a]) * rcp
    return out.type(torch.int)


# ------------------------------------------------------------------------------------------------
# third-party restatements
# ------------------------------------------------------------------------------------------------
def scatter_max(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter_max(src, index, dim=0)[0]: per-row max, empty rows 0."""
    out = torch.zeros((dim_size, src.shape[1]), dtype=src.dtype)
    idx = index.view(-1, 1).expand(-1, src.shape[1])
    out.scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    return out


def pooled_size(n: int, k: int) -> int:
    """spconv get_conv_output_size with stride == kernel, padding 0, dilation 1."""
    return (n - k) // k + 1


def sparse_max_pool_dense(feat: torch.Tensor, coors: torch.Tensor, grid_size, kernel, batch_size: int,
                          clamp_zero: bool = False, reduce: str = "amax") -> torch.Tensor:
    """SparseMaxPool3d(kernel, stride=kernel, padding=0)(SparseConvTensor(feat, coors, grid_size,
    batch_size)).dense() -> [B, C, X', Y', Z']."""
    out_shape = [pooled_size(int(g), int(k)) for g, k in zip(grid_size, kernel)]
    C = feat.shape[1]
    oc = coors.long().clone()
    for a in range(3):
        oc[:, 1 + a] = torch.div(oc[:, 1 + a], int(kernel[a]), rounding_mode="floor")
    valid = torch.ones(oc.shape[0], dtype=torch.bool)
    for a in range(3):
        valid &= oc[:, 1 + a] < out_shape[a]
    oc, f = oc[valid], feat[valid]
    ncell = batch_size * out_shape[0] * out_shape[1] * out_shape[2]
    lin = ((oc[:, 0] * out_shape[0] + oc[:, 1]) * out_shape[1] + oc[:, 2]) * out_shape[2] + oc[:, 3]
    dense = torch.zeros((ncell, C), dtype=feat.dtype)
    if lin.numel():
        dense.scatter_reduce_(0, lin.view(-1, 1).expand(-1, C), f, reduce=reduce, include_self=False)
    if clamp_zero:
        dense.clamp_(min=0)
    return dense.view(batch_size, out_shape[0], out_shape[1], out_shape[2], C).permute(0, 4, 1, 2, 3)


# ------------------------------------------------------------------------------------------------
# a3  the scatter part of PointTriplaneProjector.forward — point_triplane_projector.py:99-115
# ------------------------------------------------------------------------------------------------
def pool_kernels(grid_size, split) -> Tuple[int, int, int]:
    """int(grid/split) per axis — point_triplane_projector.py:53-58."""
    return tuple(int(grid_size[a] / split[a]) for a in range(3))


def encode_pooled(feats: torch.Tensor, cat_pt_ind: torch.Tensor, grid_size, split, batch_size: int,
                  reduce: str = "max", clamp_zero: bool = False):
    """feats [N',C], cat_pt_ind [N',4] = (b,x,y,z) int -> the three channels-last flattened tensors
    that feed mlp_xy / mlp_yz / mlp_xz, plus unq_cnt and unq (projector.py:99-115).

    reduce='max' is the reference; 'mean' is the north-star extension (sum/count per pooled cell).
    """
    grid = [int(g) for g in grid_size]
    inb = torch.ones(cat_pt_ind.shape[0], dtype=torch.bool)
    for a in range(3):  # spconv would index out of bounds; the product drops such points
        inb &= (cat_pt_ind[:, 1 + a] >= 0) & (cat_pt_ind[:, 1 + a] < grid[a])
    cat_pt_ind, feats = cat_pt_ind[inb], feats[inb]
    kx, ky, kz = pool_kernels(grid, split)
    if reduce == "max":
        unq, unq_inv, unq_cnt = torch.unique(cat_pt_ind, return_inverse=True, return_counts=True, dim=0)  # :99
        pooled = scatter_max(feats, unq_inv, unq.shape[0])  # :104
        coors = unq.int()
        red = "amax"
    else:
        # mean over the POINTS of a pooled cell (not mean of voxel means)
        unq, unq_cnt = torch.unique(cat_pt_ind, return_counts=True, dim=0)
        pooled, coors, red = feats, cat_pt_ind.int(), "mean"
    xy = sparse_max_pool_dense(pooled, coors, grid, [1, 1, kz], batch_size, clamp_zero, red)
    yz = sparse_max_pool_dense(pooled, coors, grid, [kx, 1, 1], batch_size, clamp_zero, red)
    xz = sparse_max_pool_dense(pooled, coors, grid, [1, ky, 1], batch_size, clamp_zero, red)
    xy = xy.permute(0, 2, 3, 4, 1).flatten(start_dim=3)  # :113
    yz = yz.permute(0, 3, 4, 2, 1).flatten(start_dim=3)  # :114
    xz = xz.permute(0, 2, 4, 3, 1).flatten(start_dim=3)  # :115
    return xy.contiguous(), yz.contiguous(), xz.contiguous(), unq, unq_cnt


def cat_indices(grid_ind: Sequence[torch.Tensor]) -> torch.Tensor:
    """F.pad(grid_ind[i], (1,0), value=i) + cat — projector.py:80-85."""
    return torch.cat([F.pad(g, (1, 0), "constant", value=i) for i, g in enumerate(grid_ind)], dim=0)


def cell_counts(cat_pt_ind: torch.Tensor, grid_size, split, batch_size: int):
    """points per pooled cell for the three planes, flattened in output order (xy | yz | xz)."""
    grid = [int(g) for g in grid_size]
    k = pool_kernels(grid, split)
    P = [pooled_size(grid[a], k[a]) for a in range(3)]
    ind = cat_pt_ind.long()
    inb = torch.ones(ind.shape[0], dtype=torch.bool)
    for a in range(3):
        inb &= (ind[:, 1 + a] >= 0) & (ind[:, 1 + a] < grid[a])
    ind = ind[inb]
    b, x, y, z = ind[:, 0], ind[:, 1], ind[:, 2], ind[:, 3]
    px, py, pz = x // k[0], y // k[1], z // k[2]
    outs = []
    for lin, ok, n in (
        ((((b * grid[0] + x) * grid[1] + y) * P[2] + pz), pz < P[2], batch_size * grid[0] * grid[1] * P[2]),
        ((((b * grid[1] + y) * grid[2] + z) * P[0] + px), px < P[0], batch_size * grid[1] * grid[2] * P[0]),
        ((((b * grid[0] + x) * grid[2] + z) * P[1] + py), py < P[1], batch_size * grid[0] * grid[2] * P[1]),
    ):
        outs.append(torch.bincount(lin[ok], minlength=n).to(torch.int32))
    return outs


# ------------------------------------------------------------------------------------------------
# a4  sample_points_triplane — the five variants
# ------------------------------------------------------------------------------------------------
def _sample3(p0, p1, p2, voxel_coors):
    xy = F.grid_sample(p0, voxel_coors[..., [0, 1]], mode="bilinear", padding_mode="zeros", align_corners=False)
    yz = F.grid_sample(p1, voxel_coors[..., [1, 2]], mode="bilinear", padding_mode="zeros", align_corners=False)
    xz = F.grid_sample(p2, voxel_coors[..., [0, 2]], mode="bilinear", padding_mode="zeros", align_corners=False)
    return xy + yz + xz


def sample_points_triplane_stacked(triplane: torch.Tensor, points: torch.Tensor, lo, vs) -> torch.Tensor:
    """triplane [B,3,C,H,W]; points [B,h,w,3] (triplane.py:490-514) or [B,h,w,d,3]
    (triplane_occ.py:321-348, triplane_elev.py:286-313)."""
    voxel_coors = torch.zeros_like(points)
    voxel_coors[..., 0] = (points[..., 0] - lo[0]) / vs[0]
    voxel_coors[..., 1] = (points[..., 1] - lo[1]) / vs[1]
    voxel_coors[..., 2] = (points[..., 2] - lo[2]) / vs[2]
    voxel_coors = voxel_coors / (triplane.shape[-1] / 2) - 1
    if points.dim() == 5:
        b, h, w, d, p = voxel_coors.shape
        voxel_coors = voxel_coors.view(b, h, w * d, p)
        out = _sample3(triplane[:, 0], triplane[:, 1], triplane[:, 2], voxel_coors)
        return out.view(b, -1, h, w, d)
    return _sample3(triplane[:, 0], triplane[:, 1], triplane[:, 2], voxel_coors)


def sample_points_triplane_list(triplane: Sequence[torch.Tensor], points: torch.Tensor, lo, vs,
                                grid_size) -> torch.Tensor:
    """triplane = [xy [B,C,X,Y], yz [B,C,Y,Z], xz [B,C,X,Z]]; points [B,h,w,3]
    (point_triplane.py:439-466) or [B,h,w,d,3] (point_triplane_occ.py:407-440)."""
    voxel_coors = torch.zeros_like(points)
    voxel_coors[..., 0] = (points[..., 0] - lo[0]) / vs[0]
    voxel_coors[..., 1] = (points[..., 1] - lo[1]) / vs[1]
    voxel_coors[..., 2] = (points[..., 2] - lo[2]) / vs[2]
    voxel_coors[..., 0] = voxel_coors[..., 0] / (grid_size[0] / 2) - 1
    voxel_coors[..., 1] = voxel_coors[..., 1] / (grid_size[1] / 2) - 1
    voxel_coors[..., 2] = voxel_coors[..., 2] / (grid_size[2] / 2) - 1
    if points.dim() == 5:
        b, h, w, d, p = voxel_coors.shape
        voxel_coors = voxel_coors.view(b, h, w * d, p)
        out = _sample3(triplane[0], triplane[1], triplane[2], voxel_coors)
        return out.view(b, -1, h, w, d)
    return _sample3(triplane[0], triplane[1], triplane[2], voxel_coors)


# ------------------------------------------------------------------------------------------------
# a5  query generators
# ------------------------------------------------------------------------------------------------
def roi(occ_range, voxel_size):
    """TriplaneOcc.roi — triplane_occ.py:291-318."""
    min_x = int((abs(-50 - occ_range[0]) + 0.5) / voxel_size[0])
    min_y = int((abs(-50 - occ_range[1]) + 0.5) / voxel_size[1])
    max_x = int((abs(50 - occ_range[0]) - 0.5) / voxel_size[0])
    max_y = int((abs(50 - occ_range[1]) - 0.5) / voxel_size[1])
    X = max_x - min_x + 1
    Y = max_y - min_y + 1
    Z = int((occ_range[5] - occ_range[2]) / voxel_size[2])
    xs = torch.arange(0, X).view(X, 1, 1).expand(X, Y, Z).type(torch.float32)
    ys = torch.arange(0, Y).view(1, Y, 1).expand(X, Y, Z).type(torch.float32)
    zs = torch.arange(0, Z).view(1, 1, Z).expand(X, Y, Z).type(torch.float32)
    ref_3d = torch.stack((xs, ys, zs), -1)
    ref_3d[..., 0] = (ref_3d[..., 0] + 0.5) * voxel_size[0] + occ_range[0]
    ref_3d[..., 1] = (ref_3d[..., 1] + 0.5) * voxel_size[1] + occ_range[1]
    ref_3d[..., 2] = (ref_3d[..., 2] + 0.5) * voxel_size[2] + occ_range[2]
    return (min_x, min_y, max_x, max_y), ref_3d


# ------------------------------------------------------------------------------------------------
# a2  point_to_cam — point_triplane.py:164-241
# ------------------------------------------------------------------------------------------------
def point_to_cam(points: Sequence[torch.Tensor], img_features: torch.Tensor, img_metas) -> List[torch.Tensor]:
    resize_dims = img_metas[0]["img_shape"][::-1]
    lidar2imgs = np.asarray([m["lidar2image"] for m in img_metas])
    img_augs = [m["imgs_aug"] for m in img_metas]
    lidar2imgs = points[0].new_tensor(lidar2imgs)
    cam_point_features = []
    for i, pts in enumerate(points):
        point_feature = torch.zeros((pts.shape[0], img_features.shape[2]), dtype=img_features.dtype,
                                    device=img_features.device)
        lidar2img = lidar2imgs[i]
        hom_points = torch.cat((pts[:, 0:3], torch.ones_like(pts[..., :1])), -1)
        cam_points = torch.einsum("cij, hj->chi", lidar2img, hom_points)
        cam_points = cam_points[..., 0:2] / torch.maximum(
            cam_points[..., 2:3], torch.ones_like(cam_points[..., 2:3]) * 1e-5)
        num_cam = lidar2imgs.shape[1]
        resize = [aug["resize"] for aug in img_augs[i]]
        crop = [aug["crop"] for aug in img_augs[i]]
        flip = [aug["flip"] for aug in img_augs[i]]
        for cam_it in range(num_cam):
            this_coor = cam_points[cam_it]
            H, W = resize_dims
            this_coor[:, :2] = this_coor[:, :2] * resize[cam_it]
            this_coor[:, 0] -= crop[cam_it][0]
            this_coor[:, 1] -= crop[cam_it][1]
            if flip[cam_it]:
                this_coor[:, 0] = resize_dims[1] - this_coor[:, 0]
            this_coor[:, 0] -= W / 2.0
            this_coor[:, 1] -= H / 2.0
            h = 0.0
            rot_matrix = this_coor.new_tensor([[math.cos(h), math.sin(h)], [-math.sin(h), math.cos(h)]])
            this_coor[:, :2] = torch.matmul(rot_matrix, this_coor[:, :2].T).T
            this_coor[:, 0] += W / 2.0
            this_coor[:, 1] += H / 2.0
            valid_mask = ((this_coor[:, 1] < resize_dims[0]) & (this_coor[:, 0] < resize_dims[1])
                          & (this_coor[:, 1] >= 0) & (this_coor[:, 0] >= 0))
            valid_coor = this_coor[valid_mask, :]
            valid_coor[:, [0, 1]] = valid_coor[:, [1, 0]]
            valid_coor[:, 0] = 2 * valid_coor[:, 0] / H - 1
            valid_coor[:, 1] = 2 * valid_coor[:, 1] / W - 1
            features = F.grid_sample(img_features[i][cam_it][None], valid_coor[None, :, None],
                                     mode="bilinear", padding_mode="zeros", align_corners=False
                                     ).squeeze(0).squeeze(-1)
            point_feature[valid_mask] += features.permute(1, 0).contiguous()
        cam_point_features.append(point_feature)
    return cam_point_features


# ------------------------------------------------------------------------------------------------
# Independent numpy restatement of ATen's bilinear grid_sample (Appendix B of SURVEY.md), used to
# cross-check the torch call above and to produce an fp64 ground truth for error budgets.
# ------------------------------------------------------------------------------------------------
def grid_sample_np(plane: np.ndarray, gx: np.ndarray, gy: np.ndarray, dtype=np.float64) -> np.ndarray:
    """plane [C,H,W]; gx (-> W), gy (-> H) normalised coords [Q]. Returns [C,Q]. zeros padding,
    align_corners=False, taps accumulated nw, ne, sw, se."""
    C, H, W = plane.shape
    plane = plane.astype(dtype)
    gx = gx.astype(dtype)
    gy = gy.astype(dtype)
    ix = ((gx + 1) * W - 1) / 2
    iy = ((gy + 1) * H - 1) / 2
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    w = [(x1 - ix) * (y1 - iy), (ix - x0) * (y1 - iy), (x1 - ix) * (iy - y0), (ix - x0) * (iy - y0)]
    taps = [(x0, y0), (x1, y0), (x0, y1), (x1, y1)]
    out = np.zeros((C, gx.shape[0]), dtype=dtype)
    for (tx, ty), tw in zip(taps, w):
        ok = (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
        txi = np.clip(tx, 0, W - 1).astype(np.int64)
        tyi = np.clip(ty, 0, H - 1).astype(np.int64)
        out += np.where(ok, plane[:, tyi, txi] * tw, 0)
    return out


def sample3_np(planes: Sequence[np.ndarray], points: np.ndarray, lo, vs, half, dtype=np.float64) -> np.ndarray:
    """planes: 3 x [C,H,W]; points [Q,3] -> [C,Q] (one sample)."""
    p = points.astype(dtype)
    g = [((p[:, a] - dtype(lo[a])) / dtype(vs[a])) / dtype(half[a]) - 1 for a in range(3)]
    return (grid_sample_np(planes[0], g[0], g[1], dtype) + grid_sample_np(planes[1], g[1], g[2], dtype)
            + grid_sample_np(planes[2], g[0], g[2], dtype))


# ------------------------------------------------------------------------------------------------
# SURVEY 8f #4 — gathers / scatters either side of the decode
# ------------------------------------------------------------------------------------------------
def _augmented_pixels(cam_points_cam: torch.Tensor, resize_dims, resize, crop, flip):
    """The per-camera augmentation chain shared by point_to_cam (point_triplane.py:203-227), cam_rec_feat
    (point_triplane.py:277-301) and JointEncoder.interact (joint_encoder.py:160-185): returns (this_coor, valid_mask)."""
    this_coor = cam_points_cam
    H, W = resize_dims
    this_coor[:, :2] = this_coor[:, :2] * resize
    this_coor[:, 0] -= crop[0]
    this_coor[:, 1] -= crop[1]
    if flip:
        this_coor[:, 0] = resize_dims[1] - this_coor[:, 0]
    this_coor[:, 0] -= W / 2.0
    this_coor[:, 1] -= H / 2.0
    h = 0.0
    rot_matrix = this_coor.new_tensor([[math.cos(h), math.sin(h)], [-math.sin(h), math.cos(h)]])
    this_coor[:, :2] = torch.matmul(rot_matrix, this_coor[:, :2].T).T
    this_coor[:, 0] += W / 2.0
    this_coor[:, 1] += H / 2.0
    valid_mask = ((this_coor[:, 1] < resize_dims[0]) & (this_coor[:, 0] < resize_dims[1])
                  & (this_coor[:, 1] >= 0) & (this_coor[:, 0] >= 0))
    return this_coor, valid_mask


def cam_proj_feat(range_proj_feat: torch.Tensor, range_cam_coors: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """TriplaneMAE.forward, triplane.py:379-390 (inline code): scatter range-image features into camera images.
    torch-CPU index_put keeps the LAST duplicate in list order; that order is the parity definition."""
    B, N = range_cam_coors.shape[:2]
    range_cam_coors = range_cam_coors.long()
    out = torch.zeros(B, N, range_proj_feat.shape[1], H, W, device=range_proj_feat.device)
    for b in range(B):
        for cam_it in range(N):
            cam_coors = range_cam_coors[b, cam_it]
            cam_coors_valid = cam_coors[..., 0] > 0
            proj_feat = range_proj_feat[b]
            cam_coors = cam_coors[cam_coors_valid, :]
            proj_feat = proj_feat[:, cam_coors_valid]
            out[b, cam_it][:, cam_coors[:, 0], cam_coors[:, 1]] = proj_feat
    return out


def cam_rec_feat(points: torch.Tensor, points_feat: torch.Tensor, img_metas) -> torch.Tensor:
    """PointTriplane.cam_rec_feat, point_triplane.py:243-309: points [N,3], points_feat [C,N], img_metas one dict."""
    resize_dims = img_metas["img_shape"][::-1]
    lidar2img = points.new_tensor(np.asarray(img_metas["lidar2image"]))
    img_augs = img_metas["imgs_aug"]
    num_cam = lidar2img.shape[0]
    out = torch.zeros((num_cam, points_feat.shape[0], resize_dims[0], resize_dims[1]), device=points_feat.device)
    hom_points = torch.cat((points, torch.ones_like(points[..., :1])), -1)
    cam_points = torch.einsum("cij, hj->chi", lidar2img, hom_points)
    cam_points = cam_points[..., 0:2] / torch.maximum(cam_points[..., 2:3], torch.ones_like(cam_points[..., 2:3]) * 1e-5)
    for cam_it in range(num_cam):
        aug = img_augs[cam_it]
        this_coor, valid_mask = _augmented_pixels(cam_points[cam_it], resize_dims, aug["resize"], aug["crop"], aug["flip"])
        valid_coor = this_coor[valid_mask, :].type(torch.long)
        valid_coor[:, [0, 1]] = valid_coor[:, [1, 0]]
        out[cam_it][:, valid_coor[:, 0], valid_coor[:, 1]] = points_feat[:, valid_mask]
    return out


def interact(img_features: torch.Tensor, range_image: torch.Tensor, img_metas, range_points: torch.Tensor, position_encoder):
    """JointEncoder.interact, joint_encoder.py:97-215. img_features is updated in place, as in the reference."""
    resize_dims = img_metas[0]["img_shape"][::-1]
    lidar2imgs = range_points.new_tensor(np.asarray([m["lidar2image"] for m in img_metas]))
    img_augs = [m["imgs_aug"] for m in img_metas]
    hom_points = torch.cat((range_points, torch.ones_like(range_points[..., :1])), -1)
    cam_points = torch.einsum("bcij, bhwj->bchwi", lidar2imgs, hom_points)
    cam_points = cam_points[..., 0:2] / torch.maximum(cam_points[..., 2:3], torch.ones_like(cam_points[..., 2:3]) * 1e-5)
    num_cam = lidar2imgs.shape[1]
    batch_range_mask = range_image > 0
    batch_size = range_points.shape[0]
    batch_no_point_mask = torch.ones_like(range_image).squeeze(1)
    batch_no_point_mask[(range_points == 0).sum(dim=3) == 3] = 0
    batch_no_point_mask = batch_no_point_mask.type(torch.bool)
    cam_range_features = torch.zeros(batch_size, img_features.shape[2], range_image.shape[2], range_image.shape[3],
                                     device=range_image.device)
    range_cam_coors = torch.zeros(batch_size, num_cam, range_image.shape[2], range_image.shape[3], 2,
                                  device=range_image.device) - 1
    for i in range(batch_size):
        this_coors = cam_points[i]
        range_mask = batch_range_mask[i].squeeze(0)
        no_point_mask = batch_no_point_mask[i]
        no_point_range_mask = range_mask[no_point_mask]
        for cam_it in range(num_cam):
            aug = img_augs[i][cam_it]
            this_coor, valid_mask = _augmented_pixels(this_coors[cam_it][no_point_mask, :], resize_dims, aug["resize"],
                                                      aug["crop"], aug["flip"])
            valid_coor = this_coor[valid_mask, :]
            valid_coor[:, [0, 1]] = valid_coor[:, [1, 0]]
            range_valid_mask = no_point_mask.clone()
            range_valid_mask[no_point_mask] = valid_mask
            range_cam_coors[i, cam_it][range_valid_mask] = valid_coor
            valid_mask = valid_mask & no_point_range_mask
            valid_points = range_points[i, no_point_mask, :][valid_mask, :]
            valid_coor = this_coor[valid_mask, :]
            valid_coor[:, [0, 1]] = valid_coor[:, [1, 0]]
            range_valid_mask = no_point_mask.clone()
            range_valid_mask[no_point_mask] = valid_mask
            valid_coor[:, 0] = valid_coor[:, 0] * img_features.shape[-2] / resize_dims[0]
            valid_coor[:, 1] = valid_coor[:, 1] * img_features.shape[-1] / resize_dims[1]
            valid_coor = valid_coor.type(torch.long)
            cam_range_features[i, :, range_valid_mask] += img_features[i, cam_it][:, valid_coor[:, 0], valid_coor[:, 1]]
            pos_embed = position_encoder(valid_points)
            img_features[i, cam_it][:, valid_coor[:, 0], valid_coor[:, 1]] += pos_embed.permute(1, 0)
    return torch.cat((range_image, cam_range_features), dim=1), img_features, range_cam_coors


def radius(x: torch.Tensor, y: torch.Tensor, r: float, batch_x: torch.Tensor, batch_y: torch.Tensor,
           max_num_neighbors: int = 32):
    """torch_geometric.nn.radius -> torch_cluster.radius (interpnet.py:5,44,65). NOT in /root/reference and not
    installable here: restated from torch_cluster's CUDA kernel (one thread per query walks the sources of the same
    sample in index order, keeps those with squared distance < r*r, stops after max_num_neighbors) — 'parity
    unpinned'. Returns (row = query index, col = source index), queries ascending."""
    rows, cols = [], []
    for q in range(y.shape[0]):
        src = torch.nonzero(batch_x == batch_y[q]).flatten()
        d = x[src] - y[q]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        hit = src[d2 < r * r][:max_num_neighbors]
        rows.append(torch.full_like(hit, q))
        cols.append(hit)
    return torch.cat(rows), torch.cat(cols)


def contrastive_features(sample_fn, triplane, points: Sequence[torch.Tensor], pc_range, num_cam: int = 6):
    """The sampling loop of the contrastive branch, triplane.py:438-455: per sample crop, per camera the points with a
    positive SAM label; `sample_fn(triplane[i][None], coords[None, None]).squeeze().permute(1, 0)`. Returns the list
    of ([N_s, C] features, labels) in loop order, skipping subsets with <= 1 point as the reference does."""
    out = []
    for i, pts in enumerate(points):
        crop_mask = ((pts[..., 0] > pc_range[0]) & (pts[..., 0] < pc_range[3]) & (pts[..., 1] > pc_range[1]) &
                     (pts[..., 1] < pc_range[4]) & (pts[..., 2] > pc_range[2]) & (pts[..., 2] < pc_range[5]))
        pts = pts[crop_mask]
        for cam in range(num_cam):
            coords = pts[:, 0:3]
            labels = pts[:, 5 + cam]
            valid_mask = labels > 0
            coords = coords[valid_mask]
            labels = labels[valid_mask].type(torch.int)
            if labels.shape[0] > 1:
                features = sample_fn(triplane[i][None, ...], coords[None, None, ...]).squeeze()
                out.append((features.permute(1, 0), labels))
    return out
