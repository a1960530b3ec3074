"""torch.autograd.Function wrappers: forward and backward both run in libtriplane.so. They give the
drop-in modules the gradients the reference gets from autograd through torch_scatter.scatter_max +
SparseMaxPool3d (point_triplane_projector.py:104,113-115) and F.grid_sample
(sample_points_triplane, point_to_cam). Coordinates (points, queries) are data in every reference
config and receive no gradient; asking for one raises."""
from __future__ import annotations

from typing import Sequence, Union

import torch

from . import ops
from ._lib import TriplaneError


class _EncodeMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, grid_ind, offsets, grid_size, split, clamp_zero):
        xy, yz, xz = ops.encode(feats, offsets, [0] * 6, (1, 1, 1), grid_size, split, grid_ind=grid_ind, reduce="max",
                                clamp_zero=clamp_zero)
        ctx.save_for_backward(feats, grid_ind, offsets, xy, yz, xz)
        ctx.geom = (list(grid_size), list(split), bool(clamp_zero))
        return xy, yz, xz

    @staticmethod
    def backward(ctx, g_xy, g_yz, g_xz):
        feats, grid_ind, offsets, xy, yz, xz = ctx.saved_tensors
        grid_size, split, clamp_zero = ctx.geom
        gf = ops.encode_backward((g_xy, g_yz, g_xz), feats, grid_ind, offsets, grid_size, split, outs=(xy, yz, xz),
                                 reduce="max", clamp_zero=clamp_zero)
        return gf, None, None, None, None, None


class _EncodeMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, grid_ind, offsets, grid_size, split):
        xy, yz, xz, cnt = ops.encode(feats, offsets, [0] * 6, (1, 1, 1), grid_size, split, grid_ind=grid_ind,
                                     reduce="mean", want_counts=True)
        ctx.save_for_backward(grid_ind, offsets, cnt)
        ctx.geom = (list(grid_size), list(split), feats.shape[1])
        return xy, yz, xz

    @staticmethod
    def backward(ctx, g_xy, g_yz, g_xz):
        grid_ind, offsets, cnt = ctx.saved_tensors
        grid_size, split, channels = ctx.geom
        gf = ops.encode_backward((g_xy, g_yz, g_xz), None, grid_ind, offsets, grid_size, split, counts=cnt,
                                 reduce="mean", channels=channels)
        return gf, None, None, None, None


def encode_max_autograd(feats, grid_ind, offsets, grid_size, split, clamp_zero=False):
    """Differentiable (w.r.t. feats) fused scatter-max encode: (xy, yz, xz) as ops.encode."""
    return _EncodeMax.apply(feats, grid_ind, offsets, grid_size, split, clamp_zero)


def encode_mean_autograd(feats, grid_ind, offsets, grid_size, split):
    return _EncodeMean.apply(feats, grid_ind, offsets, grid_size, split)


class _Sample3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, queries, lo, vs, half, arith, dims, p0, p1, p2):
        out = ops.sample3([p0, p1, p2], queries, lo, vs, half, arith=arith, grid_dims=dims)
        ctx.save_for_backward(queries)
        ctx.cfg = (lo, vs, half, arith, [tuple(p.shape[-2:]) for p in (p0, p1, p2)], dims)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (queries,) = ctx.saved_tensors
        lo, vs, half, arith, shapes, dims = ctx.cfg
        g0, g1, g2 = ops.sample3_backward(grad_out, queries, shapes, lo, vs, half, arith=arith, grid_dims=dims)
        need = ctx.needs_input_grad
        return (None, None, None, None, None, None, g0 if need[6] else None, g1 if need[7] else None,
                g2 if need[8] else None)


def sample3_autograd(triplane: Union[torch.Tensor, Sequence[torch.Tensor]], queries: torch.Tensor, lo, vs, half,
                     arith: str = "cuda", dims=None) -> torch.Tensor:
    """Differentiable (w.r.t. the planes) fused 3-plane sample + sum. queries [B,Q,3] -> [B,C,Q]."""
    if queries.requires_grad:
        raise TriplaneError("sample_points_triplane: gradients w.r.t. the query points are not implemented "
                            "(points are data in every reference config)")
    if isinstance(triplane, torch.Tensor):
        planes = [triplane[:, 0], triplane[:, 1], triplane[:, 2]]
    else:
        planes = list(triplane)
    return _Sample3.apply(queries, list(lo), list(vs), list(half), arith, dims, *planes)


class _LiftCam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_features, points, offsets, cams, resize_dims, arith):
        out = ops.lift_cam(points, offsets, img_features, cams, resize_dims, arith=arith)
        ctx.save_for_backward(points, offsets, cams)
        ctx.cfg = (tuple(img_features.shape), tuple(resize_dims), arith)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        points, offsets, cams = ctx.saved_tensors
        shape, resize_dims, arith = ctx.cfg
        g = ops.lift_cam_backward(grad_out, points, offsets, shape, cams, resize_dims, arith=arith)
        return g, None, None, None, None, None


def lift_cam_autograd(points, offsets, img_features, cams, resize_dims, arith: str = "cuda") -> torch.Tensor:
    """Differentiable (w.r.t. img_features) fused point_to_cam: [N, Cf]."""
    if points.requires_grad:
        raise TriplaneError("point_to_cam: gradients w.r.t. the points are not implemented")
    return _LiftCam.apply(img_features, points, offsets, cams, resize_dims, arith)


class _Sample3Segments(torch.autograd.Function):
    @staticmethod
    def forward(ctx, queries, seg_offsets, seg_batch, lo, vs, half, arith, p0, p1, p2):
        out = ops.sample3_segments([p0, p1, p2], queries, seg_offsets, seg_batch, lo, vs, half, arith=arith)
        ctx.save_for_backward(queries, seg_offsets, seg_batch)
        ctx.cfg = (lo, vs, half, arith, p0.shape[0], [tuple(p.shape[-2:]) for p in (p0, p1, p2)])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        queries, seg_offsets, seg_batch = ctx.saved_tensors
        lo, vs, half, arith, batch, shapes = ctx.cfg
        g0, g1, g2 = ops.sample3_segments_backward(grad_out, queries, seg_offsets, seg_batch, batch, shapes, lo, vs, half,
                                                   arith=arith)
        need = ctx.needs_input_grad
        return (None, None, None, None, None, None, None, g0 if need[7] else None, g1 if need[8] else None,
                g2 if need[9] else None)


def sample3_segments_autograd(triplane, queries, seg_offsets, seg_batch, lo, vs, half, arith: str = "cuda") -> torch.Tensor:
    """Differentiable (w.r.t. the planes) ragged-subset decode: [T, C] point-major."""
    if queries.requires_grad:
        raise TriplaneError("sample_points_triplane: gradients w.r.t. the query points are not implemented")
    planes = [triplane[:, 0], triplane[:, 1], triplane[:, 2]] if isinstance(triplane, torch.Tensor) else list(triplane)
    return _Sample3Segments.apply(queries, seg_offsets, seg_batch, list(lo), list(vs), list(half), arith, *planes)


class _WinnerGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, winner, imgs_per_feat, layout, row0):
        out = ops.winner_gather(winner, feat, imgs_per_feat, layout=layout, row0=row0)
        ctx.save_for_backward(winner, row0)
        ctx.cfg = (tuple(feat.shape), imgs_per_feat, layout)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        winner, row0 = ctx.saved_tensors
        shape, imgs_per_feat, layout = ctx.cfg
        return ops.winner_gather_backward(winner, grad_out, shape, imgs_per_feat, layout=layout, row0=row0), None, None, None, None


def winner_gather_autograd(feat, winner, imgs_per_feat, layout="bcn", row0=None):
    return _WinnerGather.apply(feat, winner, imgs_per_feat, layout, row0)


class _RangeGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_features, fidx):
        ctx.save_for_backward(fidx)
        ctx.shape = tuple(img_features.shape)
        return ops.range_gather(fidx, img_features)

    @staticmethod
    def backward(ctx, grad_out):
        (fidx,) = ctx.saved_tensors
        return ops.range_gather_backward(fidx, grad_out, ctx.shape), None


def range_gather_autograd(img_features, fidx):
    return _RangeGather.apply(img_features, fidx)


class _PosEmbedScatter(torch.autograd.Function):
    """img_features + scatter(pos_embed): the forward updates a COPY when gradients are tracked (autograd forbids the
    reference's in-place write into a tensor other ops have saved); without gradients ops.posembed_scatter_ is used."""

    @staticmethod
    def forward(ctx, img_features, pos_embed, winner):
        ctx.save_for_backward(winner)
        out = img_features.contiguous().clone()
        return ops.posembed_scatter_(out, winner, pos_embed)

    @staticmethod
    def backward(ctx, grad_img):
        (winner,) = ctx.saved_tensors
        need = ctx.needs_input_grad
        gpe = ops.posembed_scatter_backward(grad_img, winner) if need[1] else None
        return (grad_img if need[0] else None), gpe, None


def posembed_scatter_autograd(img_features, pos_embed, winner):
    return _PosEmbedScatter.apply(img_features, pos_embed, winner)
