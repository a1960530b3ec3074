"""B200-native triplane hot path of charyyev/efficient_multimodal_perception.

Encode (points + camera-lifted features -> xy/yz/xz planes by fused voxel-index + scatter-max/mean)
and decode (per-query bilinear sample of the three planes + sum) as hand-written sm_100a CUDA
kernels in ``libtriplane.so`` (C ABI: include/triplane.h), behind the reference's module and
method names. There is no CPU or PyTorch fallback: a missing library raises TriplaneError.
"""
from ._lib import LIB_PATH, TriplaneError, lib
from . import ops, synth
from .modules import (Mlp, PointTriplaneProjector, TriplaneHotPathMixin, cam_proj_feat, cam_rec_feat, interact,
                      point_to_cam, radius_search, register_with_mmdet, roi, sam_subsets, sample_and_decode,
                      sample_points_triplane, sample_points_triplane_segments, sample_roi_triplane, voxelize_points)

__all__ = ["LIB_PATH", "TriplaneError", "lib", "ops", "synth", "Mlp", "PointTriplaneProjector",
           "TriplaneHotPathMixin", "cam_proj_feat", "cam_rec_feat", "interact", "point_to_cam", "radius_search",
           "register_with_mmdet", "roi", "sam_subsets", "sample_and_decode", "sample_points_triplane",
           "sample_points_triplane_segments", "sample_roi_triplane", "voxelize_points"]

register_with_mmdet()
