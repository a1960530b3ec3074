"""ctypes binding of libtriplane.so (include/triplane.h). There is no fallback: if the library is
missing or a call fails, a TriplaneError is raised."""
from __future__ import annotations

import ctypes as C
import os

LIB_NAME = "libtriplane.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

TP_ARITH_TORCH_CUDA = 0
TP_ARITH_TORCH_CPU = 1
TP_ARITH_TORCH_CUDA_NOFMA = 2  # diagnostic: CUDA formula without the fma contraction (decode only)
TP_REDUCE_MAX, TP_REDUCE_MEAN, TP_REDUCE_SUM, TP_REDUCE_MAX_PARTIAL = 0, 1, 2, 3


class TriplaneError(RuntimeError):
    pass


class tp_geom(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("hi", C.c_float * 3), ("vs", C.c_float * 3),
                ("grid", C.c_int32 * 3), ("pool", C.c_int32 * 3)]


class tp_plane(C.Structure):
    _fields_ = [("data", C.c_void_p), ("batch_stride", C.c_int64), ("H", C.c_int32), ("W", C.c_int32)]


class tp_sample_geom(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("vs", C.c_float * 3), ("half", C.c_float * 3)]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

# name -> (restype, argtypes); must list every TP_API symbol of include/triplane.h
SIGNATURES = {
    "tp_last_error": (C.c_char_p, []),
    "tp_version": (C.c_int, []),
    "tp_sm_count": (C.c_int, []),
    "tp_voxelize_workspace_bytes": (_i64, [_i64]),
    "tp_voxelize_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, C.POINTER(tp_geom), _i32, _vp, _vp, _vp,
                                  _vp, _i64, _vp]),
    "tp_voxel_index_f32": (C.c_int, [_vp, _i64, _i32, C.POINTER(tp_geom), _i32, _vp, _vp, _vp]),
    "tp_encode_cells": (_i64, [C.POINTER(tp_geom), _i32, C.POINTER(_i64 * 3)]),
    "tp_encode_workspace_bytes": (_i64, [C.POINTER(tp_geom), _i32, _i64]),
    "tp_encode_workspace_init": (C.c_int, [_vp, _i64, _vp]),
    "tp_encode_f32": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _i64, _vp, _i32, C.POINTER(tp_geom), _i32,
                                _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "tp_projector_sparse_workspace_bytes": (_i64, [C.POINTER(tp_geom), _i32, _i32]),
    "tp_projector_sparse_f32": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _i64, _vp, _i32, C.POINTER(tp_geom), _i32, _i32,
                                          C.POINTER(_vp * 3), C.POINTER(_vp * 3), _i32, C.POINTER(_vp * 3), _vp, _i64, _vp]),
    "tp_encode_finalize_mean_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "tp_encode_finalize_max_f32": (C.c_int, [_vp, _i64, _i32, _vp]),
    "tp_voxel_counts_i32": (C.c_int, [_vp, _i64, _vp, _i32, C.POINTER(tp_geom), _vp, _vp]),
    "tp_planes_nchw_to_nhwc_f32": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tp_planes3_nchw_to_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), C.POINTER(_vp * 3), _i32, _i32, _vp]),
    "tp_sample3_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, _i64, _i32,
                                      C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_sample3_nchw_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, _i64, _i32,
                                      C.POINTER(tp_sample_geom), _i32, _vp, _vp, _i64, _vp]),
    "tp_sample3_grid_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, C.POINTER(_i32 * 3), _i32,
                                           C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_sample3_grid_nchw_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, C.POINTER(_i32 * 3), _i32,
                                           C.POINTER(tp_sample_geom), _i32, _vp, _vp, _i64, _vp]),
    "tp_sample3_lattice_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, C.POINTER(_i32 * 3), C.POINTER(C.c_float * 3),
                                              C.POINTER(C.c_float * 3), _i32, C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_sample3_lattice_nchw_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, C.POINTER(_i32 * 3), C.POINTER(C.c_float * 3),
                                              C.POINTER(C.c_float * 3), _i32, C.POINTER(tp_sample_geom), _i32, _vp, _vp, _i64,
                                              _vp]),
    "tp_sample3_seg_nchw_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, _i64, _vp, _vp, _i32, _i32,
                                          C.POINTER(tp_sample_geom), _i32, _vp, _vp, _i64, _vp]),
    "tp_sample3_grid_head_nchw_tf32": (C.c_int, [C.POINTER(tp_plane * 3), _vp, C.POINTER(_i32 * 3), _i32,
                                                 C.POINTER(tp_sample_geom), _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "tp_sample3_seg_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, _i64, _vp, _vp, _i32, _i32,
                                          C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_sample3_seg_backward_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, _i64, _vp, _vp, _i32, _i32,
                                                   C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_pixel_winner_coors_i32": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "tp_pixel_winner_points_i32": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _vp, _i32, C.c_float, C.c_float, _vp, _vp]),
    "tp_winner_gather_f32": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "tp_winner_gather_backward_f32": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "tp_range_project_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i32, C.c_float, C.c_float, _i32, _i32, _i32, _vp, _vp,
                                       _vp, _vp]),
    "tp_range_gather_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "tp_range_gather_backward_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "tp_posembed_scatter_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "tp_posembed_scatter_backward_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "tp_radius_i32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, C.c_float, _i32, _vp, _vp, _vp]),
    "tp_lift_cam_f32": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, C.c_float, C.c_float,
                                  _i32, _vp, _vp]),
    "tp_sample3_grid_backward_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, C.POINTER(_i32 * 3), _i32,
                                                    C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_sample3_backward_nhwc_f32": (C.c_int, [C.POINTER(tp_plane * 3), _i32, _vp, _i64, _i32,
                                               C.POINTER(tp_sample_geom), _i32, _vp, _vp]),
    "tp_encode_backward_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _i32, C.POINTER(tp_geom), _i32, _i32,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tp_lift_cam_backward_f32": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _vp, C.c_float,
                                           C.c_float, _i32, _vp, _vp, _vp]),
    "tp_mlp_head_tf32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "tp_sample3_grid_head_tf32": (C.c_int, [C.POINTER(tp_plane * 3), _vp, C.POINTER(_i32 * 3), _i32,
                                            C.POINTER(tp_sample_geom), _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "tp_route_points_f32": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _i64, C.POINTER(tp_geom), _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _i64, _vp]),
    "tp_comm_unique_id": (C.c_int, [_vp]),
    "tp_comm_init": (C.c_int, [C.POINTER(_vp), _i32, _i32, _vp]),
    "tp_comm_destroy": (C.c_int, [_vp]),
    "tp_allreduce_planes": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "tp_sample3_host_f32": (C.c_int, [C.POINTER(_vp * 3), C.POINTER(_i32 * 6), C.POINTER(_i64 * 3), _i32, _vp,
                                      _i64, _i32, C.POINTER(tp_sample_geom), _i32, _vp]),
    "tp_sample3_grid_host_f32": (C.c_int, [C.POINTER(_vp * 3), C.POINTER(_i32 * 6), C.POINTER(_i64 * 3), _i32, _vp,
                                           C.POINTER(_i32 * 3), _i32, C.POINTER(tp_sample_geom), _i32, _vp]),
    "tp_encode_host_f32": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _vp, _i32, C.POINTER(tp_geom), _i32, _i32,
                                     _i32, _vp, _vp, _vp]),
    "tp_host_arena_release": (None, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load libtriplane.so once. Raises TriplaneError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TriplaneError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C efficient_multimodal_perception_b200/csrc`. There is no CPU/PyTorch fallback.")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:
            raise TriplaneError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:
                raise TriplaneError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().tp_last_error()
        raise TriplaneError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def make_geom(pc_range, voxel_size, grid_size, pool) -> tp_geom:
    g = tp_geom()
    for a in range(3):
        g.lo[a] = float(pc_range[a])
        g.hi[a] = float(pc_range[3 + a])
        g.vs[a] = float(voxel_size[a])
        g.grid[a] = int(grid_size[a])
        g.pool[a] = int(pool[a])
    return g


def make_sample_geom(lo, vs, half) -> tp_sample_geom:
    s = tp_sample_geom()
    for a in range(3):
        s.lo[a] = float(lo[a])
        s.vs[a] = float(vs[a])
        s.half[a] = float(half[a])
    return s
