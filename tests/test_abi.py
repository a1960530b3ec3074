"""CPU: libtriplane.so loads, exports every TP_API symbol of include/triplane.h, and rejects bad
arguments before touching a device (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from efficient_multimodal_perception_b200 import _lib as L


def header_symbols():
    text = open(os.path.join(ROOT, "include", "triplane.h")).read()
    return re.findall(r"TP_API\s+[\w\s\*]+?\b(tp_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in triplane.h but not exported"
    assert set(syms) == set(L.SIGNATURES), "ctypes SIGNATURES out of sync with include/triplane.h"
    assert lib.tp_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(L.tp_geom) == 9 * 4 + 6 * 4
    assert C.sizeof(L.tp_plane) == 24
    assert C.sizeof(L.tp_sample_geom) == 36


def test_argument_errors_are_reported_without_a_device():
    lib = L.lib()
    sg = L.make_sample_geom([0, 0, 0], [1, 1, 1], [1, 1, 1])
    planes = (L.tp_plane * 3)()
    rc = lib.tp_sample3_nhwc_f32(C.byref(planes), 32, None, 10, 1, C.byref(sg), 0, None, None)
    assert rc == -1 and b"null" in lib.tp_last_error()
    rc = lib.tp_sample3_nhwc_f32(C.byref(planes), 30, 1, 10, 1, C.byref(sg), 0, 1, None)
    assert rc == -2 and b"multiple of 4" in lib.tp_last_error()
    geom = L.make_geom([-1, -1, -1, 1, 1, 1], (0.0, 1, 1), (4, 4, 4), (1, 1, 1))
    rc = lib.tp_voxel_index_f32(1, 10, 3, C.byref(geom), 0, 1, 1, None)
    assert rc == -2 and b"geometry" in lib.tp_last_error()
    geom = L.make_geom([-1, -1, -1, 1, 1, 1], (1, 1, 1), (4, 4, 4), (1, 1, 1))
    rc = lib.tp_voxel_index_f32(1, 10, 3, C.byref(geom), 7, 1, 1, None)
    assert rc == -3
    rc = lib.tp_encode_f32(16, 8, 8, None, None, 0, 5, 1, 1, C.byref(geom), 0, 0, 0, None, None, None, None, 16,
                           1 << 30, None)
    assert rc == -1  # neither idx nor points
    rc = lib.tp_encode_f32(16, 8, 8, 16, None, 0, 5, 1, 1, C.byref(geom), 0, 0, 0, None, None, None, None, 16, 8,
                           None)
    assert rc == -4 and b"workspace" in lib.tp_last_error()


def test_new_entry_points_validate_arguments_without_a_device():
    lib = L.lib()
    sg = L.make_sample_geom([0, 0, 0], [1, 1, 1], [1, 1, 1])
    planes = (L.tp_plane * 3)()
    dims = (C.c_int32 * 3)(4, 4, 16)
    rc = lib.tp_sample3_grid_nhwc_f32(C.byref(planes), 32, 16, C.byref(dims), 1, C.byref(sg), 0, 16, None)
    assert rc == -1 and b"plane 0 is null" in lib.tp_last_error()
    rc = lib.tp_sample3_grid_nhwc_f32(C.byref(planes), 30, 16, C.byref(dims), 1, C.byref(sg), 0, 16, None)
    assert rc == -2
    bad = (C.c_int32 * 3)(4, -1, 16)
    rc = lib.tp_sample3_grid_nhwc_f32(C.byref(planes), 32, 16, C.byref(bad), 1, C.byref(sg), 0, 16, None)
    assert rc == -2 and b"dims" in lib.tp_last_error()
    # lift: camera slots, channel multiple, null pointers
    rc = lib.tp_lift_cam_f32(16, 3, 10, 16, 1, 16, 9, 16, 32, 768, 16, 512.0, 256.0, 0, 16, None)
    assert rc == -2 and b"ncam" in lib.tp_last_error()
    rc = lib.tp_lift_cam_f32(16, 3, 10, 16, 1, 16, 6, 16, 32, 770, 16, 512.0, 256.0, 0, 16, None)
    assert rc == -2 and b"Cf" in lib.tp_last_error()
    rc = lib.tp_lift_cam_f32(None, 3, 10, 16, 1, 16, 6, 16, 32, 768, 16, 512.0, 256.0, 0, 16, None)
    assert rc == -1
    assert lib.tp_lift_cam_f32(None, 3, 0, None, 1, None, 6, 16, 32, 768, None, 512.0, 256.0, 0, None, None) == 0  # no points
    rc = lib.tp_lift_cam_f32(16, 3, 10, 16, 1, 16, 6, 16, 32, 768, 16, 512.0, 256.0, 5, 16, None)
    assert rc == -3
    # backward entry points
    geom = L.make_geom([-1, -1, -1, 1, 1, 1], (1, 1, 1), (4, 4, 4), (1, 1, 1))
    rc = lib.tp_encode_backward_f32(16, 8, 8, 16, 5, 16, 1, C.byref(geom), 2, 0, None, None, None, None, None, None,
                                    None, 16, None)
    assert rc == -3 and b"MAX or MEAN" in lib.tp_last_error()
    rc = lib.tp_encode_backward_f32(16, 8, 8, 16, 5, 16, 1, C.byref(geom), 1, 0, None, None, None, None, None, None,
                                    None, 16, None)
    assert rc == -1 and b"cell_count" in lib.tp_last_error()
    rc = lib.tp_sample3_backward_nhwc_f32(C.byref(planes), 32, 16, 10, 1, C.byref(sg), 0, 16, None)
    assert rc == -1
    rc = lib.tp_sample3_backward_nhwc_f32(C.byref(planes), 32, 16, 10, 1, C.byref(sg), 9, 16, None)
    assert rc == -3
    # decode + head: class count, lattice depth, null planes, arithmetic mode
    rc = lib.tp_sample3_grid_head_tf32(C.byref(planes), 16, C.byref(dims), 1, C.byref(sg), 0, 16, 16, 16, 17, 16, None)
    assert rc == -2 and b"num_classes" in lib.tp_last_error()
    odd = (C.c_int32 * 3)(4, 4, 6)
    rc = lib.tp_sample3_grid_head_tf32(C.byref(planes), 16, C.byref(odd), 1, C.byref(sg), 0, 16, 16, 16, 5, 16, None)
    assert rc == -2 and b"multiple of 4" in lib.tp_last_error()
    rc = lib.tp_sample3_grid_head_tf32(C.byref(planes), 16, C.byref(dims), 1, C.byref(sg), 0, 16, 16, 16, 5, 16, None)
    assert rc == -1 and b"plane 0 is null" in lib.tp_last_error()
    rc = lib.tp_sample3_grid_head_tf32(C.byref(planes), 16, C.byref(dims), 1, C.byref(sg), 7, 16, 16, 16, 5, 16, None)
    assert rc == -3
    empty = (C.c_int32 * 3)(0, 4, 16)
    assert lib.tp_sample3_grid_head_tf32(C.byref(planes), None, C.byref(empty), 1, C.byref(sg), 0, None, None, None, 5, None, None) == 0


def test_cell_and_workspace_sizes():
    lib = L.lib()
    geom = L.make_geom([-25, -25, -5, 25, 25, 3], (0.4, 0.4, 0.1), (128, 128, 80), (5, 5, 4))
    cells = (C.c_int64 * 3)()
    total = lib.tp_encode_cells(C.byref(geom), 1, C.byref(cells))
    # 327 680 + 256 000 + 256 000 = 839 680 pooled cells per sample at geometry A
    assert list(cells) == [327680, 256000, 256000] and total == 839680
    geom_b = L.make_geom([-50, -50, -5, 50, 50, 3], (0.5, 0.5, 0.5), (200, 200, 16), (8, 8, 1))
    assert lib.tp_encode_cells(C.byref(geom_b), 2, C.byref(cells)) == 2 * 800000
    # tile counters + CSR offsets for 16-cell tiles (C = 512 worst case) + rank and entry lists
    assert lib.tp_encode_workspace_bytes(C.byref(geom), 1, 1000) >= 2 * 4 * (839680 // 16) + 1000 * 36
    assert lib.tp_voxelize_workspace_bytes(0) > 0


def test_ops_refuse_cpu_tensors():
    import torch
    from efficient_multimodal_perception_b200 import TriplaneError, ops
    with pytest.raises(TriplaneError, match="CUDA tensor"):
        ops.sample3(torch.zeros(1, 3, 4, 8, 8), torch.zeros(1, 5, 3), [0] * 3, [1] * 3, [4] * 3)
    with pytest.raises(TriplaneError, match="CUDA tensor"):
        ops.voxel_index(torch.zeros(5, 3), [0] * 6, [1] * 3)
    # the round-2 decode entry points: no silent CPU path either
    tri, lo, vs, half = torch.zeros(1, 3, 32, 8, 8), [0] * 3, [1] * 3, [4] * 3
    with pytest.raises(TriplaneError, match="CUDA tensor"):
        ops.sample3_lattice(tri, (4, 4, 16), [0] * 3, [1] * 3, lo, vs, half)
    with pytest.raises(TriplaneError, match="CUDA tensor"):
        ops.sample3_segments(tri, torch.zeros(5, 3), torch.tensor([0, 5]), None, lo, vs, half)
    with pytest.raises(TriplaneError, match="CUDA tensor"):
        ops.sample3_head(tri, torch.zeros(1, 256, 3), lo, vs, half, torch.zeros(64, 32), torch.zeros(32, 64),
                         torch.zeros(5, 32), grid_dims=(4, 4, 16))


def test_missing_library_fails_loudly(monkeypatch):
    from efficient_multimodal_perception_b200 import TriplaneError
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libtriplane.so")
    with pytest.raises(TriplaneError, match="no CPU/PyTorch fallback"):
        L.lib()


def test_reference_layout_entry_points_validate_workspace_without_a_device():
    """The *_nchw_* forms (conversion + decode in one call) reject a missing / short workspace and bad shapes before
    any launch: every decode entry point has one."""
    lib = L.lib()
    sg = L.make_sample_geom([0, 0, 0], [1, 1, 1], [1, 1, 1])
    planes = (L.tp_plane * 3)()
    for k in range(3):
        planes[k].data, planes[k].batch_stride, planes[k].H, planes[k].W = 256, 32 * 8 * 8, 8, 8
    dims = (C.c_int32 * 3)(4, 4, 16)
    org, stp = (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(1, 1, 1)
    need = 3 * 32 * 8 * 8
    # no workspace
    assert lib.tp_sample3_nchw_f32(C.byref(planes), 32, 256, 10, 1, C.byref(sg), 0, 256, None, need, None) == -1
    assert lib.tp_sample3_grid_nchw_f32(C.byref(planes), 32, 256, C.byref(dims), 1, C.byref(sg), 0, 256, None, need, None) == -1
    # short workspace
    for call in (
        lambda ws: lib.tp_sample3_nchw_f32(C.byref(planes), 32, 256, 10, 1, C.byref(sg), 0, 256, 256, ws, None),
        lambda ws: lib.tp_sample3_grid_nchw_f32(C.byref(planes), 32, 256, C.byref(dims), 1, C.byref(sg), 0, 256, 256, ws, None),
        lambda ws: lib.tp_sample3_lattice_nchw_f32(C.byref(planes), 32, C.byref(dims), C.byref(org), C.byref(stp), 1, C.byref(sg), 0,
                                                   256, 256, ws, None),
        lambda ws: lib.tp_sample3_seg_nchw_f32(C.byref(planes), 32, 256, 10, 256, None, 1, 1, C.byref(sg), 0, 256, 256, ws, None),
        lambda ws: lib.tp_sample3_grid_head_nchw_tf32(C.byref(planes), 256, C.byref(dims), 1, C.byref(sg), 0, 256, 256, 256, 5, 256,
                                                      256, ws, None),
    ):
        assert call(need - 1) == -4 and b"workspace" in lib.tp_last_error()
    # empty problems return before touching anything
    assert lib.tp_sample3_seg_nchw_f32(C.byref(planes), 32, None, 0, None, None, 1, 1, C.byref(sg), 0, None, None, 0, None) == 0
    zero = (C.c_int32 * 3)(0, 4, 16)
    assert lib.tp_sample3_grid_head_nchw_tf32(C.byref(planes), None, C.byref(zero), 1, C.byref(sg), 0, None, None, None, 5, None,
                                              None, 0, None) == 0
    # lattice form: null origin
    assert lib.tp_sample3_lattice_nchw_f32(C.byref(planes), 32, C.byref(dims), None, C.byref(stp), 1, C.byref(sg), 0, 256, 256,
                                           need, None) == -1
