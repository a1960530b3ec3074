"""Tensor-level wrappers over the C ABI (include/triplane.h). PyTorch is used for device memory and
streams only; every arithmetic step runs in libtriplane.so. All functions require CUDA tensors and
raise (never fall back) otherwise."""
from __future__ import annotations

import ctypes as C
import functools
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _lib as L
from ._lib import (TP_ARITH_TORCH_CPU, TP_ARITH_TORCH_CUDA, TP_REDUCE_MAX, TP_REDUCE_MEAN, TP_REDUCE_SUM,
                   TriplaneError)

_REDUCE = {"max": TP_REDUCE_MAX, "mean": TP_REDUCE_MEAN, "sum": TP_REDUCE_SUM, "max_partial": L.TP_REDUCE_MAX_PARTIAL}
_ARITH = {"cuda": TP_ARITH_TORCH_CUDA, "cpu": TP_ARITH_TORCH_CPU, "cuda_nofma": L.TP_ARITH_TORCH_CUDA_NOFMA}

#: count of libtriplane kernel launches issued through this module (bench.py's gpu_launches)
launch_count = 0


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _tensors(objs):
    for o in objs:
        if isinstance(o, torch.Tensor):
            yield o
        elif isinstance(o, (list, tuple)):
            yield from _tensors(o)


def _on_device(fn):
    """Every wrapper hands raw pointers and the tensors' current stream to libtriplane, which launches on the
    CURRENT device: make the tensors' device current for the call (model.to('cuda:1') without set_device,
    several GPUs in one process) and refuse arguments spread over several devices."""

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        devs = {t.device for t in _tensors(list(args) + list(kwargs.values())) if t.is_cuda}
        if len(devs) > 1:
            raise TriplaneError(f"{fn.__name__}: tensor arguments live on several devices: {sorted(map(str, devs))}")
        if not devs:
            return fn(*args, **kwargs)  # the wrapper raises its own 'expected a CUDA tensor' error
        with torch.cuda.device(next(iter(devs))):
            return fn(*args, **kwargs)

    return wrapped


def _need_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TriplaneError(f"{name}: expected a CUDA tensor (no CPU path exists), got "
                            f"{type(t).__name__}{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}")
    if t.dtype != dtype:
        raise TriplaneError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pool_kernels(grid_size, split) -> Tuple[int, int, int]:
    """int(grid/split), as point_triplane_projector.py:53-58 computes the pooling kernels."""
    return tuple(int(grid_size[a] / split[a]) for a in range(3))


def pooled_sizes(grid_size, pool) -> Tuple[int, int, int]:
    return tuple((int(grid_size[a]) - int(pool[a])) // int(pool[a]) + 1 for a in range(3))


# ------------------------------------------------------------------------------------------------
# a1
# ------------------------------------------------------------------------------------------------
@_on_device
def voxelize(points: torch.Tensor, offsets: torch.Tensor, pc_range, voxel_size, grid_size=(1, 1, 1),
             ncols: Optional[int] = None, arith: str = "cuda"):
    """points [N, D] (samples concatenated), offsets [B+1] int64 ->
    (cropped [N', ncols], grid_ind [N', 3] int32, out_offsets [B+1] int64 on device).
    One host sync (reading N') — the reference syncs once per sample (point_triplane.py:152)."""
    global launch_count
    _need_cuda(points, "points")
    _need_cuda(offsets, "offsets", torch.int64)
    if points.dim() != 2 or points.shape[1] < 3:
        raise TriplaneError(f"points must be [N, >=3], got {tuple(points.shape)}")
    points = points.contiguous()
    n, stride = points.shape
    ncols = stride if ncols is None else ncols
    batch = offsets.numel() - 1
    geom = L.make_geom(pc_range, voxel_size, grid_size, (1, 1, 1))
    lib = L.lib()
    ws_bytes = lib.tp_voxelize_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points.device)
    out_pts = torch.empty((n, ncols), dtype=torch.float32, device=points.device)
    out_idx = torch.empty((n, 3), dtype=torch.int32, device=points.device)
    out_off = torch.empty(batch + 1, dtype=torch.int64, device=points.device)
    L.check(lib.tp_voxelize_f32(points.data_ptr(), n, stride, ncols, offsets.data_ptr(), batch, C.byref(geom),
                                _ARITH[arith], out_pts.data_ptr(), out_idx.data_ptr(), out_off.data_ptr(),
                                ws.data_ptr(), ws_bytes, _stream(points)), "tp_voxelize_f32")
    launch_count += 4 if n else 0
    n_kept = int(out_off[-1].item())
    return out_pts[:n_kept], out_idx[:n_kept], out_off


@_on_device
def voxel_index(points: torch.Tensor, pc_range, voxel_size, arith: str = "cuda"):
    """Uncompacted crop mask [N] (uint8) and int32 voxel index [N,3] for every raw point."""
    global launch_count
    _need_cuda(points, "points")
    points = points.contiguous()
    n, stride = points.shape
    geom = L.make_geom(pc_range, voxel_size, (1, 1, 1), (1, 1, 1))
    keep = torch.empty(n, dtype=torch.uint8, device=points.device)
    idx = torch.empty((n, 3), dtype=torch.int32, device=points.device)
    L.check(L.lib().tp_voxel_index_f32(points.data_ptr(), n, stride, C.byref(geom), _ARITH[arith],
                                       keep.data_ptr(), idx.data_ptr(), _stream(points)), "tp_voxel_index_f32")
    launch_count += 1 if n else 0
    return keep, idx


# ------------------------------------------------------------------------------------------------
# a3
# ------------------------------------------------------------------------------------------------
class _EncodeWorkspace:
    """Per (device, stream, geometry, batch) scratch whose tile counters are left zero by every successful call.
    Keyed by stream: two encodes of one geometry on different streams must not share counters / CSR lists. An
    entry is evicted when its call fails (the counters may be dirty), see encode()."""
    cache = {}

    @classmethod
    def clear(cls) -> None:
        cls.cache.clear()

    @classmethod
    def evict(cls, k) -> None:
        cls.cache.pop(k, None)

    @classmethod
    def key(cls, device, stream: int, key, batch: int):
        return (device.index, stream, key, batch)

    @classmethod
    def get(cls, device, geom: L.tp_geom, key, batch: int, n: int, stream: int) -> Tuple[torch.Tensor, int]:
        lib = L.lib()
        k = cls.key(device, stream, key, batch)
        need = lib.tp_encode_workspace_bytes(C.byref(geom), batch, n)
        if need < 0:
            raise TriplaneError("tp_encode_workspace_bytes: bad geometry")
        ent = cls.cache.get(k)
        if ent is None or ent[1] < n:
            cap_n = max(n, 1024) if ent is None else max(n, 2 * ent[1])
            nbytes = lib.tp_encode_workspace_bytes(C.byref(geom), batch, cap_n)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            L.check(lib.tp_encode_workspace_init(ws.data_ptr(), nbytes, stream), "tp_encode_workspace_init")
            ent = (ws, cap_n, nbytes)
            cls.cache[k] = ent
        return ent[0], ent[2]


@_on_device
def encode(feats: torch.Tensor, offsets: torch.Tensor, pc_range, voxel_size, grid_size, split, *,
           grid_ind: Optional[torch.Tensor] = None, points: Optional[torch.Tensor] = None,
           reduce: str = "max", clamp_zero: bool = False, want_counts: bool = False, arith: str = "cuda",
           planes: Sequence[bool] = (True, True, True), pool: Optional[Sequence[int]] = None):
    """Fused triplane encode (point_triplane_projector.py:99-115).

    feats [N, C]; offsets [B+1] int64; either grid_ind [N,3] int32 (reference signature) or raw
    points [N, >=3] (crop + index fused in-kernel). Returns (xy [B,X,Y,Zp*C], yz [B,Y,Z,Xp*C],
    xz [B,X,Z,Yp*C][, counts int32 [cells]]). pool overrides int(grid_size / split) (a slab of a larger grid keeps the
    full grid's pooling kernels: dist.encode_point_sharded, strategy 'owner')."""
    global launch_count
    _need_cuda(feats, "feats")
    _need_cuda(offsets, "offsets", torch.int64)
    if feats.dim() != 2:
        raise TriplaneError(f"feats must be [N, C], got {tuple(feats.shape)}")
    if feats.stride(1) != 1 or feats.stride(0) % 4 or feats.data_ptr() % 16:
        feats = feats.contiguous()
    n, Cch = feats.shape
    if (grid_ind is None) == (points is None):
        raise TriplaneError("encode: pass exactly one of grid_ind / points")
    if grid_ind is not None:
        _need_cuda(grid_ind, "grid_ind", torch.int32)
        grid_ind = grid_ind.contiguous()
        if tuple(grid_ind.shape) != (n, 3):
            raise TriplaneError(f"grid_ind must be [{n}, 3], got {tuple(grid_ind.shape)}")
    else:
        _need_cuda(points, "points")
        points = points.contiguous()
        if points.shape[0] != n or points.shape[1] < 3:
            raise TriplaneError(f"points must be [{n}, >=3], got {tuple(points.shape)}")
    batch = offsets.numel() - 1
    pool = pool_kernels(grid_size, split) if pool is None else tuple(int(v) for v in pool)
    P = pooled_sizes(grid_size, pool)
    X, Y, Z = (int(g) for g in grid_size)
    geom = L.make_geom(pc_range, voxel_size, grid_size, pool)
    key = (tuple(float(v) for v in pc_range), tuple(float(v) for v in voxel_size), (X, Y, Z), pool)
    dev = feats.device
    stream = _stream(feats)
    ws, ws_bytes = _EncodeWorkspace.get(dev, geom, key, batch, n, stream)
    shapes = [(batch, X, Y, P[2] * Cch), (batch, Y, Z, P[0] * Cch), (batch, X, Z, P[1] * Cch)]
    outs = [torch.empty(s, dtype=torch.float32, device=dev) if use else None for s, use in zip(shapes, planes)]
    counts = None
    if want_counts:
        ncell = batch * (X * Y * P[2] + Y * Z * P[0] + X * Z * P[1])
        counts = torch.zeros(ncell, dtype=torch.int32, device=dev)
    try:
        L.check(L.lib().tp_encode_f32(feats.data_ptr(), feats.stride(0), Cch, _ptr(grid_ind), _ptr(points),
                                      0 if points is None else points.shape[1], n, offsets.data_ptr(), batch,
                                      C.byref(geom), _ARITH[arith], _REDUCE[reduce], int(bool(clamp_zero)),
                                      _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(counts),
                                      ws.data_ptr(), ws_bytes, stream), "tp_encode_f32")
    except BaseException:
        # a failure between the passes can leave tile counters non-zero: never reuse this workspace
        _EncodeWorkspace.evict(_EncodeWorkspace.key(dev, stream, key, batch))
        raise
    launch_count += 4 if n else 1
    return (outs[0], outs[1], outs[2], counts) if want_counts else (outs[0], outs[1], outs[2])


def clear_workspaces() -> None:
    """Drop every cached encode workspace (frees the device memory once in-flight work has finished)."""
    _EncodeWorkspace.clear()
    _sparse_ws.clear()


_sparse_ws = {}


def sparse_projector_supported(C_: int, grid_size, split) -> bool:
    pool = pool_kernels(grid_size, split)
    return C_ % 4 == 0 and 4 <= C_ <= 128 and all(1 <= p <= 64 for p in pooled_sizes(grid_size, pool))


@_on_device
def projector_sparse(feats: torch.Tensor, offsets: torch.Tensor, pc_range, voxel_size, grid_size, split, weights, biases, *,
                     grid_ind: Optional[torch.Tensor] = None, points: Optional[torch.Tensor] = None, relu: bool = True,
                     clamp_zero: bool = False, arith: str = "cuda"):
    """Scatter-max + the first per-plane Linear of PointTriplaneProjector (point_triplane_projector.py:99-115, 60-64)
    over the OCCUPIED pooled cells only (tp_projector_sparse_f32): the dense [B,X,Y,Zp*C] tensors never exist.
    weights = (mlp_xy[0].weight [C, Zp*C], mlp_yz[0].weight [C, Xp*C], mlp_xz[0].weight [C, Yp*C]), biases likewise.
    Returns (h_xy [B,X,Y,C], h_yz [B,Y,Z,C], h_xz [B,X,Z,C]) = act(Linear(dense pooled tensor))."""
    global launch_count
    _need_cuda(feats, "feats")
    _need_cuda(offsets, "offsets", torch.int64)
    if feats.dim() != 2:
        raise TriplaneError(f"feats must be [N, C], got {tuple(feats.shape)}")
    if feats.stride(1) != 1 or feats.stride(0) % 4 or feats.data_ptr() % 16:
        feats = feats.contiguous()
    n, Cc = feats.shape
    if (grid_ind is None) == (points is None):
        raise TriplaneError("projector_sparse: pass exactly one of grid_ind / points")
    if grid_ind is not None:
        _need_cuda(grid_ind, "grid_ind", torch.int32)
        grid_ind = grid_ind.contiguous()
    else:
        _need_cuda(points, "points")
        points = points.contiguous()
    if not sparse_projector_supported(Cc, grid_size, split):
        raise TriplaneError(f"projector_sparse: C={Cc} / pooled sizes unsupported (C % 4 == 0, C <= 128, <= 64 pooled cells per axis)")
    batch = offsets.numel() - 1
    pool = pool_kernels(grid_size, split)
    P = pooled_sizes(grid_size, pool)
    X, Y, Z = (int(g) for g in grid_size)
    groups = (P[2], P[0], P[1])
    geom = L.make_geom(pc_range, voxel_size, grid_size, pool)
    dev = feats.device
    stream = _stream(feats)
    wts, bs = [], []
    for k, (w, b, G) in enumerate(zip(weights, biases, groups)):
        _need_cuda(w, f"weight {k}")
        _need_cuda(b, f"bias {k}")
        if tuple(w.shape) != (Cc, G * Cc) or tuple(b.shape) != (Cc,):
            raise TriplaneError(f"projector_sparse: plane {k}: weight {tuple(w.shape)} / bias {tuple(b.shape)} do not form "
                                f"Linear({G}*{Cc} -> {Cc})")
        wts.append(w.detach().view(Cc, G, Cc).permute(1, 2, 0).contiguous())  # [g][k][n]
        bs.append(b.detach().contiguous())
    lib = L.lib()
    need = lib.tp_projector_sparse_workspace_bytes(C.byref(geom), batch, Cc)
    key = (dev.index, stream)
    ws = _sparse_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _sparse_ws[key] = ws
    hidden = [torch.empty((batch, X, Y, Cc), dtype=torch.float32, device=dev),
              torch.empty((batch, Y, Z, Cc), dtype=torch.float32, device=dev),
              torch.empty((batch, X, Z, Cc), dtype=torch.float32, device=dev)]
    arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])  # noqa: E731
    wa, ba, ha = arr(wts), arr(bs), arr(hidden)
    L.check(lib.tp_projector_sparse_f32(feats.data_ptr(), feats.stride(0), Cc, _ptr(grid_ind), _ptr(points),
                                        0 if points is None else points.shape[1], n, offsets.data_ptr(), batch, C.byref(geom),
                                        _ARITH[arith], int(bool(clamp_zero)), C.byref(wa), C.byref(ba), int(bool(relu)),
                                        C.byref(ha), ws.data_ptr(), ws.numel(), stream), "tp_projector_sparse_f32")
    launch_count += 5 if n else 1
    return tuple(hidden)


@_on_device
def finalize_mean(planes: torch.Tensor, counts: torch.Tensor, channels: int) -> torch.Tensor:
    """In place: planes[cell, :] /= max(count[cell], 1) — after a point-sharded SUM all-reduce."""
    global launch_count
    _need_cuda(planes, "planes")
    _need_cuda(counts, "counts", torch.int32)
    if not planes.is_contiguous():
        raise TriplaneError("finalize_mean: planes must be contiguous")
    cells = planes.numel() // channels
    L.check(L.lib().tp_encode_finalize_mean_f32(planes.data_ptr(), counts.data_ptr(), cells, channels,
                                                _stream(planes)), "tp_encode_finalize_mean_f32")
    launch_count += 1
    return planes


@_on_device
def finalize_max(planes: torch.Tensor, clamp_zero: bool = False) -> torch.Tensor:
    """In place: -inf -> 0 after the all-reduce(max) of reduce='max_partial' planes."""
    global launch_count
    _need_cuda(planes, "planes")
    if not planes.is_contiguous():
        raise TriplaneError("finalize_max: planes must be contiguous")
    L.check(L.lib().tp_encode_finalize_max_f32(planes.data_ptr(), planes.numel(), int(bool(clamp_zero)),
                                               _stream(planes)), "tp_encode_finalize_max_f32")
    launch_count += 1
    return planes


@_on_device
def voxel_counts(grid_ind: torch.Tensor, offsets: torch.Tensor, grid_size) -> torch.Tensor:
    """Dense [B,X,Y,Z] int32 histogram of points per voxel (unq_cnt of projector.py:99, densified)."""
    global launch_count
    _need_cuda(grid_ind, "grid_ind", torch.int32)
    _need_cuda(offsets, "offsets", torch.int64)
    batch = offsets.numel() - 1
    X, Y, Z = (int(g) for g in grid_size)
    geom = L.make_geom([0] * 6, (1, 1, 1), grid_size, (1, 1, 1))
    counts = torch.zeros((batch, X, Y, Z), dtype=torch.int32, device=grid_ind.device)
    grid_ind = grid_ind.contiguous()
    L.check(L.lib().tp_voxel_counts_i32(grid_ind.data_ptr(), grid_ind.shape[0], offsets.data_ptr(), batch,
                                        C.byref(geom), counts.data_ptr(), _stream(grid_ind)), "tp_voxel_counts_i32")
    launch_count += 1 if grid_ind.shape[0] else 0
    return counts


# ------------------------------------------------------------------------------------------------
# a4
# ------------------------------------------------------------------------------------------------
def _plane_array(planes: Sequence[torch.Tensor]):
    arr = (L.tp_plane * 3)()
    for k, p in enumerate(planes):
        arr[k].data = p.data_ptr()
        arr[k].batch_stride = p.stride(0)
        arr[k].H, arr[k].W = p.shape[-2], p.shape[-1]
    return arr


@_on_device
def planes_to_channels_last(planes: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """NCHW planes (the reference layout; views of the stacked [B,3,C,H,W] are fine) -> contiguous
    channels-last [B,H,W,C] copies for the gather kernel."""
    global launch_count
    if len(planes) != 3:
        raise TriplaneError("expected three planes")
    srcs, out = [], []
    for k, p in enumerate(planes):
        _need_cuda(p, f"plane {k}")
        if p.dim() != 4:
            raise TriplaneError(f"plane {k} must be [B,C,H,W], got {tuple(p.shape)}")
        B, Cc, H, W = p.shape
        if p.stride(3) != 1 or p.stride(2) != W or p.stride(1) != H * W:
            p = p.contiguous()
        srcs.append(p)
        out.append(torch.empty((B, H, W, Cc), dtype=torch.float32, device=p.device))
    B, Cc = srcs[0].shape[0], srcs[0].shape[1]
    if any(p.shape[0] != B or p.shape[1] != Cc for p in srcs):
        raise TriplaneError("planes must share batch and channel sizes")
    arr = _plane_array(srcs)
    dsts = (C.c_void_p * 3)(*[o.data_ptr() for o in out])
    L.check(L.lib().tp_planes3_nchw_to_nhwc_f32(C.byref(arr), C.byref(dsts), B, Cc, _stream(srcs[0])),
            "tp_planes3_nchw_to_nhwc_f32")
    launch_count += 1
    return out


def _check_out(out, B, Cc, Q, device):
    if out is None:
        return torch.empty((B, Cc, Q), dtype=torch.float32, device=device)
    if tuple(out.shape) != (B, Cc, Q) or not out.is_contiguous() or not out.is_cuda:
        raise TriplaneError("sample3: bad `out`")
    return out


@_on_device
def sample3(planes: Union[torch.Tensor, Sequence[torch.Tensor]], queries: torch.Tensor, lo, vs, half, *,
            arith: str = "cuda", channels_last: bool = False, out: Optional[torch.Tensor] = None,
            grid_dims: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Fused 3-plane bilinear sample + sum.

    planes: stacked [B,3,C,H,W] or a list of three NCHW planes [B,C,H_p,W_p] (channels_last=True:
    already-converted [B,H_p,W_p,C] copies from planes_to_channels_last()). queries [B,Q,3].
    half[a] = S_a / 2 (triplane_occ.py:337 / point_triplane.py:455-458). Returns [B,C,Q].
    grid_dims=(h, w, d) with h*w*d == Q: the queries are a flattened [B,h,w,d,3] tensor (the 5-D
    callers); same result bit for bit, through tp_sample3_grid_nhwc_f32 which evaluates lattice-
    structured blocks once per index pair instead of once per query."""
    global launch_count
    if isinstance(planes, torch.Tensor):
        if planes.dim() != 5 or planes.shape[1] != 3:
            raise TriplaneError(f"stacked triplane must be [B,3,C,H,W], got {tuple(planes.shape)}")
        planes = [planes[:, 0], planes[:, 1], planes[:, 2]]
    if len(planes) != 3:
        raise TriplaneError("expected three planes")
    _need_cuda(queries, "queries")
    if queries.dim() != 3 or queries.shape[-1] != 3:
        raise TriplaneError(f"queries must be [B,Q,3], got {tuple(queries.shape)}")
    queries = queries.contiguous()
    B, Q, _ = queries.shape
    sg = L.make_sample_geom(lo, vs, half)
    dims = None
    if grid_dims is not None:
        h, w, d = (int(v) for v in grid_dims)
        if h * w * d != Q:
            raise TriplaneError(f"sample3: grid_dims {tuple(grid_dims)} do not multiply to Q={Q}")
        dims = (C.c_int32 * 3)(h, w, d)
    if not channels_last:
        # reference layout in: ONE C-ABI call = conversion kernel + gather kernel
        srcs = []
        for k, p in enumerate(planes):
            _need_cuda(p, f"plane {k}")
            if p.dim() != 4:
                raise TriplaneError(f"plane {k} must be [B,C,H,W], got {tuple(p.shape)}")
            if p.stride(3) != 1 or p.stride(2) != p.shape[3] or p.stride(1) != p.shape[2] * p.shape[3]:
                p = p.contiguous()
            srcs.append(p)
        Cc = srcs[0].shape[1]
        if any(p.shape[0] != B or p.shape[1] != Cc for p in srcs):
            raise TriplaneError("planes must share the queries' batch size and one channel count")
        arr = _plane_array(srcs)
        ws_floats = sum(B * Cc * p.shape[2] * p.shape[3] for p in srcs)
        ws = torch.empty(ws_floats, dtype=torch.float32, device=queries.device)
        out = _check_out(out, B, Cc, Q, queries.device)
        if Q:
            if dims is not None:
                L.check(L.lib().tp_sample3_grid_nchw_f32(C.byref(arr), Cc, queries.data_ptr(), C.byref(dims), B, C.byref(sg),
                                                         _ARITH[arith], out.data_ptr(), ws.data_ptr(), ws_floats,
                                                         _stream(queries)), "tp_sample3_grid_nchw_f32")
                launch_count += 2
            else:
                L.check(L.lib().tp_sample3_nchw_f32(C.byref(arr), Cc, queries.data_ptr(), Q, B, C.byref(sg), _ARITH[arith],
                                                    out.data_ptr(), ws.data_ptr(), ws_floats, _stream(queries)),
                        "tp_sample3_nchw_f32")
                launch_count += 2
        return out
    nhwc = list(planes)
    Cc = nhwc[0].shape[-1]
    for k, p in enumerate(nhwc):
        _need_cuda(p, f"plane {k}")
        if p.shape[0] != B or p.shape[-1] != Cc or not p.is_contiguous():
            raise TriplaneError(f"plane {k}: expected contiguous [B={B},H,W,C={Cc}], got {tuple(p.shape)}")
    arr = (L.tp_plane * 3)()
    for k, p in enumerate(nhwc):
        arr[k].data = p.data_ptr()
        arr[k].batch_stride = p.stride(0)
        arr[k].H, arr[k].W = p.shape[1], p.shape[2]
    out = _check_out(out, B, Cc, Q, queries.device)
    if dims is not None:
        L.check(L.lib().tp_sample3_grid_nhwc_f32(C.byref(arr), Cc, queries.data_ptr(), C.byref(dims), B,
                                                 C.byref(sg), _ARITH[arith], out.data_ptr(), _stream(queries)),
                "tp_sample3_grid_nhwc_f32")
    else:
        L.check(L.lib().tp_sample3_nhwc_f32(C.byref(arr), Cc, queries.data_ptr(), Q, B, C.byref(sg),
                                            _ARITH[arith], out.data_ptr(), _stream(queries)), "tp_sample3_nhwc_f32")
    launch_count += 1 if Q else 0
    return out


# ------------------------------------------------------------------------------------------------
# a2
# ------------------------------------------------------------------------------------------------
def pack_cameras(img_metas, device) -> torch.Tensor:
    """img_metas (the reference's per-sample dicts: 'lidar2image' [ncam,4,4], 'imgs_aug' list of
    dict(resize, crop, flip)) -> [B, ncam, 20] fp32 on `device` for tp_lift_cam_f32. Host-side packing
    of ~1 KB; the reference uploads the same matrices every call (point_triplane.py:183-184)."""
    import numpy as np
    rows = []
    for meta in img_metas:
        l2i = np.asarray(meta["lidar2image"], dtype=np.float64)
        augs = meta["imgs_aug"]
        per = np.zeros((l2i.shape[0], 20), dtype=np.float32)
        per[:, :16] = l2i.reshape(l2i.shape[0], 16).astype(np.float32)  # new_tensor(float64) -> fp32
        for c, aug in enumerate(augs):
            per[c, 16] = np.float32(aug["resize"])
            per[c, 17] = np.float32(aug["crop"][0])
            per[c, 18] = np.float32(aug["crop"][1])
            per[c, 19] = 1.0 if aug["flip"] else 0.0
        rows.append(per)
    return torch.from_numpy(np.stack(rows)).to(device, non_blocking=True)


@_on_device
def features_to_channels_last(img_features: torch.Tensor) -> torch.Tensor:
    """[B, ncam, Cf, Hf, Wf] (camera encoder output) -> contiguous [B, ncam, Hf, Wf, Cf]."""
    global launch_count
    _need_cuda(img_features, "img_features")
    if img_features.dim() != 5:
        raise TriplaneError(f"img_features must be [B,ncam,C,H,W], got {tuple(img_features.shape)}")
    img_features = img_features.contiguous()
    B, ncam, Cf, Hf, Wf = img_features.shape
    out = torch.empty((B, ncam, Hf, Wf, Cf), dtype=torch.float32, device=img_features.device)
    L.check(L.lib().tp_planes_nchw_to_nhwc_f32(img_features.data_ptr(), Cf * Hf * Wf, out.data_ptr(), B * ncam, Cf, Hf,
                                               Wf, _stream(img_features)), "tp_planes_nchw_to_nhwc_f32")
    launch_count += 1
    return out


@_on_device
def lift_cam(points: torch.Tensor, offsets: torch.Tensor, img_features: torch.Tensor, cams: torch.Tensor,
             resize_dims, *, arith: str = "cuda", channels_last: bool = False) -> torch.Tensor:
    """Fused point_to_cam (point_triplane.py:164-241): points [N, >=3] (samples concatenated, offsets
    [B+1] int64), img_features [B,ncam,Cf,Hf,Wf] (or the [B,ncam,Hf,Wf,Cf] copy with
    channels_last=True), cams = pack_cameras(img_metas). Returns [N, Cf]."""
    global launch_count
    _need_cuda(points, "points")
    _need_cuda(offsets, "offsets", torch.int64)
    _need_cuda(cams, "cams")
    points = points.contiguous()
    feats = img_features if channels_last else features_to_channels_last(img_features)
    _need_cuda(feats, "img_features")
    if feats.dim() != 5 or not feats.is_contiguous():
        raise TriplaneError("lift_cam: channels-last features must be contiguous [B,ncam,Hf,Wf,Cf]")
    B, ncam, Hf, Wf, Cf = feats.shape
    if tuple(cams.shape) != (B, ncam, 20) or offsets.numel() != B + 1:
        raise TriplaneError(f"lift_cam: cams must be [{B},{ncam},20] and offsets [{B + 1}]")
    cams = cams.contiguous()
    n = points.shape[0]
    out = torch.empty((n, Cf), dtype=torch.float32, device=points.device)
    L.check(L.lib().tp_lift_cam_f32(points.data_ptr(), points.shape[1], n, offsets.data_ptr(), B, feats.data_ptr(),
                                    ncam, Hf, Wf, Cf, cams.data_ptr(), float(resize_dims[0]), float(resize_dims[1]),
                                    _ARITH[arith], out.data_ptr(), _stream(points)), "tp_lift_cam_f32")
    launch_count += 1 if n else 0
    return out


# ------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------
@_on_device
def channels_last_to_nchw(x: torch.Tensor) -> torch.Tensor:
    """[N, H, W, C] -> contiguous [N, C, H, W] (the transpose kernel with the roles of C and H*W swapped)."""
    global launch_count
    _need_cuda(x, "x")
    N, H, W, Cc = x.shape
    x = x.contiguous()
    out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    # source viewed as [N, "C"=H*W, "HW"=C] NCHW-style -> destination [N, "HW"=C, "C"=H*W]
    L.check(L.lib().tp_planes_nchw_to_nhwc_f32(x.data_ptr(), H * W * Cc, out.data_ptr(), N, H * W, 1, Cc, _stream(x)),
            "tp_planes_nchw_to_nhwc_f32")
    launch_count += 1
    return out


@_on_device
def sample3_backward(grad_out: torch.Tensor, queries: torch.Tensor, plane_shapes, lo, vs, half, *,
                     arith: str = "cuda", grid_dims: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
    """Gradient of sample3 w.r.t. the three planes. grad_out [B,C,Q], queries [B,Q,3], plane_shapes = three
    (H, W). Returns three NCHW gradients [B,C,H,W]. grid_dims=(h, w, d): the queries are a flattened [B,h,w,d,3]
    tensor; lattice blocks are reduced per index pair before the scatter (tp_sample3_grid_backward_nhwc_f32)."""
    global launch_count
    _need_cuda(grad_out, "grad_out")
    _need_cuda(queries, "queries")
    grad_out, queries = grad_out.contiguous(), queries.contiguous()
    B, Cc, Q = grad_out.shape
    g_nhwc = [torch.zeros((B, H, W, Cc), dtype=torch.float32, device=grad_out.device) for H, W in plane_shapes]
    arr = (L.tp_plane * 3)()
    for k, p in enumerate(g_nhwc):
        arr[k].data = p.data_ptr()
        arr[k].batch_stride = p.stride(0)
        arr[k].H, arr[k].W = p.shape[1], p.shape[2]
    sg = L.make_sample_geom(lo, vs, half)
    if grid_dims is not None:
        h, w, d = (int(v) for v in grid_dims)
        if h * w * d != Q:
            raise TriplaneError(f"sample3_backward: grid_dims {tuple(grid_dims)} do not multiply to Q={Q}")
        dims = (C.c_int32 * 3)(h, w, d)
        L.check(L.lib().tp_sample3_grid_backward_nhwc_f32(C.byref(arr), Cc, queries.data_ptr(), C.byref(dims), B, C.byref(sg),
                                                          _ARITH[arith], grad_out.data_ptr(), _stream(grad_out)),
                "tp_sample3_grid_backward_nhwc_f32")
    else:
        L.check(L.lib().tp_sample3_backward_nhwc_f32(C.byref(arr), Cc, queries.data_ptr(), Q, B, C.byref(sg), _ARITH[arith],
                                                     grad_out.data_ptr(), _stream(grad_out)), "tp_sample3_backward_nhwc_f32")
    launch_count += 1 if Q else 0
    return [channels_last_to_nchw(g) for g in g_nhwc]


@_on_device
def encode_backward(grads, feats: Optional[torch.Tensor], grid_ind: torch.Tensor, offsets: torch.Tensor, grid_size,
                    split, *, outs=None, counts: Optional[torch.Tensor] = None, reduce: str = "max",
                    clamp_zero: bool = False, channels: Optional[int] = None) -> torch.Tensor:
    """Gradient of encode w.r.t. the point features: grads = (g_xy, g_yz, g_xz) (None to skip a plane),
    outs = the forward outputs (reduce='max'), counts = the forward cell counts (reduce='mean')."""
    global launch_count
    _need_cuda(grid_ind, "grid_ind", torch.int32)
    _need_cuda(offsets, "offsets", torch.int64)
    n = grid_ind.shape[0]
    Cc = feats.shape[1] if feats is not None else int(channels)
    if feats is not None:
        _need_cuda(feats, "feats")
        if feats.stride(1) != 1 or feats.stride(0) % 4 or feats.data_ptr() % 16:
            feats = feats.contiguous()
    grads = [None if g is None else _need_cuda(g, "grad").contiguous() for g in grads]
    outs = [None, None, None] if outs is None else [None if o is None else o.contiguous() for o in outs]
    pool = pool_kernels(grid_size, split)
    geom = L.make_geom([0] * 6, (1, 1, 1), grid_size, pool)
    gf = torch.empty((n, Cc), dtype=torch.float32, device=grid_ind.device)
    L.check(L.lib().tp_encode_backward_f32(_ptr(feats), 0 if feats is None else feats.stride(0), Cc,
                                           grid_ind.contiguous().data_ptr(), n, offsets.data_ptr(), offsets.numel() - 1,
                                           C.byref(geom), _REDUCE[reduce], int(bool(clamp_zero)), _ptr(outs[0]),
                                           _ptr(outs[1]), _ptr(outs[2]), _ptr(counts), _ptr(grads[0]), _ptr(grads[1]),
                                           _ptr(grads[2]), gf.data_ptr(), _stream(grid_ind)), "tp_encode_backward_f32")
    launch_count += 1 if n else 0
    return gf


@_on_device
def lift_cam_backward(grad_out: torch.Tensor, points: torch.Tensor, offsets: torch.Tensor, feat_shape, cams,
                      resize_dims, *, arith: str = "cuda") -> torch.Tensor:
    """Gradient of lift_cam w.r.t. img_features: grad_out [N,Cf] -> [B,ncam,Cf,Hf,Wf]."""
    global launch_count
    _need_cuda(grad_out, "grad_out")
    grad_out, points = grad_out.contiguous(), points.contiguous()
    B, ncam, Cf, Hf, Wf = feat_shape
    g = torch.zeros((B * ncam, Hf, Wf, Cf), dtype=torch.float32, device=grad_out.device)
    n = points.shape[0]
    L.check(L.lib().tp_lift_cam_backward_f32(points.data_ptr(), points.shape[1], n, offsets.data_ptr(), B, ncam, Hf, Wf,
                                             Cf, cams.contiguous().data_ptr(), float(resize_dims[0]),
                                             float(resize_dims[1]), _ARITH[arith], grad_out.data_ptr(), g.data_ptr(),
                                             _stream(grad_out)), "tp_lift_cam_backward_f32")
    launch_count += 1 if n else 0
    return channels_last_to_nchw(g).view(B, ncam, Cf, Hf, Wf)


# ------------------------------------------------------------------------------------------------
# occupancy head
# ------------------------------------------------------------------------------------------------
@_on_device
def sample3_head(planes, queries: torch.Tensor, lo, vs, half, w1: torch.Tensor, w2: torch.Tensor, w3: torch.Tensor, *,
                 grid_dims: Sequence[int], arith: str = "cuda", channels_last: bool = False) -> torch.Tensor:
    """TriplaneOcc's decode + occupancy head in one kernel (triplane_occ.py:182-186 after :321-348): the [B,32,Q]
    feature tensor never exists. planes / queries / grid_dims as in sample3 (C = 32), weights as in mlp_head.
    Returns logits [B, ncls, Q], equal to mlp_head(sample3(...))."""
    global launch_count
    if isinstance(planes, torch.Tensor):
        if planes.dim() != 5 or planes.shape[1] != 3:
            raise TriplaneError(f"stacked triplane must be [B,3,C,H,W], got {tuple(planes.shape)}")
        planes = [planes[:, 0], planes[:, 1], planes[:, 2]]
    _need_cuda(queries, "queries")
    queries = queries.contiguous()
    B, Q, _ = queries.shape
    h, w, d = (int(v) for v in grid_dims)
    if h * w * d != Q:
        raise TriplaneError(f"sample3_head: grid_dims {tuple(grid_dims)} do not multiply to Q={Q}")
    ws = []
    for k, wt in enumerate((w1, w2, w3)):
        _need_cuda(wt, f"w{k + 1}")
        ws.append(wt.reshape(wt.shape[0], wt.shape[1]).contiguous())
    ncls = ws[2].shape[0]
    sg = L.make_sample_geom(lo, vs, half)
    dims = (C.c_int32 * 3)(h, w, d)
    if not channels_last:
        arr, nws, nws_floats, Bp, Cc, keep = _nchw_planes(planes, "sample3_head")
    else:
        nhwc = list(planes)
        Cc, Bp = nhwc[0].shape[-1], nhwc[0].shape[0]
    if Bp != B or Cc != 32 or tuple(ws[0].shape) != (64, 32) or tuple(ws[1].shape) != (32, 64) or ws[2].shape[1] != 32:
        raise TriplaneError(f"sample3_head: this build fuses C=32 -> 64 -> 32 -> ncls on B={B} samples; got B={Bp}, C={Cc}, "
                            f"weights {[tuple(x.shape) for x in ws]}")
    out = torch.empty((B, ncls, Q), dtype=torch.float32, device=queries.device)
    if not channels_last:
        if Q:
            L.check(L.lib().tp_sample3_grid_head_nchw_tf32(C.byref(arr), queries.data_ptr(), C.byref(dims), B, C.byref(sg),
                                                           _ARITH[arith], ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(),
                                                           ncls, out.data_ptr(), nws.data_ptr(), nws_floats, _stream(queries)),
                    "tp_sample3_grid_head_nchw_tf32")
            launch_count += 2
        return out
    arr = (L.tp_plane * 3)()
    for k, p in enumerate(nhwc):
        _need_cuda(p, f"plane {k}")
        if p.shape[0] != B or p.shape[-1] != Cc or not p.is_contiguous():
            raise TriplaneError(f"plane {k}: expected contiguous [B={B},H,W,C={Cc}], got {tuple(p.shape)}")
        arr[k].data = p.data_ptr()
        arr[k].batch_stride = p.stride(0)
        arr[k].H, arr[k].W = p.shape[1], p.shape[2]
    L.check(L.lib().tp_sample3_grid_head_tf32(C.byref(arr), queries.data_ptr(), C.byref(dims), B, C.byref(sg), _ARITH[arith],
                                              ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ncls, out.data_ptr(),
                                              _stream(queries)), "tp_sample3_grid_head_tf32")
    launch_count += 1 if Q else 0
    return out


@_on_device
def mlp_head(feats: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor, w3: torch.Tensor) -> torch.Tensor:
    """The reference's Mlp head (dense_heads/mlp.py:57-70) on decode output: feats [B, C, Q] (or [B,C,X,Y,Z]),
    Conv3d weights w1 [2C,C,1,1,1], w2 [C,2C,1,1,1], w3 [ncls,C,1,1,1] (or already 2-D) -> logits [B, ncls, ...].
    Three layers fused on the tensor cores (TF32 inputs, fp32 accumulation, like cuDNN's default for Conv3d)."""
    global launch_count
    _need_cuda(feats, "feats")
    shape = feats.shape
    B, Cc = shape[0], shape[1]
    f = feats.reshape(B, Cc, -1).contiguous()
    ws = []
    for k, w in enumerate((w1, w2, w3)):
        _need_cuda(w, f"w{k + 1}")
        ws.append(w.reshape(w.shape[0], w.shape[1]).contiguous())
    ncls = ws[2].shape[0]
    if tuple(ws[0].shape) != (2 * Cc, Cc) or tuple(ws[1].shape) != (Cc, 2 * Cc) or ws[2].shape[1] != Cc:
        raise TriplaneError(f"mlp_head: weights {[tuple(w.shape) for w in ws]} do not form C -> 2C -> C -> ncls with C={Cc}")
    Q = f.shape[2]
    out = torch.empty((B, ncls, Q), dtype=torch.float32, device=f.device)
    L.check(L.lib().tp_mlp_head_tf32(f.data_ptr(), Q, B, Cc, ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ncls,
                                     out.data_ptr(), _stream(f)), "tp_mlp_head_tf32")
    launch_count += 1 if Q else 0
    return out.view(B, ncls, *shape[2:])


# ------------------------------------------------------------------------------------------------
# a4, ragged subsets and generated lattices
# ------------------------------------------------------------------------------------------------
def _nhwc_planes(planes, B: Optional[int], channels_last: bool, who: str):
    """-> (list of contiguous [B,H,W,C] planes, tp_plane array)"""
    if isinstance(planes, torch.Tensor):
        if planes.dim() != 5 or planes.shape[1] != 3:
            raise TriplaneError(f"{who}: stacked triplane must be [B,3,C,H,W], got {tuple(planes.shape)}")
        planes = [planes[:, 0], planes[:, 1], planes[:, 2]]
    if len(planes) != 3:
        raise TriplaneError(f"{who}: expected three planes")
    nhwc = list(planes) if channels_last else planes_to_channels_last(planes)
    Cc = nhwc[0].shape[-1]
    arr = (L.tp_plane * 3)()
    for k, p in enumerate(nhwc):
        _need_cuda(p, f"plane {k}")
        if (B is not None and p.shape[0] != B) or p.shape[-1] != Cc or p.dim() != 4 or not p.is_contiguous():
            raise TriplaneError(f"{who}: plane {k}: expected contiguous [B,H,W,C={Cc}], got {tuple(p.shape)}")
        arr[k].data = p.data_ptr()
        arr[k].batch_stride = p.stride(0)
        arr[k].H, arr[k].W = p.shape[1], p.shape[2]
    return nhwc, arr


def _nchw_planes(planes, who: str):
    """Reference-layout planes for the *_nchw_* entry points (conversion + decode chained inside ONE C-ABI call by
    programmatic dependent launch) -> (tp_plane array, workspace, workspace floats, B, C, keep-alive list)"""
    if isinstance(planes, torch.Tensor):
        if planes.dim() != 5 or planes.shape[1] != 3:
            raise TriplaneError(f"{who}: stacked triplane must be [B,3,C,H,W], got {tuple(planes.shape)}")
        planes = [planes[:, 0], planes[:, 1], planes[:, 2]]
    if len(planes) != 3:
        raise TriplaneError(f"{who}: expected three planes")
    srcs = []
    for k, p in enumerate(planes):
        _need_cuda(p, f"plane {k}")
        if p.dim() != 4:
            raise TriplaneError(f"{who}: plane {k} must be [B,C,H,W], got {tuple(p.shape)}")
        if p.stride(3) != 1 or p.stride(2) != p.shape[3] or p.stride(1) != p.shape[2] * p.shape[3]:
            p = p.contiguous()
        srcs.append(p)
    B, Cc = srcs[0].shape[0], srcs[0].shape[1]
    if any(p.shape[0] != B or p.shape[1] != Cc for p in srcs):
        raise TriplaneError(f"{who}: planes must share batch and channel sizes")
    ws_floats = sum(B * Cc * p.shape[2] * p.shape[3] for p in srcs)
    ws = torch.empty(ws_floats, dtype=torch.float32, device=srcs[0].device)
    return _plane_array(srcs), ws, ws_floats, B, Cc, srcs


@_on_device
def sample3_lattice(planes, dims: Sequence[int], origin, step, lo, vs, half, *, arith: str = "cuda",
                    channels_last: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sample3 on roi()'s voxel-centre lattice ((i, j, k) + 0.5) * step + origin (triplane_occ.py:311-316) without the
    [B,h,w,d,3] tensor: the coordinates are generated in the kernel. Returns [B, C, h*w*d], bit-identical to
    sample3(planes, roi_points, grid_dims=dims)."""
    global launch_count
    h, w, d = (int(v) for v in dims)
    Q = h * w * d
    sg = L.make_sample_geom(lo, vs, half)
    cd = (C.c_int32 * 3)(h, w, d)
    org = (C.c_float * 3)(*[float(v) for v in origin[:3]])
    stp = (C.c_float * 3)(*[float(v) for v in step[:3]])
    if not channels_last:
        arr, ws, ws_floats, B, Cc, keep = _nchw_planes(planes, "sample3_lattice")
        out = _check_out(out, B, Cc, Q, keep[0].device)
        if Q:
            L.check(L.lib().tp_sample3_lattice_nchw_f32(C.byref(arr), Cc, C.byref(cd), C.byref(org), C.byref(stp), B,
                                                        C.byref(sg), _ARITH[arith], out.data_ptr(), ws.data_ptr(), ws_floats,
                                                        _stream(keep[0])), "tp_sample3_lattice_nchw_f32")
            launch_count += 2
        return out
    nhwc, arr = _nhwc_planes(planes, None, True, "sample3_lattice")
    B, Cc = nhwc[0].shape[0], nhwc[0].shape[-1]
    out = _check_out(out, B, Cc, Q, nhwc[0].device)
    L.check(L.lib().tp_sample3_lattice_nhwc_f32(C.byref(arr), Cc, C.byref(cd), C.byref(org), C.byref(stp), B, C.byref(sg),
                                                _ARITH[arith], out.data_ptr(), _stream(nhwc[0])), "tp_sample3_lattice_nhwc_f32")
    launch_count += 1 if Q else 0
    return out


@_on_device
def sample3_segments(planes, queries: torch.Tensor, seg_offsets: torch.Tensor, seg_batch: Optional[torch.Tensor], lo, vs,
                     half, *, arith: str = "cuda", channels_last: bool = False) -> torch.Tensor:
    """Ragged subsets in one launch: queries [T,3] (segments concatenated), seg_offsets [S+1] int64, seg_batch [S] int32
    (None: segment s reads sample s) -> point-major [T, C]."""
    global launch_count
    _need_cuda(queries, "queries")
    _need_cuda(seg_offsets, "seg_offsets", torch.int64)
    if seg_batch is not None:
        _need_cuda(seg_batch, "seg_batch", torch.int32)
        seg_batch = seg_batch.contiguous()
    if queries.dim() != 2 or queries.shape[1] != 3:
        raise TriplaneError(f"queries must be [T,3], got {tuple(queries.shape)}")
    queries, seg_offsets = queries.contiguous(), seg_offsets.contiguous()
    nseg = seg_offsets.numel() - 1
    if nseg < 1 or (seg_batch is not None and seg_batch.numel() != nseg):
        raise TriplaneError("sample3_segments: seg_offsets must be [S+1] with S >= 1 and seg_batch [S]")
    T = queries.shape[0]
    sg = L.make_sample_geom(lo, vs, half)
    if not channels_last:
        arr, ws, ws_floats, B, Cc, keep = _nchw_planes(planes, "sample3_segments")
        out = torch.empty((T, Cc), dtype=torch.float32, device=queries.device)
        if T:
            L.check(L.lib().tp_sample3_seg_nchw_f32(C.byref(arr), Cc, queries.data_ptr(), T, seg_offsets.data_ptr(),
                                                    _ptr(seg_batch), nseg, B, C.byref(sg), _ARITH[arith], out.data_ptr(),
                                                    ws.data_ptr(), ws_floats, _stream(queries)), "tp_sample3_seg_nchw_f32")
            launch_count += 2
        return out
    nhwc, arr = _nhwc_planes(planes, None, True, "sample3_segments")
    B, Cc = nhwc[0].shape[0], nhwc[0].shape[-1]
    out = torch.empty((T, Cc), dtype=torch.float32, device=queries.device)
    L.check(L.lib().tp_sample3_seg_nhwc_f32(C.byref(arr), Cc, queries.data_ptr(), T, seg_offsets.data_ptr(), _ptr(seg_batch),
                                            nseg, B, C.byref(sg), _ARITH[arith], out.data_ptr(), _stream(queries)),
            "tp_sample3_seg_nhwc_f32")
    launch_count += 1 if T else 0
    return out


@_on_device
def sample3_segments_backward(grad_out: torch.Tensor, queries: torch.Tensor, seg_offsets: torch.Tensor,
                              seg_batch: Optional[torch.Tensor], batch: int, plane_shapes, lo, vs, half, *,
                              arith: str = "cuda") -> List[torch.Tensor]:
    """Gradient of sample3_segments w.r.t. the three planes: grad_out [T,C] -> three NCHW gradients [B,C,H,W]."""
    global launch_count
    _need_cuda(grad_out, "grad_out")
    grad_out, queries = grad_out.contiguous(), queries.contiguous()
    T, Cc = grad_out.shape
    g_nhwc = [torch.zeros((batch, H, W, Cc), dtype=torch.float32, device=grad_out.device) for H, W in plane_shapes]
    arr = (L.tp_plane * 3)()
    for k, p in enumerate(g_nhwc):
        arr[k].data = p.data_ptr()
        arr[k].batch_stride = p.stride(0)
        arr[k].H, arr[k].W = p.shape[1], p.shape[2]
    sg = L.make_sample_geom(lo, vs, half)
    nseg = seg_offsets.numel() - 1
    L.check(L.lib().tp_sample3_seg_backward_nhwc_f32(C.byref(arr), Cc, queries.data_ptr(), T, seg_offsets.data_ptr(),
                                                     _ptr(seg_batch), nseg, batch, C.byref(sg), _ARITH[arith],
                                                     grad_out.data_ptr(), _stream(grad_out)), "tp_sample3_seg_backward_nhwc_f32")
    launch_count += 1 if T else 0
    return [channels_last_to_nchw(g) for g in g_nhwc]


# ------------------------------------------------------------------------------------------------
# 8f #4: nearest-pixel gathers / scatters around the decode
# ------------------------------------------------------------------------------------------------
@_on_device
def pixel_winner_from_coors(coors: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """coors [..., npix, 2] fp32 (row, col; range_cam_coors of JointEncoder.interact, leading dims = images) ->
    winner [n_images, H, W] int32: the highest source pixel whose long() coordinates land there (-1: none)."""
    global launch_count
    _need_cuda(coors, "coors")
    if coors.dim() < 2 or coors.shape[-1] != 2:
        raise TriplaneError(f"coors must be [..., npix, 2], got {tuple(coors.shape)}")
    coors = coors.contiguous()
    npix = coors.shape[-2]
    M = coors.numel() // (2 * npix) if npix else 0
    winner = torch.empty((M, H, W), dtype=torch.int32, device=coors.device)
    L.check(L.lib().tp_pixel_winner_coors_i32(coors.data_ptr(), M, npix, H, W, winner.data_ptr(), _stream(coors)),
            "tp_pixel_winner_coors_i32")
    launch_count += 1
    return winner


@_on_device
def pixel_winner_from_points(points: torch.Tensor, offsets: torch.Tensor, cams: torch.Tensor, resize_dims) -> torch.Tensor:
    """points [N, >=3] (samples concatenated, offsets [B+1]), cams [B,ncam,20] -> winner [B*ncam, R0, R1] int32 holding
    the in-sample index of the last point projected onto each pixel (point_triplane.py:263-307)."""
    global launch_count
    _need_cuda(points, "points")
    _need_cuda(offsets, "offsets", torch.int64)
    _need_cuda(cams, "cams")
    points, cams = points.contiguous(), cams.contiguous()
    B, ncam = cams.shape[0], cams.shape[1]
    H, W = int(resize_dims[0]), int(resize_dims[1])
    winner = torch.empty((B * ncam, H, W), dtype=torch.int32, device=points.device)
    L.check(L.lib().tp_pixel_winner_points_i32(points.data_ptr(), points.shape[1], points.shape[0], offsets.data_ptr(), B,
                                               cams.data_ptr(), ncam, float(resize_dims[0]), float(resize_dims[1]),
                                               winner.data_ptr(), _stream(points)), "tp_pixel_winner_points_i32")
    launch_count += 1
    return winner


def _feat_addressing(feat: torch.Tensor, layout: str):
    """-> (C, batch stride, channel stride, source stride) of a contiguous feature tensor."""
    if layout == "bcn":  # [Bf, C, N] channel-major decode output
        if feat.dim() != 3:
            raise TriplaneError(f"feat must be [B,C,N], got {tuple(feat.shape)}")
        return feat.shape[1], feat.shape[1] * feat.shape[2], feat.shape[2], 1
    if layout == "nc":   # [N, C] point-major rows, per-sample first rows given separately
        if feat.dim() != 2:
            raise TriplaneError(f"feat must be [N,C], got {tuple(feat.shape)}")
        return feat.shape[1], 0, 1, feat.shape[1]
    raise TriplaneError(f"unknown feature layout {layout!r}")


@_on_device
def winner_gather(winner: torch.Tensor, feat: torch.Tensor, imgs_per_feat: int, *, layout: str = "bcn",
                  row0: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out [M, C, H, W] = feat[m // imgs_per_feat][:, winner[m]] (zeros where winner < 0). layout 'bcn': feat
    [Bf, C, N]; 'nc': feat [N, C] with row0 [Bf+1] int64 = first row of each sample."""
    global launch_count
    _need_cuda(winner, "winner", torch.int32)
    _need_cuda(feat, "feat")
    winner, feat = winner.contiguous(), feat.contiguous()
    M, H, W = winner.shape
    Cc, bs, cs, ns = _feat_addressing(feat, layout)
    if row0 is not None:
        _need_cuda(row0, "row0", torch.int64)
    out = torch.empty((M, Cc, H, W), dtype=torch.float32, device=feat.device)
    L.check(L.lib().tp_winner_gather_f32(winner.data_ptr(), M, H * W, Cc, imgs_per_feat, feat.data_ptr(), bs, cs, ns,
                                         _ptr(row0), out.data_ptr(), _stream(feat)), "tp_winner_gather_f32")
    launch_count += 1 if M else 0
    return out


@_on_device
def winner_gather_backward(winner: torch.Tensor, grad_out: torch.Tensor, feat_shape, imgs_per_feat: int, *,
                           layout: str = "bcn", row0: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launch_count
    grad_out, winner = grad_out.contiguous(), winner.contiguous()
    M, H, W = winner.shape
    gfeat = torch.zeros(tuple(feat_shape), dtype=torch.float32, device=grad_out.device)
    Cc, bs, cs, ns = _feat_addressing(gfeat, layout)
    L.check(L.lib().tp_winner_gather_backward_f32(winner.data_ptr(), M, H * W, Cc, imgs_per_feat, grad_out.data_ptr(),
                                                  gfeat.data_ptr(), bs, cs, ns, _ptr(row0), _stream(grad_out)),
            "tp_winner_gather_backward_f32")
    launch_count += 1 if M else 0
    return gfeat


@_on_device
def range_project(range_points: torch.Tensor, range_image: torch.Tensor, cams: torch.Tensor, resize_dims, feat_hw, *,
                  arith: str = "cuda"):
    """JointEncoder.interact's projection (joint_encoder.py:125-205): range_points [B,Hr,Wr,3], range_image
    [B,1,Hr,Wr] (masked), cams [B,ncam,20] -> (range_cam_coors [B,ncam,Hr,Wr,2], fidx [B,ncam,Hr*Wr] int32, winner
    [B*ncam, Hf*Wf] int32)."""
    global launch_count
    _need_cuda(range_points, "range_points")
    _need_cuda(range_image, "range_image")
    _need_cuda(cams, "cams")
    B, Hr, Wr, _ = range_points.shape
    range_points, range_image, cams = range_points.contiguous(), range_image.contiguous(), cams.contiguous()
    if range_image.numel() != B * Hr * Wr:
        raise TriplaneError(f"range_image {tuple(range_image.shape)} does not match range_points {tuple(range_points.shape)}")
    ncam = cams.shape[1]
    Hf, Wf = int(feat_hw[0]), int(feat_hw[1])
    npix = Hr * Wr
    dev = range_points.device
    coors = torch.empty((B, ncam, Hr, Wr, 2), dtype=torch.float32, device=dev)
    fidx = torch.empty((B, ncam, npix), dtype=torch.int32, device=dev)
    winner = torch.empty((B * ncam, Hf * Wf), dtype=torch.int32, device=dev)
    L.check(L.lib().tp_range_project_f32(range_points.data_ptr(), range_image.data_ptr(), npix, B, cams.data_ptr(), ncam,
                                         float(resize_dims[0]), float(resize_dims[1]), Hf, Wf, _ARITH[arith],
                                         coors.data_ptr(), fidx.data_ptr(), winner.data_ptr(), _stream(range_points)),
            "tp_range_project_f32")
    launch_count += 1
    return coors, fidx, winner


@_on_device
def range_gather(fidx: torch.Tensor, img_features: torch.Tensor) -> torch.Tensor:
    """cam_range_features [B, C, npix] = sum over cameras of img_features[b, cam, :, fidx] (joint_encoder.py:208)."""
    global launch_count
    _need_cuda(fidx, "fidx", torch.int32)
    _need_cuda(img_features, "img_features")
    B, ncam, Cc, Hf, Wf = img_features.shape
    img_features = img_features.contiguous()
    npix = fidx.shape[-1]
    out = torch.empty((B, Cc, npix), dtype=torch.float32, device=img_features.device)
    L.check(L.lib().tp_range_gather_f32(fidx.contiguous().data_ptr(), npix, B, ncam, img_features.data_ptr(), Cc, Hf * Wf,
                                        out.data_ptr(), _stream(img_features)), "tp_range_gather_f32")
    launch_count += 1 if npix else 0
    return out


@_on_device
def range_gather_backward(fidx: torch.Tensor, grad_out: torch.Tensor, img_shape) -> torch.Tensor:
    global launch_count
    B, ncam, Cc, Hf, Wf = img_shape
    grad_out = grad_out.contiguous()
    g = torch.zeros(tuple(img_shape), dtype=torch.float32, device=grad_out.device)
    npix = fidx.shape[-1]
    L.check(L.lib().tp_range_gather_backward_f32(fidx.contiguous().data_ptr(), npix, B, ncam, grad_out.data_ptr(), Cc, Hf * Wf,
                                                 g.data_ptr(), _stream(grad_out)), "tp_range_gather_backward_f32")
    launch_count += 1 if npix else 0
    return g


@_on_device
def posembed_scatter_(img_features: torch.Tensor, winner: torch.Tensor, pos_embed: torch.Tensor) -> torch.Tensor:
    """In place: img_features [B,ncam,C,Hf,Wf][m, :, p] += pos_embed [M, Hf*Wf, C][m, p, :] where winner [M, Hf*Wf] >= 0."""
    global launch_count
    _need_cuda(img_features, "img_features")
    _need_cuda(pos_embed, "pos_embed")
    _need_cuda(winner, "winner", torch.int32)
    if not img_features.is_contiguous():
        raise TriplaneError("posembed_scatter_: img_features must be contiguous (it is updated in place)")
    B, ncam, Cc, Hf, Wf = img_features.shape
    pos_embed = pos_embed.contiguous()
    L.check(L.lib().tp_posembed_scatter_f32(winner.contiguous().data_ptr(), B * ncam, Cc, Hf * Wf, img_features.data_ptr(),
                                            pos_embed.data_ptr(), _stream(img_features)), "tp_posembed_scatter_f32")
    launch_count += 1
    return img_features


@_on_device
def posembed_scatter_backward(grad_img: torch.Tensor, winner: torch.Tensor) -> torch.Tensor:
    global launch_count
    B, ncam, Cc, Hf, Wf = grad_img.shape
    grad_img = grad_img.contiguous()
    g = torch.empty((B * ncam, Hf * Wf, Cc), dtype=torch.float32, device=grad_img.device)
    L.check(L.lib().tp_posembed_scatter_backward_f32(winner.contiguous().data_ptr(), B * ncam, Cc, Hf * Wf, grad_img.data_ptr(),
                                                     g.data_ptr(), _stream(grad_img)), "tp_posembed_scatter_backward_f32")
    launch_count += 1
    return g


@_on_device
def radius(x: torch.Tensor, x_offsets: torch.Tensor, y: torch.Tensor, y_offsets: torch.Tensor, r: float,
           max_num_neighbors: int = 32):
    """torch_cluster.radius semantics (interpnet.py:65): for every query y, the first max_num_neighbors sources x of the
    same sample (index order) with squared distance < r^2. Returns (col [Ny, max] int32 with -1 padding, count [Ny])."""
    global launch_count
    _need_cuda(x, "x")
    _need_cuda(y, "y")
    _need_cuda(x_offsets, "x_offsets", torch.int64)
    _need_cuda(y_offsets, "y_offsets", torch.int64)
    x, y = x.contiguous(), y.contiguous()
    if x.dim() != 2 or x.shape[1] != 3 or y.dim() != 2 or y.shape[1] != 3:
        raise TriplaneError("radius: x and y must be [N,3]")
    ny = y.shape[0]
    B = y_offsets.numel() - 1
    if x_offsets.numel() != B + 1:
        raise TriplaneError("radius: x_offsets and y_offsets must describe the same number of samples")
    col = torch.empty((ny, max_num_neighbors), dtype=torch.int32, device=y.device)
    cnt = torch.empty((ny,), dtype=torch.int32, device=y.device)
    L.check(L.lib().tp_radius_i32(x.data_ptr(), x_offsets.data_ptr(), y.data_ptr(), y_offsets.data_ptr(), ny, B, float(r),
                                  int(max_num_neighbors), col.data_ptr(), cnt.data_ptr(), _stream(y)), "tp_radius_i32")
    launch_count += 1 if ny else 0
    return col, cnt
