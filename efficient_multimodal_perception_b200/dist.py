"""Multi-GPU plumbing for the triplane hot path: one process per GPU,
torch.distributed (NCCL over NVLink 5 / NVSwitch; gloo in CPU tests).

* decode — queries are independent: shard Q (or the batch) across ranks, no collective.
* encode, sample-sharded — planes are per sample: shard the batch across ranks, no collective (this is
  what the reference's DDP does, tools/euler_train.sh:3-11).
* encode, point-sharded — the one real exchange step, two strategies with the same result:
  - "planes": every rank scatters ITS points into full-size partial planes (max: empty cells are -inf, the
    identity of max; mean: sums + counts), the partial planes are combined with all-reduce (MAX, or SUM on
    sums and counts), then finalised (-inf -> 0, or sum / count). The message is the full dense pooled
    tensor (430 MB per sample at the config geometry).
  - "points": the ranks all-gather their point shards (12 + 4C bytes per point: 18 MB for one sweep, 183 MB
    for a 350k-point sample at C = 128) and every rank runs the complete fused encode. No partial planes, no
    finalise pass; it moves 2-20x fewer bytes over NVLink whenever a sample has fewer than ~800k points.
  max is exact under any reduction / gathering order; mean is within rounding. Point sharding only pays when
  there are fewer samples than GPUs — bench.py reports both strategies next to sample sharding.

The reference has no counterpart for the point-sharded path (its only collectives are the scalar loss
all-reduce, point_triplane.py:529-531, and DDP's gradient all-reduce).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


#: point-sharded encode strategies (encode_point_sharded): name -> what crosses NVLink
STRATEGIES = {
    "planes": "partial planes (-inf empties) -> NCCL all-reduce(max) of the dense pooled planes -> finalise; every rank holds the full planes",
    "points": "NCCL all-gather of the point shards (12 + 4C bytes per point) -> full fused encode on every rank; every rank holds the full planes",
    "owner": "every rank pushes its points (12 B index + 4C B features, to at most two owners) into the owners' receive buffers over NVLink peer "
             "memory in one kernel (tp_route_points_f32, torch symmetric memory, two device-side barriers) -> each rank encodes ITS slab of the "
             "three planes; the planes stay distributed (rank r: x-slab of xy / xz, y-slab of yz), nothing is replicated",
}


def planes_equal(got, ref, strategy: str, rank: int, world: int, rtol: float = 0.0, slab_bounds=None) -> bool:
    """Does this rank's result of encode_point_sharded equal the single-GPU planes `ref` = (xy, yz, xz)? 'planes' and
    'points' replicate the full planes on every rank; 'owner' leaves rank r with the x-slab of xy / xz and the y-slab of yz
    (slab_bounds = (xb, yb) as passed to encode_point_sharded, default equal widths)."""
    if strategy == "owner":
        X, Y = ref[0].shape[1], ref[1].shape[1]
        if slab_bounds is None:
            (x0, x1), (y0, y1) = shard_bounds(X, rank, world), shard_bounds(Y, rank, world)
        else:
            (x0, x1), (y0, y1) = slab_bounds[0][rank:rank + 2], slab_bounds[1][rank:rank + 2]
        ref = (ref[0][:, x0:x1], ref[1][:, y0:y1], ref[2][:, x0:x1])
    if rtol == 0.0:
        return all(a.shape == b.shape and torch.equal(a, b) for a, b in zip(got, ref))
    return all(a.shape == b.shape and float((a - b).abs().max()) <= rtol * float(b.abs().max()) for a, b in zip(got, ref))


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of range(n): the first n % world shards get one extra element."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_samples(items: Sequence, rank: int, world: int) -> list:
    lo, hi = shard_bounds(len(items), rank, world)
    return list(items[lo:hi])


def shard_points(points: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    """Per-sample contiguous slice of every sample's points for this rank (order preserved)."""
    out = []
    for p in points:
        lo, hi = shard_bounds(p.shape[0], rank, world)
        out.append(p[lo:hi])
    return out


def shard_queries(queries: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """[B, Q, 3] -> this rank's [B, Q_r, 3] slice along Q."""
    lo, hi = shard_bounds(queries.shape[1], rank, world)
    return queries[:, lo:hi]


def all_reduce_planes(planes: Sequence[torch.Tensor], reduce: str, counts: Optional[torch.Tensor] = None,
                      group=None) -> None:
    """In-place combine of partial planes across ranks: reduce='max' -> MAX; 'sum' -> SUM on planes and
    on the int32 point counts. Works on CUDA (NCCL) and CPU (gloo) tensors."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    op = dist.ReduceOp.MAX if reduce == "max" else dist.ReduceOp.SUM
    for p in planes:
        if p is not None:
            dist.all_reduce(p, op=op, group=group)
    if counts is not None:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)


def gather_point_shards(feats: torch.Tensor, points: torch.Tensor, offsets: torch.Tensor, group=None):
    """All-gather of the ranks' point shards. Every rank passes feats [N_r, C], points [N_r, D] and the
    offsets [B+1] of ITS shard (the same B everywhere) and receives (feats [N, C], points [N, D], offsets
    [B+1]) of the whole batch, sample-major; inside a sample the shards follow each other in rank order.
    Shards are padded to the largest one for the collective (one small all-gather of sizes first)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return feats, points, offsets
    world = dist.get_world_size(group)
    dev = feats.device
    offs = [torch.empty_like(offsets) for _ in range(world)]
    dist.all_gather(offs, offsets.contiguous(), group=group)
    offs = torch.stack(offs).cpu()                      # [world, B+1], one host sync
    sizes = offs[:, -1].tolist()
    nmax = max(max(sizes), 1)

    even = all(sz == nmax for sz in sizes)

    def gather(x):
        if even:  # the usual case (shard_bounds splits differ by at most one row): no padding copy
            pad = x.contiguous()
        else:
            pad = torch.zeros((nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
            pad[:x.shape[0]] = x
        flat = torch.empty((world * nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
        dist.all_gather_into_tensor(flat, pad, group=group)
        return flat

    f_flat, p_flat = gather(feats), gather(points)
    B = offsets.numel() - 1
    ends = [0]
    for b in range(B):
        ends.append(ends[-1] + int((offs[:, b + 1] - offs[:, b]).sum()))
    off_all = torch.tensor(ends, dtype=torch.int64, device=dev)
    if even and B == 1:
        return f_flat, p_flat, off_all  # rank order IS sample-major order: nothing to move
    f_parts = [f_flat[r * nmax:(r + 1) * nmax] for r in range(world)]
    p_parts = [p_flat[r * nmax:(r + 1) * nmax] for r in range(world)]
    f_out, p_out = [], []
    for b in range(B):
        for r in range(world):
            lo, hi = int(offs[r, b]), int(offs[r, b + 1])
            if hi > lo:
                f_out.append(f_parts[r][lo:hi])
                p_out.append(p_parts[r][lo:hi])
    cat = lambda xs, like: torch.cat(xs) if xs else like[:0]  # noqa: E731
    return cat(f_out, feats).contiguous(), cat(p_out, points).contiguous(), off_all


def balanced_slab_bounds(points: torch.Tensor, pc_range, voxel_size, grid_size, group=None, min_width=(1, 1)):
    """Slab boundaries for strategy 'owner' that give every rank about the same number of points instead of the same
    width: LiDAR density falls off with range, so with equal-width x / y slabs of the +-25 m grid the two centre ranks
    of eight receive 34 % and 27 % of a 10-sweep cloud and the outer ones 1 % (measured: their NVLink ingress and their
    encode are the step). One all-reduce of two voxel-index histograms + one host sync: call it once per stream of
    samples (the density profile of a sensor does not change from frame to frame), not per step.
    points: this rank's shard [N_r, >=3]. Returns (xb, yb): world + 1 increasing voxel indices each, every x / y slab
    at least min_width[0] / [1] voxel rows wide (the slab encode needs at least one pooling window:
    ops.pool_kernels(grid_size, split)[:2])."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    X, Y = int(grid_size[0]), int(grid_size[1])
    hist = torch.zeros(X + Y, dtype=torch.int64, device=points.device)
    if points.shape[0]:
        lo = torch.tensor([float(v) for v in pc_range[:3]], device=points.device)
        hi = torch.tensor([float(v) for v in pc_range[3:6]], device=points.device)
        p = points[:, :3]
        keep = ((p > lo) & (p < hi)).all(1)
        vs = torch.tensor([float(voxel_size[0]), float(voxel_size[1])], device=points.device)
        ij = ((p[keep, :2] - lo[:2]) / vs).long()
        ij[:, 0].clamp_(0, X - 1)
        ij[:, 1].clamp_(0, Y - 1)
        hist += torch.bincount(torch.cat((ij[:, 0], ij[:, 1] + X)), minlength=X + Y)
    if world > 1:
        dist.all_reduce(hist, group=group)
    hist = hist.cpu()

    def cut(h, minw):
        n = h.numel()
        if n < world * minw:
            raise ValueError(f"{n} voxel rows cannot be cut into {world} slabs of at least {minw}")
        cum = torch.cumsum(h, 0)
        total = int(cum[-1])
        b = [0]
        for r in range(1, world):
            target = total * r / world
            k = int(torch.searchsorted(cum, torch.tensor(target, dtype=cum.dtype))) + 1 if total else n * r // world
            k = max(k, b[-1] + minw)              # at least minw rows per slab ...
            k = min(k, n - (world - r) * minw)    # ... and room for the slabs that follow
            b.append(k)
        return b + [n]

    return cut(hist[:X], int(min_width[0])), cut(hist[X:], int(min_width[1]))


class _OwnerExchange:
    """Symmetric-memory receive buffers of the 'owner' strategy, one per (device, group, capacity, C): every rank can
    address every other rank's buffers (NVLink peer mappings). Layout of a rank's buffer (bytes):
    [cnt int32 x 64 | idx_x int32 [cap,3] | idx_y int32 [cap,3] | feat_x f32 [cap,C] | feat_y f32 [cap,C]]."""
    cache = {}

    def __init__(self, device, group, cap: int, C: int):
        import torch.distributed._symmetric_memory as symm_mem
        self.cap, self.C = cap, C
        pad = lambda b: (b + 255) // 256 * 256  # noqa: E731
        self.off_idx_x = 256
        self.off_idx_y = self.off_idx_x + pad(cap * 12)
        self.off_feat_x = self.off_idx_y + pad(cap * 12)
        self.off_feat_y = self.off_feat_x + pad(cap * C * 4)
        nbytes = self.off_feat_y + pad(cap * C * 4)
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.world, self.rank = self.hdl.world_size, self.hdl.rank
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        import ctypes as C_
        arr = lambda off: (C_.c_void_p * self.world)(*[p + off for p in ptrs])  # noqa: E731
        self.p_cnt, self.p_idx_x, self.p_idx_y = arr(0), arr(self.off_idx_x), arr(self.off_idx_y)
        self.p_feat_x, self.p_feat_y = arr(self.off_feat_x), arr(self.off_feat_y)
        view = lambda off, n, dt: self.buf[off:off + n * 4].view(dt)  # noqa: E731
        self.head = self.buf[:self.off_feat_x]                       # counters + both index regions: reset every call
        self.head32 = self.head.view(torch.int32)
        self.idx_x = view(self.off_idx_x, cap * 3, torch.int32).view(cap, 3)
        self.idx_y = view(self.off_idx_y, cap * 3, torch.int32).view(cap, 3)
        self.feat_x = view(self.off_feat_x, cap * C, torch.float32).view(cap, C)
        self.feat_y = view(self.off_feat_y, cap * C, torch.float32).view(cap, C)
        self.offsets = torch.tensor([0, cap], dtype=torch.int64, device=device)

    @classmethod
    def get(cls, device, group, cap: int, C: int) -> "_OwnerExchange":
        key = (device.index, id(group), C)
        ex = cls.cache.get(key)
        if ex is None or ex.cap < cap:
            ex = cls(device, group, cap, C)
            cls.cache[key] = ex
        return ex


def _encode_owner(feats, points, offsets, pc_range, voxel_size, grid_size, split, reduce, clamp_zero, arith, group, capacity,
                  slab_bounds=None):
    import ctypes as C_
    from . import _lib as L
    from . import ops
    if offsets.numel() != 2:
        raise ValueError("strategy 'owner' handles one sample per call (rows of several samples cannot be told apart on arrival)")
    if not feats.is_cuda:
        raise ValueError("strategy 'owner' needs CUDA tensors (NVLink peer memory)")
    dev = feats.device
    n, Cc = feats.shape
    if capacity is None:  # global point count: one small all-reduce + host sync; pass capacity= to avoid it
        t = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(t, group=group)
        capacity = int(t.item())
    ex = _OwnerExchange.get(dev, group, max(int(capacity), 1), Cc)
    world, rank = ex.world, ex.rank
    X, Y, Z = (int(g) for g in grid_size)
    if slab_bounds is None:
        xb = [shard_bounds(X, r, world)[0] for r in range(world)] + [X]
        yb = [shard_bounds(Y, r, world)[0] for r in range(world)] + [Y]
    else:
        xb, yb = [int(v) for v in slab_bounds[0]], [int(v) for v in slab_bounds[1]]
        for bnd, extent in ((xb, X), (yb, Y)):
            if len(bnd) != world + 1 or bnd[0] != 0 or bnd[-1] != extent or any(bnd[i] >= bnd[i + 1] for i in range(world)):
                raise ValueError(f"slab_bounds must be {world + 1} increasing voxel indices from 0 to {extent}, got {bnd}")
    pool = ops.pool_kernels(grid_size, split)
    if min(xb[r + 1] - xb[r] for r in range(world)) < pool[0] or min(yb[r + 1] - yb[r] for r in range(world)) < pool[1]:
        raise ValueError(f"strategy 'owner': every x / y slab must span at least one pooling window {pool[:2]} "
                         f"(got x {xb}, y {yb}): fewer ranks, or balanced_slab_bounds(..., min_width=pool[:2])")
    geom = L.make_geom(pc_range, voxel_size, grid_size, pool)
    stream = torch.cuda.current_stream(dev).cuda_stream
    feats = feats if (feats.stride(1) == 1 and feats.stride(0) % 4 == 0 and feats.data_ptr() % 16 == 0) else feats.contiguous()
    points = points.contiguous()
    with torch.cuda.device(dev):
        ex.head32.fill_(-1)                # index regions = -1 (rows never written are dropped by the encode) ...
        ex.head32[:64].zero_()             # ... counters = 0
        ex.hdl.barrier(channel=0)          # every rank has reset before anybody pushes
        xba, yba = (C_.c_int32 * (world + 1))(*xb), (C_.c_int32 * (world + 1))(*yb)
        L.check(L.lib().tp_route_points_f32(points.data_ptr(), points.shape[1], feats.data_ptr(), feats.stride(0), Cc, n,
                                            C_.byref(geom), ops._ARITH[arith], world, xba, yba, ex.p_cnt, ex.p_idx_x, ex.p_feat_x,
                                            ex.p_idx_y, ex.p_feat_y, ex.cap, stream), "tp_route_points_f32")
        ops.launch_count += 1 if n else 0
        ex.hdl.barrier(channel=1)          # every push has landed before the owners read
    xs, ys = xb[rank + 1] - xb[rank], yb[rank + 1] - yb[rank]
    kw = dict(reduce=reduce, clamp_zero=clamp_zero, arith=arith, pool=pool)
    xy, _, xz = ops.encode(ex.feat_x, ex.offsets, [0] * 6, (1, 1, 1), (xs, Y, Z), split, grid_ind=ex.idx_x, planes=(True, False, True), **kw)
    _, yz, _ = ops.encode(ex.feat_y, ex.offsets, [0] * 6, (1, 1, 1), (X, ys, Z), split, grid_ind=ex.idx_y, planes=(False, True, False), **kw)
    return xy, yz, xz


def encode_point_sharded(feats: torch.Tensor, points: torch.Tensor, offsets: torch.Tensor, pc_range, voxel_size,
                         grid_size, split, reduce: str = "max", clamp_zero: bool = False, arith: str = "cuda",
                         group=None, strategy: str = "planes", capacity: Optional[int] = None, slab_bounds=None):
    """Each rank passes ITS shard of the points (feats [N_r, C], raw points [N_r, >=3], offsets [B+1] of
    the shard). strategy "planes" (all-reduce of partial planes) and "points" (all-gather of the shards, then the full
    encode on every rank) return the complete planes (xy, yz, xz) on every rank; "owner" (push over NVLink peer memory,
    see STRATEGIES) returns this rank's slabs: xy[:, x0:x1], yz[:, y0:y1], xz[:, x0:x1] with (x0, x1) =
    shard_bounds(X, rank, world), (y0, y1) = shard_bounds(Y, rank, world) — or the slabs named by slab_bounds = (xb, yb)
    (balanced_slab_bounds: equal point counts instead of equal widths). capacity (owner): the global point count."""
    from . import ops
    if strategy == "owner":
        return _encode_owner(feats, points[:, :3], offsets, pc_range, voxel_size, grid_size, split, reduce, clamp_zero, arith,
                             group, capacity, slab_bounds)
    if strategy == "points":
        f_all, p_all, off_all = gather_point_shards(feats, points[:, :3].contiguous(), offsets, group)
        return ops.encode(f_all, off_all, pc_range, voxel_size, grid_size, split, points=p_all, reduce=reduce,
                          clamp_zero=clamp_zero, arith=arith)
    if strategy != "planes":
        raise ValueError(f"strategy must be 'planes', 'points' or 'owner', got {strategy!r}")
    if reduce == "max":
        xy, yz, xz = ops.encode(feats, offsets, pc_range, voxel_size, grid_size, split, points=points,
                                reduce="max_partial", arith=arith)
        all_reduce_planes((xy, yz, xz), "max", group=group)
        for p in (xy, yz, xz):
            ops.finalize_max(p, clamp_zero)
        return xy, yz, xz
    if reduce == "mean":
        xy, yz, xz, cnt = ops.encode(feats, offsets, pc_range, voxel_size, grid_size, split, points=points,
                                     reduce="sum", want_counts=True, arith=arith)
        all_reduce_planes((xy, yz, xz), "sum", counts=cnt, group=group)
        C = feats.shape[1]
        o = 0
        for p in (xy, yz, xz):
            n = p.numel() // C
            ops.finalize_mean(p, cnt[o:o + n], C)
            o += n
        return xy, yz, xz
    raise ValueError(f"reduce must be 'max' or 'mean', got {reduce!r}")
