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
}


def planes_equal(got, ref, strategy: str, rank: int, world: int) -> bool:
    """Does this rank's result of encode_point_sharded equal the single-GPU planes `ref`? Both current strategies
    replicate the full planes on every rank."""
    return all(torch.equal(a, b) for a, b in zip(got, ref))


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


def encode_point_sharded(feats: torch.Tensor, points: torch.Tensor, offsets: torch.Tensor, pc_range, voxel_size,
                         grid_size, split, reduce: str = "max", clamp_zero: bool = False, arith: str = "cuda",
                         group=None, strategy: str = "planes"):
    """Each rank passes ITS shard of the points (feats [N_r, C], raw points [N_r, >=3], offsets [B+1] of
    the shard); every rank returns the complete planes (xy, yz, xz). strategy: "planes" (all-reduce of
    partial planes) or "points" (all-gather of the shards, then the full encode on every rank)."""
    from . import ops
    if strategy == "points":
        f_all, p_all, off_all = gather_point_shards(feats, points[:, :3].contiguous(), offsets, group)
        return ops.encode(f_all, off_all, pc_range, voxel_size, grid_size, split, points=p_all, reduce=reduce,
                          clamp_zero=clamp_zero, arith=arith)
    if strategy != "planes":
        raise ValueError(f"strategy must be 'planes' or 'points', got {strategy!r}")
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
