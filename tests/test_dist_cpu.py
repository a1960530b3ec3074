"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (sharding + the partial-plane exchange).
The per-rank scatter is done by the oracle here; the GPU version of the same exchange is
tests/multigpu_check.py (torchrun, NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from efficient_multimodal_perception_b200 import dist as tpd
from efficient_multimodal_perception_b200 import synth
from oracle import triplane_oracle as O

GRID, SPLIT, C = [16, 16, 8], [4, 4, 2], 8


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 34720, 350001):
        for world in (1, 2, 3, 8):
            cuts = [tpd.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    q = torch.arange(2 * 11 * 3).view(2, 11, 3)
    parts = [tpd.shard_queries(q, r, 3) for r in range(3)]
    assert torch.equal(torch.cat(parts, 1), q)
    assert tpd.shard_samples(list(range(5)), 1, 2) == [3, 4]


def _case():
    g = torch.Generator().manual_seed(77)
    inds, feats = [], []
    for n in (900, 0, 700):
        inds.append(torch.stack([torch.randint(0, GRID[a], (n,), generator=g) for a in range(3)], 1).int())
        feats.append(torch.randn(n, C, generator=g) - 0.5)  # plenty of all-negative cells
    return inds, feats


def _partial(feats, inds, reduce):
    """What ops.encode(reduce='max_partial' | 'sum') produces for one rank's shard."""
    cat = O.cat_indices(inds)
    f = torch.cat(feats)
    cnts = O.cell_counts(cat, GRID, SPLIT, len(inds))
    if reduce == "max":
        xy, yz, xz, _, _ = O.encode_pooled(f, cat, GRID, SPLIT, len(inds))
        outs = []
        for p, c in zip((xy, yz, xz), cnts):
            p = p.reshape(-1, C).clone()
            p[c == 0] = float("-inf")
            outs.append(p)
        return outs, None
    xy, yz, xz, _, _ = O.encode_pooled(f.double(), cat, GRID, SPLIT, len(inds), reduce="mean")
    outs = [(p.reshape(-1, C) * c.view(-1, 1)).float() for p, c in zip((xy, yz, xz), cnts)]  # mean * count = sum
    return outs, torch.cat(cnts)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inds, feats = _case()
        my_i, my_f = tpd.shard_points(inds, rank, world), tpd.shard_points(feats, rank, world)
        full = O.encode_pooled(torch.cat(feats), O.cat_indices(inds), GRID, SPLIT, len(inds))
        # max: partial planes with -inf empties -> all-reduce MAX -> -inf to 0: bit-exact
        parts, _ = _partial(my_f, my_i, "max")
        tpd.all_reduce_planes(parts, "max")
        ok = True
        for p, ref in zip(parts, full[:3]):
            p = torch.where(torch.isinf(p) & (p < 0), torch.zeros_like(p), p)
            ok &= torch.equal(p, ref.reshape(-1, C))
        # mean: partial sums + counts -> all-reduce SUM -> divide
        parts, cnt = _partial(my_f, my_i, "mean")
        tpd.all_reduce_planes(parts, "sum", counts=cnt)
        refm = O.encode_pooled(torch.cat(feats).double(), O.cat_indices(inds), GRID, SPLIT, len(inds), reduce="mean")
        o = 0
        for p, ref in zip(parts, refm[:3]):
            n = p.shape[0]
            m = p / cnt[o:o + n].clamp(min=1).view(-1, 1)
            o += n
            ok &= float((m.double() - ref.reshape(-1, C)).abs().max()) <= 1e-5 * float(ref.abs().max())
        ok &= int(cnt.sum()) == 3 * 1600  # every point counted once per plane (grid divisible: nothing dropped)
        # "points" strategy: all-gather of the shards -> every rank holds the whole batch, sample-major
        my_off = synth.batch_offsets([f.shape[0] for f in my_f])
        f_all, i_all, off_all = tpd.gather_point_shards(torch.cat(my_f), torch.cat(my_i).float(), my_off)
        ok &= off_all.tolist() == synth.batch_offsets([f.shape[0] for f in feats]).tolist()
        ok &= torch.equal(f_all, torch.cat(feats)) and torch.equal(i_all.int(), torch.cat(inds))
        g_inds = [i_all[off_all[b]:off_all[b + 1]].int() for b in range(len(inds))]
        g = O.encode_pooled(f_all, O.cat_indices(g_inds), GRID, SPLIT, len(inds))
        ok &= all(torch.equal(a, b) for a, b in zip(g[:3], full[:3]))
        # owner-strategy slab boundaries by point count: the same on every rank, increasing, covering the grid, and the
        # x-slabs of a centre-heavy cloud hold about the same number of points
        gen = torch.Generator().manual_seed(5)
        cloud = torch.randn(4000, 3, generator=gen) * torch.tensor([6.0, 9.0, 1.0])
        lo_r, hi_r = tpd.shard_bounds(cloud.shape[0], rank, world)
        rng, vs, grid = [-25.0, -25.0, -5.0, 25.0, 25.0, 3.0], (0.4, 0.4, 0.1), [128, 128, 80]
        xb, yb = tpd.balanced_slab_bounds(cloud[lo_r:hi_r], rng, vs, grid)
        both = [None, None]
        dist.all_gather_object(both, (xb, yb))
        ok &= both[0] == both[1]
        for b in (xb, yb):
            ok &= len(b) == world + 1 and b[0] == 0 and b[-1] == 128 and all(b[i] < b[i + 1] for i in range(world))
        ix = ((cloud[:, 0] + 25.0) / 0.4).long()
        inside = (cloud[:, 0].abs() < 25) & (cloud[:, 1].abs() < 25) & (cloud[:, 2] > -5) & (cloud[:, 2] < 3)
        left = int(((ix < xb[1]) & inside).sum())
        ok &= abs(left - int(inside.sum()) / 2) < 0.1 * int(inside.sum())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_point_sharded_exchange_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) is True and ret.get(1) is True
