"""Property tests (hypothesis) of the oracle's own invariants — SURVEY §7 step 1: permutation invariance of the
scatter-max encode, zero padding and partition of unity of the bilinear decode, index range of voxelize. CPU only:
these pin the CHECKER; the CUDA path is compared against it in the -m gpu tests."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import triplane_oracle as O

GRID, SPLIT = [16, 12, 8], [4, 4, 2]


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 200), st.integers(0, 2 ** 31 - 1))
def test_encode_is_permutation_invariant_and_bounded(n, seed):
    g = torch.Generator().manual_seed(seed)
    ind = torch.stack([torch.randint(0, GRID[a], (n,), generator=g) for a in range(3)], 1).int()
    feats = torch.randn(n, 4, generator=g)
    perm = torch.randperm(n, generator=g)
    a = O.encode_pooled(feats, O.cat_indices([ind]), GRID, SPLIT, 1)
    b = O.encode_pooled(feats[perm], O.cat_indices([ind[perm]]), GRID, SPLIT, 1)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    # every output value is 0 (empty cell) or one of the inputs, and never exceeds the global maximum
    vals = torch.cat([x.reshape(-1) for x in a[:3]])
    assert float(vals.max()) <= max(float(feats.max()), 0.0)
    assert bool(torch.isin(vals[vals != 0], feats.reshape(-1)).all())
    # counts: every in-grid point is counted once per plane whose pooled extent covers it
    cnt = O.cell_counts(O.cat_indices([ind]), GRID, SPLIT, 1)
    assert all(int(c.sum()) <= n for c in cnt)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_decode_zero_padding_and_partition_of_unity(seed):
    g = torch.Generator().manual_seed(seed)
    lo, vs = [-4.0, -4.0, -1.0], (0.5, 0.5, 0.25)
    ones = torch.ones(1, 3, 2, 16, 16)
    far = torch.rand(1, 1, 50, 3, generator=g) * 100 + 50  # far outside every plane
    assert float(O.sample_points_triplane_stacked(ones, far, lo, vs).abs().max()) == 0
    # interior of all three planes: weights of each bilinear footprint sum to one
    inner = torch.rand(1, 1, 50, 3, generator=g) * torch.tensor([6.0, 6.0, 3.0]) + torch.tensor([-3.0, -3.0, -0.5])
    out = O.sample_points_triplane_stacked(ones, inner, lo, vs)
    assert float((out - 3).abs().max()) < 1e-5
    # linear in the planes
    t1, t2 = torch.randn(1, 3, 2, 16, 16, generator=g), torch.randn(1, 3, 2, 16, 16, generator=g)
    s = O.sample_points_triplane_stacked(t1 + t2, inner, lo, vs)
    assert torch.allclose(s, O.sample_points_triplane_stacked(t1, inner, lo, vs) +
                          O.sample_points_triplane_stacked(t2, inner, lo, vs), atol=1e-5)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 300), st.integers(0, 2 ** 31 - 1))
def test_voxelize_indices_in_range_and_order_kept(n, seed):
    g = torch.Generator().manual_seed(seed)
    rng, vs = [-25.0, -25.0, -5.0, 25.0, 25.0, 3.0], (0.4, 0.4, 0.1)
    pts = (torch.rand(n, 5, generator=g) - 0.5) * torch.tensor([70.0, 70.0, 12.0, 1.0, 1.0])
    cropped, ind = O.voxelize_points([pts], rng, vs)
    c, i = cropped[0], ind[0]
    assert i.dtype == torch.int32 and c.shape[0] == i.shape[0]
    if c.shape[0]:
        assert int(i.min()) >= 0 and int(i[:, 0].max()) <= 125 and int(i[:, 1].max()) <= 125 and int(i[:, 2].max()) <= 80
        # stable: the kept rows appear in their original order
        keep = ((pts[:, 0] > rng[0]) & (pts[:, 0] < rng[3]) & (pts[:, 1] > rng[1]) & (pts[:, 1] < rng[4]) &
                (pts[:, 2] > rng[2]) & (pts[:, 2] < rng[5]))
        assert torch.equal(c, pts[keep])
