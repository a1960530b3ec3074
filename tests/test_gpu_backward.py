"""GPU parity of the backward kernels against what autograd produces for the
reference's op sequence: torch-CUDA grid_sample backward for decode / lift, and the CPU oracle's
scatter-max / max-pool autograd for encode. All through the C ABI via the autograd.Function wrappers
the drop-in modules use when gradients are required."""
import pytest
import torch

from conftest import normwise
from efficient_multimodal_perception_b200 import (PointTriplaneProjector, point_to_cam, sample_points_triplane, synth)
from efficient_multimodal_perception_b200.autograd import encode_max_autograd, encode_mean_autograd
from oracle import triplane_oracle as O
from test_gpu_parity import _torch_cuda_sample

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def cu(t):
    return t.to(DEV)


@pytest.mark.parametrize("kind", ["stacked4d", "stacked5d_lattice", "list4d"])
def test_sample_backward_vs_torch_cuda_autograd(kind):
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    B, C = 2, 32
    if kind == "list4d":
        grid = [128, 128, 80]
        planes = [cu(p).requires_grad_() for p in synth.triplane_list(B, 96, grid, seed=3)]
        half, arg, C = [g / 2 for g in grid], planes, 96
        pts = cu(torch.stack([synth.uniform_queries(4000, seed=s) for s in (1, 2)]))[:, :, None] * 1.05
        pts = pts.reshape(B, 50, 80, 3)
    else:
        tri = cu(synth.triplane_stacked(B, C, 128, seed=4)).requires_grad_()
        planes, half, arg, grid = [tri[:, 0], tri[:, 1], tri[:, 2]], [64.0] * 3, tri, None
        if kind == "stacked4d":
            pts = cu(synth.range_image_points(B, seed=7)[:, :, ::8])
        else:
            pts = cu(synth.lattice((40, 36, 16), (0.5, 0.5, 0.5), (-12.0, -30.0, -5.0)))[None].repeat(B, 1, 1, 1, 1)
    out = sample_points_triplane(arg, pts, lo, vs, grid_size=grid)
    assert out.requires_grad
    wgt = torch.randn(out.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(5))
    leaves = planes if kind == "list4d" else [arg]
    got = torch.autograd.grad((out * wgt).sum(), leaves)
    ref_out = _torch_cuda_sample(planes, pts.reshape(B, -1, 3), lo, vs, half)
    ref = torch.autograd.grad((ref_out * wgt.reshape(ref_out.shape)).sum(), leaves)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and normwise(a, b) <= TOL
        assert float(b.abs().max()) > 0


@pytest.mark.parametrize("grid,split,C,n_per", [
    ([16, 16, 8], [4, 4, 2], 8, (300, 200)),
    ([13, 13, 9], [4, 4, 2], 12, (400, 0, 250)),
    ([128, 128, 80], [25, 25, 20], 128, (20000,)),
])
def test_encode_backward_vs_oracle_autograd(grid, split, C, n_per):
    g = torch.Generator().manual_seed(sum(n_per))
    inds = [torch.stack([torch.randint(0, grid[a], (n,), generator=g) for a in range(3)], 1).int() for n in n_per]
    feats = torch.randn(sum(n_per), C, generator=g)
    off = cu(synth.batch_offsets(n_per))
    cat = cu(torch.cat(inds))
    for mode in ("max", "mean"):
        f_ref = feats.clone().requires_grad_()
        ref_out = O.encode_pooled(f_ref, O.cat_indices(inds), grid, split, len(n_per), reduce=mode)[:3]
        ws = [torch.randn(o.shape, generator=g) for o in ref_out]
        (g_ref,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(ref_out, ws)), [f_ref])
        f = cu(feats).requires_grad_()
        fn = encode_max_autograd if mode == "max" else encode_mean_autograd
        outs = fn(f, cat, off, grid, split)
        (g_got,) = torch.autograd.grad(sum((o * cu(w)).sum() for o, w in zip(outs, ws)), [f])
        if mode == "max":  # routing is exact: same adds of the same numbers in plane order xy, yz, xz
            assert normwise(g_got.cpu(), g_ref) <= 1e-6
            assert torch.equal(g_got.cpu() != 0, g_ref != 0)
        else:
            assert normwise(g_got.cpu(), g_ref) <= TOL


def test_encode_backward_ties_documented_behaviour():
    """Exact ties (duplicated rows in one voxel: repeated returns in accumulated sweeps). torch_scatter.scatter_max
    routes a voxel's gradient to ONE arg-max point (which one is unspecified on CUDA), spconv's pool backward fans out
    to every tied voxel; this library gives every point that attains the pooled maximum the cell's gradient (documented
    in include/triplane.h). The sum over the tied points is therefore multiplicity x the reference's: asserted here so
    a change of the rule is a conscious one. Untied points are unaffected."""
    grid, split = [16, 16, 8], [4, 4, 2]
    g = torch.Generator().manual_seed(11)
    ind = torch.stack([torch.randint(0, grid[a], (300,), generator=g) for a in range(3)], 1).int()
    f = torch.randn(300, 8, generator=g)
    ind2, f2 = torch.cat([ind, ind[:40]]), torch.cat([f, f[:40]])     # rows 300..339 duplicate rows 0..39
    fa, fb = cu(f).requires_grad_(), cu(f2).requires_grad_()
    oa = encode_max_autograd(fa, cu(ind), cu(synth.batch_offsets([300])), grid, split)
    ob = encode_max_autograd(fb, cu(ind2), cu(synth.batch_offsets([340])), grid, split)
    for x, y in zip(oa, ob):
        assert torch.equal(x, y)                                      # duplicates do not change the forward
    ws = [torch.randn(o.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(12)) for o in oa]
    (ga,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(oa, ws)), [fa])
    (gb,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(ob, ws)), [fb])
    assert torch.equal(gb[:300], ga) and torch.equal(gb[300:], ga[:40])  # each copy receives the full cell gradient


def test_encode_backward_clamp_zero_blocks_negative_maxima():
    grid, split = [16, 16, 8], [4, 4, 2]
    g = torch.Generator().manual_seed(3)
    ind = cu(torch.stack([torch.randint(0, grid[a], (600,), generator=g) for a in range(3)], 1).int())
    f = cu(torch.randn(600, 8, generator=g)).requires_grad_()
    off = cu(synth.batch_offsets([600]))
    outs = encode_max_autograd(f, ind, off, grid, split, True)
    (gr,) = torch.autograd.grad(sum(o.sum() for o in outs), [f])
    assert float(gr[f.detach() <= 0].abs().max()) == 0 and float(gr[f.detach() > 0].abs().max()) > 0


def test_lift_backward_vs_torch_cuda_autograd():
    rig = synth.camera_rig(1004)
    metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug),
             dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(),
                  imgs_aug=[dict(a, flip=(i % 2 == 0)) for i, a in enumerate(rig.imgs_aug)])]
    pts = [cu(synth.lidar_sweep(6000, seed=50)[:, :5]), cu(synth.lidar_sweep(2500, seed=51)[:, :5])]
    feats = cu(torch.randn(2, 6, 64, 16, 32, generator=torch.Generator().manual_seed(52))).requires_grad_()
    out = point_to_cam(pts, feats, metas)
    ws = [torch.randn(o.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(6)) for o in out]
    (got,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(out, ws)), [feats])
    ref_out = O.point_to_cam(pts, feats, metas)
    (ref,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(ref_out, ws)), [feats])
    assert got.shape == ref.shape and normwise(got, ref) <= TOL and float(ref.abs().max()) > 0


def test_projector_module_trains():
    """The registered module with gradients on: loss.backward() reaches point_mlp / reduce_cam_channels /
    mlp_* through the CUDA encode, and the feature gradient equals the oracle pipeline's."""
    torch.manual_seed(0)
    grid, split, C = [16, 16, 8], [4, 4, 2], 8
    m = PointTriplaneProjector(grid, in_channels=5, out_channels=C, base_channels=C, split=split).to(DEV).train()
    g = torch.Generator().manual_seed(1)
    pts = [cu(torch.randn(n, 11, generator=g)) for n in (300, 200)]
    ind = [cu(torch.stack([torch.randint(0, grid[a], (n,), generator=g) for a in range(3)], 1).int()) for n in (300, 200)]
    cam = [cu(torch.randn(n, 768, generator=g)) for n in (300, 200)]
    planes = m(pts, ind, cam)
    loss = sum(p.square().mean() for p in planes)
    loss.backward()
    for name, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert float(m.point_mlp[1].weight.grad.abs().max()) > 0 and float(m.reduce_cam_channels.weight.grad.abs().max()) > 0
    # the same loss through the CPU oracle's scatter path, sharing the module's weights
    mc = PointTriplaneProjector(grid, in_channels=5, out_channels=C, base_channels=C, split=split).train()
    mc.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    feats = mc.point_features([p.cpu() for p in pts], [c.cpu() for c in cam])
    xy, yz, xz = O.encode_pooled(feats, O.cat_indices([i.cpu() for i in ind]), grid, split, 2)[:3]
    ref_planes = [mc.mlp_xy(xy).permute(0, 3, 1, 2), mc.mlp_yz(yz).permute(0, 3, 1, 2), mc.mlp_xz(xz).permute(0, 3, 1, 2)]
    sum(p.square().mean() for p in ref_planes).backward()
    for (name, p), (_, pc) in zip(m.named_parameters(), mc.named_parameters()):
        if float(pc.grad.abs().max()) < 1e-7:  # a bias feeding a BatchNorm: the gradient is rounding noise
            assert float(p.grad.abs().max()) < 1e-6, name
            continue
        assert normwise(p.grad.cpu(), pc.grad) <= 1e-3, name  # cuBLAS vs CPU GEMMs + BatchNorm statistics


@pytest.mark.parametrize("seed", list(range(8)))
def test_sample_backward_lattice_kernel_matches_per_query_kernel(seed):
    """tp_sample3_grid_backward_nhwc_f32 (sum over the free lattice index, then one scatter per index pair) vs the
    per-query scatter kernel on random lattices / plane overlaps / shapes, incl. jittered (non-lattice) blocks:
    same gradient planes to fp32 rounding (the association of the sums differs)."""
    import numpy as np
    from efficient_multimodal_perception_b200 import ops
    rs = np.random.RandomState(500 + seed)
    g = torch.Generator().manual_seed(600 + seed)
    B = int(rs.randint(1, 3))
    C_ = int(rs.choice([4, 32, 96]))
    h, w = int(rs.randint(1, 30)), int(rs.randint(1, 30))
    d = int(rs.choice([4, 8, 16, 20]))
    shapes = [(int(rs.choice([16, 128])), int(rs.choice([16, 80, 128]))) for _ in range(3)]
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    half = [float(rs.choice([8.0, 40.0, 64.0])) for _ in range(3)]
    origin, step = [], []
    for a in range(3):
        span = 2 * half[a] * vs[a]
        n = (h, w, d)[a]
        mode = rs.randint(0, 3)
        o, s = [(lo[a] + 0.1 * span, 0.7 * span / n), (lo[a] + 1.5 * span, 0.5 * span / n), (lo[a] - 0.5 * span, 2.0 * span / n)][mode]
        origin.append(float(o)); step.append(float(s))
    q = synth.lattice((h, w, d), step, origin).unsqueeze(0).repeat(B, 1, 1, 1, 1)
    if seed % 4 == 1:
        q = q + 0.03 * torch.randn(q.shape, generator=g)        # nothing is a lattice
    if seed % 4 == 2:
        q[:, : max(1, h // 2)] += 0.03 * torch.randn(q[:, : max(1, h // 2)].shape, generator=g)  # some blocks are
    qd = cu(q.reshape(B, -1, 3))
    gout = cu(torch.randn(B, C_, h * w * d, generator=g))
    arith = "cpu" if seed % 2 else "cuda"
    flat = ops.sample3_backward(gout, qd, shapes, lo, vs, half, arith=arith)
    grid = ops.sample3_backward(gout, qd, shapes, lo, vs, half, arith=arith, grid_dims=(h, w, d))
    for a, b in zip(grid, flat):
        assert a.shape == b.shape
        if float(b.abs().max()) > 0:
            assert normwise(a, b) <= TOL
        else:
            assert float(a.abs().max()) == 0
