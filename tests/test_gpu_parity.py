"""GPU parity: libtriplane (through the C ABI) vs the reference-generated goldens, the CPU oracle and
torch-CUDA's own ops. Bars: bit-exact for crop mask, indices, counts, scatter-max;
normwise max|a-b| <= 1e-5 * max|b| for bilinear sampling and scatter-mean."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, normwise
from efficient_multimodal_perception_b200 import (PointTriplaneProjector, ops, point_to_cam, sample_points_triplane,
                                                  synth, voxelize_points)
from oracle import triplane_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5  # normwise, north star "within 1e-5 relative for fp32 scatter-mean and bilinear sampling"


def cu(t):
    return t.to(DEV)


# ---------------------------------------------------------------------------------------------
# a1 voxelize
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["voxelize_A", "voxelize_B"])
def test_voxelize_golden_bit_exact_cpu_arith(name):
    """arith='cpu' (true division) must reproduce the reference run on torch-CPU bit for bit,
    including the rows on / one ulp around the crop bounds, NaN and inf rows and an empty sample."""
    g = load_golden(name)
    pts = [cu(g.t(f"points{i}")) for i in range(3)]
    cropped, ind = voxelize_points(pts, g["pc_range"].tolist(), g["voxel_size"].tolist(), arith="cpu")
    for i in range(3):
        assert torch.equal(cropped[i].cpu(), g.t(f"cropped{i}"))
        assert torch.equal(ind[i].cpu(), g.t(f"ind{i}"))
        assert ind[i].dtype == torch.int32


def test_voxelize_matches_torch_cuda_bit_exact():
    """arith='cuda' vs the reference's expression evaluated by torch on the same GPU (multiply by the
    fp32 reciprocal); 350k points = the S5 sample size."""
    G = synth.GEOM_A
    pts = cu(synth.multi_sweep(10, 35000, seed=1005))
    rng, vs = G["pc_range"], G["voxel_size"]
    mask = ((pts[:, 0] > rng[0]) & (pts[:, 0] < rng[3]) & (pts[:, 1] > rng[1]) & (pts[:, 1] < rng[4])
            & (pts[:, 2] > rng[2]) & (pts[:, 2] < rng[5]))
    ref_pts = pts[mask]
    vi = torch.zeros((ref_pts.shape[0], 3), device=DEV)
    for a in range(3):
        vi[:, a] = (ref_pts[:, a] - rng[a]) / vs[a]
    ref_ind = vi.type(torch.int)
    cropped, ind = voxelize_points([pts], rng, vs, arith="cuda")
    assert torch.equal(cropped[0], ref_pts)
    assert torch.equal(ind[0], ref_ind)
    keep, idx = ops.voxel_index(pts, rng, vs, arith="cuda")
    assert torch.equal(keep.bool(), mask)
    assert torch.equal(idx[mask], ref_ind)


def test_voxelize_config_sweep_and_batch_offsets():
    g = load_golden("voxelize_S1")
    pts = synth.lidar_sweep(int(g["n"]), seed=int(g["seed"]))
    batch = [cu(pts), cu(pts[:0]), cu(pts[5000:9000]), cu(pts[:1])]
    cropped, ind = voxelize_points(batch, g["pc_range"].tolist(), g["voxel_size"].tolist(), arith="cpu")
    assert cropped[0].shape[0] == int(g["n_kept"])
    assert torch.equal(ind[0].cpu().to(torch.int16), g.t("ind"))
    assert torch.equal(cropped[0].cpu().double().sum(0), g.t("cropped_checksum"))
    assert cropped[1].shape[0] == 0
    ref = O.voxelize_points([pts[5000:9000], pts[:1]], g["pc_range"].tolist(), g["voxel_size"].tolist())
    assert torch.equal(cropped[2].cpu(), ref[0][0]) and torch.equal(ind[2].cpu(), ref[1][0])
    assert torch.equal(cropped[3].cpu(), ref[0][1])


# ---------------------------------------------------------------------------------------------
# a3 encode
# ---------------------------------------------------------------------------------------------
def _rand_encode_case(n_per, grid, C, seed, hot=False):
    g = torch.Generator().manual_seed(seed)
    inds, feats = [], []
    for n in n_per:
        if hot:  # many points per cell: long lists
            ind = torch.stack([torch.randint(0, max(1, grid[a] // 8), (n,), generator=g) for a in range(3)], 1)
        else:
            ind = torch.stack([torch.randint(0, grid[a], (n,), generator=g) for a in range(3)], 1)
        inds.append(ind.int())
        feats.append(torch.randn(n, C, generator=g))
    return inds, torch.cat(feats)


@pytest.mark.parametrize("grid,split,C,n_per,hot", [
    ([16, 16, 8], [4, 4, 2], 8, (300, 200), False),
    ([13, 13, 9], [4, 4, 2], 12, (400, 0, 250), False),      # ragged: idx 12 falls off the pooled extent
    ([128, 128, 80], [25, 25, 20], 128, (28000,), False),    # geometry A, one sweep
    ([200, 200, 16], [25, 25, 16], 128, (9000, 11000), False),  # geometry B
    ([32, 32, 16], [4, 4, 4], 32, (20000,), True),           # ~150 points per cell
    ([16, 16, 8], [4, 4, 2], 132, (500,), False),            # C not a multiple of 128: 2 float4 per lane
])
def test_encode_max_bit_exact_vs_oracle(grid, split, C, n_per, hot):
    inds, feats = _rand_encode_case(n_per, grid, C, seed=sum(n_per) + C, hot=hot)
    B = len(n_per)
    ref = O.encode_pooled(feats, O.cat_indices(inds), grid, split, B)
    off = cu(synth.batch_offsets(n_per))
    for rep in range(2):  # second call proves the head table was left clean
        xy, yz, xz, cnt = ops.encode(cu(feats), off, [0] * 6, (1, 1, 1), grid, split,
                                     grid_ind=cu(torch.cat(inds)), want_counts=True)
        assert torch.equal(xy.cpu(), ref[0]) and torch.equal(yz.cpu(), ref[1]) and torch.equal(xz.cpu(), ref[2])
        ref_cnt = torch.cat(O.cell_counts(O.cat_indices(inds), grid, split, B))
        assert torch.equal(cnt.cpu(), ref_cnt)


def test_encode_negative_features_and_clamp_zero():
    grid, split = [16, 16, 8], [4, 4, 2]
    inds, feats = _rand_encode_case((700,), grid, 16, seed=5)
    feats = -feats.abs() - 0.5
    off = cu(synth.batch_offsets([700]))
    ref = O.encode_pooled(feats, O.cat_indices(inds), grid, split, 1)
    out = ops.encode(cu(feats), off, [0] * 6, (1, 1, 1), grid, split, grid_ind=cu(inds[0]))
    assert torch.equal(out[0].cpu(), ref[0]) and float(out[0].max()) <= 0 and float(out[0].min()) < 0
    out = ops.encode(cu(feats), off, [0] * 6, (1, 1, 1), grid, split, grid_ind=cu(inds[0]), clamp_zero=True)
    assert float(out[0].abs().max()) == 0 and float(out[1].abs().max()) == 0


def test_encode_out_of_grid_indices_are_dropped():
    """z index 80 (one ulp below the upper bound) is outside spatial_shape: dropped."""
    grid, split = [128, 128, 80], [25, 25, 20]
    ind = torch.tensor([[3, 4, 80], [3, 4, 79], [125, 4, 5], [-1, 0, 0], [127, 127, 0]], dtype=torch.int32)
    feats = torch.arange(5 * 8, dtype=torch.float32).view(5, 8) + 1
    ref = O.encode_pooled(feats, O.cat_indices([ind]), grid, split, 1)
    out = ops.encode(cu(feats), cu(synth.batch_offsets([5])), [0] * 6, (1, 1, 1), grid, split, grid_ind=cu(ind))
    for a, b in zip(out, ref[:3]):
        assert torch.equal(a.cpu(), b)
    # x index 125..127: kept by pool_xy, outside the 25 pooled cells of pool_yz
    assert float(out[0][0, 125, 4].abs().sum()) > 0 and float(out[0][0, 127, 127].abs().sum()) > 0
    assert float(out[1].abs().sum()) == float(ref[1].abs().sum())


def test_encode_fused_crop_index_equals_two_step():
    """idx == NULL path: crop + voxel index inside the link kernel == voxelize_points then encode."""
    G = synth.GEOM_A
    raw = [synth.lidar_sweep(20000, seed=71), synth.lidar_sweep(15000, seed=72)]
    feats = torch.randn(35000, 32, generator=torch.Generator().manual_seed(73))
    off = cu(synth.batch_offsets([20000, 15000]))
    for arith in ("cuda", "cpu"):
        fused = ops.encode(cu(feats), off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"],
                           points=cu(torch.cat(raw)[:, :3].contiguous()), arith=arith)
        cropped, ind = voxelize_points([cu(r) for r in raw], G["pc_range"], G["voxel_size"], arith=arith)
        keep, _ = ops.voxel_index(cu(torch.cat(raw)), G["pc_range"], G["voxel_size"], arith=arith)
        f2 = cu(feats)[keep.bool()]
        off2 = cu(synth.batch_offsets([c.shape[0] for c in cropped]))
        two = ops.encode(f2, off2, [0] * 6, (1, 1, 1), G["grid_size"], G["split"], grid_ind=torch.cat(ind))
        for a, b in zip(fused, two):
            assert torch.equal(a, b)
    # and against the CPU oracle end to end (arith='cpu' == torch-CPU indices)
    cr, gi = O.voxelize_points(raw, G["pc_range"], G["voxel_size"])
    mask = torch.cat([(r[:, 0] > -25) & (r[:, 0] < 25) & (r[:, 1] > -25) & (r[:, 1] < 25) & (r[:, 2] > -5) & (r[:, 2] < 3)
                      for r in raw])
    ref = O.encode_pooled(feats[mask], O.cat_indices(gi), G["grid_size"], G["split"], 2)
    for a, b in zip(fused, ref[:3]):
        assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("hot", [False, True])
def test_encode_mean_within_tolerance(hot):
    grid, split, C = [32, 32, 16], [4, 4, 4], 32
    inds, feats = _rand_encode_case((6000, 5000), grid, C, seed=17, hot=hot)
    ref = O.encode_pooled(feats.double(), O.cat_indices(inds), grid, split, 2, reduce="mean")
    off = cu(synth.batch_offsets([6000, 5000]))
    out = ops.encode(cu(feats), off, [0] * 6, (1, 1, 1), grid, split, grid_ind=cu(torch.cat(inds)), reduce="mean")
    for a, b in zip(out, ref[:3]):
        assert normwise(a.cpu(), b) <= TOL
    # SUM + counts + finalize == MEAN (the point-sharded multi-GPU path)
    xy, yz, xz, cnt = ops.encode(cu(feats), off, [0] * 6, (1, 1, 1), grid, split, grid_ind=cu(torch.cat(inds)),
                                 reduce="sum", want_counts=True)
    n0, n1 = xy.numel() // C, yz.numel() // C
    ops.finalize_mean(xy, cnt[:n0], C)
    ops.finalize_mean(yz, cnt[n0:n0 + n1], C)
    ops.finalize_mean(xz, cnt[n0 + n1:], C)
    for a, b in zip((xy, yz, xz), out):
        # sums are accumulated with shared-memory atomics: the order of a cell's points is not fixed,
        # so two runs agree to rounding (like torch's own CUDA scatter_reduce / index_add), not bitwise
        assert normwise(a, b) <= TOL


def test_voxel_counts_match_unique():
    grid = [16, 16, 8]
    inds, _ = _rand_encode_case((3000, 2000), grid, 4, seed=23)
    cat = O.cat_indices(inds)
    unq, cnt = torch.unique(cat, return_counts=True, dim=0)
    dense = ops.voxel_counts(cu(torch.cat(inds)), cu(synth.batch_offsets([3000, 2000])), grid).cpu()
    assert int(dense.sum()) == 5000
    got = dense[unq[:, 0].long(), unq[:, 1].long(), unq[:, 2].long(), unq[:, 3].long()]
    assert torch.equal(got.long(), cnt) and int((dense > 0).sum()) == unq.shape[0]


@pytest.mark.parametrize("name", ["projector_small", "projector_ragged"])
def test_projector_forward_vs_reference_golden(name):
    """The drop-in module with the reference's state_dict vs the reference class's forward."""
    g = load_golden(name)
    ns = int(g["nsamples"])
    m = PointTriplaneProjector(g["grid"].tolist(), in_channels=5, out_channels=int(g["C"]),
                               base_channels=int(g["C"]), split=g["split"].tolist())
    m.load_state_dict({k[3:]: g.t(k) for k in g if k.startswith("sd.")})
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m([cu(g.t(f"points{i}")) for i in range(ns)], [cu(g.t(f"ind{i}")) for i in range(ns)],
                [cu(g.t(f"cam{i}")) for i in range(ns)])
    # the Linear layers run in cuBLAS here and MKL in the golden: tolerance, not bits
    for a, key in zip(out, ("tpv_xy", "tpv_yz", "tpv_xz")):
        assert a.shape == g[key].shape
        assert normwise(a.cpu(), g.t(key)) < 1e-4
    # the kernel itself on the golden's features: bit-exact vs the oracle's dense tensors
    feats = g.t("feats")
    inds = [g.t(f"ind{i}") for i in range(ns)]
    ref = O.encode_pooled(feats, O.cat_indices(inds), g["grid"].tolist(), g["split"].tolist(), ns)
    got = ops.encode(cu(feats), cu(synth.batch_offsets([i.shape[0] for i in inds])), [0] * 6, (1, 1, 1),
                     g["grid"].tolist(), g["split"].tolist(), grid_ind=cu(torch.cat(inds)))
    for a, b in zip(got, ref[:3]):
        assert torch.equal(a.cpu(), b)


# ---------------------------------------------------------------------------------------------
# a2 camera -> point lift
# ---------------------------------------------------------------------------------------------
def _lift_metas(g, flips):
    return [dict(img_shape=tuple(int(v) for v in g["img_shape"]), lidar2image=g["lidar2image"],
                 imgs_aug=[dict(resize=float(r), crop=tuple(int(c) for c in cr), flip=bool(f))
                           for r, cr, f in zip(g["resize"], g["crop"], fl)]) for fl in flips]


def _seen_mask(pts, metas, b):
    """which points the oracle's projection puts inside at least one image, with a margin: points within
    1e-3 px of an image border may legitimately flip between devices (einsum rounding)."""
    H, W = metas[0]["img_shape"][::-1]
    l2i = torch.as_tensor(metas[b]["lidar2image"], dtype=torch.float64)
    hom = torch.cat([pts[:, :3].double(), torch.ones(pts.shape[0], 1, dtype=torch.float64)], 1)
    cam = torch.einsum("cij,hj->chi", l2i, hom)
    xy = cam[..., :2] / torch.clamp(cam[..., 2:3], min=1e-5)
    border = torch.zeros(pts.shape[0], dtype=torch.bool)
    for c, aug in enumerate(metas[b]["imgs_aug"]):
        x = xy[c, :, 0] * aug["resize"] - aug["crop"][0]
        y = xy[c, :, 1] * aug["resize"] - aug["crop"][1]
        if aug["flip"]:
            x = W - x
        for v, lim in ((x, W), (y, H)):
            border |= (v.abs() < 1e-3) | ((v - lim).abs() < 1e-3)
    return ~border


def test_lift_vs_reference_golden():
    """point_to_cam executed from the reference's own source on torch-CPU (tests/golden/make_golden.py)
    vs the fused kernel with arith='cpu'. The projection is a K=4 matmul whose accumulation order is
    library-defined, so this is a tolerance test: normwise 1e-5 on points away from image borders."""
    g = load_golden("lift")
    metas = _lift_metas(g, [g["flip0"], g["flip1"]])
    pts = [g.t("points0"), g.t("points1")]
    out = point_to_cam([cu(p) for p in pts], cu(g.t("img_features")), metas, arith="cpu")
    for b in range(2):
        ref = g.t(f"out{b}")
        assert out[b].shape == ref.shape
        keep = _seen_mask(pts[b], metas, b)
        assert int(keep.sum()) > 0.95 * keep.numel()
        assert normwise(out[b].cpu()[keep], ref[keep]) <= TOL
        seen_ref, seen_out = ref.abs().sum(1) > 0, out[b].cpu().abs().sum(1) > 0
        assert torch.equal(seen_ref[keep], seen_out[keep]) and int(seen_ref.sum()) > 20


@pytest.mark.parametrize("Cf,n", [(768, 30000), (32, 5000), (260, 3000)])
def test_lift_vs_torch_cuda_chain(Cf, n):
    """Config-size lift (6 cameras, 16x32 maps, 768 channels, one sweep) vs the reference's op sequence
    run by torch on the same GPU (oracle.point_to_cam moved to CUDA tensors)."""
    rig = synth.camera_rig(1004)
    metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug),
             dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(),
                  imgs_aug=[dict(a, flip=(i % 2 == 0)) for i, a in enumerate(rig.imgs_aug)])]
    pts = [synth.lidar_sweep(n, seed=50)[:, :5], synth.lidar_sweep(n // 3, seed=51)[:, :5]]
    feats = torch.randn(2, 6, Cf, 16, 32, generator=torch.Generator().manual_seed(52))
    ref = O.point_to_cam([cu(p) for p in pts], cu(feats), metas)  # torch-CUDA executes the reference chain
    out = point_to_cam([cu(p) for p in pts], cu(feats), metas, arith="cuda")
    for b in range(2):
        keep = cu(_seen_mask(pts[b], metas, b))
        assert normwise(out[b][keep], ref[b][keep]) <= TOL
        same = float((out[b][keep] == ref[b][keep]).float().mean())
        print(f"\n[lift Cf={Cf} b={b}] bitwise-equal to the torch-CUDA chain: {same:.6f}; "
              f"points seen by a camera: {int((ref[b].abs().sum(1) > 0).sum())}/{ref[b].shape[0]}")
        assert same > 0.99
        assert torch.equal((out[b].abs().sum(1) > 0)[keep], (ref[b].abs().sum(1) > 0)[keep])


def test_lift_edge_cases():
    rig = synth.camera_rig(7)
    metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug)]
    feats = cu(torch.ones(1, 6, 8, 16, 32))
    # no points
    assert point_to_cam([torch.zeros(0, 5, device=DEV)], feats, metas)[0].shape == (0, 8)
    # NaN / inf / behind-every-camera points -> zero rows, like the reference (masks are all False)
    bad = torch.tensor([[float("nan"), 0, 0, 0, 0], [float("inf"), 1, 1, 0, 0], [0, 0, 1e6, 0, 0]], device=DEV)
    out = point_to_cam([bad], feats, metas)[0]
    ref = O.point_to_cam([bad], feats, metas)[0]
    assert torch.equal(out, ref) and float(out.abs().max()) == 0
    # constant feature maps: interior points get (#cameras seeing them) * 1 up to edge taps
    pts = cu(synth.lidar_sweep(2000, seed=9)[:, :5])
    out = point_to_cam([pts], feats, metas)[0]
    ref = O.point_to_cam([pts], feats, metas)[0]
    assert normwise(out, ref) <= TOL and float(out.max()) <= 2.0 + 1e-5


# ---------------------------------------------------------------------------------------------
# a4 decode
# ---------------------------------------------------------------------------------------------
def _torch_cuda_sample(planes, q, lo, vs, half):
    """The reference's op sequence evaluated by torch on the GPU: [B,Q,3] -> [B,C,Q]."""
    v = torch.zeros_like(q)
    for a in range(3):
        v[..., a] = (q[..., a] - lo[a]) / vs[a]
    for a in range(3):
        v[..., a] = v[..., a] / half[a] - 1
    v = v[:, None]
    xy = F.grid_sample(planes[0], v[..., [0, 1]], mode="bilinear", padding_mode="zeros", align_corners=False)
    yz = F.grid_sample(planes[1], v[..., [1, 2]], mode="bilinear", padding_mode="zeros", align_corners=False)
    xz = F.grid_sample(planes[2], v[..., [0, 2]], mode="bilinear", padding_mode="zeros", align_corners=False)
    return (xy + yz + xz)[:, :, 0]


@pytest.mark.parametrize("name", ["sample_stacked4d", "sample_stacked5d_TriplaneOcc", "sample_stacked5d_TriplaneElev"])
@pytest.mark.parametrize("arith", ["cuda", "cpu"])
def test_sample_stacked_vs_reference_golden(name, arith):
    g = load_golden(name)
    tri = synth.triplane_stacked(int(g["batch"]), int(g["channels"]), int(g["size"]), seed=int(g["seed"]))
    out = sample_points_triplane(cu(tri), cu(g.t("points")), g["lo"].tolist(), g["vs"].tolist(), arith=arith)
    assert out.shape == g["out"].shape
    # The goldens were produced by the reference on torch-CPU: arith='cpu' replays that op chain and
    # must meet the bar. arith='cuda' replays torch-CUDA's chain (x * fp32(1/d) instead of x / d), which
    # torch itself puts ~1.4e-5 (normwise) away from its own CPU result; its 1e-5 bar is checked against
    # torch-CUDA in test_sample_vs_torch_cuda_grid_sample.
    assert normwise(out.cpu(), g.t("out")) <= (TOL if arith == "cpu" else 5 * TOL)


@pytest.mark.parametrize("name", ["sample_list4d", "sample_list5d"])
def test_sample_list_vs_reference_golden(name):
    g = load_golden(name)
    grid = g["grid"].tolist()
    planes = synth.triplane_list(int(g["batch"]), int(g["channels"]), grid, seed=int(g["seed"]))
    out = sample_points_triplane([cu(p) for p in planes], cu(g.t("points")), g["lo"].tolist(), g["vs"].tolist(),
                                 grid_size=grid, arith="cpu")
    assert out.shape == g["out"].shape
    assert normwise(out.cpu(), g.t("out")) <= TOL


def test_sample_config_exact_occ_decode():
    """configs/triplane_occ.py shapes: C=32, 3x128x128, the 99x99x16 roi() lattice."""
    c = load_golden("sample_occ_config")
    tri = cu(synth.triplane_stacked(1, 32, 128, seed=int(c["seed"])))
    ref3d = cu(synth.roi_lattice())[None]
    out = sample_points_triplane(tri, ref3d, c["lo"].tolist(), c["vs"].tolist(), arith="cpu")
    assert out.shape == (1, 32, 99, 99, 16)
    flat = out.reshape(1, 32, -1).cpu()
    assert normwise(flat[:, :, :: int(c["stride"])], c.t("out_strided")) <= TOL
    assert normwise(flat.double().sum(-1), c.t("out_sum")) <= 1e-6


@pytest.mark.parametrize("C,grid", [(32, None), (96, [128, 128, 80]), (4, None), (36, [20, 24, 12])])
def test_sample_vs_torch_cuda_grid_sample(C, grid):
    """Same GPU, same op chain: report the bitwise-equal fraction, assert the normwise bar and that we
    are no further from fp64 truth than torch's own fp32 result."""
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    B, Q = 2, 50000
    if grid is None:
        tri = cu(synth.triplane_stacked(B, C, 128, seed=C))
        planes, half = [tri[:, 0], tri[:, 1], tri[:, 2]], [64.0] * 3
        arg = tri
    else:
        vs = tuple(50.0 / grid[0] for _ in range(2)) + (8.0 / grid[2],)
        planes = [cu(p) for p in synth.triplane_list(B, C, grid, seed=C)]
        half, arg = [grid[a] / 2 for a in range(3)], planes
    q = torch.stack([synth.uniform_queries(Q, seed=40 + b) for b in range(B)])
    q[:, :5000] *= 1.2  # out-of-range share
    q = cu(q)
    ref = _torch_cuda_sample(planes, q, lo, vs, half)
    out = ops.sample3(arg, q, lo, vs, half)
    assert normwise(out, ref) <= TOL
    same = float((out == ref).float().mean())
    print(f"\n[sample C={C}] bitwise-equal to torch-CUDA grid_sample: {same:.6f}")
    out_nofma = ops.sample3(arg, q, lo, vs, half, arith="cuda_nofma")
    print(f"[sample C={C}] no-fma variant bitwise-equal: {float((out_nofma == ref).float().mean()):.6f}")
    f64 = _torch_cuda_sample([p.double() for p in planes], q.double(), lo, vs, half)
    assert normwise(out, f64) <= max(TOL, 1.05 * normwise(ref, f64))


def test_sample_edge_cases():
    lo, vs, half = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
    tri = cu(synth.triplane_stacked(1, 8, 128, seed=3))
    # Q = 0
    out = ops.sample3(tri, torch.zeros(1, 0, 3, device=DEV), lo, vs, half)
    assert out.shape == (1, 8, 0)
    # ragged Q (not a multiple of 32), all far outside -> exact zeros
    q = torch.full((1, 77, 3), 1e6, device=DEV)
    assert float(ops.sample3(tri, q, lo, vs, half).abs().max()) == 0
    # constant planes, interior queries -> 3 * constant (partition of unity)
    ones = torch.ones(1, 3, 8, 128, 128, device=DEV)
    qi = cu(synth.uniform_queries(1000, seed=8)) * 0.9
    qi[..., 2] = qi[..., 2].clamp(-4.5, 2.0)
    out = ops.sample3(ones, qi[None], lo, vs, half)
    assert float((out - 3).abs().max()) < 1e-5
    # linearity
    t2 = cu(synth.triplane_stacked(1, 8, 128, seed=4))
    a, b = ops.sample3(tri, qi[None], lo, vs, half), ops.sample3(t2, qi[None], lo, vs, half)
    ab = ops.sample3(tri + t2, qi[None], lo, vs, half)
    assert normwise(ab, a + b) < 1e-5
    # NaN query -> NaN result, like grid_sample
    qn = qi[None, :4].clone()
    qn[0, 0, 0] = float("nan")
    assert torch.isnan(ops.sample3(tri, qn, lo, vs, half)[0, :, 0]).all()


def test_sample_baseline_size_640k():
    """BASELINE.json config[1]: 640k queries. Checked against torch-CUDA on the same inputs plus the
    size-independent property that out-of-plane queries are exactly zero."""
    lo, vs, half = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
    tri = cu(synth.triplane_stacked(1, 32, 128, seed=1002))
    lat = cu(synth.occ_gt_lattice().reshape(1, -1, 3))
    out = ops.sample3(tri, lat, lo, vs, half)
    ref = _torch_cuda_sample([tri[:, 0], tri[:, 1], tri[:, 2]], lat, lo, vs, half)
    assert out.shape == (1, 32, 640000) and normwise(out, ref) <= TOL
    # a query misses all three planes only if BOTH x and y are outside (z is always inside here);
    # the 128-px planes span [-25, 26.2] m, half a pixel of bilinear support beyond that
    xy = lat[0, :, :2]
    far = ((xy < -25.5) | (xy > 26.5)).all(1)
    assert float(out[0][:, far].abs().max()) == 0 and int(far.sum()) > 100000
    inside = (xy.abs() < 24.7).all(1)
    assert int(inside.sum()) > 150000 and bool((out[0][:, inside].abs().sum(0) > 0).all())
    rnd = cu(synth.uniform_queries(640000, seed=1002))[None]
    out = ops.sample3(tri, rnd, lo, vs, half)
    ref = _torch_cuda_sample([tri[:, 0], tri[:, 1], tri[:, 2]], rnd, lo, vs, half)
    assert normwise(out, ref) <= TOL


def _grid_case(dims, C, B, seed, lattice_vs=(0.5, 0.5, 0.5), origin=(-30.0, -28.0, -5.0), planes_grid=None):
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    pts = synth.lattice(dims, lattice_vs, origin)[None].repeat(B, 1, 1, 1, 1)
    if planes_grid is None:
        tri = cu(synth.triplane_stacked(B, C, 128, seed=seed))
        return tri, cu(pts), lo, vs, [64.0] * 3
    planes = [cu(p) for p in synth.triplane_list(B, C, planes_grid, seed=seed)]
    vs = (50.0 / planes_grid[0], 50.0 / planes_grid[1], 8.0 / planes_grid[2])
    return planes, cu(pts), lo, vs, [g / 2 for g in planes_grid]


@pytest.mark.parametrize("dims,C,B,planes_grid", [
    ((200, 200, 16), 32, 1, None),          # BASELINE.json config[1]: the 640k occupancy-GT lattice
    ((99, 99, 16), 32, 2, None),            # configs/triplane_occ.py roi() lattice, ragged blocks in h and w
    ((100, 100, 80), 32, 1, None),          # triplane_elev.py get_reference_points volume: five k blocks
    ((37, 21, 20), 96, 2, [128, 128, 80]),  # list-of-planes variant, 3 channel chunks, ragged in h, w, d
    ((9, 5, 4), 36, 1, [20, 24, 12]),       # runtime C, partial channel chunk, d < 16
    ((16, 16, 6), 32, 1, None),             # d % 4 != 0: routed to the per-query kernel
])
@pytest.mark.parametrize("arith", ["cuda", "cpu"])
def test_sample_grid_bit_identical_to_flat(dims, C, B, planes_grid, arith, monkeypatch):
    """tp_sample3_grid_nhwc_f32 == tp_sample3_nhwc_f32 bit for bit: on the reference's lattices
    (separable blocks), with every block-shape configuration, and on inputs where only some blocks
    are lattices (in-launch fallback)."""
    planes, pts, lo, vs, half = _grid_case(dims, C, B, seed=11 + C, planes_grid=planes_grid)
    q = pts.reshape(B, -1, 3)
    flat = ops.sample3(planes, q, lo, vs, half, arith=arith)
    for tile in ("", "0", "1"):
        if tile:
            monkeypatch.setenv("TP_GRID_TILE", tile)
        got = ops.sample3(planes, q, lo, vs, half, arith=arith, grid_dims=dims)
        assert torch.equal(got, flat), f"lattice path differs (TP_GRID_TILE={tile!r})"
    monkeypatch.delenv("TP_GRID_TILE", raising=False)
    # break separability in a few places: those blocks must take the fallback, the rest the tables
    g = torch.Generator().manual_seed(5)
    jit = pts.clone()
    n = max(1, q.shape[1] // 500)
    idx = torch.randint(0, q.shape[1], (n,), generator=g).to(DEV)
    jit.view(B, -1, 3)[:, idx] += 0.173
    jit.view(B, -1, 3)[0, 0, 2] = float("nan")
    qj = jit.reshape(B, -1, 3)
    flat = ops.sample3(planes, qj, lo, vs, half, arith=arith)
    got = ops.sample3(planes, qj, lo, vs, half, arith=arith, grid_dims=dims)
    same = (got == flat) | (torch.isnan(got) & torch.isnan(flat))
    assert bool(same.all())
    # no lattice structure at all
    rnd = cu(torch.stack([synth.uniform_queries(q.shape[1], seed=70 + b) for b in range(B)])) * 1.1
    assert torch.equal(ops.sample3(planes, rnd, lo, vs, half, arith=arith, grid_dims=dims),
                       ops.sample3(planes, rnd, lo, vs, half, arith=arith))


def test_sample_grid_through_module_signature_matches_torch_cuda():
    """sample_points_triplane with 5-D points (the TriplaneOcc call) goes through the lattice entry
    point and is bitwise what torch-CUDA's grid_sample chain gives on >= 99.9 % of elements."""
    tri = cu(synth.triplane_stacked(2, 32, 128, seed=1002))
    pts = cu(synth.roi_lattice())[None].repeat(2, 1, 1, 1, 1)
    lo, vs = synth.OCC["triplane_range"], synth.OCC["triplane_voxel_size"]
    before = ops.launch_count
    out = sample_points_triplane(tri, pts, lo, vs)
    assert out.shape == (2, 32, 99, 99, 16) and ops.launch_count == before + 2
    ref = _torch_cuda_sample([tri[:, 0], tri[:, 1], tri[:, 2]], pts.reshape(2, -1, 3), lo, vs, [64.0] * 3)
    assert normwise(out.reshape(2, 32, -1), ref) <= TOL
    assert float((out.reshape(2, 32, -1) == ref).float().mean()) > 0.999


def test_host_buffer_entry_points():
    """tp_sample3_host_f32 / tp_encode_host_f32 (what a non-PyTorch caller binds)."""
    import ctypes as C
    from efficient_multimodal_perception_b200 import _lib as L
    lib = L.lib()
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    tri = synth.triplane_stacked(2, 8, 128, seed=3).contiguous()
    q = torch.stack([synth.uniform_queries(3000, seed=s) for s in (1, 2)]).contiguous()
    out = torch.empty(2, 8, 3000)
    ptrs = (C.c_void_p * 3)(*[tri[:, k].data_ptr() for k in range(3)])
    hw = (C.c_int32 * 6)(128, 128, 128, 128, 128, 128)
    bs = (C.c_int64 * 3)(*[tri.stride(0)] * 3)
    sg = L.make_sample_geom(lo, vs, [64.0] * 3)
    L.check(lib.tp_sample3_host_f32(C.byref(ptrs), C.byref(hw), C.byref(bs), 8, q.data_ptr(), 3000, 2,
                                    C.byref(sg), 1, out.data_ptr()), "tp_sample3_host_f32")
    ref = O.sample_points_triplane_stacked(tri, q[:, None], lo, vs)[:, :, 0]
    assert normwise(out, ref) <= TOL
    G = synth.GEOM_A
    raw = synth.lidar_sweep(5000, seed=4)[:, :3].contiguous()
    feats = torch.randn(5000, 16, generator=torch.Generator().manual_seed(5))
    geom = L.make_geom(G["pc_range"], G["voxel_size"], G["grid_size"], (5, 5, 4))
    outs = [torch.empty(1, 128, 128, 20 * 16), torch.empty(1, 128, 80, 25 * 16), torch.empty(1, 128, 80, 25 * 16)]
    off = torch.tensor([0, 5000], dtype=torch.int64)
    for _ in range(2):
        L.check(lib.tp_encode_host_f32(feats.data_ptr(), 16, raw.data_ptr(), 3, 5000, off.data_ptr(), 1,
                                       C.byref(geom), 1, 0, 0, outs[0].data_ptr(), outs[1].data_ptr(),
                                       outs[2].data_ptr()), "tp_encode_host_f32")
    cr, gi = O.voxelize_points([raw], G["pc_range"], G["voxel_size"])
    mask = (raw[:, 0].abs() < 25) & (raw[:, 1].abs() < 25) & (raw[:, 2] > -5) & (raw[:, 2] < 3)
    ref = O.encode_pooled(feats[mask], O.cat_indices(gi), G["grid_size"], G["split"], 1)
    for a, b in zip(outs, ref[:3]):
        assert torch.equal(a, b)
    lib.tp_host_arena_release()


@pytest.mark.parametrize("dims", [(100, 100, 16), None])
def test_host_entry_points_chunked_pipeline(dims):
    """Large enough for the host entry points to pipeline the query range in 4 chunks (H2D / kernel / D2H
    overlapped on two streams): the result must be bitwise what the device entry points give, including a
    ragged last chunk (Q = 100003 point list) and batch > 1."""
    import ctypes as C
    from efficient_multimodal_perception_b200 import _lib as L
    lib = L.lib()
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    B, Cc = 2, 32
    tri = synth.triplane_stacked(B, Cc, 128, seed=21).contiguous().pin_memory()
    if dims:
        q = synth.lattice(dims, (0.5, 0.5, 0.5), (-26.0, -24.0, -5.0)).reshape(1, -1, 3).repeat(B, 1, 1).contiguous()
    else:
        q = torch.stack([synth.uniform_queries(100003, seed=s) for s in (1, 2)]).contiguous() * 1.05
    Q = q.shape[1]
    q = q.pin_memory()
    out = torch.empty(B, Cc, Q).pin_memory()
    ptrs = (C.c_void_p * 3)(*[tri[:, k].data_ptr() for k in range(3)])
    hw = (C.c_int32 * 6)(*[128] * 6)
    bs = (C.c_int64 * 3)(*[tri.stride(0)] * 3)
    sg = L.make_sample_geom(lo, vs, [64.0] * 3)
    for _ in range(2):  # second call reuses the arena
        if dims:
            cd = (C.c_int32 * 3)(*dims)
            L.check(lib.tp_sample3_grid_host_f32(C.byref(ptrs), C.byref(hw), C.byref(bs), Cc, q.data_ptr(), C.byref(cd), B,
                                                 C.byref(sg), 0, out.data_ptr()), "tp_sample3_grid_host_f32")
        else:
            L.check(lib.tp_sample3_host_f32(C.byref(ptrs), C.byref(hw), C.byref(bs), Cc, q.data_ptr(), Q, B, C.byref(sg), 0,
                                            out.data_ptr()), "tp_sample3_host_f32")
    ref = ops.sample3(cu(tri), cu(q), lo, vs, [64.0] * 3)
    assert torch.equal(out, ref.cpu())
    lib.tp_host_arena_release()


@pytest.mark.parametrize("B,Q,ncls", [(1, 156816, 5), (2, 1000, 5), (1, 77, 16), (1, 128, 1), (3, 1283, 7), (2, 40004, 5),
                                      (5, 30001, 3)])
def test_mlp_head_tensor_core_kernel_vs_torch(B, Q, ncls):
    """tp_mlp_head_tf32 (tcgen05, TF32 inputs / fp32 accumulate) vs the reference head's three bias-free 1x1x1 convs
    (dense_heads/mlp.py:57-70) in fp32: within TF32 rounding, and no worse than torch's own TF32 path."""
    g = torch.Generator().manual_seed(B * 1000 + Q + ncls)
    C_ = 32
    x = cu(torch.randn(B, C_, Q, generator=g))
    w1 = cu(torch.randn(2 * C_, C_, 1, 1, 1, generator=g) / C_ ** 0.5)
    w2 = cu(torch.randn(C_, 2 * C_, 1, 1, 1, generator=g) / (2 * C_) ** 0.5)
    w3 = cu(torch.randn(ncls, C_, 1, 1, 1, generator=g) / C_ ** 0.5)

    def chain(xx, prec):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = prec == "tf32"
        try:
            h = torch.relu(torch.einsum("oc,bcq->boq", w1.view(2 * C_, C_).to(xx.dtype), xx))
            h = torch.relu(torch.einsum("oc,bcq->boq", w2.view(C_, 2 * C_).to(xx.dtype), h))
            return torch.einsum("oc,bcq->boq", w3.view(ncls, C_).to(xx.dtype), h)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old

    ref64 = chain(x.double(), "fp64")
    got = ops.mlp_head(x, w1, w2, w3)
    assert got.shape == (B, ncls, Q)
    err = normwise(got, ref64)
    err_tf32 = normwise(chain(x, "tf32"), ref64)
    print(f"\n[mlp head B={B} Q={Q} ncls={ncls}] normwise error vs fp64: ours {err:.2e}, torch TF32 {err_tf32:.2e}")
    assert err <= 3e-3 and err <= max(2 * err_tf32, 1e-3)
    # 5-D input as the detector passes it (triplane_occ.py:182-186)
    if Q == 1000:
        got5 = ops.mlp_head(x.view(B, C_, 10, 10, 10), w1, w2, w3)
        assert got5.shape == (B, ncls, 10, 10, 10) and torch.equal(got5.reshape(B, ncls, Q), got)


def test_mlp_module_eval_uses_the_fused_kernel_and_trains_with_torch():
    from efficient_multimodal_perception_b200 import Mlp
    torch.manual_seed(0)
    m = Mlp(32, 5).to(DEV)
    x = torch.randn(2, 32, 9, 9, 16, device=DEV)
    ref = m.conv3(m.conv2(m.conv1(x))).detach()
    before = ops.launch_count
    with torch.no_grad():
        got = m.eval()(x)
    assert ops.launch_count == before + 1 and got.shape == ref.shape
    assert normwise(got, ref) <= 3e-3
    out = m.train()(x)  # gradients required: PyTorch path, trainable
    assert ops.launch_count == before + 1
    out.square().mean().backward()
    assert m.conv1[0].weight.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize("B,dims,ncls,kind", [
    (1, (200, 200, 16), 5, "occ640k"),      # BASELINE lattice: 3/4 of it outside the planes
    (2, (99, 99, 16), 5, "roi"),            # config-exact roi lattice, ragged blocks in h and w
    (1, (10, 13, 8), 16, "small"),          # d < 16, 16 classes
    (2, (7, 9, 20), 3, "ragged_d"),         # two k-blocks, the second partial
    (1, (12, 16, 16), 5, "jitter"),         # not a lattice: every block takes the per-query path
    (1, (12, 16, 16), 5, "mixed"),          # one corner jittered: lattice and per-query blocks in one launch
])
def test_decode_plus_head_fused_equals_two_kernels(B, dims, ncls, kind):
    """tp_sample3_grid_head_tf32 == tp_mlp_head_tf32(tp_sample3_grid_nhwc_f32(...)) (triplane_occ.py:182-186 after
    :321-348): same features bit for bit, same TF32 tensor-core chain; and within TF32 rounding of the fp64 head."""
    from efficient_multimodal_perception_b200 import synth
    g = torch.Generator().manual_seed(77 + dims[0])
    h, w, d = dims
    lo, vs, half = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
    if kind == "occ640k":
        q = synth.occ_gt_lattice()
    elif kind == "roi":
        q = synth.roi_lattice()
    else:
        q = synth.lattice(dims, (0.9, 0.7, 0.35), (-4.0, -5.0, -4.5))
    assert tuple(q.shape[:3]) == dims
    q = q.unsqueeze(0).repeat(B, 1, 1, 1, 1)
    if kind == "jitter":
        q = q + 0.05 * torch.randn(q.shape, generator=g)
    if kind == "mixed":
        q[:, :4, :8] += 0.05 * torch.randn(q[:, :4, :8].shape, generator=g)
    tri = cu(torch.randn(B, 3, 32, 128, 128, generator=g))
    w1 = cu(torch.randn(64, 32, 1, 1, 1, generator=g) / 32 ** 0.5)
    w2 = cu(torch.randn(32, 64, 1, 1, 1, generator=g) / 64 ** 0.5)
    w3 = cu(torch.randn(ncls, 32, 1, 1, 1, generator=g) / 32 ** 0.5)
    qd = cu(q.reshape(B, -1, 3))
    feats = ops.sample3(tri, qd, lo, vs, half, grid_dims=dims)
    two = ops.mlp_head(feats, w1, w2, w3)
    before = ops.launch_count
    one = ops.sample3_head(tri, qd, lo, vs, half, w1, w2, w3, grid_dims=dims)
    assert ops.launch_count == before + 2  # layout conversion + the fused kernel
    assert one.shape == two.shape == (B, ncls, h * w * d)
    ref64 = torch.einsum("oc,bcq->boq", w3.view(ncls, 32).double(), torch.relu(torch.einsum(
        "oc,bcq->boq", w2.view(32, 64).double(), torch.relu(torch.einsum("oc,bcq->boq", w1.view(64, 32).double(), feats.double())))))
    assert normwise(one, ref64) <= 3e-3
    assert torch.equal(one, two), f"fused vs two-kernel logits differ: max abs {float((one - two).abs().max())}"


@pytest.mark.gpu
def test_mixin_sample_and_decode_matches_reference_two_step():
    """TriplaneOcc's `pred = self.decoder(self.sample_points_triplane(triplane, ref_3d))` (triplane_occ.py:182-186)
    through TriplaneHotPathMixin.sample_and_decode: one kernel without gradients, the two calls with them; both match
    the oracle's grid_sample chain followed by the reference head in fp32 within TF32 rounding."""
    from efficient_multimodal_perception_b200 import Mlp, TriplaneHotPathMixin, synth

    class Occ(TriplaneHotPathMixin, torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.triplane_range = [-25.0, -25.0, -5.0, 25.0, 25.0, 3.0]
            self.triplane_voxel_size = (0.4, 0.4, 0.1)
            self.occ_range, self.voxel_size = [-25.0, -25.0, -5.0, 25.0, 25.0, 3.0], (0.5, 0.5, 0.5)
            self.decoder = Mlp(32, 5)

    torch.manual_seed(3)
    m = Occ().cuda()
    tri = cu(torch.randn(2, 3, 32, 128, 128))
    _, ref_3d = m.roi()
    pts = cu(ref_3d).unsqueeze(0).repeat(2, 1, 1, 1, 1)
    with torch.no_grad():
        before = ops.launch_count
        pred = m.sample_and_decode(tri, pts)
        assert ops.launch_count == before + 2  # layout conversion + fused kernel
        two = m.decoder(m.sample_points_triplane(tri, pts))
    assert pred.shape == (2, 5, 99, 99, 16) and torch.equal(pred, two)
    feats = O.sample_points_triplane_stacked(tri.cpu(), pts.cpu(), m.triplane_range, m.triplane_voxel_size)
    with torch.no_grad():
        ref = m.decoder.cpu()(feats.float())
    m.decoder.cuda()
    assert normwise(pred.cpu(), ref) <= 3e-3
    tri.requires_grad_(True)
    out = m.sample_and_decode(tri, pts)  # training path: autograd through both steps
    out.square().mean().backward()
    assert tri.grad is not None and normwise(out.detach().cpu(), ref) <= 3e-3


@pytest.mark.gpu
def test_decode_plus_head_concurrent_streams_and_graph_replay():
    """The fused kernel hands out blocks from a device-side ticket counter that resets itself: launches that overlap
    on different streams must not share a counter, and a captured launch must be replayable."""
    from efficient_multimodal_perception_b200 import synth
    g = torch.Generator().manual_seed(11)
    lo, vs, half = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
    q = cu(synth.roi_lattice().reshape(1, -1, 3))
    dims = (99, 99, 16)
    w1 = cu(torch.randn(64, 32, generator=g) / 32 ** 0.5)
    w2 = cu(torch.randn(32, 64, generator=g) / 8)
    w3 = cu(torch.randn(5, 32, generator=g) / 32 ** 0.5)
    tris = [cu(torch.randn(1, 3, 32, 128, 128, generator=g)) for _ in range(3)]
    nhwc = [ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]]) for t in tris]
    want = [ops.mlp_head(ops.sample3(n, q, lo, vs, half, channels_last=True, grid_dims=dims), w1, w2, w3) for n in nhwc]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    got = [[] for _ in range(3)]
    for rep in range(8):
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                got[k].append(ops.sample3_head(nhwc[k], q, lo, vs, half, w1, w2, w3, grid_dims=dims, channels_last=True))
    torch.cuda.synchronize()
    for k in range(3):
        for o in got[k]:
            assert torch.equal(o, want[k])
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ops.sample3_head(nhwc[0], q, lo, vs, half, w1, w2, w3, grid_dims=dims, channels_last=True)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = [ops.sample3_head(nhwc[k], q, lo, vs, half, w1, w2, w3, grid_dims=dims, channels_last=True) for k in range(3)]
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    for k in range(3):
        assert torch.equal(outs[k], want[k])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(16)))
def test_lattice_fuzz_grid_and_fused_equal_flat(seed):
    """Random lattices that overlap the planes in random ways (fully inside, fully outside, only the x range inside,
    only y, a corner, ...), random plane sizes / channel counts / batch: the lattice kernel (plane skip, zero fast
    path) must equal the per-query kernel bit for bit, and for C = 32 the fused decode + head the two-kernel path."""
    from efficient_multimodal_perception_b200 import synth
    rs = np.random.RandomState(1000 + seed)
    g = torch.Generator().manual_seed(2000 + seed)
    B = int(rs.randint(1, 4))
    C_ = int(rs.choice([4, 32, 32, 96]))
    h, w = int(rs.randint(1, 41)), int(rs.randint(1, 41))
    d = int(rs.choice([4, 8, 12, 16, 20, 32]))
    sizes = [(int(rs.choice([16, 25, 128])), int(rs.choice([16, 80, 128]))) for _ in range(3)]  # (H, W) per plane
    lo, vs = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1)
    half = [float(rs.choice([8.0, 40.0, 64.0])) for _ in range(3)]
    # lattice origin and step per axis: in range is [lo, lo + 2 * half * vs)
    origin, step = [], []
    for a in range(3):
        span = 2 * half[a] * vs[a]
        mode = rs.randint(0, 4)
        n = (h, w, d)[a]
        if mode == 0:    # inside
            o, s = lo[a] + 0.1 * span, 0.7 * span / max(n, 1)
        elif mode == 1:  # entirely below the range
            o, s = lo[a] - 3.0 * span, 0.5 * span / max(n, 1)
        elif mode == 2:  # entirely above
            o, s = lo[a] + 1.5 * span, 0.5 * span / max(n, 1)
        else:            # straddling both ends
            o, s = lo[a] - 0.5 * span, 2.0 * span / max(n, 1)
        origin.append(float(o)); step.append(float(s))
    q = synth.lattice((h, w, d), step, origin).unsqueeze(0).repeat(B, 1, 1, 1, 1).reshape(B, -1, 3)
    planes = [cu(torch.randn(B, C_, H, W, generator=g)) for H, W in sizes]
    qd = cu(q)
    arith = "cpu" if seed % 4 == 3 else "cuda"  # both replayed rounding chains
    flat = ops.sample3(planes, qd, lo, vs, half, arith=arith)
    grid = ops.sample3(planes, qd, lo, vs, half, grid_dims=(h, w, d), arith=arith)
    assert torch.equal(grid, flat), f"seed {seed}: lattice kernel differs from the per-query kernel"
    if C_ == 32:
        w1 = cu(torch.randn(64, 32, generator=g) / 32 ** 0.5)
        w2 = cu(torch.randn(32, 64, generator=g) / 8)
        w3 = cu(torch.randn(7, 32, generator=g) / 32 ** 0.5)
        two = ops.mlp_head(flat, w1, w2, w3)
        one = ops.sample3_head(planes, qd, lo, vs, half, w1, w2, w3, grid_dims=(h, w, d), arith=arith)
        assert torch.equal(one, two), f"seed {seed}: fused decode + head differs from the two-kernel path"
