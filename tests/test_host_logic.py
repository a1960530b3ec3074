"""CPU: host-side mirror of the reference interface (names, shapes, parameter keys, geometry)."""
import torch

from conftest import load_golden
from efficient_multimodal_perception_b200 import PointTriplaneProjector, ops, roi, synth
from efficient_multimodal_perception_b200.modules import TriplaneHotPathMixin


def test_projector_parameter_names_match_reference_state_dict():
    g = load_golden("projector_small")
    ref_keys = {k[3:]: g[k].shape for k in g if k.startswith("sd.")}
    m = PointTriplaneProjector(g["grid"].tolist(), in_channels=5, out_channels=int(g["C"]),
                               base_channels=int(g["C"]), split=g["split"].tolist())
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert set(mine) == set(ref_keys)
    for k, shp in ref_keys.items():
        assert mine[k] == tuple(shp), k
    m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g if k.startswith("sd.")}, strict=True)


def test_projector_config_shapes():
    G = synth.GEOM_A
    m = PointTriplaneProjector(G["grid_size"], in_channels=5, out_channels=128, base_channels=128, split=G["split"])
    assert m.mlp_xy[0].in_features == 2560 and m.mlp_yz[0].in_features == 3200 and m.mlp_xz[0].in_features == 3200
    assert ops.pool_kernels(G["grid_size"], G["split"]) == (5, 5, 4)
    assert ops.pooled_sizes(G["grid_size"], (5, 5, 4)) == (25, 25, 20)
    assert ops.pool_kernels(synth.GEOM_B["grid_size"], synth.GEOM_B["split"]) == (8, 8, 1)


def test_roi_matches_reference():
    g = load_golden("roi")
    bounds, ref = roi(synth.OCC["occ_range"], synth.OCC["voxel_size"])
    assert list(bounds) == g["bounds"].tolist()
    assert torch.equal(ref, g.t("ref_3d"))


def test_mixin_geometry_resolution():
    class Occ(TriplaneHotPathMixin):
        triplane_range, triplane_voxel_size, voxel_size = [1] * 6, (2, 2, 2), (3, 3, 3)

    class Elev(TriplaneHotPathMixin):  # triplane_elev.py:298 mixes triplane_range with voxel_size
        triplane_range, voxel_size, pc_range = [1] * 6, (3, 3, 3), [9] * 6

    class Mae(TriplaneHotPathMixin):
        pc_range, voxel_size = [9] * 6, (3, 3, 3)

    assert Occ()._tp_geometry() == ([1] * 6, (2, 2, 2))
    assert Elev()._tp_geometry() == ([1] * 6, (3, 3, 3))
    assert Mae()._tp_geometry() == ([9] * 6, (3, 3, 3))


def test_mixin_voxelize_uses_the_class_own_geometry(monkeypatch):
    """PointTriplane crops with pc_range / voxel_size, its twin PointTriplaneOcc with triplane_range /
    triplane_voxel_size (point_triplane.py:148-156 vs point_triplane_occ.py:147-155)."""
    from efficient_multimodal_perception_b200 import modules
    seen = []
    monkeypatch.setattr(modules, "voxelize_points", lambda pts, rng, vs, arith: seen.append((rng, vs)) or ([], []))

    class Pre(TriplaneHotPathMixin):
        pc_range, voxel_size = [9] * 6, (3, 3, 3)

    class Occ(TriplaneHotPathMixin):
        pc_range, voxel_size, triplane_range, triplane_voxel_size = [9] * 6, (3, 3, 3), [1] * 6, (2, 2, 2)

    Pre().voxelize_points([])
    Occ().voxelize_points([])
    assert seen == [([9] * 6, (3, 3, 3)), ([1] * 6, (2, 2, 2))]


def test_synth_shapes():
    assert synth.lidar_sweep(1000, 1).shape == (1000, 11)
    assert synth.occ_gt_lattice().reshape(-1, 3).shape[0] == 640000
    assert synth.roi_lattice().shape == (99, 99, 16, 3)
    assert synth.range_image_points(2).shape == (2, 32, 1024, 3)
    lat = synth.occ_gt_lattice().reshape(-1, 3)
    inside = ((lat[:, :2].abs() < 25).all(1)).float().mean()
    assert abs(float(inside) - 0.25) < 0.01  # 3/4 of BASELINE's 640k lattice is outside the planes


def test_mlp_head_module_keeps_reference_parameter_names():
    """dense_heads/mlp.py:25-53: conv1 / conv2 / conv3 are nn.Sequential(Conv3d[, ReLU]) -> keys convN.0.weight."""
    from efficient_multimodal_perception_b200 import Mlp
    m = Mlp(32, 5)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {
        "conv1.0.weight": (64, 32, 1, 1, 1), "conv2.0.weight": (32, 64, 1, 1, 1), "conv3.0.weight": (5, 32, 1, 1, 1)}
    x = torch.randn(1, 32, 3, 4, 5)
    assert m(x).shape == (1, 5, 3, 4, 5)  # CPU tensors: the reference's own convolutions
    loss = m.loss(m(x), torch.randint(0, 5, (1, 3, 4, 5)))
    assert set(loss) == {"loss"} and loss["loss"].requires_grad
