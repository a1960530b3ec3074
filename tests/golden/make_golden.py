"""Generate tests/golden/*.npz by executing the REFERENCE's own code on torch-CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; the GPU box (no /root/reference) only reads them.

Every case stores inputs (or the seed that regenerates them through
efficient_multimodal_perception_b200.synth) and the reference's outputs. Source of each output:
  voxelize_*   PointTriplane.voxelize_points        mmdet3d/models/detectors/point_triplane.py:133-161
  sample_*     the five sample_points_triplane       triplane.py:490, triplane_occ.py:321,
                                                     triplane_elev.py:286, point_triplane.py:439,
                                                     point_triplane_occ.py:407
  roi          TriplaneOcc.roi                       triplane_occ.py:291-318
  lift         PointTriplane.point_to_cam            point_triplane.py:164-241
  projector_*  PointTriplaneProjector.forward        point_triplane_projector.py:66-117, with
               torch_scatter / spconv replaced by the oracle's restatements (oracle/ref_extract.py)
  interact     JointEncoder.interact                 mmdet3d/models/backbones/joint_encoder.py:97-215
  cam_proj_feat  the inline camera-pixel scatter     triplane.py:380-390 (statements cut out of TriplaneMAE.forward)
  cam_rec_feat PointTriplane.cam_rec_feat            point_triplane.py:243-309
  contrastive  the contrastive sampling loop         triplane.py:438-455 around TriplaneMAE.sample_points_triplane
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from efficient_multimodal_perception_b200 import synth  # noqa: E402
from oracle import ref_extract as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DET = "mmdet3d/models/detectors/"


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        conv[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(f"  {name}.npz  " + ", ".join(f"{k}{list(v.shape)}" for k, v in conv.items()))


def edge_points(pc_range, n_cols=11):
    """Rows that sit on / one ulp around the crop bounds (idx 125 / z idx 80 cases)."""
    lo, hi = np.float32(pc_range[:3]), np.float32(pc_range[3:])
    rows = []
    mid = (lo + hi) / 2
    for a in range(3):
        for v in (lo[a], np.nextafter(lo[a], np.float32(np.inf)), np.nextafter(lo[a], np.float32(-np.inf)),
                  hi[a], np.nextafter(hi[a], np.float32(-np.inf)), np.nextafter(hi[a], np.float32(np.inf))):
            r = mid.copy()
            r[a] = v
            rows.append(r)
    rows.append(np.array([np.nan, 0, 0], np.float32))
    rows.append(np.array([0, np.inf, 0], np.float32))
    pts = np.zeros((len(rows), n_cols), np.float32)
    pts[:, :3] = np.stack(rows)
    pts[:, 3:] = np.arange(len(rows) * (n_cols - 3), dtype=np.float32).reshape(len(rows), -1)
    return torch.from_numpy(pts)


def gen_voxelize():
    fn = R.load_method(DET + "point_triplane.py", "PointTriplane", "voxelize_points")
    for name, G in (("A", synth.GEOM_A), ("B", synth.GEOM_B)):
        self = SimpleNamespace(pc_range=G["pc_range"], voxel_size=G["voxel_size"])
        pts = [torch.cat([synth.lidar_sweep(700, seed=11), edge_points(G["pc_range"])]),
               synth.lidar_sweep(500, seed=12), torch.zeros(0, 11)]
        cropped, grid_ind = fn(self, pts)
        save(f"voxelize_{name}", pc_range=G["pc_range"], voxel_size=G["voxel_size"],
             points0=pts[0], points1=pts[1], points2=pts[2],
             cropped0=cropped[0], cropped1=cropped[1], cropped2=cropped[2],
             ind0=grid_ind[0], ind1=grid_ind[1], ind2=grid_ind[2])
    # config-sized sweep: store only the seed and the reference's indices
    G = synth.GEOM_A
    self = SimpleNamespace(pc_range=G["pc_range"], voxel_size=G["voxel_size"])
    pts = synth.lidar_sweep(34720, seed=1001)
    cropped, grid_ind = fn(self, [pts])
    save("voxelize_S1", seed=1001, n=34720, pc_range=G["pc_range"], voxel_size=G["voxel_size"],
         n_kept=cropped[0].shape[0], ind=grid_ind[0].to(torch.int16),
         cropped_checksum=cropped[0].double().sum(0))


def queries_mixed(B, h, w, lo, hi, seed, d=None):
    g = torch.Generator().manual_seed(seed)
    shape = (B, h, w, 3) if d is None else (B, h, w, d, 3)
    u = torch.rand(shape, generator=g)
    lo_t, hi_t = torch.tensor(lo), torch.tensor(hi)
    span = hi_t - lo_t
    q = lo_t - 0.15 * span + u * 1.3 * span  # ~25 % outside: exercises zero padding
    flat = q.view(-1, 3)
    flat[0] = lo_t  # exact corners / borders
    flat[1] = hi_t
    flat[2] = (lo_t + hi_t) / 2
    flat[3] = torch.tensor([0.0, 0.0, 0.0])  # empty range-image pixel
    return q


def gen_sample():
    lo, hi, vs = [-25.0, -25.0, -5.0], [25.0, 25.0, 3.0], (0.4, 0.4, 0.1)
    # stacked 4-D: TriplaneMAE (pc_range / voxel_size)
    fn = R.load_method(DET + "triplane.py", "TriplaneMAE", "sample_points_triplane")
    tri = synth.triplane_stacked(2, 8, 128, seed=21)
    pts = queries_mixed(2, 6, 50, lo, hi, 22)
    out = fn(SimpleNamespace(pc_range=lo + hi, voxel_size=vs), tri, pts)
    save("sample_stacked4d", seed=21, batch=2, channels=8, size=128, lo=lo, vs=vs, points=pts, out=out)
    # stacked 5-D: TriplaneOcc and TriplaneElev (triplane_range / triplane_voxel_size)
    for cls, f in (("TriplaneOcc", "triplane_occ.py"), ("TriplaneElev", "triplane_elev.py")):
        fn = R.load_method(DET + f, cls, "sample_points_triplane")
        tri = synth.triplane_stacked(2, 8, 128, seed=23)
        pts = queries_mixed(2, 5, 7, lo, hi, 24, d=9)
        # TriplaneElev reads self.voxel_size (triplane_elev.py:298), TriplaneOcc self.triplane_voxel_size
        out = fn(SimpleNamespace(triplane_range=lo + hi, triplane_voxel_size=vs, voxel_size=vs), tri, pts)
        save(f"sample_stacked5d_{cls}", seed=23, batch=2, channels=8, size=128, lo=lo, vs=vs, points=pts, out=out)
    # list variants with the config's non-square planes (X,Y,Z) = (128,128,80); C = 12 (multiple of 4, not of 8)
    grid = [128, 128, 80]
    proj = SimpleNamespace(grid_size=grid)
    fn = R.load_method(DET + "point_triplane.py", "PointTriplane", "sample_points_triplane")
    planes = synth.triplane_list(2, 12, grid, seed=25)
    pts = queries_mixed(2, 1, 300, lo, hi, 26)
    out = fn(SimpleNamespace(pc_range=lo + hi, voxel_size=vs, point_triplane_projector=proj), planes, pts)
    save("sample_list4d", seed=25, batch=2, channels=12, grid=grid, lo=lo, vs=vs, points=pts, out=out)
    fn = R.load_method(DET + "point_triplane_occ.py", "PointTriplaneOcc", "sample_points_triplane")
    pts = queries_mixed(2, 4, 5, lo, hi, 27, d=11)
    out = fn(SimpleNamespace(triplane_range=lo + hi, triplane_voxel_size=vs, point_triplane_projector=proj),
             planes, pts)
    save("sample_list5d", seed=25, batch=2, channels=12, grid=grid, lo=lo, vs=vs, points=pts, out=out)
    # config-exact occupancy decode: C=32, roi() lattice (156 816 queries); keep every 97th query
    roi_fn = R.load_method(DET + "triplane_occ.py", "TriplaneOcc", "roi")
    self = SimpleNamespace(occ_range=synth.OCC["occ_range"], voxel_size=synth.OCC["voxel_size"],
                           triplane_range=synth.OCC["triplane_range"],
                           triplane_voxel_size=synth.OCC["triplane_voxel_size"])
    bounds, ref_3d = roi_fn(self)
    save("roi", bounds=np.array(bounds), ref_3d=ref_3d)
    fn = R.load_method(DET + "triplane_occ.py", "TriplaneOcc", "sample_points_triplane")
    tri = synth.triplane_stacked(1, 32, 128, seed=1002)
    out = fn(self, tri, ref_3d[None])
    flat = out.reshape(1, 32, -1)
    save("sample_occ_config", seed=1002, stride=97, out_strided=flat[:, :, ::97],
         out_sum=flat.double().sum(-1), lo=self.triplane_range[:3], vs=self.triplane_voxel_size)


def gen_lift():
    fn = R.load_method(DET + "point_triplane.py", "PointTriplane", "point_to_cam")
    rig = synth.camera_rig(31)
    metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug)
             for _ in range(2)]
    metas[1] = dict(metas[1], imgs_aug=[dict(a, flip=(i % 2 == 1)) for i, a in enumerate(rig.imgs_aug)])
    pts = [synth.lidar_sweep(400, seed=32)[:, :5], synth.lidar_sweep(300, seed=33)[:, :5]]
    feats = torch.randn(2, 6, 16, 16, 32, generator=torch.Generator().manual_seed(34))
    out = fn(None, [p.clone() for p in pts], feats, metas)
    save("lift", points0=pts[0], points1=pts[1], img_features=feats, lidar2image=rig.lidar2image,
         resize=[a["resize"] for a in rig.imgs_aug], crop=[a["crop"] for a in rig.imgs_aug],
         flip0=[a["flip"] for a in metas[0]["imgs_aug"]], flip1=[a["flip"] for a in metas[1]["imgs_aug"]],
         img_shape=rig.img_shape, out0=out[0], out1=out[1])


def gen_projector():
    Proj = R.load_projector_class()
    cases = {
        # divisible grid
        "small": dict(grid=[16, 16, 8], split=[4, 4, 2], C=8, n=(300, 200)),
        # the config's non-divisible pattern: k=int(13/4)=3 -> index 12 falls outside the pooled extent
        "ragged": dict(grid=[13, 13, 9], split=[4, 4, 2], C=8, n=(400, 0, 250)),
    }
    for name, c in cases.items():
        torch.manual_seed(41)
        m = Proj(c["grid"], in_channels=5, out_channels=c["C"], base_channels=c["C"], split=c["split"]).eval()
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.running_mean.uniform_(-0.5, 0.5)
                    mod.running_var.uniform_(0.5, 1.5)
        g = torch.Generator().manual_seed(42)
        pts, ind, cam = [], [], []
        for n in c["n"]:
            pts.append(torch.randn(n, 11, generator=g))
            ind.append(torch.stack([torch.randint(0, c["grid"][a], (n,), generator=g) for a in range(3)], 1).int())
            cam.append(torch.randn(n, 768, generator=g))
        # spconv takes batch_size from the last coordinate row; an empty LAST sample would shrink it,
        # so cases keep the last sample non-empty.
        with torch.no_grad(), R.cpu_randperm():
            feats = m.point_mlp(torch.cat([p[:, :5] for p in pts])) + m.reduce_cam_channels(torch.cat(cam))
            out = m(pts, ind, cam)
        arrays = {f"sd.{k}": v for k, v in m.state_dict().items()}
        for i in range(len(pts)):
            arrays[f"points{i}"], arrays[f"ind{i}"], arrays[f"cam{i}"] = pts[i], ind[i], cam[i]
        save(f"projector_{name}", grid=c["grid"], split=c["split"], C=c["C"], nsamples=len(pts), feats=feats,
             tpv_xy=out[0], tpv_yz=out[1], tpv_xz=out[2], **arrays)


def small_rig(seed, flips=(False,) * 6):
    """synth.camera_rig scaled to 64x128 (H x W) images (resize 0.525/4, crop (41, 54)) so dense image fixtures stay
    small; img_shape is PIL's (W, H)."""
    rig = synth.camera_rig(seed)
    augs = [dict(resize=0.525 / 4, crop=(41, 54), flip=bool(f)) for f in flips]
    return dict(img_shape=(128, 64), lidar2image=rig.lidar2image.numpy(), imgs_aug=augs)


def meta_arrays(metas, prefix="meta"):
    out = {}
    for i, m in enumerate(metas):
        out[f"{prefix}{i}.lidar2image"] = np.asarray(m["lidar2image"], dtype=np.float64)
        out[f"{prefix}{i}.resize"] = [a["resize"] for a in m["imgs_aug"]]
        out[f"{prefix}{i}.crop"] = [a["crop"] for a in m["imgs_aug"]]
        out[f"{prefix}{i}.flip"] = [a["flip"] for a in m["imgs_aug"]]
        out[f"{prefix}{i}.img_shape"] = m["img_shape"]
    return out


def gen_interact():
    """JointEncoder.interact (joint_encoder.py:97-215) executed as-is, then the inline camera-pixel scatter of
    TriplaneMAE.forward (triplane.py:380-390) executed on ITS range_cam_coors."""
    fn = R.load_method("mmdet3d/models/backbones/joint_encoder.py", "JointEncoder", "interact")
    B, C, Hf, Wf = 2, 8, 8, 16
    metas = [small_rig(51), small_rig(52, flips=(False, True, False, False, True, False))]
    rp = synth.range_image_points(B, seed=53)[:, ::2, ::4].contiguous()       # [2,16,256,3], ~30 % empty pixels
    g = torch.Generator().manual_seed(54)
    rng_img = rp.norm(dim=-1)[:, None].clone()                                # range image; 0 where no point
    rng_img[torch.rand(rng_img.shape, generator=g) < 0.4] = 0.0               # MAE-masked pixels
    img = torch.randn(B, 6, C, Hf, Wf, generator=g)
    torch.manual_seed(55)
    pe = torch.nn.Sequential(torch.nn.Linear(3, 32), torch.nn.ReLU(), torch.nn.Linear(32, C))
    self = SimpleNamespace(position_encoder=pe)
    with torch.no_grad():
        cat, img_out, coors = fn(self, img.clone(), rng_img, metas, rp)
    arrays = {f"pe.{k}": v for k, v in pe.state_dict().items()}
    arrays.update(meta_arrays(metas))
    save("interact", range_points=rp, range_image=rng_img, img_features=img, out_cat=cat, out_img=img_out,
         range_cam_coors=coors, **arrays)
    # triplane.py:380-390 on the reference's own coordinates
    run = R.load_statements(DET + "triplane.py", "TriplaneMAE", "forward",
                            "cam_proj_feat[b, cam_it][:, cam_coors[:, 0], cam_coors[:, 1]] = proj_feat", before=2)
    assert run.lines == (380, 390), run.lines
    feat = torch.randn(B, 4, rp.shape[1], rp.shape[2], generator=g)
    H, W = metas[0]["img_shape"][::-1]
    ns = run(dict(range_cam_coors=coors.clone(), B=B, N=6, range_proj_feat=feat, H=H, W=W, img=feat))
    save("cam_proj_feat", range_proj_feat=feat, range_cam_coors=coors, H=H, W=W, out=ns["cam_proj_feat"])


def gen_cam_rec():
    fn = R.load_method(DET + "point_triplane.py", "PointTriplane", "cam_rec_feat")
    meta = small_rig(61, flips=(False, False, True, False, False, False))
    pts = synth.lidar_sweep(3000, seed=62)[:, :3].contiguous()
    feat = torch.randn(4, pts.shape[0], generator=torch.Generator().manual_seed(63))
    out = fn(None, pts.clone(), feat, meta)
    save("cam_rec_feat", points=pts, points_feat=feat, out=out, **meta_arrays([meta]))


def gen_contrastive():
    """The contrastive sampling loop (triplane.py:438-455) with the reference's own sample_points_triplane."""
    from oracle import triplane_oracle as O
    fn = R.load_method(DET + "triplane.py", "TriplaneMAE", "sample_points_triplane")
    G = synth.GEOM_A
    self = SimpleNamespace(pc_range=G["pc_range"], voxel_size=G["voxel_size"])
    tri = synth.triplane_stacked(3, 8, 32, seed=71)
    pts = [synth.lidar_sweep(900, seed=72), synth.lidar_sweep(700, seed=73), synth.lidar_sweep(40, seed=74)]
    pts[2][:, 5:] = 0          # a sample whose subsets are all empty
    pts[2][:1, 5] = 3.0        # ... except one single-point subset (skipped: labels.shape[0] > 1 fails)
    res = O.contrastive_features(lambda t, c: fn(self, t, c), tri, pts, G["pc_range"])
    arrays = {}
    for k, (f, lab) in enumerate(res):
        arrays[f"feat{k}"], arrays[f"label{k}"] = f, lab
    save("contrastive", triplane=tri, points0=pts[0], points1=pts[1], points2=pts[2], pc_range=G["pc_range"],
         voxel_size=G["voxel_size"], nsubsets=len(res), **arrays)


if __name__ == "__main__":
    if not R.available():
        sys.exit("needs /root/reference (build container only)")
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    for fn in (gen_voxelize, gen_sample, gen_lift, gen_projector, gen_interact, gen_cam_rec, gen_contrastive):
        if only and fn.__name__ not in only:
            continue
        print(fn.__name__)
        fn()
