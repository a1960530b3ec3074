#!/usr/bin/env python
"""Benchmark of the triplane hot path (BASELINE.json metric: triplane encode points/s + query-sample
queries/s on B200, % of HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--queries ...]

Headline (`value`): occupancy decode of configs/triplane_occ.py at BASELINE.json's size — 640 000
voxel queries sampled from three fp32 128x128 triplanes with C=32 (`configs[1]`) — in queries/s,
inputs resident in HBM. A step = one pass of the decode path over one batch: the NCHW->NHWC
conversion of the three planes (one launch) + the fused 3-plane gather kernel (one launch). Steps rotate over `nsets` disjoint
buffer sets whose total footprint exceeds 3x the 126 MB L2, and are replayed from CUDA graphs so the
Python launch cost is not what is measured. The same JSON line also carries the encode leg
(`encode`: points/s for one synthetic nuScenes sweep at the config-exact geometry; `variants_kernel_only`:
the other query sets of SURVEY 8(d) S2, including 640k uniform-random in-range queries = worst-case locality), `roofline`
(dominant kernel, algorithmic bytes / CUDA-event time, vs MEASURED_PEAKS.json), `cpu_baseline`
(the oracle on the box's host cores), `e2e` (host buffers through the C ABI) and `clocks`.

Under torchrun (N > 1) every rank runs the same per-GPU workload on its own shard of queries /
samples (weak scaling, no data-path collective: decode queries and encode samples are independent);
time is the max over ranks, value the sum of work over ranks / that time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent

OCC_LO, OCC_VS, OCC_HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
C_DEC, PLANE = 32, 128


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def decode_queries(kind: str):
    from efficient_multimodal_perception_b200 import synth
    if kind == "lattice640k":
        return synth.occ_gt_lattice().reshape(1, -1, 3).contiguous()
    if kind == "uniform640k":
        return synth.uniform_queries(640000, seed=1002).reshape(1, -1, 3).contiguous()
    if kind == "roi":
        return synth.roi_lattice().reshape(1, -1, 3).contiguous()
    raise SystemExit(f"unknown --queries {kind}")


#: [h, w, d] of the query tensors the reference passes as [B,h,w,d,3] (5-D callers); None = a flat point list
QUERY_DIMS = {"lattice640k": (200, 200, 16), "roi": (99, 99, 16), "uniform640k": None}


def ncu_traffic(kernel: str, key: str):
    """dram bytes per launch from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            return json.load(fh)[kernel][key]
    except Exception:
        return None


def decode_bytes(Q: int, C_: int = C_DEC) -> int:
    """Q*(12 + 4C) + 4C*sum(HW) per sample."""
    return Q * (12 + 4 * C_) + 4 * C_ * 3 * PLANE * PLANE


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed regions run."""

    def __init__(self, index: int, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.samples, self.reasons, self.max_mhz = index, uuid, [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:  # CUDA_VISIBLE_DEVICES may renumber devices: prefer the UUID
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                if self.active.is_set():
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.004)
        except Exception as e:  # NVML missing: report nulls rather than invent numbers
            self.error = str(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# b200 arm
# --------------------------------------------------------------------------------------------------
class DecodeSets:
    """nsets disjoint (planes, queries, out) buffer sets + CUDA graphs of the 2-launch step."""

    def __init__(self, q_host: torch.Tensor, nsets: int, dev, dims=None):
        from efficient_multimodal_perception_b200 import ops, synth
        self.ops, self.dev, self.nsets, self.dims = ops, dev, nsets, dims
        self.Q = q_host.shape[1]
        self.sets = []
        for s in range(nsets):
            tri = synth.triplane_stacked(1, C_DEC, PLANE, seed=1002 + s).to(dev)
            # every set has its own query buffer; point lists are also rotated, lattices keep their [h,w,d] order
            q = q_host.to(dev).clone() if (s == 0 or dims is not None) else q_host.roll(s * 1013, 1).to(dev)
            out = torch.empty(1, C_DEC, self.Q, device=dev)
            self.sets.append((tri, q, out))
        self.graph_all = self.graph_one = None

    def step(self, s: int):
        tri, q, out = self.sets[s % self.nsets]
        self.ops.sample3(tri, q, OCC_LO, OCC_VS, OCC_HALF, out=out, grid_dims=self.dims)

    def capture(self):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for s in range(self.nsets):
                self.step(s)
        torch.cuda.synchronize()
        self.graph_all = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_all):
            for s in range(self.nsets):
                self.step(s)
        self.graph_one = []
        for s in range(self.nsets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.step(s)
            self.graph_one.append(g)
        torch.cuda.synchronize()

    def run_steps(self, k: int):
        full, rem = divmod(k, self.nsets)
        for _ in range(full):
            self.graph_all.replay()
        for s in range(rem):
            self.graph_one[s].replay()


def time_region(fn, barrier):
    """CUDA-event time (ms) of fn() on the current stream, barrier + synchronize on both sides."""
    barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    barrier()
    return a.elapsed_time(b)


def kernel_time_ms(launch, reps: int, group: int):
    """Average duration of ONE launch of a kernel: `group` back-to-back launches (one per rotating buffer set)
    are captured in a CUDA graph, CUDA events bracket each replay, duration = elapsed / group. Events around a
    single ~20 us launch would add the event/launch latency (~4 us) to every sample."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(group):
            launch(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(group):
            launch(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    evs = []
    for _ in range(max(3, reps // group)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) / group for a, b in evs)
    return sum(ts) / len(ts), ts[len(ts) // 2], ts[0]


def bench_decode_device(args, dev, barrier, sampler):
    q_host = decode_queries(args.queries)
    Q = q_host.shape[1]
    per_set = decode_bytes(Q) + 4 * C_DEC * 3 * PLANE * PLANE  # + the channels-last copy
    nsets = max(4, -(-3 * L2_BYTES // per_set))
    dims = QUERY_DIMS[args.queries]
    sets = DecodeSets(q_host, nsets, dev, dims)
    sets.capture()
    sets.run_steps(args.warmup)
    sampler.active.set()
    ms = time_region(lambda: sets.run_steps(args.steps), barrier)
    # per-kernel duration of the dominant kernel (the gather), channels-last copies prepared outside
    ops = sets.ops
    nhwc = [ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]]) for t, _, _ in sets.sets]
    reps = min(max(args.steps, 20), 400)

    def launch(i):
        _, q, out = sets.sets[i % nsets]
        ops.sample3(nhwc[i % nsets], q, OCC_LO, OCC_VS, OCC_HALF, channels_last=True, out=out, grid_dims=dims)

    for i in range(8):
        launch(i)
    torch.cuda.synchronize()
    k_avg, k_med, k_min = kernel_time_ms(launch, reps, nsets)
    # the other query sets of SURVEY 8(d) S2, gather kernel only, same rotation of plane / output sets
    variants = {}
    for kind in ("lattice640k", "uniform640k", "roi"):
        if kind == args.queries:
            continue
        qv = decode_queries(kind)
        vdims = QUERY_DIMS[kind]
        vsets = max(nsets, -(-3 * L2_BYTES // decode_bytes(qv.shape[1])))
        qs = [qv.to(dev).clone() if vdims is not None else qv.roll(s * 1013, 1).to(dev) for s in range(vsets)]
        outs = [torch.empty(1, C_DEC, qv.shape[1], device=dev) for s in range(vsets)]

        def launch_v(i, qs=qs, outs=outs, vdims=vdims, vsets=vsets):
            ops.sample3(nhwc[i % nsets], qs[i % vsets], OCC_LO, OCC_VS, OCC_HALF, channels_last=True,
                        out=outs[i % vsets], grid_dims=vdims)

        for i in range(8):
            launch_v(i)
        torch.cuda.synchronize()
        va, vm, vmin = kernel_time_ms(launch_v, min(reps, 200), vsets)
        vb = decode_bytes(qv.shape[1])
        variants[kind] = {"Q": qv.shape[1], "kernel_ms_avg": va, "kernel_ms_median": vm, "queries_per_s": qv.shape[1] / (va * 1e-3),
                          "algorithmic_bytes": vb, "achieved_gbs": vb / (va * 1e-3) / 1e9}
        del qs, outs
    # range-image points of configs/triplane_surf_sam.py, bs=8: [8,32,1024,3] (30 % empty pixels at
    # the origin), per-query kernel, one stacked triplane per sample
    from efficient_multimodal_perception_b200 import synth
    Br = 8
    rq = synth.range_image_points(Br, seed=1003).reshape(Br, -1, 3).contiguous()
    rbytes = Br * decode_bytes(rq.shape[1])
    rsets = max(4, -(-3 * L2_BYTES // rbytes))
    r_tri = [ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]])
             for t in (synth.triplane_stacked(Br, C_DEC, PLANE, seed=1003 + s).to(dev) for s in range(rsets))]
    r_q = [rq.to(dev).clone() for _ in range(rsets)]
    r_out = [torch.empty(Br, C_DEC, rq.shape[1], device=dev) for _ in range(rsets)]

    def launch_r(i):
        ops.sample3(r_tri[i % rsets], r_q[i % rsets], OCC_LO, OCC_VS, OCC_HALF, channels_last=True, out=r_out[i % rsets])

    for i in range(rsets):
        launch_r(i)
    torch.cuda.synchronize()
    ra, rm, _ = kernel_time_ms(launch_r, min(reps, 200), rsets)
    variants["range_points_bs8"] = {"Q": Br * rq.shape[1], "kernel_ms_avg": ra, "kernel_ms_median": rm,
                                    "queries_per_s": Br * rq.shape[1] / (ra * 1e-3), "algorithmic_bytes": rbytes,
                                    "achieved_gbs": rbytes / (ra * 1e-3) / 1e9}
    del r_tri, r_q, r_out
    sampler.active.clear()
    return dict(Q=Q, nsets=nsets, ms_total=ms, kernel_ms_avg=k_avg, kernel_ms_med=k_med, kernel_ms_min=k_min,
                launches=args.steps * 2, sets=sets, variants=variants)


def torch_cuda_decode(q_dev, tri_dev, reps=5):
    """The reference's own op sequence (normalise + 3 x F.grid_sample + 2 adds, triplane_occ.py:332-346) run by
    torch on the same GPU: the GPU baseline the kernels replace. Returns ms per pass."""
    import torch.nn.functional as F

    def chain():
        v = torch.zeros_like(q_dev)
        for a in range(3):
            v[..., a] = (q_dev[..., a] - OCC_LO[a]) / OCC_VS[a]
        v = v / OCC_HALF[0] - 1
        v = v[:, None]
        xy = F.grid_sample(tri_dev[:, 0], v[..., [0, 1]], mode="bilinear", padding_mode="zeros", align_corners=False)
        yz = F.grid_sample(tri_dev[:, 1], v[..., [1, 2]], mode="bilinear", padding_mode="zeros", align_corners=False)
        xz = F.grid_sample(tri_dev[:, 2], v[..., [0, 2]], mode="bilinear", padding_mode="zeros", align_corners=False)
        return xy + yz + xz

    for _ in range(2):
        chain()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        chain()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_decode_e2e(args, Q_host, barrier):
    """Host buffers through the C ABI (tp_sample3_host_f32): H2D of planes + queries, conversion,
    gather, D2H of the full [C,Q] result, synchronised, every step."""
    from efficient_multimodal_perception_b200 import _lib as L
    from efficient_multimodal_perception_b200 import synth
    lib = L.lib()
    Q = Q_host.shape[1]
    tri = synth.triplane_stacked(1, C_DEC, PLANE, seed=1002).pin_memory()
    q = Q_host.clone().pin_memory()
    out = torch.empty(1, C_DEC, Q).pin_memory()
    ptrs = (C.c_void_p * 3)(*[tri[:, k].data_ptr() for k in range(3)])
    hw = (C.c_int32 * 6)(*[PLANE] * 6)
    bs = (C.c_int64 * 3)(*[tri.stride(0)] * 3)
    sg = L.make_sample_geom(OCC_LO, OCC_VS, OCC_HALF)

    dims = QUERY_DIMS[args.queries]
    cdims = (C.c_int32 * 3)(*dims) if dims else None

    def call():
        if dims:
            L.check(lib.tp_sample3_grid_host_f32(C.byref(ptrs), C.byref(hw), C.byref(bs), C_DEC, q.data_ptr(),
                                                 C.byref(cdims), 1, C.byref(sg), L.TP_ARITH_TORCH_CUDA,
                                                 out.data_ptr()), "tp_sample3_grid_host_f32")
        else:
            L.check(lib.tp_sample3_host_f32(C.byref(ptrs), C.byref(hw), C.byref(bs), C_DEC, q.data_ptr(), Q, 1,
                                            C.byref(sg), L.TP_ARITH_TORCH_CUDA, out.data_ptr()), "tp_sample3_host_f32")

    steps = max(5, min(args.steps, 100))
    for _ in range(3):
        call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = time.perf_counter() - t0
    barrier()
    h2d = tri.numel() * 4 + q.numel() * 4
    d2h = out.numel() * 4
    return dict(qps=Q * steps / dt, steps=steps, h2d=h2d, d2h=d2h, ms=dt / steps * 1e3, out=out)


def bench_encode_device(args, dev, barrier, sampler):
    """Encode leg: configs/point_triplane.py geometry (128x128x80, pool 5/5/4, C=128), one synthetic
    sweep (34 720 raw points), crop + index fused into the encode (raw points in)."""
    from efficient_multimodal_perception_b200 import ops, synth
    G = synth.GEOM_A
    n, Cc = 34720, G["channels"]
    pts = synth.lidar_sweep(n, seed=1001)
    xyz = pts[:, :3].contiguous().to(dev)
    feats = synth.point_features(n, Cc, seed=1001).to(dev)
    off = synth.batch_offsets([n]).to(dev)
    inside = int(((pts[:, 0].abs() < 25) & (pts[:, 1].abs() < 25) & (pts[:, 2] > -5) & (pts[:, 2] < 3)).sum())
    cells = 128 * 128 * 20 + 2 * 128 * 80 * 25

    def step():
        return ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)

    for _ in range(max(3, min(args.warmup, 10))):
        outs = step()
    steps = max(10, min(args.steps, 200))
    sampler.active.set()
    def run():
        for _ in range(steps):
            step()  # outputs are dropped each step: the caching allocator hands the same blocks back

    ms = time_region(run, barrier)
    sampler.active.clear()
    bytes_alg = n * 12 + inside * 4 * Cc + 4 * Cc * cells
    return dict(n=n, inside=inside, cells=cells, steps=steps, ms_per_step=ms / steps, bytes=bytes_alg,
                launches=4 * steps)


def bench_encode_dense(args, dev, barrier, sampler):
    """BASELINE.json configs[4] shape per GPU: samples of 10 accumulated sweeps (~350k points each), sample-sharded
    (bs=64 over 8 GPUs = 8 per GPU; 4 per GPU here to bound the footprint), one batched encode call."""
    from efficient_multimodal_perception_b200 import ops, synth
    G = synth.GEOM_A
    B, Cc = 4, G["channels"]
    pts = [synth.multi_sweep(10, 35000, seed=1005 + b) for b in range(B)]
    sizes = [p.shape[0] for p in pts]
    xyz = torch.cat([p[:, :3] for p in pts]).contiguous().to(dev)
    feats = synth.point_features(sum(sizes), Cc, seed=1005).to(dev)
    off = synth.batch_offsets(sizes).to(dev)
    inside = int(sum(int(((p[:, 0].abs() < 25) & (p[:, 1].abs() < 25) & (p[:, 2] > -5) & (p[:, 2] < 3)).sum()) for p in pts))
    cells = B * (128 * 128 * 20 + 2 * 128 * 80 * 25)

    def step():
        return ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)

    for _ in range(3):
        step()
    steps = max(5, min(args.steps, 20))
    sampler.active.set()

    def run():
        for _ in range(steps):
            step()

    ms = time_region(run, barrier)
    sampler.active.clear()
    n = sum(sizes)
    return dict(n=n, B=B, inside=inside, steps=steps, ms_per_step=ms / steps,
                bytes=n * 12 + inside * 4 * Cc + 4 * Cc * cells, launches=4 * steps)


def bench_lift(args, dev, barrier, sampler):
    """Camera -> point lift (point_to_cam) at the PointTriplane config: 6 cameras x [768,16,32] feature maps per
    sample, bs=2 sweeps of 34 720 points. A step = channels-last copy of the maps + the fused lift kernel."""
    from efficient_multimodal_perception_b200 import ops, synth
    B, ncam, Cf, Hf, Wf, n = 2, 6, 768, 16, 32, 34720
    rig = synth.camera_rig(1004)
    metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug) for _ in range(B)]
    pts = torch.cat([synth.lidar_sweep(n, seed=1004 + b)[:, :3] for b in range(B)]).contiguous().to(dev)
    off = synth.batch_offsets([n] * B).to(dev)
    feats = torch.randn(B, ncam, Cf, Hf, Wf, generator=torch.Generator().manual_seed(1004)).to(dev)
    cams = ops.pack_cameras(metas, dev)
    dims = rig.img_shape[::-1]

    def step():
        return ops.lift_cam(pts, off, feats, cams, dims)

    for _ in range(3):
        step()
    nhwc = ops.features_to_channels_last(feats)
    steps = max(10, min(args.steps, 100))
    sampler.active.set()

    def run():
        for _ in range(steps):
            step()

    ms = time_region(run, barrier) / steps

    def run_k():
        for _ in range(steps):
            ops.lift_cam(pts, off, nhwc, cams, dims, channels_last=True)

    ms_k = time_region(run_k, barrier) / steps
    sampler.active.clear()
    # N'*(12 + 4*768) + the feature maps once
    return dict(n=B * n, steps=steps, ms_per_step=ms, kernel_ms=ms_k, launches=2 * steps,
                bytes=B * n * (12 + 4 * Cf) + B * ncam * Cf * Hf * Wf * 4)


def bench_occ_head(args, dev, barrier, sampler):
    """configs/triplane_occ.py occupancy pipeline on the BASELINE lattice: decode (conversion + gather) followed by the
    Mlp head (dense_heads/mlp.py: 32 -> 64 -> 32 -> 5, three bias-free 1x1x1 convs) as ONE tensor-core kernel."""
    from efficient_multimodal_perception_b200 import ops, synth
    q = synth.occ_gt_lattice().reshape(1, -1, 3).contiguous().to(dev)
    tri = synth.triplane_stacked(1, C_DEC, PLANE, seed=1002).to(dev)
    g = torch.Generator().manual_seed(7)
    w1 = (torch.randn(2 * C_DEC, C_DEC, 1, 1, 1, generator=g) / C_DEC ** 0.5).to(dev)
    w2 = (torch.randn(C_DEC, 2 * C_DEC, 1, 1, 1, generator=g) / (2 * C_DEC) ** 0.5).to(dev)
    w3 = (torch.randn(5, C_DEC, 1, 1, 1, generator=g) / C_DEC ** 0.5).to(dev)
    dims = QUERY_DIMS["lattice640k"]
    nsets = 4  # 4 x (82 MB features + 6 MB planes + 13 MB logits) > 3 x L2
    tris = [synth.triplane_stacked(1, C_DEC, PLANE, seed=1002 + i).to(dev) for i in range(nsets)]
    feats = [ops.sample3(t, q, OCC_LO, OCC_VS, OCC_HALF, grid_dims=dims) for t in tris]

    def head(i):
        return ops.mlp_head(feats[i % nsets], w1, w2, w3)

    def both(i):
        ops.sample3(tris[i % nsets], q, OCC_LO, OCC_VS, OCC_HALF, grid_dims=dims, out=feats[i % nsets])
        return ops.mlp_head(feats[i % nsets], w1, w2, w3)

    def fused(i):
        return ops.sample3_head(tris[i % nsets], q, OCC_LO, OCC_VS, OCC_HALF, w1, w2, w3, grid_dims=dims)

    assert torch.equal(fused(0), both(0)), "fused decode + head differs from the two-kernel path"
    steps = max(16, min(args.steps, 200))
    out = {}
    sampler.active.set()
    for name, fn in (("head", head), ("decode_plus_head", both), ("fused", fused)):
        barrier()
        out[name] = kernel_time_ms(fn, steps, 2 * nsets)[0]  # CUDA-graph replays of 8 steps, events around each replay
    barrier()
    sampler.active.clear()
    Q = q.shape[1]
    return dict(Q=Q, steps=steps, head_ms=out["head"], both_ms=out["decode_plus_head"], fused_ms=out["fused"],
                head_bytes=Q * (4 * C_DEC + 4 * 5), flops=2 * Q * (C_DEC * 2 * C_DEC * 2 + C_DEC * 5))


def bench_encode_point_sharded(args, dev, barrier, rank, world):
    """N > 1 only: ONE 10-sweep sample (350 000 raw points) point-sharded over the ranks:
    partial planes (-inf empties) -> NCCL all-reduce(max) of the 430 MB dense planes -> finalise."""
    from efficient_multimodal_perception_b200 import dist as tpd
    from efficient_multimodal_perception_b200 import synth
    G = synth.GEOM_A
    pts = synth.multi_sweep(10, 35000, seed=1005)[:, :3].contiguous()
    feats = synth.point_features(pts.shape[0], G["channels"], seed=1005)
    lo, hi = tpd.shard_bounds(pts.shape[0], rank, world)
    my_pts, my_feats = pts[lo:hi].to(dev), feats[lo:hi].contiguous().to(dev)
    off = synth.batch_offsets([hi - lo]).to(dev)

    out = {}
    for strategy in ("planes", "points"):
        def step(strategy=strategy):
            return tpd.encode_point_sharded(my_feats, my_pts, off, G["pc_range"], G["voxel_size"], G["grid_size"],
                                            G["split"], reduce="max", strategy=strategy)

        for _ in range(3):
            step()
        steps = max(5, min(args.steps, 30))

        def run(step=step, steps=steps):
            for _ in range(steps):
                step()

        out[strategy] = time_region(run, barrier) / steps
    return dict(n=pts.shape[0], ms_per_step=out["planes"], ms_points=out["points"], steps=steps,
                allreduce_bytes=4 * G["channels"] * 839680, gather_bytes=pts.shape[0] * (12 + 4 * G["channels"]))


def cpu_baseline_decode(q_host, budget_s=15.0):
    """The oracle (torch-CPU restatement of the reference, kind='port') on all host cores."""
    from efficient_multimodal_perception_b200 import synth
    from oracle import triplane_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    tri = synth.triplane_stacked(1, C_DEC, PLANE, seed=1002)
    pts = q_host.view(1, 1, -1, 3)
    O.sample_points_triplane_stacked(tri, pts, OCC_LO, OCC_VS)  # warm-up
    ts, t_start = [], time.perf_counter()
    while len(ts) < 10 and (time.perf_counter() - t_start < budget_s or len(ts) < 2):
        t0 = time.perf_counter()
        ref = O.sample_points_triplane_stacked(tri, pts, OCC_LO, OCC_VS)
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return dict(value=q_host.shape[1] / med, unit="queries/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{len(ts)} full passes of the {q_host.shape[1]}-query workload, median "
                       f"(oracle.sample_points_triplane_stacked = the reference's 3 x F.grid_sample path on torch-CPU)"), ref


def bind_to_gpu_numa_node(local: int):
    """Pin this rank's threads (and therefore its pinned host buffers, first-touch) to the NUMA node of its GPU:
    the e2e leg moves ~96 MB per step over PCIe per rank and 8 ranks on the wrong socket share one inter-socket
    link. Best effort: returns the node or None."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_b200(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        barrier = lambda: dist.barrier()  # noqa: E731
    else:
        barrier = lambda: None  # noqa: E731
    import efficient_multimodal_perception_b200 as emp
    emp.lib()
    peak, peak_src = measured_hbm_peak()
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.start()

    dec = bench_decode_device(args, dev, barrier, sampler)
    enc = bench_encode_device(args, dev, barrier, sampler)
    encd = bench_encode_dense(args, dev, barrier, sampler)
    lift = bench_lift(args, dev, barrier, sampler)
    occ = bench_occ_head(args, dev, barrier, sampler)
    eps = bench_encode_point_sharded(args, dev, barrier, rank, world) if world > 1 else None
    # max over ranks (device time)
    t = torch.tensor([dec["ms_total"], enc["ms_per_step"], dec["kernel_ms_avg"], eps["ms_per_step"] if eps else 0.0,
                      encd["ms_per_step"], lift["ms_per_step"], lift["kernel_ms"], eps["ms_points"] if eps else 0.0],
                     device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, enc_ms, k_avg, eps_ms, encd_ms, lift_ms, lift_k_ms, eps_pts_ms = (float(x) for x in t.tolist())
    clocks = sampler.stop()

    q_host = decode_queries(args.queries)
    e2e = bench_decode_e2e(args, q_host, barrier)
    te = torch.tensor([e2e["ms"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())

    if rank == 0:
        Q = dec["Q"]
        ms_per_step = ms_total / args.steps
        qps = world * Q * args.steps / (ms_total * 1e-3)
        kbytes = decode_bytes(Q)
        achieved = kbytes / (k_avg * 1e-3) / 1e9
        kernel_name = ("tp::sample3_grid_kernel<0,8,8>" if QUERY_DIMS[args.queries] else "tp::sample3_kernel<0,8>")
        cpu, parity, torch_ms = None, None, None
        if world == 1:
            cpu, ref = cpu_baseline_decode(q_host)
            # live parity check of what was just timed (device result of set 0 and the e2e result)
            tri0, q0, out0 = dec["sets"].sets[0]
            dec["sets"].step(0)
            torch.cuda.synchronize()
            scale = float(ref.abs().max())
            torch_ms = torch_cuda_decode(q0, tri0)
            out_cpu_arith = dec["sets"].ops.sample3(tri0, q0, OCC_LO, OCC_VS, OCC_HALF, arith="cpu",
                                                    grid_dims=QUERY_DIMS[args.queries])
            parity = {"note": "oracle = torch-CPU op chain; arith='cpu' replays it, the timed arith='cuda' replays "
                              "torch-CUDA's (x * fp32(1/vs)) and is checked bitwise-level against torch-CUDA in tests/",
                      "device_cpu_arith_vs_oracle_normwise": float((out_cpu_arith.cpu() - ref[:, :, 0]).abs().max()) / scale,
                      "device_vs_oracle_normwise": float((out0.cpu() - ref[:, :, 0]).abs().max()) / scale,
                      "e2e_vs_oracle_normwise": float((e2e["out"] - ref[:, :, 0]).abs().max()) / scale, "bar": 1e-5}
        line = {
            "metric": "triplane decode queries/s (3-plane bilinear sample + sum, triplane_occ occupancy decode)",
            "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs/triplane_occ.py occupancy decode: {Q} voxel queries ({args.queries}) "
                                   f"x C={C_DEC} from 3 fp32 {PLANE}x{PLANE} triplanes, bs=1 per GPU",
                       "queries": args.queries, "Q": Q, "C": C_DEC, "planes": [PLANE, PLANE],
                       "query_tensor": (f"[1,{','.join(map(str, QUERY_DIMS[args.queries]))},3] through the 5-D entry point "
                                        "tp_sample3_grid_nhwc_f32 (per-block lattice detection on the device)"
                                        if QUERY_DIMS[args.queries] else "[1,Q,3] point list through tp_sample3_nhwc_f32"),
                       "step": "NCHW->NHWC conversion of the 3 planes (1 launch) + fused gather kernel (1 launch), CUDA-graph replay",
                       "l2": f"{dec['nsets']} rotating buffer sets, total footprint "
                             f"{dec['nsets'] * (kbytes + 4 * C_DEC * 3 * PLANE * PLANE) / 1e6:.0f} MB > 3x L2 (no flush kernel)",
                       "parallelism": f"queries sharded over {world} GPU(s), no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(kernel_name.split("<")[0].replace("tp::", ""), args.queries), "kernel": kernel_name,
                         "algorithmic_bytes": kbytes,
                         "kernel_ms_avg": k_avg, "kernel_ms_median": dec["kernel_ms_med"],
                         "kernel_ms_min": dec["kernel_ms_min"], "peak_source": peak_src,
                         "step_frac": kbytes / (ms_per_step * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "torch_cuda_reference": (None if torch_ms is None else {
                "value": Q / (torch_ms * 1e-3), "unit": "queries/s", "ms_per_step": torch_ms,
                "what": "the reference's op sequence (normalise + 3 x F.grid_sample + sum) run by torch-CUDA on this GPU, "
                        "same queries and planes; reported, not a target"}),
            "e2e": {"value": world * Q / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e_ms,
                    "steps": e2e["steps"], "numa_node_rank0": numa, "api": "tp_sample3_grid_host_f32 / tp_sample3_host_f32 (C ABI, pinned host buffers; H2D planes+queries, "
                                                  "D2H full result, synchronised every step)"},
            "gpu_launches": dec["launches"],
            "variants_kernel_only": {k: dict(v, frac=v["achieved_gbs"] / peak) for k, v in dec["variants"].items()},
            "clocks": clocks,
            "parity": parity,
            "encode": {"metric": "triplane encode points/s (fused crop+index+scatter-max into dense pooled planes)",
                       "value": world * enc["n"] / (enc_ms * 1e-3), "unit": "points/s", "ms_per_step": enc_ms,
                       "workload": f"configs/point_triplane.py geometry 128x128x80 pool 5/5/4 C=128, 1 sweep "
                                   f"{enc['n']} raw pts ({enc['inside']} in range), bs=1 per GPU, "
                                   f"{enc['cells']} pooled cells dense out",
                       "roofline": {"bound": "hbm", "achieved": enc["bytes"] / (enc_ms * 1e-3) / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": enc["bytes"] / (enc_ms * 1e-3) / 1e9 / peak,
                                    "algorithmic_bytes": enc["bytes"],
                                    "traffic": ncu_traffic("encode_reduce_kernel", "S1_geomA_C128"),
                                    "note": "whole step: count + scan + fill + reduce (4 launches)"},
                       "steps": enc["steps"], "gpu_launches": enc["launches"]},
        }
        line["encode_dense"] = {
            "metric": "triplane encode points/s, 10-sweep samples (BASELINE.json configs[4] shape, sample-sharded)",
            "value": world * encd["n"] / (encd_ms * 1e-3), "unit": "points/s", "ms_per_step": encd_ms,
            "workload": f"{encd['B']} samples x ~350k raw pts per GPU ({encd['n']} pts, {encd['inside']} in range), geometry "
                        f"128x128x80 C=128, one batched call, dense pooled output",
            "roofline": {"bound": "hbm", "achieved": encd["bytes"] / (encd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": encd["bytes"] / (encd_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": encd["bytes"]},
            "steps": encd["steps"], "gpu_launches": encd["launches"]}
        line["lift"] = {
            "metric": "camera->point lift points/s (point_to_cam: 6 cameras x [768,16,32] maps, bilinear gather + camera sum)",
            "value": world * lift["n"] / (lift_ms * 1e-3), "unit": "points/s", "ms_per_step": lift_ms,
            "kernel_ms": lift_k_ms, "workload": f"bs=2 x 34720 pts per GPU, Cf=768; step = channels-last copy + lift kernel",
            "roofline": {"bound": "hbm", "achieved": lift["bytes"] / (lift_k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": lift["bytes"] / (lift_k_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": lift["bytes"],
                         "note": "lift kernel alone (events around the loop of launches)"},
            "steps": lift["steps"], "gpu_launches": lift["launches"]}
        line["occupancy_head"] = {
            "metric": "Mlp occupancy head queries/s (32 -> 64 -> 32 -> 5 per query, fused on the tensor cores: tcgen05 "
                      "kind::tf32, TMEM accumulators) and decode + head",
            "value": world * occ["Q"] / (occ["head_ms"] * 1e-3), "unit": "queries/s", "ms_per_step": occ["head_ms"],
            "decode_plus_head_ms": occ["both_ms"], "decode_plus_head_queries_per_s": world * occ["Q"] / (occ["both_ms"] * 1e-3),
            "fused_decode_head_ms": occ["fused_ms"],
            "fused_decode_head_queries_per_s": world * occ["Q"] / (occ["fused_ms"] * 1e-3),
            "fused_note": "tp_sample3_grid_head_tf32: layout conversion + ONE kernel from planes and queries to logits (the "
                          "[B,32,Q] features never reach HBM); bit-identical to decode_plus_head (asserted in this run); "
                          "every leg: CUDA-graph replays over 4 rotating buffer sets (> 3x L2), layout conversion included",
            "workload": f"{occ['Q']} queries (640k lattice), C=32, 5 classes; per-rank numbers, no cross-rank max",
            "roofline": {"bound": "hbm", "achieved": occ["head_bytes"] / (occ["head_ms"] * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": occ["head_bytes"] / (occ["head_ms"] * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes": occ["head_bytes"],
                         "tensor_tflops": occ["flops"] / (occ["head_ms"] * 1e-3) / 1e12,
                         "note": "HBM-bound by a wide margin: 8.5 kflop per 148 bytes; the tensor cores are there to keep "
                                 "the 2C / C wide intermediates on the SM, not for their peak"},
            "steps": occ["steps"], "gpu_launches": occ["steps"]}
        if eps:
            line["encode_point_sharded"] = {
                "workload": f"ONE 10-sweep sample ({eps['n']} raw pts) point-sharded over {world} GPUs, geometry "
                            f"128x128x80 C=128; strategy 'planes': partial planes -> NCCL all-reduce(max) of {eps['allreduce_bytes'] / 1e6:.0f} MB "
                            f"-> finalise; strategy 'points': NCCL all-gather of the point shards -> full encode on every rank "
                            f"(strong scaling of one sample; value = the faster; the sample-sharded path above needs no collective)",
                "value": eps["n"] / (min(eps_ms, eps_pts_ms) * 1e-3), "unit": "points/s",
                "ms_per_step": min(eps_ms, eps_pts_ms), "steps": eps["steps"],
                "strategy_planes": {"ms_per_step": eps_ms, "allreduce_bytes_per_step": eps["allreduce_bytes"]},
                "strategy_points": {"ms_per_step": eps_pts_ms, "allgather_bytes_per_step": eps["gather_bytes"],
                                    "what": "all-gather of the point shards (12 + 4C bytes per point), full fused encode on "
                                            "every rank: same planes, no partial planes / finalise pass"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from efficient_multimodal_perception_b200 import synth
    from oracle import triplane_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    q_host = decode_queries(args.queries)
    Q = q_host.shape[1]
    tri = synth.triplane_stacked(1, C_DEC, PLANE, seed=1002)
    t0 = time.perf_counter()
    O.sample_points_triplane_stacked(tri, q_host.view(1, 1, -1, 3), OCC_LO, OCC_VS)
    t_full = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    Qs = int(max(4096, min(Q, Q * budget / max(t_full, 1e-6))))
    sample = q_host[:, torch.randperm(Q, generator=torch.Generator().manual_seed(0))[:Qs]].contiguous().view(1, 1, -1, 3)
    for _ in range(args.warmup):
        O.sample_points_triplane_stacked(tri, sample, OCC_LO, OCC_VS)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.sample_points_triplane_stacked(tri, sample, OCC_LO, OCC_VS)
    dt = time.perf_counter() - t0
    qps = Qs * args.steps / dt
    desc = (f"each step = {Qs} of the {Q} queries (random subset, seed 0) through "
            f"oracle.sample_points_triplane_stacked (the reference's normalise + 3 x F.grid_sample + sum on torch-CPU)")
    print(json.dumps({
        "impl": "reference",
        "metric": "triplane decode queries/s (3-plane bilinear sample + sum, triplane_occ occupancy decode)",
        "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs/triplane_occ.py occupancy decode: {Q} voxel queries ({args.queries}) "
                               f"x C={C_DEC} from 3 fp32 {PLANE}x{PLANE} triplanes, bs=1", "queries": args.queries,
                   "Q": Q, "C": C_DEC, "sample_per_step": Qs},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": desc},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", default="lattice640k", choices=["lattice640k", "uniform640k", "roi"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
