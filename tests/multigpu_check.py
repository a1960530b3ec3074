"""Multi-GPU parity check (run under torchrun, NCCL; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
Point-sharded encode (partial planes + all-reduce + finalise) and query-sharded decode must equal
the single-GPU result: bit-exact for max and for decode, <= 1e-5 normwise for mean."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import dist as tpd  # noqa: E402
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G = synth.GEOM_A
    C = 32
    raw = [synth.lidar_sweep(30000, seed=5)[:, :3].contiguous(), synth.lidar_sweep(20000, seed=6)[:, :3].contiguous()]
    feats = [synth.point_features(30000, C, seed=7) - 0.3, synth.point_features(20000, C, seed=8) - 0.3]
    full_off = synth.batch_offsets([30000, 20000]).to(dev)
    my_raw, my_feats = tpd.shard_points(raw, rank, world), tpd.shard_points(feats, rank, world)
    my_off = synth.batch_offsets([p.shape[0] for p in my_raw]).to(dev)
    ok = True
    for reduce in ("max", "mean"):
        ref = ops.encode(torch.cat(feats).to(dev), full_off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"],
                         points=torch.cat(raw).to(dev), reduce=reduce)
        for strategy in ("planes", "points"):
            got = tpd.encode_point_sharded(torch.cat(my_feats).to(dev), torch.cat(my_raw).to(dev), my_off, G["pc_range"],
                                           G["voxel_size"], G["grid_size"], G["split"], reduce=reduce, strategy=strategy)
            for a, b in zip(got, ref):
                if reduce == "max":
                    good = torch.equal(a, b)
                else:
                    good = float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
                ok &= bool(good)
                if not good:
                    print(f"[rank {rank}] {reduce}/{strategy}: mismatch, max diff {float((a - b).abs().max())}", flush=True)
        # owner exchange over NVLink peer memory: one sample per call; every rank ends up with its slabs of the planes
        ref1 = ops.encode(feats[0].to(dev), synth.batch_offsets([30000]).to(dev), G["pc_range"], G["voxel_size"], G["grid_size"],
                          G["split"], points=raw[0].to(dev), reduce=reduce)
        bal = tpd.balanced_slab_bounds(my_raw[0].to(dev), G["pc_range"], G["voxel_size"], G["grid_size"],
                                       min_width=ops.pool_kernels(G["grid_size"], G["split"])[:2])
        for rep in range(3):  # twice with equal-width slabs (the buffers are reused and must be reset correctly), once balanced
            sb = bal if rep == 2 else None
            got = tpd.encode_point_sharded(my_feats[0].to(dev), my_raw[0].to(dev), synth.batch_offsets([my_raw[0].shape[0]]).to(dev),
                                           G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], reduce=reduce, strategy="owner",
                                           slab_bounds=sb)
            good = tpd.planes_equal(got, ref1, "owner", rank, world, rtol=0.0 if reduce == "max" else 1e-5, slab_bounds=sb)
            ok &= bool(good)
            if not good:
                (x0, x1), (y0, y1) = (sb[0][rank:rank + 2], sb[1][rank:rank + 2]) if sb else (tpd.shard_bounds(128, rank, world),) * 2
                refs = (ref1[0][:, x0:x1], ref1[1][:, y0:y1], ref1[2][:, x0:x1])
                info = []
                for a, b in zip(got, refs):
                    d = (a != b).nonzero()
                    info.append((tuple(a.shape), int(d.shape[0]), d[:2].tolist(), d[-1:].tolist()))
                print(f"[rank {rank}] {reduce}/owner (pass {rep}, bounds {sb}): mismatch {info}", flush=True)
    # C-ABI collective hooks (tp_comm_*, tp_allreduce_planes): same result as torch.distributed
    import ctypes as C_
    from efficient_multimodal_perception_b200 import _lib as L
    lib = L.lib()
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        L.check(lib.tp_comm_unique_id(idbuf.data_ptr()), "tp_comm_unique_id")
    idd = idbuf.to(dev)
    dist.broadcast(idd, 0)
    idbuf = idd.cpu()
    comm = C_.c_void_p()
    L.check(lib.tp_comm_init(C_.byref(comm), world, rank, idbuf.data_ptr()), "tp_comm_init")
    part = torch.randn(1 << 20, device=dev, generator=torch.Generator(dev).manual_seed(100 + rank))
    cnts = torch.full((4096,), rank + 1, dtype=torch.int32, device=dev)
    want_max, want_cnt = part.clone(), cnts.clone()
    dist.all_reduce(want_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(want_cnt, op=dist.ReduceOp.SUM)
    L.check(lib.tp_allreduce_planes(comm, part.data_ptr(), part.numel(), L.TP_REDUCE_MAX_PARTIAL, cnts.data_ptr(), cnts.numel(),
                                    torch.cuda.current_stream().cuda_stream), "tp_allreduce_planes")
    torch.cuda.synchronize()
    ok &= torch.equal(part, want_max) and torch.equal(cnts, want_cnt)
    L.check(lib.tp_comm_destroy(comm), "tp_comm_destroy")
    # decode: each rank samples its slice of the queries; gathered result == full result
    tri = synth.triplane_stacked(1, 32, 128, seed=9).to(dev)
    q = synth.uniform_queries(100003, seed=10)[None].to(dev)
    lo, vs, half = synth.OCC["triplane_range"][:3], synth.OCC["triplane_voxel_size"], [64.0] * 3
    full = ops.sample3(tri, q, lo, vs, half)
    mine = ops.sample3(tri, tpd.shard_queries(q, rank, world).contiguous(), lo, vs, half)
    lo_i, hi_i = tpd.shard_bounds(q.shape[1], rank, world)
    ok &= torch.equal(mine, full[:, :, lo_i:hi_i])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if int(flag) == 1 else "FAIL", f"world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
