"""Stage timing of the 'owner' point-sharded encode (run under torchrun): reset / barrier / route / barrier / encodes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C_
import torch
import torch.distributed as dist
from efficient_multimodal_perception_b200 import _lib as L, dist as tpd, ops, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
G = synth.GEOM_A
C = G["channels"]
pts = synth.multi_sweep(10, 35000, seed=1005)[:, :3].contiguous()
feats = synth.point_features(pts.shape[0], C, seed=1005)
lo, hi = tpd.shard_bounds(pts.shape[0], rank, world)
my_p, my_f = pts[lo:hi].to(dev), feats[lo:hi].contiguous().to(dev)
off = synth.batch_offsets([hi - lo]).to(dev)
cap = pts.shape[0]

BAL = len(sys.argv) > 1 and sys.argv[1] == "balanced"
bounds = tpd.balanced_slab_bounds(my_p, G["pc_range"], G["voxel_size"], G["grid_size"],
                                  min_width=ops.pool_kernels(G["grid_size"], G["split"])[:2]) if BAL else None


def step():
    return tpd.encode_point_sharded(my_f, my_p, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], strategy="owner",
                                    capacity=cap, slab_bounds=bounds)

for _ in range(5):
    step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    step()
host = (time.perf_counter() - t0) / 20
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 20
# stages, by hand
ex = tpd._OwnerExchange.get(dev, None, cap, C)
X, Y, Z = G["grid_size"]
xb = bounds[0] if BAL else [tpd.shard_bounds(X, r, world)[0] for r in range(world)] + [X]
yb = bounds[1] if BAL else [tpd.shard_bounds(Y, r, world)[0] for r in range(world)] + [Y]
pool = ops.pool_kernels(G["grid_size"], G["split"])
geom = L.make_geom(G["pc_range"], G["voxel_size"], G["grid_size"], pool)
xba, yba = (C_.c_int32 * (world + 1))(*xb), (C_.c_int32 * (world + 1))(*yb)
stream = torch.cuda.current_stream().cuda_stream
# stage times from events recorded in a free-running loop (the host runs ahead: no launch gaps inside the stages)
NIT = 20
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(NIT)]
torch.cuda.synchronize(); dist.barrier()
for it in range(NIT):
    ev = evs[it]
    ev[0].record()
    ex.head32.fill_(-1); ex.head32[:64].zero_()
    ev[1].record()
    ex.hdl.barrier(channel=0)
    ev[2].record()
    L.check(L.lib().tp_route_points_f32(my_p.data_ptr(), 3, my_f.data_ptr(), C, C, hi - lo, C_.byref(geom), 0, world, xba, yba, ex.p_cnt,
                                        ex.p_idx_x, ex.p_feat_x, ex.p_idx_y, ex.p_feat_y, ex.cap, stream), "route")
    ev[3].record()
    ex.hdl.barrier(channel=1)
    ev[4].record()
    xs, ys = xb[rank + 1] - xb[rank], yb[rank + 1] - yb[rank]
    a = ops.encode(ex.feat_x, ex.offsets, [0] * 6, (1, 1, 1), (xs, Y, Z), G["split"], grid_ind=ex.idx_x, planes=(True, False, True), pool=pool)
    ev[5].record()
    b = ops.encode(ex.feat_y, ex.offsets, [0] * 6, (1, 1, 1), (X, ys, Z), G["split"], grid_ind=ex.idx_y, planes=(False, True, False), pool=pool)
    ev[6].record()
torch.cuda.synchronize()
acc = [sum(evs[it][k].elapsed_time(evs[it][k + 1]) for it in range(5, NIT)) / (NIT - 5) for k in range(6)]
cnt = ex.buf[:8].view(torch.int32).tolist()
print(f"rank {rank}: host-issue {host*1e3:.3f} ms, wall {wall*1e3:.3f} ms per step; stages ms: reset {acc[0]:.3f} barrier {acc[1]:.3f} route {acc[2]:.3f} "
      f"barrier {acc[3]:.3f} encode_x {acc[4]:.3f} encode_y {acc[5]:.3f}; received rows x/y {cnt[:2]}", flush=True)
dist.destroy_process_group()
