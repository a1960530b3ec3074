// Point-sharded encode, "owner" exchange (encode by point): every rank PUSHES each of its points straight
// into the receive buffers of the ranks that own the point's output rows, over NVLink peer memory, in one kernel:
//   * rank r owns the x-slab [xb[r], xb[r+1]) of the xy and xz planes and the y-slab [yb[r], yb[r+1]) of the yz plane,
//     so a point has at most two destinations and every pooled cell has exactly one owner: no partial planes, no
//     all-reduce of 430 MB, no replicated output;
//   * a destination slot comes from a system-scope atomicAdd on the owner's counter (peer memory), the row (voxel
//     index relative to the slab + the C feature floats) is written with plain peer stores;
//   * the owners then run the ordinary fused encode on what they received (rows never written keep index -1 and are
//     dropped), each producing its slab of the three planes.
// The host side (dist.py) allocates the buffers as torch symmetric memory, resets them, and brackets this kernel with
// two device-side barriers. The reference has no counterpart (data parallel only, tools/euler_train.sh:3-11).
#include "tp_common.cuh"

namespace tp {

constexpr int kRouteMaxWorld = 8;

struct RouteParams {
  const float* points;
  const float* feats;
  int64_t n, feat_stride, cap;
  int point_stride, C4, world;
  GeomDev g;
  int xb[kRouteMaxWorld + 1], yb[kRouteMaxWorld + 1];
  int32_t* cnt[kRouteMaxWorld];     // [2]: rows received for the x-slab planes, for the y-slab plane
  int32_t* idx_x[kRouteMaxWorld];   // [cap, 3]
  float* feat_x[kRouteMaxWorld];    // [cap, C]
  int32_t* idx_y[kRouteMaxWorld];
  float* feat_y[kRouteMaxWorld];
};

__device__ __forceinline__ int slab_owner(const int* b, int world, int i) {
  int r = 0;
  while (r + 1 < world && i >= b[r + 1]) ++r;
  return r;
}

// One warp per 32 points. Slots are claimed with ONE system-scope atomic per (warp, destination) — a remote atomic is
// a multi-microsecond NVLink round trip, so per-point atomics would serialise the kernel — then the warp copies the 32
// feature rows (all lanes on one 4*C-byte row at a time: full-width peer stores).
template <int ARITH>
__global__ void __launch_bounds__(256)
route_points_kernel(const RouteParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int C = P.C4 * 4;
  for (int64_t i0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; i0 < P.n; i0 += nwarp * 32) {
    const int64_t i = i0 + lane;
    int ix = 0, iy = 0, iz = 0;
    bool keep = false;
    if (i < P.n) {
      const float* p = P.points + i * P.point_stride;
      keep = tp_crop_index<ARITH>(P.g, __ldg(p), __ldg(p + 1), __ldg(p + 2), ix, iy, iz);
      keep = keep & (ix >= 0) & (ix < P.g.grid[0]) & (iy >= 0) & (iy < P.g.grid[1]) & (iz >= 0) & (iz < P.g.grid[2]);
    }
    const int dx = keep ? slab_owner(P.xb, P.world, ix) : -1, dy = keep ? slab_owner(P.yb, P.world, iy) : -1;
    // slots: lane r claims the x-slab rows of destination r, lane 8 + r its y-slab rows — all (at most 16) remote atomics
    // of the warp are in flight at once (issued one destination after the other they were up to 16 dependent NVLink
    // round trips per 32 points)
    int pre_x = 0, pre_y = 0, mine = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int r = 0; r < P.world; ++r) {
      const unsigned mx = __ballot_sync(0xffffffffu, dx == r), my = __ballot_sync(0xffffffffu, dy == r);
      if (dx == r) pre_x = __popc(mx & lt);
      if (dy == r) pre_y = __popc(my & lt);
      if (lane == r) mine = __popc(mx);
      if (lane == kRouteMaxWorld + r) mine = __popc(my);
    }
    int base = 0;
    if (mine) base = atomicAdd_system(P.cnt[lane & (kRouteMaxWorld - 1)] + (lane >= kRouteMaxWorld ? 1 : 0), mine);
    const int bx = __shfl_sync(0xffffffffu, base, dx < 0 ? 0 : dx);
    const int by = __shfl_sync(0xffffffffu, base, kRouteMaxWorld + (dy < 0 ? 0 : dy));
    int sx = dx >= 0 ? bx + pre_x : -1, sy = dy >= 0 ? by + pre_y : -1;
    if (sx >= P.cap) sx = -1;  // capacity is the global point count: cannot happen
    if (sy >= P.cap) sy = -1;
    if (sx >= 0) {
      int32_t* d = P.idx_x[dx] + (int64_t)sx * 3;
      d[0] = ix - P.xb[dx]; d[1] = iy; d[2] = iz;
    }
    if (sy >= 0) {
      int32_t* d = P.idx_y[dy] + (int64_t)sy * 3;
      d[0] = ix; d[1] = iy - P.yb[dy]; d[2] = iz;
    }
    // feature rows: the warp walks its kept points four at a time (four row loads in flight); 32 lanes x 16 B per row
    // and destination
    unsigned m = __ballot_sync(0xffffffffu, keep);
    while (m) {
      int src[4];
      float4* ox[4];
      float4* oy[4];
      const float4* frow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        src[j] = m ? __ffs(m) - 1 : -1;
        m &= m - 1;  // 0 stays 0
        const int sl = src[j] < 0 ? 0 : src[j];
        const int pdx = __shfl_sync(0xffffffffu, dx, sl), pdy = __shfl_sync(0xffffffffu, dy, sl);
        const int psx = __shfl_sync(0xffffffffu, sx, sl), psy = __shfl_sync(0xffffffffu, sy, sl);
        frow[j] = reinterpret_cast<const float4*>(P.feats + (i0 + sl) * P.feat_stride);
        ox[j] = (src[j] >= 0 && psx >= 0) ? reinterpret_cast<float4*>(P.feat_x[pdx] + (int64_t)psx * C) : nullptr;
        oy[j] = (src[j] >= 0 && psy >= 0) ? reinterpret_cast<float4*>(P.feat_y[pdy] + (int64_t)psy * C) : nullptr;
      }
      for (int v = lane; v < P.C4; v += 32) {
        float4 f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (src[j] >= 0) f[j] = __ldg(frow[j] + v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ox[j]) ox[j][v] = f[j];
          if (oy[j]) oy[j][v] = f[j];
        }
      }
    }
  }
}

}  // namespace tp

using namespace tp;

extern "C" int tp_route_points_f32(const float* points, int32_t point_stride, const float* feats, int64_t feat_stride,
                                   int32_t C, int64_t n, const tp_geom* geom, int32_t arith, int32_t world,
                                   const int32_t* x_bounds, const int32_t* y_bounds, void* const* peer_cnt,
                                   void* const* peer_idx_x, void* const* peer_feat_x, void* const* peer_idx_y,
                                   void* const* peer_feat_y, int64_t capacity, void* stream) {
  if (!geom || !x_bounds || !y_bounds || !peer_cnt || !peer_idx_x || !peer_feat_x || !peer_idx_y || !peer_feat_y)
    return fail(TP_E_NULL, "tp_route_points_f32: null argument");
  if (world < 1 || world > kRouteMaxWorld) return fail(TP_E_SHAPE, "tp_route_points_f32: world=%d must be in 1..%d", world, kRouteMaxWorld);
  if (C <= 0 || (C & 3) || n < 0 || capacity <= 0 || point_stride < 3) return fail(TP_E_SHAPE, "tp_route_points_f32: bad C=%d n=%lld", C, (long long)n);
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_route_points_f32: unknown arith %d", arith);
  for (int a = 0; a < 3; ++a)
    if (!(geom->vs[a] > 0.f) || geom->grid[a] <= 0) return fail(TP_E_SHAPE, "tp_route_points_f32: bad geometry on axis %d", a);
  if (n == 0) return 0;
  if (!points || !feats || (feat_stride & 3) || ((uintptr_t)feats & 15)) return fail(TP_E_SHAPE, "tp_route_points_f32: feats must be 16-byte aligned rows");
  RouteParams P;
  P.points = points; P.feats = feats; P.n = n; P.feat_stride = feat_stride; P.cap = capacity;
  P.point_stride = point_stride; P.C4 = C / 4; P.world = world;
  P.g = make_geom_dev(*geom);
  for (int r = 0; r <= world; ++r) { P.xb[r] = x_bounds[r]; P.yb[r] = y_bounds[r]; }
  if (P.xb[0] != 0 || P.xb[world] != geom->grid[0] || P.yb[0] != 0 || P.yb[world] != geom->grid[1])
    return fail(TP_E_SHAPE, "tp_route_points_f32: slab bounds must cover [0, X) and [0, Y)");
  for (int r = 0; r < world; ++r) {
    if (!peer_cnt[r] || !peer_idx_x[r] || !peer_feat_x[r] || !peer_idx_y[r] || !peer_feat_y[r])
      return fail(TP_E_NULL, "tp_route_points_f32: peer %d has a null buffer", r);
    P.cnt[r] = (int32_t*)peer_cnt[r];
    P.idx_x[r] = (int32_t*)peer_idx_x[r]; P.feat_x[r] = (float*)peer_feat_x[r];
    P.idx_y[r] = (int32_t*)peer_idx_y[r]; P.feat_y[r] = (float*)peer_feat_y[r];
  }
  const int64_t wb = (n + 255) / 256;   // 8 warps per CTA, 32 points per warp
  const int grid = (int)(wb < (int64_t)kSMs * 8 ? wb : (int64_t)kSMs * 8);
  if (arith == TP_ARITH_TORCH_CUDA) route_points_kernel<TP_ARITH_TORCH_CUDA><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  else route_points_kernel<TP_ARITH_TORCH_CPU><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("route_points_kernel");
  return 0;
}
