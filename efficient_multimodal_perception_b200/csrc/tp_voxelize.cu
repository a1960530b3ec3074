// a1: crop + voxel index + stable stream compaction, batched over samples.
// Replaces PointTriplane.voxelize_points (point_triplane.py:133-161): per sample ~10 elementwise
// launches + a boolean-mask gather (a device sync each). Here: three launches for the whole batch
// and no sync; the caller reads out_offsets[B] once when it needs N'.
#include "tp_common.cuh"

namespace tp {

constexpr int kVoxBlock = 256;
constexpr int kVoxPerThread = 8;
constexpr int kVoxTile = kVoxBlock * kVoxPerThread;  // points per CTA, contiguous

template <int ARITH>
__global__ void __launch_bounds__(kVoxBlock)
voxel_index_kernel(const float* __restrict__ pts, int64_t n, int stride, GeomDev g,
                   uint8_t* __restrict__ keep, int32_t* __restrict__ idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = pts + i * stride;
    int ix, iy, iz;
    bool k = tp_crop_index<ARITH>(g, __ldg(p), __ldg(p + 1), __ldg(p + 2), ix, iy, iz);
    keep[i] = k ? 1 : 0;
    idx[i * 3 + 0] = ix;
    idx[i * 3 + 1] = iy;
    idx[i * 3 + 2] = iz;
  }
}

// pass 1: kept points per tile
__global__ void __launch_bounds__(kVoxBlock)
vox_count_kernel(const float* __restrict__ pts, int64_t n, int stride, GeomDev g,
                 int32_t* __restrict__ tile_count) {
  __shared__ int s_warp[kVoxBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kVoxTile;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < kVoxPerThread; ++j) {
    int64_t i = base + j * kVoxBlock + threadIdx.x;
    if (i < n) {
      const float* p = pts + i * stride;
      float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
      cnt += (x > g.lo[0]) & (x < g.hi[0]) & (y > g.lo[1]) & (y < g.hi[1]) & (z > g.lo[2]) &
             (z < g.hi[2]);
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kVoxBlock / 32; ++w) t += s_warp[w];
    tile_count[blockIdx.x] = t;
  }
}

// pass 2: exclusive scan of tile counts (single CTA, int64 running total) + per-sample offsets.
__global__ void __launch_bounds__(1024)
vox_scan_kernel(const int32_t* __restrict__ tile_count, int64_t ntiles,
                int64_t* __restrict__ tile_off /* [ntiles+1] */) {
  __shared__ int64_t s_part[32];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < ntiles; base += 1024) {
    int64_t i = base + threadIdx.x;
    int64_t v = i < ntiles ? tile_count[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_part[lane];
      int64_t wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      s_part[lane] = wi - w;  // exclusive warp prefix
    }
    __syncthreads();
    int64_t excl = s_carry + s_part[warp] + incl - v;
    if (i < ntiles) tile_off[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_off[ntiles] = s_carry;
}

// pass 3: stable write. Warp w of a tile owns rows [w*256, (w+1)*256) and walks them 32 at a time,
// so loads are coalesced and the rank of a kept row is
//   tile_off + kept in earlier warps + kept in earlier 32-row steps + kept lanes below me.
template <int ARITH>
__global__ void __launch_bounds__(kVoxBlock)
vox_write_kernel(const float* __restrict__ pts, int64_t n, int stride, int ncols, GeomDev g,
                 const int64_t* __restrict__ tile_off, float* __restrict__ out_pts,
                 int32_t* __restrict__ out_idx) {
  __shared__ int s_warp[kVoxBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wfirst = (int64_t)blockIdx.x * kVoxTile + (int64_t)warp * (32 * kVoxPerThread);
  int ix[kVoxPerThread], iy[kVoxPerThread], iz[kVoxPerThread];
  unsigned ballots[kVoxPerThread];
  int wtotal = 0;
#pragma unroll
  for (int j = 0; j < kVoxPerThread; ++j) {
    int64_t i = wfirst + j * 32 + lane;
    bool k = false;
    if (i < n) {
      const float* p = pts + i * stride;
      k = tp_crop_index<ARITH>(g, __ldg(p), __ldg(p + 1), __ldg(p + 2), ix[j], iy[j], iz[j]);
    }
    ballots[j] = __ballot_sync(0xffffffffu, k);
    wtotal += __popc(ballots[j]);
  }
  if (lane == 0) s_warp[warp] = wtotal;
  __syncthreads();
  int wpre = 0;
#pragma unroll
  for (int w = 0; w < kVoxBlock / 32; ++w) wpre += (w < warp) ? s_warp[w] : 0;
  int64_t rank = tile_off[blockIdx.x] + wpre;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < kVoxPerThread; ++j) {
    if (ballots[j] & (1u << lane)) {
      const int64_t i = wfirst + j * 32 + lane;
      const int64_t r = rank + __popc(ballots[j] & lt);
      const float* p = pts + i * stride;
      float* o = out_pts + r * ncols;
      for (int c = 0; c < ncols; ++c) o[c] = __ldg(p + c);
      out_idx[r * 3 + 0] = ix[j];
      out_idx[r * 3 + 1] = iy[j];
      out_idx[r * 3 + 2] = iz[j];
    }
    rank += __popc(ballots[j]);
  }
}

// compacted sample boundaries: out_offsets[b] = kept rows before raw row in_offsets[b].
// One CTA per boundary recounts the (< kVoxTile) rows between the tile start and the boundary.
__global__ void __launch_bounds__(kVoxBlock)
vox_sample_offsets_kernel(const float* __restrict__ pts, int64_t n, int stride, GeomDev g,
                          const int64_t* __restrict__ in_offsets,
                          const int64_t* __restrict__ tile_off, int64_t ntiles,
                          int64_t* __restrict__ out_offsets) {
  __shared__ int s_warp[kVoxBlock / 32];
  int64_t off = in_offsets[blockIdx.x];
  if (off < 0) off = 0;
  if (off >= n) {
    if (threadIdx.x == 0) out_offsets[blockIdx.x] = tile_off[ntiles];
    return;
  }
  const int64_t t = off / kVoxTile;
  int cnt = 0;
  for (int64_t i = t * kVoxTile + threadIdx.x; i < off; i += kVoxBlock) {
    const float* p = pts + i * stride;
    float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    cnt += (x > g.lo[0]) & (x < g.hi[0]) & (y > g.lo[1]) & (y < g.hi[1]) & (z > g.lo[2]) &
           (z < g.hi[2]);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < kVoxBlock / 32; ++w) tot += s_warp[w];
    out_offsets[blockIdx.x] = tile_off[t] + tot;
  }
}

}  // namespace tp

using namespace tp;

static int check_geom(const tp_geom* g, const char* who) {
  if (!g) return fail(TP_E_NULL, "%s: geom is null", who);
  for (int a = 0; a < 3; ++a) {
    if (!(g->vs[a] > 0.f) || g->grid[a] <= 0 || g->pool[a] <= 0 || g->pool[a] > g->grid[a])
      return fail(TP_E_SHAPE, "%s: bad geometry on axis %d (vs=%g grid=%d pool=%d)", who, a,
                  (double)g->vs[a], g->grid[a], g->pool[a]);
  }
  return 0;
}

extern "C" int tp_voxel_index_f32(const float* points, int64_t n, int32_t stride,
                                  const tp_geom* geom, int32_t arith, uint8_t* keep, int32_t* idx,
                                  void* stream) {
  if (int rc = check_geom(geom, "tp_voxel_index_f32")) return rc;
  if (n == 0) return 0;
  if (!points || !keep || !idx) return fail(TP_E_NULL, "tp_voxel_index_f32: null argument");
  if (stride < 3 || n < 0) return fail(TP_E_SHAPE, "tp_voxel_index_f32: stride=%d n=%lld", stride, (long long)n);
  GeomDev g = make_geom_dev(*geom);
  int64_t blocks = (n + kVoxBlock - 1) / kVoxBlock;
  int grid = (int)(blocks < kSMs * 8 ? blocks : kSMs * 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (arith == TP_ARITH_TORCH_CUDA)
    voxel_index_kernel<TP_ARITH_TORCH_CUDA><<<grid, kVoxBlock, 0, s>>>(points, n, stride, g, keep, idx);
  else if (arith == TP_ARITH_TORCH_CPU)
    voxel_index_kernel<TP_ARITH_TORCH_CPU><<<grid, kVoxBlock, 0, s>>>(points, n, stride, g, keep, idx);
  else
    return fail(TP_E_ENUM, "tp_voxel_index_f32: unknown arith %d", arith);
  TP_LAUNCH_CHECK("voxel_index_kernel");
  return 0;
}

static int64_t vox_ntiles(int64_t n) { return (n + kVoxTile - 1) / kVoxTile; }

extern "C" int64_t tp_voxelize_workspace_bytes(int64_t n_total) {
  int64_t nt = vox_ntiles(n_total < 0 ? 0 : n_total);
  // tile_off int64 [nt+1] then tile_count int32 [nt], 256-byte aligned regions
  int64_t a = ((nt + 1) * 8 + 255) / 256 * 256;
  int64_t b = (nt * 4 + 255) / 256 * 256;
  return a + b + 256;
}

extern "C" int tp_voxelize_f32(const float* points, int64_t n, int32_t stride, int32_t ncols,
                               const int64_t* in_offsets, int32_t batch, const tp_geom* geom,
                               int32_t arith, float* out_points, int32_t* out_idx,
                               int64_t* out_offsets, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  if (int rc = check_geom(geom, "tp_voxelize_f32")) return rc;
  if (!in_offsets || !out_offsets) return fail(TP_E_NULL, "tp_voxelize_f32: offsets are null");
  if (batch <= 0 || batch > 1023) return fail(TP_E_SHAPE, "tp_voxelize_f32: batch=%d not in [1,1023]", batch);
  if (n < 0 || stride < 3 || ncols < 0 || ncols > stride)
    return fail(TP_E_SHAPE, "tp_voxelize_f32: n=%lld stride=%d ncols=%d", (long long)n, stride, ncols);
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU)
    return fail(TP_E_ENUM, "tp_voxelize_f32: unknown arith %d", arith);
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    TP_CUDA(cudaMemsetAsync(out_offsets, 0, (size_t)(batch + 1) * 8, s));
    return 0;
  }
  if (!points || !out_points || !out_idx || !workspace) return fail(TP_E_NULL, "tp_voxelize_f32: null argument");
  if (workspace_bytes < tp_voxelize_workspace_bytes(n))
    return fail(TP_E_WORKSPACE, "tp_voxelize_f32: workspace %lld < %lld bytes", (long long)workspace_bytes,
                (long long)tp_voxelize_workspace_bytes(n));
  const int64_t nt = vox_ntiles(n);
  if (nt > 0x7fffffff) return fail(TP_E_SHAPE, "tp_voxelize_f32: too many points");
  int64_t* tile_off = reinterpret_cast<int64_t*>(workspace);
  int32_t* tile_cnt = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) +
                                                 ((nt + 1) * 8 + 255) / 256 * 256);
  GeomDev g = make_geom_dev(*geom);
  vox_count_kernel<<<(int)nt, kVoxBlock, 0, s>>>(points, n, stride, g, tile_cnt);
  TP_LAUNCH_CHECK("vox_count_kernel");
  vox_scan_kernel<<<1, 1024, 0, s>>>(tile_cnt, nt, tile_off);
  TP_LAUNCH_CHECK("vox_scan_kernel");
  if (arith == TP_ARITH_TORCH_CUDA)
    vox_write_kernel<TP_ARITH_TORCH_CUDA><<<(int)nt, kVoxBlock, 0, s>>>(
        points, n, stride, ncols, g, tile_off, out_points, out_idx);
  else
    vox_write_kernel<TP_ARITH_TORCH_CPU><<<(int)nt, kVoxBlock, 0, s>>>(
        points, n, stride, ncols, g, tile_off, out_points, out_idx);
  TP_LAUNCH_CHECK("vox_write_kernel");
  vox_sample_offsets_kernel<<<batch + 1, kVoxBlock, 0, s>>>(points, n, stride, g, in_offsets, tile_off,
                                                           nt, out_offsets);
  TP_LAUNCH_CHECK("vox_sample_offsets_kernel");
  return 0;
}
