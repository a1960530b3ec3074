// a3: triplane encode. One pass links every point into the (pooled cell -> points) lists of the three
// planes with a single 4-byte atomicExch per (point, plane); a second pass streams over the dense
// channels-last outputs ONCE, writing zeros for empty cells and the gathered max / mean for
// occupied ones.
//
// Replaces point_triplane_projector.py:99-115:
//   torch.unique(dim=0) -> torch_scatter.scatter_max -> 3 x spconv.SparseMaxPool3d -> .dense()
//   -> permute(...).flatten(3)
// (a sort, three hash builds and ~6 passes over 430 MB per sample) with N'x3 integer atomics on an
// L2-resident 4 B/cell table and exactly one write of every output byte. max over the points of a
// voxel followed by max over the voxels of a pooling window == max over the points of the pooled
// cell, so the intermediate voxel tensor never exists. No floating-point atomics are issued.
#include <atomic>

#include "tp_common.cuh"

namespace tp {

struct EncodeParams {
  GeomDev g;
  int batch;
  int C, C4;
  int64_t n;
  // plane k: cells per sample, and the dims (D0, D1, P) of out_k [B, D0, D1, P*C]
  int64_t cells_per_sample[3];
  int64_t head_base[3];  // offset of plane k's head entries (sample-major inside the plane)
  const int32_t* idx;
  const float* points;
  int point_stride;
  const int64_t* offsets;
  const float* feats;
  int64_t feat_stride;
  int32_t* head;
  int32_t* next;  // [3][n]
  float* out[3];
  int32_t* cell_count;  // optional, order xy | yz | xz, each [B, cells_per_sample]
  int64_t count_base[3];
  int clamp_zero;
};

// ---- pass 1: link ----------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(256)
encode_link_kernel(const EncodeParams P) {
  const GeomDev& g = P.g;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    bool keep;
    if (P.idx) {
      ix = __ldg(P.idx + i * 3);
      iy = __ldg(P.idx + i * 3 + 1);
      iz = __ldg(P.idx + i * 3 + 2);
      keep = true;
    } else {
      const float* p = P.points + i * P.point_stride;
      keep = tp_crop_index<ARITH>(g, __ldg(p), __ldg(p + 1), __ldg(p + 2), ix, iy, iz);
    }
    // spconv would index out of bounds for idx outside the grid; we drop the point.
    keep = keep & (ix >= 0) & (ix < g.grid[0]) & (iy >= 0) & (iy < g.grid[1]) & (iz >= 0) &
           (iz < g.grid[2]);
    if (!keep) continue;
    const int64_t b = tp_find_batch(P.offsets, P.batch, i);
    const int px = ix / g.pool[0], py = iy / g.pool[1], pz = iz / g.pool[2];
    // plane xy: [B, X, Y, Zp]   (pool_xy(...).dense().permute(0,2,3,4,1), projector.py:113)
    if (P.out[0] && pz < g.pooled[2]) {
      int64_t cell = P.head_base[0] + b * P.cells_per_sample[0] +
                     ((int64_t)ix * g.grid[1] + iy) * g.pooled[2] + pz;
      P.next[i] = atomicExch(P.head + cell, (int)i);
    }
    // plane yz: [B, Y, Z, Xp]   (permute(0,3,4,2,1), projector.py:114)
    if (P.out[1] && px < g.pooled[0]) {
      int64_t cell = P.head_base[1] + b * P.cells_per_sample[1] +
                     ((int64_t)iy * g.grid[2] + iz) * g.pooled[0] + px;
      P.next[P.n + i] = atomicExch(P.head + cell, (int)i);
    }
    // plane xz: [B, X, Z, Yp]   (permute(0,2,4,3,1), projector.py:115)
    if (P.out[2] && py < g.pooled[1]) {
      int64_t cell = P.head_base[2] + b * P.cells_per_sample[2] +
                     ((int64_t)ix * g.grid[2] + iz) * g.pooled[1] + py;
      P.next[2 * P.n + i] = atomicExch(P.head + cell, (int)i);
    }
  }
}

__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(max_nan(a.x, b.x), max_nan(a.y, b.y), max_nan(a.z, b.z), max_nan(a.w, b.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z),
                     __fadd_rn(a.w, b.w));
}

// ---- pass 2: materialise ------------------------------------------------------------------------
// Work unit = kUnitCells consecutive cells of one (sample, plane); units are ordered sample-major
// (xy, yz, xz of sample 0, then sample 1, ...) so the three planes of a sample re-read its feature
// rows while they are still L2-resident. Warps pull units from a global counter: a unit with
// occupied cells costs one dependent load chain per point, an empty unit is 4 KB of streaming
// stores, so a static split leaves the few dense units as a long tail (ncu, profiles/).
// VPL = float4 per lane per cell (C <= 128*VPL).
constexpr int kUnitCells = 8;
constexpr int kUnitsPerGrab = 4;
constexpr int kEncSchedSlots = 256;
__device__ unsigned long long g_enc_sched[kEncSchedSlots][2];  // [next unit, finished warps]

template <int REDUCE, int VPL>
__global__ void __launch_bounds__(256)
encode_materialize_kernel(const EncodeParams P, int64_t units_per_sample0, int64_t units_per_sample1,
                          int64_t units_per_sample2, int sched_slot) {
  const int lane = threadIdx.x & 31;
  const int64_t ups[3] = {units_per_sample0, units_per_sample1, units_per_sample2};
  const int64_t ups_all = ups[0] + ups[1] + ups[2];
  const int64_t total = ups_all * P.batch;
  const int C4 = P.C4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long* sched = g_enc_sched[sched_slot];

  for (;;) {
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(&sched[0], (unsigned long long)kUnitsPerGrab);
    const int64_t u0 = (int64_t)__shfl_sync(0xffffffffu, g, 0);
    if (u0 >= total) break;
    const int64_t u1 = min(total, u0 + kUnitsPerGrab);
    for (int64_t u = u0; u < u1; ++u) {
      const int64_t b = u / ups_all;
      int64_t r = u - b * ups_all;
      int k = 0;
      if (r >= ups[0]) { r -= ups[0]; k = 1; if (r >= ups[1]) { r -= ups[1]; k = 2; } }
      const int64_t cps = P.cells_per_sample[k];
      const int64_t c0 = r * kUnitCells;  // first cell of the unit inside the sample
      const int ncell = (int)min((int64_t)kUnitCells, cps - c0);
      int32_t* hp = P.head + P.head_base[k] + b * cps + c0;
      int h = -1;
      if (lane < ncell) {
        h = hp[lane];
        if (h >= 0) hp[lane] = -1;  // leave the table clean for the next call
      }
      float4* o = reinterpret_cast<float4*>(P.out[k]) + (b * cps + c0) * C4;
      const int32_t* nxt = P.next + (int64_t)k * P.n;
      const unsigned occ = __ballot_sync(0xffffffffu, h >= 0);
      int mycount = 0;

      if (occ == 0) {  // common case: ncell*C4 contiguous float4 of zeros
        const int nvec = ncell * C4;
        for (int v = lane; v < nvec; v += 32) st_cs_f4(o + v, zero);
      } else {
        for (int j = 0; j < ncell; ++j) {
          int p = __shfl_sync(0xffffffffu, h, j);
          float4 acc[VPL];
#pragma unroll
          for (int t = 0; t < VPL; ++t) acc[t] = zero;
          int cnt = 0;
          if (p >= 0) {
            if (REDUCE == TP_REDUCE_MAX) {
              bool first = true;
              while (p >= 0) {
                const float4* f = reinterpret_cast<const float4*>(P.feats + (int64_t)p * P.feat_stride);
                const int pn = __ldg(nxt + p);
#pragma unroll
                for (int t = 0; t < VPL; ++t) {
                  const int v = lane + 32 * t;
                  if (v < C4) {
                    float4 x = __ldg(f + v);
                    acc[t] = first ? x : max4(acc[t], x);
                  }
                }
                first = false;
                ++cnt;
                p = pn;
              }
              if (P.clamp_zero) {
#pragma unroll
                for (int t = 0; t < VPL; ++t) acc[t] = max4(acc[t], zero);
              }
            } else {
              // SUM / MEAN: accumulate in ascending point order inside each run of <= 32 list
              // entries so the result does not depend on the (racy) insertion order for the
              // common case of <= 32 points per cell.
              while (p >= 0) {
                int mine = 0x7fffffff;
                int m = 0;
                for (; m < 32 && p >= 0; ++m) {
                  if (lane == m) mine = p;
                  p = __ldg(nxt + p);
                }
                for (int s = 0; s < m; ++s) {
                  const int id = __reduce_min_sync(0xffffffffu, (unsigned)mine);
                  if (mine == id) mine = 0x7fffffff;
                  const float4* f = reinterpret_cast<const float4*>(P.feats + (int64_t)id * P.feat_stride);
#pragma unroll
                  for (int t = 0; t < VPL; ++t) {
                    const int v = lane + 32 * t;
                    if (v < C4) acc[t] = add4(acc[t], __ldg(f + v));
                  }
                }
                cnt += m;
              }
              if (REDUCE == TP_REDUCE_MEAN) {
                const float d = (float)cnt;
#pragma unroll
                for (int t = 0; t < VPL; ++t)
                  acc[t] = make_float4(__fdiv_rn(acc[t].x, d), __fdiv_rn(acc[t].y, d),
                                       __fdiv_rn(acc[t].z, d), __fdiv_rn(acc[t].w, d));
              }
            }
          }
          if (lane == j) mycount = cnt;
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int v = lane + 32 * t;
            if (v < C4) st_cs_f4(o + (int64_t)j * C4 + v, acc[t]);
          }
        }
      }
      if (P.cell_count && lane < ncell) P.cell_count[P.count_base[k] + b * cps + c0 + lane] = mycount;
    }
  }
  if (lane == 0) {  // last warp out resets the scheduler slot
    __threadfence();
    if (atomicAdd(&sched[1], 1ull) == (unsigned long long)gridDim.x * (blockDim.x >> 5) - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(256)
finalize_mean_kernel(float* __restrict__ planes, const int32_t* __restrict__ cnt, int64_t cells, int C4) {
  const int64_t total = cells * C4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int c = cnt[v / C4];
    if (c > 1) {
      float4* p = reinterpret_cast<float4*>(planes) + v;
      float4 x = *p;
      const float d = (float)c;
      *p = make_float4(__fdiv_rn(x.x, d), __fdiv_rn(x.y, d), __fdiv_rn(x.z, d), __fdiv_rn(x.w, d));
    }
  }
}

__global__ void __launch_bounds__(256)
voxel_counts_kernel(const int32_t* __restrict__ idx, int64_t n, const int64_t* __restrict__ offsets,
                    int batch, GeomDev g, int32_t* __restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ix = __ldg(idx + i * 3), iy = __ldg(idx + i * 3 + 1), iz = __ldg(idx + i * 3 + 2);
    if ((ix >= 0) & (ix < g.grid[0]) & (iy >= 0) & (iy < g.grid[1]) & (iz >= 0) & (iz < g.grid[2])) {
      int64_t b = tp_find_batch(offsets, batch, i);
      atomicAdd(counts + ((b * g.grid[0] + ix) * g.grid[1] + iy) * (int64_t)g.grid[2] + iz, 1);
    }
  }
}

static void plane_cells(const GeomDev& g, int64_t cps[3]) {
  cps[0] = (int64_t)g.grid[0] * g.grid[1] * g.pooled[2];
  cps[1] = (int64_t)g.grid[1] * g.grid[2] * g.pooled[0];
  cps[2] = (int64_t)g.grid[0] * g.grid[2] * g.pooled[1];
}

static int check_geom_e(const tp_geom* g, const char* who) {
  if (!g) return fail(TP_E_NULL, "%s: geom is null", who);
  for (int a = 0; a < 3; ++a)
    if (!(g->vs[a] > 0.f) || g->grid[a] <= 0 || g->pool[a] <= 0 || g->pool[a] > g->grid[a])
      return fail(TP_E_SHAPE, "%s: bad geometry on axis %d (vs=%g grid=%d pool=%d)", who, a,
                  (double)g->vs[a], g->grid[a], g->pool[a]);
  return 0;
}

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_encode_cells(const tp_geom* geom, int32_t batch, int64_t cells_per_plane[3]) {
  if (!geom || batch < 0) return -1;
  GeomDev g = make_geom_dev(*geom);
  int64_t cps[3];
  plane_cells(g, cps);
  int64_t tot = 0;
  for (int k = 0; k < 3; ++k) {
    if (cells_per_plane) cells_per_plane[k] = cps[k] * batch;
    tot += cps[k] * batch;
  }
  return tot;
}

static int64_t head_bytes(const tp_geom* geom, int32_t batch) {
  int64_t cells = tp_encode_cells(geom, batch, nullptr);
  return (cells * 4 + 255) / 256 * 256;
}

extern "C" int64_t tp_encode_workspace_bytes(const tp_geom* geom, int32_t batch, int64_t n_total) {
  if (!geom || batch <= 0 || n_total < 0) return -1;
  return head_bytes(geom, batch) + (3 * n_total * 4 + 255) / 256 * 256;
}

extern "C" int tp_encode_workspace_init(void* workspace, int64_t workspace_bytes, void* stream) {
  if (!workspace) return fail(TP_E_NULL, "tp_encode_workspace_init: null workspace");
  TP_CUDA(cudaMemsetAsync(workspace, 0xFF, (size_t)workspace_bytes, (cudaStream_t)stream));
  return 0;
}

extern "C" int tp_encode_f32(const float* feats, int64_t feat_stride, int32_t C, const int32_t* idx,
                             const float* points, int32_t point_stride, int64_t n,
                             const int64_t* offsets, int32_t batch, const tp_geom* geom,
                             int32_t arith, int32_t reduce, int32_t clamp_zero, float* out_xy,
                             float* out_yz, float* out_xz, int32_t* cell_count, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  if (int rc = check_geom_e(geom, "tp_encode_f32")) return rc;
  if (C <= 0 || (C & 3) || C > 512) return fail(TP_E_SHAPE, "tp_encode_f32: C=%d must be a multiple of 4 in [4,512]", C);
  if (batch <= 0 || n < 0 || n >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_encode_f32: batch=%d n=%lld", batch, (long long)n);
  if (!offsets || !workspace) return fail(TP_E_NULL, "tp_encode_f32: offsets/workspace null");
  if (n > 0 && !feats) return fail(TP_E_NULL, "tp_encode_f32: feats null");
  if (n > 0 && !idx && !points) return fail(TP_E_NULL, "tp_encode_f32: need idx or points");
  if (!idx && point_stride < 3 && n > 0) return fail(TP_E_SHAPE, "tp_encode_f32: point_stride=%d", point_stride);
  if (feat_stride < C || (feat_stride & 3) || ((uintptr_t)feats & 15))
    return fail(TP_E_SHAPE, "tp_encode_f32: feats must be 16-byte aligned rows (stride=%lld)", (long long)feat_stride);
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_encode_f32: unknown arith %d", arith);
  if (reduce < 0 || reduce > 2) return fail(TP_E_ENUM, "tp_encode_f32: unknown reduce %d", reduce);
  const int64_t need = tp_encode_workspace_bytes(geom, batch, n);
  if (workspace_bytes < need) return fail(TP_E_WORKSPACE, "tp_encode_f32: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  for (float* o : {out_xy, out_yz, out_xz})
    if (o && ((uintptr_t)o & 15)) return fail(TP_E_SHAPE, "tp_encode_f32: outputs must be 16-byte aligned");

  EncodeParams P;
  P.g = make_geom_dev(*geom);
  P.batch = batch;
  P.C = C;
  P.C4 = C / 4;
  P.n = n;
  int64_t cps[3];
  plane_cells(P.g, cps);
  int64_t hb = 0;
  for (int k = 0; k < 3; ++k) {
    P.cells_per_sample[k] = cps[k];
    P.head_base[k] = hb;
    P.count_base[k] = hb;
    hb += cps[k] * batch;
  }
  P.idx = idx;
  P.points = points;
  P.point_stride = point_stride;
  P.offsets = offsets;
  P.feats = feats;
  P.feat_stride = feat_stride;
  P.head = reinterpret_cast<int32_t*>(workspace);
  P.next = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + head_bytes(geom, batch));
  P.out[0] = out_xy;
  P.out[1] = out_yz;
  P.out[2] = out_xz;
  P.cell_count = cell_count;
  P.clamp_zero = clamp_zero;
  cudaStream_t s = (cudaStream_t)stream;

  if (n > 0) {
    int64_t blocks = (n + 255) / 256;
    int grid = (int)(blocks < (int64_t)kSMs * 8 ? blocks : (int64_t)kSMs * 8);
    if (arith == TP_ARITH_TORCH_CUDA) encode_link_kernel<TP_ARITH_TORCH_CUDA><<<grid, 256, 0, s>>>(P);
    else encode_link_kernel<TP_ARITH_TORCH_CPU><<<grid, 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("encode_link_kernel");
  }
  int64_t ups[3];
  for (int k = 0; k < 3; ++k) ups[k] = P.out[k] ? (cps[k] + kUnitCells - 1) / kUnitCells : 0;
  const int64_t units = (ups[0] + ups[1] + ups[2]) * batch;
  if (units == 0) return 0;
  const int64_t ctas = (units + 8 * kUnitsPerGrab - 1) / (8 * kUnitsPerGrab);
  const int64_t cap = (int64_t)kSMs * 8;  // persistent: 8 CTAs x 8 warps resident per SM
  const int grid = (int)(ctas < cap ? ctas : cap);
  const int vpl = (P.C4 + 31) / 32;
  static std::atomic<unsigned> next_slot{0};
  const int slot = (int)(next_slot.fetch_add(1) % kEncSchedSlots);
#define TP_MAT(R, V) encode_materialize_kernel<R, V><<<grid, 256, 0, s>>>(P, ups[0], ups[1], ups[2], slot)
#define TP_MAT_V(R)                                                              \
  switch (vpl) { case 1: TP_MAT(R, 1); break; case 2: TP_MAT(R, 2); break;        \
                 case 3: TP_MAT(R, 3); break; default: TP_MAT(R, 4); break; }
  if (reduce == TP_REDUCE_MAX) { TP_MAT_V(TP_REDUCE_MAX) }
  else if (reduce == TP_REDUCE_MEAN) { TP_MAT_V(TP_REDUCE_MEAN) }
  else { TP_MAT_V(TP_REDUCE_SUM) }
#undef TP_MAT_V
#undef TP_MAT
  TP_LAUNCH_CHECK("encode_materialize_kernel");
  return 0;
}

extern "C" int tp_encode_finalize_mean_f32(float* planes, const int32_t* cell_count, int64_t cells,
                                           int32_t C, void* stream) {
  if (!planes || !cell_count) return fail(TP_E_NULL, "tp_encode_finalize_mean_f32: null argument");
  if (C <= 0 || (C & 3) || cells < 0) return fail(TP_E_SHAPE, "tp_encode_finalize_mean_f32: C=%d cells=%lld", C, (long long)cells);
  if (cells == 0) return 0;
  int64_t blocks = (cells * (C / 4) + 255) / 256;
  int grid = (int)(blocks < (int64_t)kSMs * 16 ? blocks : (int64_t)kSMs * 16);
  finalize_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(planes, cell_count, cells, C / 4);
  TP_LAUNCH_CHECK("finalize_mean_kernel");
  return 0;
}

extern "C" int tp_voxel_counts_i32(const int32_t* idx, int64_t n, const int64_t* offsets,
                                   int32_t batch, const tp_geom* geom, int32_t* counts, void* stream) {
  if (int rc = check_geom_e(geom, "tp_voxel_counts_i32")) return rc;
  if (n == 0) return 0;
  if (!idx || !offsets || !counts) return fail(TP_E_NULL, "tp_voxel_counts_i32: null argument");
  if (batch <= 0 || n < 0) return fail(TP_E_SHAPE, "tp_voxel_counts_i32: batch=%d n=%lld", batch, (long long)n);
  int64_t blocks = (n + 255) / 256;
  int grid = (int)(blocks < (int64_t)kSMs * 8 ? blocks : (int64_t)kSMs * 8);
  voxel_counts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, n, offsets, batch, make_geom_dev(*geom), counts);
  TP_LAUNCH_CHECK("voxel_counts_kernel");
  return 0;
}
