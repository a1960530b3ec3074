// PointTriplaneProjector.forward without the dense pooled tensors (SURVEY 8f #3 i).
//
// The reference densifies three pooled tensors (point_triplane_projector.py:111-115: 430 MB per sample at the config
// geometry, < 3 % of the cells occupied) and feeds them to the first per-plane Linear (mlp_xy / mlp_yz / mlp_xz [0],
// :60-64): a [rows, G*C] x [G*C, C] GEMM that is 97 % multiplications by zero and reads the 430 MB back. With
//     hidden[row, :] = b1 + sum over the OCCUPIED pooled cells (row, g) of  W1[:, g*C:(g+1)*C] . cell(row, g)
// the same numbers come from one [C x C] product per occupied cell:
//   1. count    per point: crop + voxel index + the three pooled cells (the fused a1 of tp_encode.cu), one int32
//               counter per cell (3.4 MB per sample: L2-resident);
//   2. list     per cell: occupied cells are appended to their GROUP's list (group = pooled index g: all cells of a
//               group multiply the same C x C weight block); cells shared by several points are zeroed for pass 3;
//   3. scatter  per point: max into the cell's C-float slot (order-preserving integer keys; plain 16-byte stores when
//               the point owns the cell, the common case in one sweep);
//   4. gemm     per 64 cells of a group: D = cells . W1_g^T in fp32 FMA (exact fp32 products and accumulation: the
//               module tolerance is the reference's own fp32 Linear), written over the cells' slots;
//   5. combine  per row: b1 + the row's D vectors in ascending g, optional ReLU -> hidden [B, rows, C].
// Only occupied cells are ever touched: ~60 k x 512 B instead of 3 x 430 MB. The result is deterministic (no
// floating-point atomics: the sum over g happens in pass 5 in a fixed order).
// The cell slots live in a buffer with the dense tensors' addressing ([B, rows, G, C]); nothing but the occupied
// slots is read or written, so it costs address space, not bandwidth.
#include "tp_common.cuh"

namespace tp {

constexpr int kSpTile = 64;       // cells per GEMM tile
constexpr int kSpThreads = 256;
constexpr int kSpMaxC = 128;
constexpr int kSpPad = 4;

struct SparseParams {
  GeomDev g;
  int batch, C, C4;
  int64_t n;
  const int32_t* idx;
  const float* points;
  int point_stride;
  const int64_t* offsets;
  const float* feats;
  int64_t feat_stride;
  int arith, clamp_zero, relu;
  int G[3];            // groups per plane: Zp, Xp, Yp
  int64_t rows[3];     // rows per plane and sample: X*Y, Y*Z, X*Z
  int64_t cell0[3];    // first cell of plane p in cnt / slots (planes concatenated, batch inside)
  int64_t cells_total;
  int goff[3];         // first group counter of plane p
  int ngroups;
  int32_t* cnt;        // [cells_total] points per cell
  int32_t* lists;      // [cells_total] per group: cell ids (region of plane p, group g starts at cell0[p] + g * rows[p] * B)
  int32_t* gcount;     // [ngroups]
  float* slots;        // [cells_total * C]
  const float* w1t[3]; // [G_p][C][C]  (W1[n][g*C + k] stored as [g][k][n])
  const float* b1[3];  // [C]
  float* hidden[3];    // [B * rows_p, C]
};

__device__ __forceinline__ unsigned sp_f2key(float f) {
  unsigned u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float sp_key2f(unsigned k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// the three pooled cells of point i (global cell ids, -1 = none) — same crop / index / pooling rules as tp_encode.cu
template <int ARITH>
__device__ __forceinline__ void sp_point_cells(const SparseParams& P, int64_t i, int64_t cell[3]) {
  const GeomDev& g = P.g;
  cell[0] = cell[1] = cell[2] = -1;
  int ix, iy, iz;
  bool keep = true;
  if (P.idx) {
    ix = __ldg(P.idx + i * 3);
    iy = __ldg(P.idx + i * 3 + 1);
    iz = __ldg(P.idx + i * 3 + 2);
  } else {
    const float* p = P.points + i * P.point_stride;
    keep = tp_crop_index<ARITH>(g, __ldg(p), __ldg(p + 1), __ldg(p + 2), ix, iy, iz);
  }
  keep = keep & (ix >= 0) & (ix < g.grid[0]) & (iy >= 0) & (iy < g.grid[1]) & (iz >= 0) & (iz < g.grid[2]);
  if (!keep) return;
  const int64_t b = tp_find_batch(P.offsets, P.batch, i);
  const int px = ix / g.pool[0], py = iy / g.pool[1], pz = iz / g.pool[2];
  if (pz < g.pooled[2]) cell[0] = P.cell0[0] + ((b * P.rows[0] + (int64_t)ix * g.grid[1] + iy) * g.pooled[2] + pz);
  if (px < g.pooled[0]) cell[1] = P.cell0[1] + ((b * P.rows[1] + (int64_t)iy * g.grid[2] + iz) * g.pooled[0] + px);
  if (py < g.pooled[1]) cell[2] = P.cell0[2] + ((b * P.rows[2] + (int64_t)ix * g.grid[2] + iz) * g.pooled[1] + py);
}

template <int ARITH>
__global__ void __launch_bounds__(256)
sparse_count_kernel(const SparseParams P) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell[3];
    sp_point_cells<ARITH>(P, i, cell);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (cell[k] >= 0) atomicAdd(P.cnt + cell[k], 1);
  }
}

__device__ __forceinline__ int sp_plane_of(const SparseParams& P, int64_t c) {
  return c >= P.cell0[2] ? 2 : (c >= P.cell0[1] ? 1 : 0);
}

// One thread per cell, cells in memory order (coalesced counter reads). An occupied cell takes a slot in its group's
// list; the slots of shared cells (several points: they will be max-ed with atomics) are zeroed by the WHOLE warp, 32
// lanes x 16 B per row — one thread zeroing 512 B with 32 dependent stores was what made this pass take 97 us.
// Slots are claimed per CTA: the ~60 k occupied cells of a sample share only Zp + Xp + Yp (70) group counters, so one
// global atomic per cell is ~7 k same-address atomics per counter (79 us for a 27 MB scan at bs = 8); here a cell takes
// its rank from a shared-memory counter and one thread per group reserves the CTA's range.
constexpr int kSpMaxGroups = 256;  // = the size of the gcount region

__global__ void __launch_bounds__(256)
sparse_list_kernel(const SparseParams P) {
  __shared__ int s_cnt[kSpMaxGroups], s_base[kSpMaxGroups];
  const int lane = threadIdx.x & 31, tid = threadIdx.x;
  const int ngroups = P.ngroups;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < P.cells_total; base += (int64_t)gridDim.x * blockDim.x) {
    s_cnt[tid] = 0;
    __syncthreads();
    const int64_t c = base + tid;
    const int n = c < P.cells_total ? P.cnt[c] : 0;
    int p = 0, g = 0, r = 0;
    if (n > 0) {
      p = sp_plane_of(P, c);
      g = (int)((c - P.cell0[p]) % P.G[p]);
      r = atomicAdd(&s_cnt[P.goff[p] + g], 1);
    }
    __syncthreads();
    if (tid < ngroups && s_cnt[tid] > 0) s_base[tid] = atomicAdd(P.gcount + tid, s_cnt[tid]);
    __syncthreads();
    if (n > 0)
      P.lists[P.cell0[p] + (int64_t)g * P.rows[p] * P.batch + s_base[P.goff[p] + g] + r] = (int)(c - P.cell0[p]);
    for (unsigned m = __ballot_sync(0xffffffffu, n > 1); m; m &= m - 1) {  // shared cells: max starts from key 0
      uint4* row = reinterpret_cast<uint4*>(P.slots + (base + (tid & ~31) + __ffs(m) - 1) * P.C);
      for (int k = lane; k < P.C4; k += 32) row[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    // s_base is rewritten only after the next iteration's first two barriers; s_cnt after its first: no barrier here
  }
}

// one warp per point: the feature row is read once and folded into its three cells
template <int ARITH>
__global__ void __launch_bounds__(256)
sparse_scatter_kernel(const SparseParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < P.n; i += nwarp) {
    int64_t cell[3];
    sp_point_cells<ARITH>(P, i, cell);
    if (cell[0] < 0 && cell[1] < 0 && cell[2] < 0) continue;
    const float4* frow = reinterpret_cast<const float4*>(P.feats + i * P.feat_stride);
    for (int v = lane; v < P.C4; v += 32) {
      const float4 f = __ldg(frow + v);
      const uint4 k4 = make_uint4(sp_f2key(f.x), sp_f2key(f.y), sp_f2key(f.z), sp_f2key(f.w));
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (cell[k] < 0) continue;
        unsigned* dst = reinterpret_cast<unsigned*>(P.slots + cell[k] * P.C) + v * 4;
        if (P.cnt[cell[k]] == 1) {
          *reinterpret_cast<uint4*>(dst) = k4;
        } else {
          // max is monotone: a key that does not beat the value already there needs no atomic. Ground-plane cells of
          // the yz / xz planes collect hundreds of points; after the first few almost every update is skipped.
          const uint4 cur = __ldcg(reinterpret_cast<const uint4*>(dst));  // L2 (where the atomics resolve), not a stale L1 line
          if (k4.x > cur.x) atomicMax(dst + 0, k4.x);
          if (k4.y > cur.y) atomicMax(dst + 1, k4.y);
          if (k4.z > cur.z) atomicMax(dst + 2, k4.z);
          if (k4.w > cur.w) atomicMax(dst + 3, k4.w);
        }
      }
    }
  }
}

// D[cell, n] = sum_k A[cell, k] * W_g[k, n] for tiles of 64 cells of one group, fp32 FMA. 256 threads: thread (ty, tx)
// owns cells ty*4 .. +3 and outputs tx*8 .. +7; A tile [64][C + 4] and the group's W block [C][C] live in shared
// memory (the W block is kept while consecutive tiles stay in the same group).
__global__ void __launch_bounds__(kSpThreads, 2)
sparse_gemm_kernel(const SparseParams P) {
  extern __shared__ __align__(16) float sp_smem[];
  float* As = sp_smem;                                 // [kSpTile][kSpMaxC + kSpPad]
  float* Ws = As + kSpTile * (kSpMaxC + kSpPad);       // [kSpMaxC][kSpMaxC]
  __shared__ int s_cell[kSpTile];
  __shared__ int s_tiles[3 * 64 + 1];                  // prefix of tiles per group
  const int tid = threadIdx.x, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;
  const int C = P.C, C4 = P.C4;
  constexpr int LDA = kSpMaxC + kSpPad;

  if (tid == 0) {
    int acc = 0;
    for (int q = 0; q < P.ngroups; ++q) {
      s_tiles[q] = acc;
      acc += (P.gcount[q] + kSpTile - 1) / kSpTile;
    }
    s_tiles[P.ngroups] = acc;
  }
  // zero-pad once: k >= C columns of A and the unused part of W stay zero for every tile
  for (int i = tid; i < kSpTile * LDA; i += kSpThreads) As[i] = 0.f;
  for (int i = tid; i < kSpMaxC * kSpMaxC; i += kSpThreads) Ws[i] = 0.f;
  __syncthreads();
  const int total = s_tiles[P.ngroups];
  // contiguous chunk of tiles per CTA: consecutive tiles mostly share their group (and its weights)
  const int per = (total + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(total, t_begin + per);
  int cur_q = -1, q = 0;
  for (int t = t_begin; t < t_end; ++t) {
    while (s_tiles[q + 1] <= t) ++q;
    const int p = q >= P.goff[2] ? 2 : (q >= P.goff[1] ? 1 : 0);
    const int g = q - P.goff[p];
    const int first = (t - s_tiles[q]) * kSpTile;
    const int ncell = min(kSpTile, P.gcount[q] - first);
    __syncthreads();  // previous tile's readers are done with As / s_cell (and Ws if the group changes)
    if (q != cur_q) {
      const float4* w = reinterpret_cast<const float4*>(P.w1t[p] + (int64_t)g * C * C);
      for (int i = tid; i < C * C4; i += kSpThreads) {
        const int k = i / C4, n4 = i - k * C4;
        *reinterpret_cast<float4*>(Ws + k * kSpMaxC + n4 * 4) = __ldg(w + i);
      }
      cur_q = q;
    }
    if (tid < kSpTile)
      s_cell[tid] = tid < ncell ? P.lists[P.cell0[p] + (int64_t)g * P.rows[p] * P.batch + first + tid] : -1;
    __syncthreads();
    // gather the tile's cells: one warp reads one 4*C-byte row per instruction, keys -> floats on the way
    for (int r = tid >> 5; r < kSpTile; r += kSpThreads / 32) {
      const int c = s_cell[r];
      for (int v = lane; v < C4; v += 32) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= 0) {
          const uint4 kk = *reinterpret_cast<const uint4*>(P.slots + (P.cell0[p] + c) * C + v * 4);
          a = make_float4(sp_key2f(kk.x), sp_key2f(kk.y), sp_key2f(kk.z), sp_key2f(kk.w));
          if (P.clamp_zero) a = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
        }
        *reinterpret_cast<float4*>(As + r * LDA + v * 4) = a;
      }
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* a0 = As + (ty * 4) * LDA;
    const float* w0 = Ws + tx * 8;
    for (int k = 0; k < C; k += 4) {
      float4 av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a0 + i * LDA + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 wl = *reinterpret_cast<const float4*>(w0 + (k + kk) * kSpMaxC);
        const float4 wh = *reinterpret_cast<const float4*>(w0 + (k + kk) * kSpMaxC + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = kk == 0 ? av[i].x : (kk == 1 ? av[i].y : (kk == 2 ? av[i].z : av[i].w));
          acc[i][0] = __fmaf_rn(a, wl.x, acc[i][0]);
          acc[i][1] = __fmaf_rn(a, wl.y, acc[i][1]);
          acc[i][2] = __fmaf_rn(a, wl.z, acc[i][2]);
          acc[i][3] = __fmaf_rn(a, wl.w, acc[i][3]);
          acc[i][4] = __fmaf_rn(a, wh.x, acc[i][4]);
          acc[i][5] = __fmaf_rn(a, wh.y, acc[i][5]);
          acc[i][6] = __fmaf_rn(a, wh.z, acc[i][6]);
          acc[i][7] = __fmaf_rn(a, wh.w, acc[i][7]);
        }
      }
    }
    // D over the cells' own slots (every cell belongs to exactly one tile row; its A row is already in shared memory)
    if (tx * 8 < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = s_cell[ty * 4 + i];
        if (c < 0) continue;
        float* d = P.slots + (P.cell0[p] + c) * C + tx * 8;
        *reinterpret_cast<float4*>(d) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (tx * 8 + 4 < C) *reinterpret_cast<float4*>(d + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
  }
}

// hidden[row] = b1 + sum_g D[row, g] (ascending g), optional ReLU. One warp per row.
__global__ void __launch_bounds__(256)
sparse_combine_kernel(const SparseParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t rows_all = (P.rows[0] + P.rows[1] + P.rows[2]) * P.batch;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows_all; r += nwarp) {
    int p = 0;
    int64_t rr = r;
    if (rr >= P.rows[0] * P.batch) { rr -= P.rows[0] * P.batch; p = 1;
      if (rr >= P.rows[1] * P.batch) { rr -= P.rows[1] * P.batch; p = 2; } }
    const int G = P.G[p];
    const int64_t c0 = P.cell0[p] + rr * G;
    float4* out = reinterpret_cast<float4*>(P.hidden[p] + rr * P.C);
    const float4* bias = reinterpret_cast<const float4*>(P.b1[p]);
    // which of the row's G <= 64 cells are occupied (all lanes vote, whatever C is)
    const unsigned occ_lo = __ballot_sync(0xffffffffu, lane < G && P.cnt[c0 + lane] > 0);
    const unsigned occ_hi = __ballot_sync(0xffffffffu, 32 + lane < G && P.cnt[c0 + 32 + lane] > 0);
    for (int v = lane; v < P.C4; v += 32) {
      float4 acc = __ldg(bias + v);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        for (unsigned m = half ? occ_hi : occ_lo; m; m &= m - 1) {
          const int g = half * 32 + __ffs(m) - 1;
          const float4 d = *reinterpret_cast<const float4*>(P.slots + (c0 + g) * P.C + v * 4);
          acc.x = __fadd_rn(acc.x, d.x); acc.y = __fadd_rn(acc.y, d.y); acc.z = __fadd_rn(acc.z, d.z); acc.w = __fadd_rn(acc.w, d.w);
        }
      }
      if (P.relu) acc = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
      out[v] = acc;
    }
  }
}

static int sparse_layout(SparseParams& P, const tp_geom* geom, int batch, int C) {
  P.g = make_geom_dev(*geom);
  P.batch = batch;
  P.C = C;
  P.C4 = C / 4;
  const GeomDev& g = P.g;
  P.G[0] = g.pooled[2]; P.G[1] = g.pooled[0]; P.G[2] = g.pooled[1];
  P.rows[0] = (int64_t)g.grid[0] * g.grid[1];
  P.rows[1] = (int64_t)g.grid[1] * g.grid[2];
  P.rows[2] = (int64_t)g.grid[0] * g.grid[2];
  int64_t c = 0;
  int q = 0;
  for (int p = 0; p < 3; ++p) {
    if (P.G[p] <= 0 || P.G[p] > 64) return -1;
    P.cell0[p] = c;
    P.goff[p] = q;
    c += P.rows[p] * P.G[p] * batch;
    q += P.G[p];
  }
  P.cells_total = c;
  P.ngroups = q;
  return 0;
}

static inline int64_t sp_pad(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace tp

using namespace tp;

extern "C" int64_t tp_projector_sparse_workspace_bytes(const tp_geom* geom, int32_t batch, int32_t C) {
  if (!geom || batch <= 0 || C <= 0) return -1;
  SparseParams P;
  if (sparse_layout(P, geom, batch, C)) return -1;
  return sp_pad(P.cells_total * 4) * 2 + sp_pad(256 * 4) + sp_pad(P.cells_total * (int64_t)C * 4);
}

extern "C" int tp_projector_sparse_f32(const float* feats, int64_t feat_stride, int32_t C, const int32_t* idx,
                                       const float* points, int32_t point_stride, int64_t n, const int64_t* offsets,
                                       int32_t batch, const tp_geom* geom, int32_t arith, int32_t clamp_zero,
                                       const float* const w1t[3], const float* const b1[3], int32_t relu,
                                       float* const hidden[3], void* workspace, int64_t workspace_bytes, void* stream) {
  if (!geom) return fail(TP_E_NULL, "tp_projector_sparse_f32: null geometry");
  for (int a = 0; a < 3; ++a)
    if (!(geom->vs[a] > 0.f) || geom->grid[a] <= 0 || geom->pool[a] <= 0 || geom->pool[a] > geom->grid[a])
      return fail(TP_E_SHAPE, "tp_projector_sparse_f32: bad geometry on axis %d", a);
  if (C <= 0 || (C & 3) || C > kSpMaxC) return fail(TP_E_SHAPE, "tp_projector_sparse_f32: C=%d must be a multiple of 4 in [4,%d]", C, kSpMaxC);
  if (batch <= 0 || n < 0) return fail(TP_E_SHAPE, "tp_projector_sparse_f32: batch=%d n=%lld", batch, (long long)n);
  if (!offsets || !workspace || !w1t || !b1 || !hidden) return fail(TP_E_NULL, "tp_projector_sparse_f32: null argument");
  if (n > 0 && (!feats || (!idx && !points))) return fail(TP_E_NULL, "tp_projector_sparse_f32: null input");
  if (n > 0 && (feat_stride < C || (feat_stride & 3) || ((uintptr_t)feats & 15)))
    return fail(TP_E_SHAPE, "tp_projector_sparse_f32: feats must be 16-byte aligned rows");
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_projector_sparse_f32: unknown arith %d", arith);
  SparseParams P;
  if (sparse_layout(P, geom, batch, C)) return fail(TP_E_SHAPE, "tp_projector_sparse_f32: more than 64 pooled cells along an axis");
  if (P.cells_total >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_projector_sparse_f32: too many cells");
  const int64_t need = tp_projector_sparse_workspace_bytes(geom, batch, C);
  if (workspace_bytes < need) return fail(TP_E_WORKSPACE, "tp_projector_sparse_f32: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  for (int p = 0; p < 3; ++p) {
    if (!w1t[p] || !b1[p] || !hidden[p]) return fail(TP_E_NULL, "tp_projector_sparse_f32: plane %d has a null weight / bias / output", p);
    if (((uintptr_t)w1t[p] | (uintptr_t)b1[p] | (uintptr_t)hidden[p]) & 15) return fail(TP_E_SHAPE, "tp_projector_sparse_f32: plane %d pointers must be 16-byte aligned", p);
    P.w1t[p] = w1t[p]; P.b1[p] = b1[p]; P.hidden[p] = hidden[p];
  }
  P.n = n; P.idx = idx; P.points = points; P.point_stride = point_stride; P.offsets = offsets;
  P.feats = feats; P.feat_stride = feat_stride; P.arith = arith; P.clamp_zero = clamp_zero; P.relu = relu;
  char* ws = reinterpret_cast<char*>(workspace);
  P.cnt = reinterpret_cast<int32_t*>(ws); ws += sp_pad(P.cells_total * 4);
  P.lists = reinterpret_cast<int32_t*>(ws); ws += sp_pad(P.cells_total * 4);
  P.gcount = reinterpret_cast<int32_t*>(ws); ws += sp_pad(256 * 4);
  P.slots = reinterpret_cast<float*>(ws);
  cudaStream_t s = (cudaStream_t)stream;
  TP_CUDA(cudaMemsetAsync(P.cnt, 0, (size_t)P.cells_total * 4, s));
  TP_CUDA(cudaMemsetAsync(P.gcount, 0, 256 * 4, s));
  if (n > 0) {
    const int64_t blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (int64_t)kSMs * 8 ? blocks : (int64_t)kSMs * 8);
    if (arith == TP_ARITH_TORCH_CUDA) sparse_count_kernel<TP_ARITH_TORCH_CUDA><<<grid, 256, 0, s>>>(P);
    else sparse_count_kernel<TP_ARITH_TORCH_CPU><<<grid, 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("sparse_count_kernel");
    const int64_t cb = (P.cells_total + 255) / 256;
    sparse_list_kernel<<<(int)(cb < (int64_t)kSMs * 16 ? cb : (int64_t)kSMs * 16), 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("sparse_list_kernel");
    const int64_t wb = (n + 7) / 8;
    const int sgrid = (int)(wb < (int64_t)kSMs * 8 ? wb : (int64_t)kSMs * 8);
    if (arith == TP_ARITH_TORCH_CUDA) sparse_scatter_kernel<TP_ARITH_TORCH_CUDA><<<sgrid, 256, 0, s>>>(P);
    else sparse_scatter_kernel<TP_ARITH_TORCH_CPU><<<sgrid, 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("sparse_scatter_kernel");
    constexpr int kSmem = (kSpTile * (kSpMaxC + kSpPad) + kSpMaxC * kSpMaxC) * 4;
    TP_CUDA(opt_in_smem<sparse_gemm_kernel>(kSmem));
    sparse_gemm_kernel<<<kSMs * 2, kSpThreads, kSmem, s>>>(P);
    TP_LAUNCH_CHECK("sparse_gemm_kernel");
  }
  const int64_t rows_all = (P.rows[0] + P.rows[1] + P.rows[2]) * batch;
  const int64_t rb = (rows_all + 7) / 8;
  sparse_combine_kernel<<<(int)(rb < (int64_t)kSMs * 16 ? rb : (int64_t)kSMs * 16), 256, 0, s>>>(P);
  TP_LAUNCH_CHECK("sparse_combine_kernel");
  return 0;
}
