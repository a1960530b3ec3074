// a3: triplane encode — tile-binned scatter-reduce in shared memory, dense output written once by TMA.
//
// Replaces point_triplane_projector.py:99-115:
//   torch.unique(dim=0) -> torch_scatter.scatter_max -> 3 x spconv.SparseMaxPool3d -> .dense()
//   -> permute(...).flatten(3)
// (a sort, three hash builds and ~6 passes over 430 MB per sample). max over the points of a voxel
// followed by max over the voxels of a pooling window == max over the points of the pooled cell, so
// the intermediate voxel tensor never exists.
//
// The three channels-last outputs are cut into TILES of consecutive cells (16 KB of output each).
//   1. count:  every point computes its crop / voxel index / three pooled cells (fused a1) and
//              bumps the point counter of the three tiles it lands in (int32 atomics, L2-resident).
//   2. alloc:  every occupied tile gets its CSR range by warp-aggregated bump allocation (the lists
//              need not be in tile order, so there is no serial prefix scan); heavy tiles are listed.
//   3. fill:   every point writes (point id, cell inside tile) into its slot of each tile's list.
//   4. reduce: persistent 4-warp CTAs (about ten per SM) pull tiles from a global counter. An empty
//              tile is 16 KB of streaming zero stores. An occupied tile is reduced in the CTA's
//              shared-memory tile buffer — its point rows are gathered with independent 16-byte loads
//              (ids come from the CSR list: no pointer chasing) and folded in with shared-memory
//              atomics (max on order-preserving integer keys, or fp32 add) — finalised in place and
//              written with ONE bulk-async (TMA) store. Every output byte is written exactly once;
//              there are no global floating-point atomics.
// An earlier design (per-cell linked lists walked by one warp) is in git history; its ncu summary
// (profiles/r01_encode_v1_linkedlist_ncu.txt) shows the dependent-load chains this one removes.
#include <atomic>
#include <cstdlib>

#include "tp_common.cuh"

namespace tp {

constexpr int kTileFloats = 4096;  // 16 KB of output per tile
constexpr int kMaxCpt = 256;       // cells per tile is capped (small C): bounds the per-buffer counters
constexpr int kRedThreads = 128;   // small CTAs, ~10 resident per SM: that many tile chains in flight
#ifndef TP_RED_CTAS
#define TP_RED_CTAS 10
#endif
#ifndef TP_RED_ROWS
#define TP_RED_ROWS 4
#endif
constexpr int kRedCtasPerSm = TP_RED_CTAS;
constexpr int kRows = TP_RED_ROWS;  // rows a warp has in flight (C = 128 path)
constexpr int kGrab = 4;           // tiles taken per scheduler atomic
constexpr int kHeavy = 48;         // tiles with at least this many points are reduced first
constexpr int kSplit = 384;        // tiles with at least this many points are cut into items of kSub list entries,
constexpr int kSub = 192;          // reduced by different CTAs and combined in the output tile itself

struct EncodeParams {
  GeomDev g;
  int batch;
  int C, C4;
  int cpt;  // cells per tile = largest power of two <= min(kTileFloats / C, kMaxCpt)
  int64_t n;
  int64_t cells_per_sample[3];
  int64_t tiles_per_sample[3];  // 0 for a skipped plane
  int64_t tiles_all;            // sum over planes
  int64_t tiles_total;          // * batch
  const int32_t* idx;
  const float* points;
  int point_stride;
  const int64_t* offsets;
  const float* feats;
  int64_t feat_stride;
  int32_t* tile_cnt;    // [tiles_total]   zero on entry, zero on exit
  int32_t* tile_start;  // [tiles_total + 1]
  int32_t* heavy;       // [1 + tiles_total]: count, then the tiles with >= kHeavy points (alloc pass)
  int32_t* cursor;      // [1] bump allocator of the CSR entries (zeroed by the count pass)
  // split tiles (>= kSplit points; ground-plane tiles collect thousands): split_ctr = {#split tiles, #items}
  int32_t* split_ctr;   // [2] zeroed by the count pass
  int32_t* split_tile;  // [#split] tile index
  int2* split_item;     // [#items] (split tile slot, sub-range index)
  int32_t* split_done;  // [#split] items finished (zeroed by the fill pass)
  int32_t* split_cnt;   // [#split][kMaxCpt] points per cell, accumulated over the items (zeroed by the fill pass)
  int32_t* rank;        // [3][n]
  int2* entries;        // [3n] (point id, cell inside tile)
  float* out[3];
  int32_t* cell_count;  // optional, order xy | yz | xz, each [B, cells_per_sample]
  int64_t count_base[3];
  int clamp_zero;
  int keep_rows;  // 1: the point rows fit in L2 next to the output stream: prefetch + evict_last (else default policy)
  int partial;  // 1: TP_REDUCE_MAX_PARTIAL — empty cells are -inf (identity of max) for a cross-GPU max
};

// tile order: sample-major, then plane, then position — the three planes of a sample are reduced
// close together in time so their feature rows are re-read from L2, not HBM.
__device__ __forceinline__ int64_t tile_index(const EncodeParams& P, int64_t b, int k, int64_t cell) {
  int64_t t = b * P.tiles_all + cell / P.cpt;
  if (k >= 1) t += P.tiles_per_sample[0];
  if (k >= 2) t += P.tiles_per_sample[1];
  return t;
}

// tile -> its slice of the output: first float, optional cell counters, number of cells (the last tile
// of a plane may be short)
__device__ __forceinline__ int tile_dest(const EncodeParams& P, int64_t t, float*& gdst, int32_t*& gcount) {
  const int64_t b = t / P.tiles_all;
  int64_t r = t - b * P.tiles_all;
  int k = 0;
  if (r >= P.tiles_per_sample[0]) { r -= P.tiles_per_sample[0]; k = 1;
    if (r >= P.tiles_per_sample[1]) { r -= P.tiles_per_sample[1]; k = 2; } }
  const int64_t cps = P.cells_per_sample[k];
  const int64_t c0 = r * P.cpt;
  gdst = P.out[k] + (b * cps + c0) * P.C;
  gcount = P.cell_count ? P.cell_count + P.count_base[k] + b * cps + c0 : nullptr;
  return (int)min((int64_t)P.cpt, cps - c0);
}

// crop + voxel index + the three pooled cells of point i; cell[k] = -1 where the point does not
// contribute (outside the crop / grid, plane skipped, or beyond the pooled extent).
template <int ARITH>
__device__ __forceinline__ int64_t point_cells(const EncodeParams& P, int64_t i, int64_t cell[3]) {
  const GeomDev& g = P.g;
  cell[0] = cell[1] = cell[2] = -1;
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
  if (!keep) return -1;
  const int px = ix / g.pool[0], py = iy / g.pool[1], pz = iz / g.pool[2];
  // xy: [B, X, Y, Zp]  (pool_xy(...).dense().permute(0,2,3,4,1), projector.py:113)
  if (P.out[0] && pz < g.pooled[2]) cell[0] = ((int64_t)ix * g.grid[1] + iy) * g.pooled[2] + pz;
  // yz: [B, Y, Z, Xp]  (permute(0,3,4,2,1), projector.py:114)
  if (P.out[1] && px < g.pooled[0]) cell[1] = ((int64_t)iy * g.grid[2] + iz) * g.pooled[0] + px;
  // xz: [B, X, Z, Yp]  (permute(0,2,4,3,1), projector.py:115)
  if (P.out[2] && py < g.pooled[1]) cell[2] = ((int64_t)ix * g.grid[2] + iz) * g.pooled[1] + py;
  return tp_find_batch(P.offsets, P.batch, i);
}

// ---- pass 1: count ---------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(256)
encode_count_kernel(const EncodeParams P) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // consumed by the alloc pass
    P.heavy[0] = 0; P.cursor[0] = 0; P.split_ctr[0] = 0; P.split_ctr[1] = 0;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell[3];
    const int64_t b = point_cells<ARITH>(P, i, cell);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int r = -1;
      if (b >= 0 && cell[k] >= 0) r = atomicAdd(P.tile_cnt + tile_index(P, b, k, cell[k]), 1);
      P.rank[(int64_t)k * P.n + i] = r;
    }
  }
}

// ---- pass 2: CSR ranges + heavy-tile list (one lane per tile, one atomic per warp) ---------------
__global__ void __launch_bounds__(256)
encode_alloc_kernel(const EncodeParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < P.tiles_total;
       base += nwarp * 32) {
    const int64_t t = base + lane;
    const int c = t < P.tiles_total ? P.tile_cnt[t] : 0;
    int incl = c;  // exclusive scan of the warp's counts -> one bump of the entries cursor per warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int wbase = 0;
    if (lane == 31 && incl > 0) wbase = atomicAdd(P.cursor, incl);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    if (c > 0) P.tile_start[t] = wbase + incl - c;
    if (c >= kSplit) {  // rare: this lane lists the tile and its items by itself
      const int m = atomicAdd(P.split_ctr, 1);
      const int ns = (c + kSub - 1) / kSub;
      const int i0 = atomicAdd(P.split_ctr + 1, ns);
      P.split_tile[m] = (int)t;
      for (int j = 0; j < ns; ++j) P.split_item[i0 + j] = make_int2(m, j);
    }
    const unsigned mh = __ballot_sync(0xffffffffu, c >= kHeavy && c < kSplit);
    int bh = 0;
    if (lane == 0 && mh) bh = atomicAdd(P.heavy, __popc(mh));
    bh = __shfl_sync(0xffffffffu, bh, 0);
    if (c >= kHeavy && c < kSplit) P.heavy[1 + bh + __popc(mh & ((1u << lane) - 1u))] = (int)t;
  }
}

// ---- pass 3: fill ------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(256)
encode_fill_kernel(const EncodeParams P) {
  unsigned long long pol_keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  // split tiles are accumulated in place by several CTAs of the reduce pass: start them (and their
  // per-cell counters) from zero bits = key 0 / 0.0f
  const int nsplit = P.split_ctr[0];
  for (int m = blockIdx.x; m < nsplit; m += gridDim.x) {
    float* gdst;
    int32_t* gcount;
    const int ncell = tile_dest(P, P.split_tile[m], gdst, gcount);
    float4* g4 = reinterpret_cast<float4*>(gdst);
    for (int v = threadIdx.x; v < ncell * P.C4; v += blockDim.x) g4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = threadIdx.x; c < kMaxCpt; c += blockDim.x) P.split_cnt[(int64_t)m * kMaxCpt + c] = 0;
    if (threadIdx.x == 0) P.split_done[m] = 0;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell[3];
    const int64_t b = point_cells<ARITH>(P, i, cell);
    if (b < 0) continue;
    // Pull this point's feature row into L2 now, with evict_last priority: the reduce pass gathers it
    // up to three times while the dense output streams through L2, and a gather that misses waits in
    // the DRAM queues behind ~6 TB/s of writes.
    if (P.keep_rows && (cell[0] >= 0 || cell[1] >= 0 || cell[2] >= 0)) {
      const char* row = reinterpret_cast<const char*>(P.feats + i * P.feat_stride);
      for (int o = 0; o < P.C * 4; o += 128)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(row + o));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (cell[k] < 0) continue;
      const int64_t t = tile_index(P, b, k, cell[k]);
      const int slot = P.tile_start[t] + P.rank[(int64_t)k * P.n + i];
      // keep the CSR list in L2 while the dense output streams through it
      asm volatile("st.global.L2::cache_hint.v2.s32 [%0], {%1,%2}, %3;" ::"l"(P.entries + slot), "r"((int)i),
                   "r"((int)(cell[k] % P.cpt)), "l"(pol_keep)
                   : "memory");
    }
  }
}

// ---- pass 4: reduce + write ---------------------------------------------------------------------
// order-preserving float <-> uint32 (max on keys == max on floats; +NaN is the largest key, so a NaN
// feature propagates like torch.amax; -0 < +0)
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned u = __float_as_uint(f);
  return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);  // negative: flip all bits; else: flip the sign (shift + one LOP3)
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// L2 policies: the dense output is written once and never re-read by this kernel (evict_first), the
// point rows and CSR entries are re-read by up to three planes while 430 MB of output streams past
// them (evict_last) — without the hints the rows are evicted between planes and every gather waits in
// the DRAM queue behind the writes (ncu: lts hit rate 5 %, profiles/r01_encode_v1_linkedlist_ncu.txt).
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float ld_keep_f1(const float* p, unsigned long long pol) {
  float r;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

__device__ __forceinline__ void red_add_v4(float4* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kEncSchedSlots = 256;
__device__ unsigned long long g_enc_sched[kEncSchedSlots][4];  // [next tile, finished CTAs, next phase-1 window, -]

// Persistent small CTAs (4 warps, one 16 KB tile buffer each, ~10 per SM). The chain per tile is
// metadata -> CSR entries -> point rows, three dependent global loads (~3 us); the stream of dense
// output needs a 16 KB tile per SM every ~0.4 us, so many independent tile chains must be in flight
// per SM — measured: 3 CTAs/SM with 32 KB tiles 117 us, one 16-warp CTA per SM 210 us (profiles/).
// C128: the configs' channel count (a row = 32 lanes x 4 floats, 32 cells per tile) known at compile time.
template <int REDUCE, bool C128>
__global__ void __launch_bounds__(kRedThreads, kRedCtasPerSm)
encode_reduce_kernel(const EncodeParams P, int sched_slot) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* s_work = reinterpret_cast<float*>(smem_raw);           // kTileFloats, zero outside touched cells
  int* s_cnt = reinterpret_cast<int*>(s_work + kTileFloats);    // kMaxCpt: points per cell of the tile
  __shared__ long long s_t0;
  __shared__ int s_tile[kGrab], s_npts[kGrab], s_start[kGrab], s_split[kGrab];  // the batch of work items in hand
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kRedThreads / 32;
  const int C = C128 ? 128 : P.C, C4 = C128 ? 32 : P.C4, cpt = C128 ? 32 : P.cpt;
  unsigned long long* sched = g_enc_sched[sched_slot];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float fillv = P.partial ? __uint_as_float(0xff800000u) : 0.f;  // -inf or 0 for empty cells
  const float4 fill4 = make_float4(fillv, fillv, fillv, fillv);
  const unsigned long long pol_out = policy_evict_first();
  unsigned long long pol_in = policy_evict_last();
  if (!P.keep_rows) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_in));

  // the tile buffer starts zeroed and is returned to all-zero before each reuse by clearing only the
  // cells the previous tile touched (their counters are still set): a sparse tile costs O(points).
  for (int i = tid; i < kTileFloats / 4; i += kRedThreads) reinterpret_cast<float4*>(s_work)[i] = zero4;
  for (int i = tid; i < kMaxCpt; i += kRedThreads) s_cnt[i] = 0;
  bool in_flight = false;  // thread 0: a bulk store may still be reading s_work

  // One tile: empty -> 16 KB of streaming zero stores; occupied -> reduce in shared memory, one TMA store.
  // split_m < 0: the whole tile. split_m >= 0: one item of split tile slot split_m — npts list entries from
  // `start`; the partial result is folded into the output tile with reductions, the last item finalises it.
  auto process_tile = [&](const int64_t t, const int npts, const int start, const int split_m) {
    float* gdst;
    int32_t* gcount;
    const int ncell = tile_dest(P, t, gdst, gcount);

    if (npts == 0) {
      float4* g4 = reinterpret_cast<float4*>(gdst);
      const int nvec = ncell * C4;
#pragma unroll 4
      for (int i = tid; i < nvec; i += kRedThreads) st_stream_f4(g4 + i, fill4, pol_out);
      if (gcount) for (int c = tid; c < ncell; c += kRedThreads) gcount[c] = 0;
      return;
    }
    if (tid == 0 && in_flight) { bulk_wait_read0(); in_flight = false; }
    __syncthreads();
    for (int c = warp; c < cpt; c += kWarps) {  // rows are warp-owned here: no barrier before the reset
      if (s_cnt[c] != 0) {
        for (int v = lane; v < C4; v += 32) reinterpret_cast<float4*>(s_work + c * C)[v] = zero4;
        __syncwarp();
        if (lane == 0) s_cnt[c] = 0;
      }
    }
    __syncthreads();
    const int2* ent = P.entries + start;
    // points per cell first (one integer atomic per point): a cell with a single point — the common case
    // in one sweep — is then written with plain stores, and only shared cells pay for feature atomics
    // (the ATOMS pipe, 4 x 32 lanes per row, was the busiest unit of the previous version)
    for (int e = tid; e < npts; e += kRedThreads) atomicAdd(s_cnt + __ldg(ent + e).y, 1);
    __syncthreads();
    if (C128) {
      for (int e0 = warp * kRows; e0 < npts; e0 += kWarps * kRows) {
        int2 en[kRows];
#pragma unroll
        for (int j = 0; j < kRows; ++j) en[j] = (e0 + j < npts) ? __ldg(ent + e0 + j) : make_int2(-1, 0);
        // Lane l owns words l, l + 32, l + 64, l + 96 of the row: four coalesced 128-byte loads per row, and the four
        // shared-memory atomics (or stores) of a row are conflict-free — with 16 bytes per lane they were 4-way bank
        // conflicts (ncu: 12 M of 22 M shared wavefronts on a 10-sweep sample) and the ATOMS pipe the busiest unit.
        float x[kRows][4];
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
          if (en[j].x < 0) continue;
          const float* r = P.feats + (int64_t)en[j].x * P.feat_stride + lane;
#pragma unroll
          for (int k = 0; k < 4; ++k) x[j][k] = ld_keep_f1(r + 32 * k, pol_in);
        }
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
          if (en[j].x < 0) continue;
          float* w = s_work + en[j].y * 128 + lane;
          if (s_cnt[en[j].y] == 1) {  // warp-uniform: sole owner of the row
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (REDUCE == TP_REDUCE_MAX) reinterpret_cast<unsigned*>(w)[32 * k] = f2key(x[j][k]);
              else w[32 * k] = x[j][k];
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (REDUCE == TP_REDUCE_MAX) atomicMax(reinterpret_cast<unsigned*>(w) + 32 * k, f2key(x[j][k]));
              else atomicAdd(w + 32 * k, x[j][k]);
            }
          }
        }
      }
    } else
    for (int e0 = warp * 4; e0 < npts; e0 += kWarps * 4) {
      int2 en[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) en[j] = (e0 + j < npts) ? __ldg(ent + e0 + j) : make_int2(-1, 0);
      for (int v = lane; v < C4; v += 32) {
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (en[j].x >= 0)
            x[j] = ld_keep_f4(reinterpret_cast<const float4*>(P.feats + (int64_t)en[j].x * P.feat_stride) + v, pol_in);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (en[j].x < 0) continue;
          float* w = s_work + en[j].y * C + v * 4;
          if (s_cnt[en[j].y] == 1) {  // warp-uniform: sole owner of the row
            if (REDUCE == TP_REDUCE_MAX)
              *reinterpret_cast<uint4*>(w) = make_uint4(f2key(x[j].x), f2key(x[j].y), f2key(x[j].z), f2key(x[j].w));
            else
              *reinterpret_cast<float4*>(w) = x[j];
          } else if (REDUCE == TP_REDUCE_MAX) {
            unsigned* wk = reinterpret_cast<unsigned*>(w);
            atomicMax(wk + 0, f2key(x[j].x));
            atomicMax(wk + 1, f2key(x[j].y));
            atomicMax(wk + 2, f2key(x[j].z));
            atomicMax(wk + 3, f2key(x[j].w));
          } else {
            atomicAdd(w + 0, x[j].x);
            atomicAdd(w + 1, x[j].y);
            atomicAdd(w + 2, x[j].z);
            atomicAdd(w + 3, x[j].w);
          }
        }
      }
    }
    __syncthreads();
    if (split_m >= 0) {
      // ---- fold this item's partial rows into the output tile (keys / sums) and its cell counters ------
      int32_t* gcnt = P.split_cnt + (int64_t)split_m * kMaxCpt;
      for (int c = warp; c < ncell; c += kWarps) {
        const int cnt = s_cnt[c];
        if (cnt <= 0) continue;
        const float* row = s_work + c * C;
        float* grow = gdst + (int64_t)c * C;
        for (int v = lane; v < C4; v += 32) {
          if (REDUCE == TP_REDUCE_MAX) {
            const uint4 kk = *reinterpret_cast<const uint4*>(row + v * 4);
            unsigned* gk = reinterpret_cast<unsigned*>(grow + v * 4);
            atomicMax(gk + 0, kk.x); atomicMax(gk + 1, kk.y); atomicMax(gk + 2, kk.z); atomicMax(gk + 3, kk.w);
          } else {
            const float4 x = *reinterpret_cast<const float4*>(row + v * 4);
            red_add_v4(reinterpret_cast<float4*>(grow + v * 4), x);
          }
        }
        if (lane == 0) atomicAdd(gcnt + c, cnt);
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const int items = (P.tile_cnt[t] + kSub - 1) / kSub;
        s_last = atomicAdd(P.split_done + split_m, 1) == items - 1;
      }
      __syncthreads();
      if (!s_last) return;
      // ---- last item of the tile: keys -> floats / sums -> means, in place (the tile is in L2) ---------
      __threadfence();
      for (int c = warp; c < ncell; c += kWarps) {
        const int cnt = __ldcg(gcnt + c);
        float4* grow = reinterpret_cast<float4*>(gdst + (int64_t)c * C);
        for (int v = lane; v < C4; v += 32) {
          float4 o = fill4;
          if (cnt > 0) {
            const float4 x = __ldcg(grow + v);
            if (REDUCE == TP_REDUCE_MAX) {
              o = make_float4(key2f(__float_as_uint(x.x)), key2f(__float_as_uint(x.y)), key2f(__float_as_uint(x.z)),
                              key2f(__float_as_uint(x.w)));
              if (P.clamp_zero) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
            } else if (REDUCE == TP_REDUCE_MEAN) {
              const float d = (float)cnt;
              o = make_float4(__fdiv_rn(x.x, d), __fdiv_rn(x.y, d), __fdiv_rn(x.z, d), __fdiv_rn(x.w, d));
            } else {
              o = x;
            }
          }
          grow[v] = o;
        }
        if (gcount && lane == 0) gcount[c] = cnt;
      }
      return;  // tile_cnt[t] stays >= kSplit until the last CTA cleans up: phase 2 must keep skipping this tile
    }
    // finalise touched cells in place: keys -> floats, or sum -> mean
    for (int c = warp; c < ncell; c += kWarps) {
      const int cnt = s_cnt[c];
      if (cnt > 0 && REDUCE != TP_REDUCE_SUM) {
        float4* row = reinterpret_cast<float4*>(s_work + c * C);
        for (int v = lane; v < C4; v += 32) {
          if (REDUCE == TP_REDUCE_MAX) {
            const uint4 kk = *reinterpret_cast<const uint4*>(row + v);
            float4 o = make_float4(key2f(kk.x), key2f(kk.y), key2f(kk.z), key2f(kk.w));
            if (P.clamp_zero) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
            row[v] = o;
          } else {
            const float4 o = row[v];
            const float d = (float)cnt;
            row[v] = make_float4(__fdiv_rn(o.x, d), __fdiv_rn(o.y, d), __fdiv_rn(o.z, d), __fdiv_rn(o.w, d));
          }
        }
      }
      if (gcount && lane == 0) gcount[c] = cnt;
      if (P.partial && cnt == 0) {  // untouched cell of an occupied tile: -inf, and mark the row dirty
        float4* row = reinterpret_cast<float4*>(s_work + c * C);
        for (int v = lane; v < C4; v += 32) row[v] = fill4;
        __syncwarp();
        if (lane == 0) s_cnt[c] = -1;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) { bulk_store(gdst, s_work, (unsigned)ncell * (unsigned)C * 4u, pol_out); in_flight = true; }
  };

  // Work order. Phase 0: items of the split tiles (the longest chains of all, cut into kSub-entry pieces). Phase 1:
  // the other long tiles — a tile that collects hundreds of points (the ground plane folds onto one z row of the
  // yz / xz planes) is a ~50 us chain of gathers for one CTA; started last it is the tail of the kernel, started first
  // it hides under the streaming of the rest (the alloc pass listed them; CTAs pop them one at a time). Phase 2:
  // everything else in memory order, kGrab tiles per scheduler atomic. ONE call site of process_tile for all three
  // (inlined three times the kernel was 67 KB of SASS).
  const int nitems = P.split_ctr[1];
  const int nheavy = P.heavy[0];
  for (int phase = 0;;) {
    __syncthreads();
    int nbatch = 1;
    if (phase < 2) {
      if (tid == 0) {
        const long long i = (long long)atomicAdd(&sched[phase == 0 ? 3 : 2], 1ull);
        s_t0 = i;
        if (phase == 0 && i < nitems) {
          const int2 it = P.split_item[i];
          const int t = P.split_tile[it.x];
          const int np = P.tile_cnt[t];  // stays >= kSplit until the last CTA of the launch cleans up
          s_tile[0] = t;
          s_npts[0] = min(kSub, np - it.y * kSub);
          s_start[0] = P.tile_start[t] + it.y * kSub;
          s_split[0] = it.x;
        } else if (phase == 1 && i < nheavy) {
          const int t = P.heavy[1 + i];
          s_tile[0] = t;
          s_npts[0] = P.tile_cnt[t];  // stays >= kHeavy until the last CTA cleans up: phase 2 of other CTAs may
          s_start[0] = P.tile_start[t];  // already be running and must keep skipping it
          s_split[0] = -1;
        }
      }
      __syncthreads();
      if (s_t0 >= (phase == 0 ? nitems : nheavy)) {
        ++phase;
        continue;
      }
    } else {
      if (tid < kGrab) {
        long long t0 = 0;
        if (tid == 0) t0 = (long long)atomicAdd(&sched[0], (unsigned long long)kGrab);
        t0 = __shfl_sync((1u << kGrab) - 1u, t0, 0);
        if (tid == 0) s_t0 = t0;
        const long long t = t0 + tid;
        int np = -1, st = 0;  // -1: not this phase's (past the end, or a heavy / split tile)
        if (t < P.tiles_total) {
          np = P.tile_cnt[t];
          if (np >= kHeavy) {
            np = -1;
          } else if (np > 0) {
            st = P.tile_start[t];
            P.tile_cnt[t] = 0;  // leave the counters clean for the next call
          }
        }
        s_tile[tid] = (int)t;
        s_npts[tid] = np;
        s_start[tid] = st;
        s_split[tid] = -1;
      }
      __syncthreads();
      if (s_t0 >= P.tiles_total) break;
      nbatch = kGrab;
    }
#pragma unroll 1
    for (int gi = 0; gi < nbatch; ++gi) {
      if (s_npts[gi] < 0) continue;
      process_tile(s_tile[gi], s_npts[gi], s_start[gi], s_split[gi]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    bulk_wait0();  // shared memory must outlive the bulk stores that read it
    __threadfence();
    s_t0 = (atomicAdd(&sched[1], 1ull) == (unsigned long long)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_t0) {  // last CTA out: leave the heavy tiles' counters and the scheduler slot clean
    for (int i = tid; i < nheavy; i += kRedThreads) P.tile_cnt[P.heavy[1 + i]] = 0;
    for (int i = tid; i < P.split_ctr[0]; i += kRedThreads) P.tile_cnt[P.split_tile[i]] = 0;
    __syncthreads();
    if (tid == 0) {
      sched[0] = 0;
      sched[1] = 0;
      sched[2] = 0;
      sched[3] = 0;
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

// after the cross-GPU max of TP_REDUCE_MAX_PARTIAL planes: -inf (no point on any GPU) -> 0
__global__ void __launch_bounds__(256)
finalize_max_kernel(float4* __restrict__ planes, int64_t nvec, int clamp_zero) {
  const float ninf = __uint_as_float(0xff800000u);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
       v += (int64_t)gridDim.x * blockDim.x) {
    float4 x = planes[v];
    const float lo = clamp_zero ? 0.f : ninf;
    // a cell is empty on all GPUs iff all of its channels are -inf; per element is equivalent because
    // a touched cell has a finite (or NaN / +inf) value in every channel unless the feature itself is -inf
    x.x = (x.x == ninf) ? 0.f : fmaxf(x.x, lo);
    x.y = (x.y == ninf) ? 0.f : fmaxf(x.y, lo);
    x.z = (x.z == ninf) ? 0.f : fmaxf(x.z, lo);
    x.w = (x.w == ninf) ? 0.f : fmaxf(x.w, lo);
    planes[v] = x;
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

static int cells_per_tile(int C) {
  int c = 1;
  while (c * 2 * C <= kTileFloats && c * 2 <= kMaxCpt) c *= 2;
  return c;
}

struct EncodeLayout {
  int64_t tiles_per_sample[3], tiles_all, tiles_total;
  int64_t off_cnt, off_start, off_heavy, off_split, off_rank, off_ent, bytes, tmax, max_split, max_items;
};

static EncodeLayout encode_layout(const GeomDev& g, int batch, int64_t n, int C, const bool use[3]) {
  EncodeLayout L;
  int64_t cps[3];
  plane_cells(g, cps);
  const int cpt = cells_per_tile(C);
  L.tiles_all = 0;
  for (int k = 0; k < 3; ++k) {
    L.tiles_per_sample[k] = use[k] ? (cps[k] + cpt - 1) / cpt : 0;
    L.tiles_all += L.tiles_per_sample[k];
  }
  L.tiles_total = L.tiles_all * batch;
  // region offsets do not depend on C or on which planes are used (sized for the most tiles, C = 512,
  // all planes), so the "tile counters are zero between calls" invariant survives a change of C
  int64_t tmax = 0;
  const int cpt_min = cells_per_tile(512);
  for (int k = 0; k < 3; ++k) tmax += (cps[k] + cpt_min - 1) / cpt_min;
  tmax *= batch;
  auto pad = [](int64_t b) { return (b + 255) / 256 * 256; };
  L.tmax = tmax;
  L.off_cnt = 0;
  L.off_start = L.off_cnt + pad(tmax * 4);
  L.off_heavy = L.off_start + pad((tmax + 1) * 4);
  L.off_split = L.off_heavy + pad((tmax + 2) * 4);  // + the CSR cursor
  // split tiles: each holds >= kSplit of the 3n list entries; items: ceil(c / kSub) <= c / kSub + 1 each
  L.max_split = 3 * n / kSplit + 1;
  L.max_items = 3 * n / kSub + L.max_split + 1;
  L.off_rank = L.off_split + pad(8) + pad(L.max_split * 4) + pad(L.max_items * 8) + pad(L.max_split * 4) +
               pad(L.max_split * (int64_t)kMaxCpt * 4);
  L.off_ent = L.off_rank + pad(3 * n * 4);
  L.bytes = L.off_ent + pad(3 * n * 8);
  return L;
}

// workspace is sized for all three planes and the smallest tile (C = 4 would give the fewest,
// largest tiles; C = 512 the most), so one workspace serves every C.
extern "C" int64_t tp_encode_workspace_bytes(const tp_geom* geom, int32_t batch, int64_t n_total) {
  if (!geom || batch <= 0 || n_total < 0) return -1;
  const bool all[3] = {true, true, true};
  return encode_layout(make_geom_dev(*geom), batch, n_total, 512, all).bytes;
}

// The tile counters (the first region of the workspace) must be zero before the first call; every
// call leaves them zero.
extern "C" int tp_encode_workspace_init(void* workspace, int64_t workspace_bytes, void* stream) {
  if (!workspace) return fail(TP_E_NULL, "tp_encode_workspace_init: null workspace");
  TP_CUDA(cudaMemsetAsync(workspace, 0, (size_t)workspace_bytes, (cudaStream_t)stream));
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
  if (batch <= 0 || n < 0 || 3 * n >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_encode_f32: batch=%d n=%lld", batch, (long long)n);
  if (!offsets || !workspace) return fail(TP_E_NULL, "tp_encode_f32: offsets/workspace null");
  if (n > 0 && !feats) return fail(TP_E_NULL, "tp_encode_f32: feats null");
  if (n > 0 && !idx && !points) return fail(TP_E_NULL, "tp_encode_f32: need idx or points");
  if (!idx && point_stride < 3 && n > 0) return fail(TP_E_SHAPE, "tp_encode_f32: point_stride=%d", point_stride);
  if (feat_stride < C || (feat_stride & 3) || ((uintptr_t)feats & 15))
    return fail(TP_E_SHAPE, "tp_encode_f32: feats must be 16-byte aligned rows (stride=%lld)", (long long)feat_stride);
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_encode_f32: unknown arith %d", arith);
  if (reduce < 0 || reduce > 3) return fail(TP_E_ENUM, "tp_encode_f32: unknown reduce %d", reduce);
  const int64_t need = tp_encode_workspace_bytes(geom, batch, n);
  if (workspace_bytes < need) return fail(TP_E_WORKSPACE, "tp_encode_f32: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  for (float* o : {out_xy, out_yz, out_xz})
    if (o && ((uintptr_t)o & 15)) return fail(TP_E_SHAPE, "tp_encode_f32: outputs must be 16-byte aligned");

  EncodeParams P;
  P.g = make_geom_dev(*geom);
  P.batch = batch;
  P.C = C;
  P.C4 = C / 4;
  P.cpt = cells_per_tile(C);
  P.n = n;
  int64_t cps[3];
  plane_cells(P.g, cps);
  const bool use[3] = {out_xy != nullptr, out_yz != nullptr, out_xz != nullptr};
  const EncodeLayout L = encode_layout(P.g, batch, n, C, use);
  if (L.tiles_total >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_encode_f32: too many tiles");
  int64_t cb = 0;
  for (int k = 0; k < 3; ++k) {
    P.cells_per_sample[k] = cps[k];
    P.tiles_per_sample[k] = L.tiles_per_sample[k];
    P.count_base[k] = cb;
    cb += cps[k] * batch;
  }
  P.tiles_all = L.tiles_all;
  P.tiles_total = L.tiles_total;
  P.idx = idx;
  P.points = points;
  P.point_stride = point_stride;
  P.offsets = offsets;
  P.feats = feats;
  P.feat_stride = feat_stride;
  char* ws = reinterpret_cast<char*>(workspace);
  P.tile_cnt = reinterpret_cast<int32_t*>(ws + L.off_cnt);
  P.tile_start = reinterpret_cast<int32_t*>(ws + L.off_start);
  P.heavy = reinterpret_cast<int32_t*>(ws + L.off_heavy);
  P.cursor = P.heavy + 1 + L.tmax;
  {
    auto pad = [](int64_t b) { return (b + 255) / 256 * 256; };
    char* q = ws + L.off_split;
    P.split_ctr = reinterpret_cast<int32_t*>(q); q += pad(8);
    P.split_tile = reinterpret_cast<int32_t*>(q); q += pad(L.max_split * 4);
    P.split_item = reinterpret_cast<int2*>(q); q += pad(L.max_items * 8);
    P.split_done = reinterpret_cast<int32_t*>(q); q += pad(L.max_split * 4);
    P.split_cnt = reinterpret_cast<int32_t*>(q);
  }
  P.rank = reinterpret_cast<int32_t*>(ws + L.off_rank);
  P.entries = reinterpret_cast<int2*>(ws + L.off_ent);
  P.out[0] = out_xy;
  P.out[1] = out_yz;
  P.out[2] = out_xz;
  P.cell_count = cell_count;
  P.clamp_zero = (reduce == TP_REDUCE_MAX_PARTIAL) ? 0 : clamp_zero;  // clamp after the cross-GPU max
  P.partial = (reduce == TP_REDUCE_MAX_PARTIAL) ? 1 : 0;
  // 179 MB of rows (350k points) do not fit in the 126 MB L2: pinning them only thrashes it (550 -> 500 us)
  P.keep_rows = (n * (int64_t)C * 4 <= ((int64_t)48 << 20)) ? 1 : 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (L.tiles_total == 0) return 0;

  if (n > 0) {
    int64_t blocks = (n + 255) / 256;
    int grid = (int)(blocks < (int64_t)kSMs * 8 ? blocks : (int64_t)kSMs * 8);
    if (arith == TP_ARITH_TORCH_CUDA) encode_count_kernel<TP_ARITH_TORCH_CUDA><<<grid, 256, 0, s>>>(P);
    else encode_count_kernel<TP_ARITH_TORCH_CPU><<<grid, 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("encode_count_kernel");
    const int64_t ablocks = (L.tiles_total + 255) / 256;
    encode_alloc_kernel<<<(int)(ablocks < (int64_t)kSMs * 4 ? ablocks : (int64_t)kSMs * 4), 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("encode_alloc_kernel");
    if (arith == TP_ARITH_TORCH_CUDA) encode_fill_kernel<TP_ARITH_TORCH_CUDA><<<grid, 256, 0, s>>>(P);
    else encode_fill_kernel<TP_ARITH_TORCH_CPU><<<grid, 256, 0, s>>>(P);
    TP_LAUNCH_CHECK("encode_fill_kernel");
  } else {
    TP_CUDA(cudaMemsetAsync(P.heavy, 0, 4, s));  // no points: no heavy / split tiles (the alloc pass did not run)
    TP_CUDA(cudaMemsetAsync(P.split_ctr, 0, 8, s));
  }
  constexpr int kSmem = kTileFloats * 4 + kMaxCpt * 4;  // 17 KB
  const size_t smem = kSmem;
  const int64_t ctas = (L.tiles_total + kGrab - 1) / kGrab;
  const int64_t cap = (int64_t)kSMs * kRedCtasPerSm;  // persistent: one resident wave
  const int grid = (int)(ctas < cap ? ctas : cap);
  static std::atomic<unsigned> next_slot{0};
  const int slot = (int)(next_slot.fetch_add(1) % kEncSchedSlots);
#define TP_RED(R)                                                                              \
  if (C == 128) encode_reduce_kernel<R, true><<<grid, kRedThreads, smem, s>>>(P, slot);         \
  else encode_reduce_kernel<R, false><<<grid, kRedThreads, smem, s>>>(P, slot);
  if (reduce == TP_REDUCE_MAX || reduce == TP_REDUCE_MAX_PARTIAL) { TP_RED(TP_REDUCE_MAX) }
  else if (reduce == TP_REDUCE_MEAN) { TP_RED(TP_REDUCE_MEAN) }
  else { TP_RED(TP_REDUCE_SUM) }
#undef TP_RED
  TP_LAUNCH_CHECK("encode_reduce_kernel");
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

extern "C" int tp_encode_finalize_max_f32(float* planes, int64_t n_floats, int32_t clamp_zero, void* stream) {
  if (!planes) return fail(TP_E_NULL, "tp_encode_finalize_max_f32: null argument");
  if (n_floats < 0 || (n_floats & 3) || ((uintptr_t)planes & 15))
    return fail(TP_E_SHAPE, "tp_encode_finalize_max_f32: need a 16-byte aligned multiple of 4 floats");
  if (n_floats == 0) return 0;
  int64_t blocks = (n_floats / 4 + 255) / 256;
  int grid = (int)(blocks < (int64_t)kSMs * 16 ? blocks : (int64_t)kSMs * 16);
  finalize_max_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(planes), n_floats / 4, clamp_zero);
  TP_LAUNCH_CHECK("finalize_max_kernel");
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
