// Occupancy decode + head in ONE kernel: TriplaneOcc samples the three planes on the voxel lattice
// (sample_points_triplane, triplane_occ.py:321-348) and feeds the [B,32,X,Y,Z] features straight into the `Mlp`
// head (triplane_occ.py:182-186, dense_heads/mlp.py:57-70). Run as two kernels that is 128 B/query written and read
// again; here a lattice block's features go from the decode tables (tp_sample_grid.cuh) into the tensor-core A tile
// in shared memory and only the num_classes logits leave the SM: 12 + 4*ncls bytes per query.
//
// Per CTA (256 threads, 3 per SM): blocks of 4 x 8 x 16 queries. Phases A-C are the lattice decode's (query check,
// footprints, the three 2-D tables). Then, per lattice row of the block (= one tile of 128 queries):
//   stage : (xy + yz) + xz for 32 channels, rounded to TF32, 16-byte stores into the MN-major A tile
//   layer 1..3 : the tcgen05 chain of tp_mlp.cu (A of layers 2 / 3 in TMEM), two warps per TMEM lane quadrant
//   logits: 4*ncls bytes per query
// pipelined like tp_mlp.cu: the next row is staged while layer 2 runs, its layer 1 is issued behind layer 3.
// A block that is not a lattice (checked bit for bit, as in the decode kernel) stages its rows per query instead.
// The features fed to the head are, bit for bit, what tp_sample3_grid_nhwc_f32 writes; the logits equal
// tp_mlp_head_tf32 on that tensor.
#include <atomic>

#include "tp_sample_grid.cuh"
#include "tp_umma.cuh"

namespace tp {

constexpr int kHeadBI = 4;
constexpr int kHeadCtasPerSm = 3;
constexpr int kHeadTmemCols = 128;  // D1 [0,64) | D2 [64,96) | D3 [96,112)
using HeadCfg = GridCfg<kHeadBI>;

// shared-memory map (bytes from a 1024-aligned base)
constexpr int kHOffA = 0;                      // [32 ch x 128 q] MN-major A tile, 16 KB
constexpr int kHOffW1 = 16384;                 // K-major SWIZZLE_128B weights, as in tp_mlp.cu
constexpr int kHOffW2 = kHOffW1 + 8192;
constexpr int kHOffW3 = kHOffW2 + 8192;
constexpr int kHOffTab = kHOffW3 + 2048;       // decode tables, then the entry records
constexpr int kHTabBytes = HeadCfg::kTableWords * 4;
constexpr int kHeadSmem = kHOffTab + kHTabBytes + HeadCfg::E * (16 + 8) + 1024;

// Blocks cost between nothing (outside all planes: zero logits) and four tensor-core tile chains, and a CTA only gets
// about three of them: they are handed out by an atomic ticket counter (first block by CTA index), claimed one block
// ahead so the next block's queries can be prefetched. The CTA that draws the last ticket of the launch (every CTA
// draws exactly one ticket past the end) resets the counter; concurrent launches use different counters.
constexpr int kHeadTicketSlots = 64;
__device__ unsigned int g_head_ticket[kHeadTicketSlots];

#ifdef TP_HEAD_TRACE  // tools/micro/head_trace.cu: phase timeline of every CTA's first gathering block (not in the library build)
__device__ unsigned long long g_head_cta[16 * 1024];
__device__ __forceinline__ unsigned long long head_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define HEAD_G(n) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_head_cta[blockIdx.x * 16 + (n)] = head_gtimer(); } while (0)
#else
#define HEAD_G(n) do {} while (0)
#endif

struct HeadParams {
  GridParams G;  // G.S.out unused
  int slot;
  const float* w1;
  const float* w2;
  const float* w3;
  float* logits;  // [B, ncls, Q]
  int ncls;
};

// Per-query staging of one lattice row for a block that is not a lattice: thread -> (query m = tid >> 1, 16 channels),
// the flat kernel's arithmetic (tp_sample_dev.cuh).
template <int ARITH>
__device__ __forceinline__ void stage_per_query(const SampleParams& P, const float* qrow, int d, int nj, int nk,
                                             const float4* pl0, const float4* pl1, const float4* pl2,
                                             unsigned char* atile) {
  constexpr int C4 = 8;
  const int tid = threadIdx.x;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  const unsigned long long pol_planes = policy_evict_last();
  const int m = tid >> 1, h16 = tid & 1, mj = m >> 4, mk = m & 15;
  float4 f[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) f[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mj < nj && mk < nk) {
    const float* qp = qrow + (mj * d + mk) * 3;
    const float px = __ldg(qp), py = __ldg(qp + 1), pz = __ldg(qp + 2);
    const float gx = grid_coord<ARITH>(P, px, 0), gy = grid_coord<ARITH>(P, py, 1), gz = grid_coord<ARITH>(P, pz, 2);
    float4 wgt[3];
    int base[3], msk[3];
    plane_setup<ARITH>(gx, gy, P.W[0], P.H[0], wgt[0], base[0], msk[0]);
    plane_setup<ARITH>(gy, gz, P.W[1], P.H[1], wgt[1], base[1], msk[1]);
    plane_setup<ARITH>(gx, gz, P.W[2], P.H[2], wgt[2], base[2], msk[2]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c4 = h16 * 4 + g;
      const float4 a0 = plane_taps<true>(pl0 + c4, base[0] * C4, C4, WC4_0, wgt[0], msk[0], pol_planes);
      const float4 a1 = plane_taps<true>(pl1 + c4, base[1] * C4, C4, WC4_1, wgt[1], msk[1], pol_planes);
      const float4 a2 = plane_taps<true>(pl2 + c4, base[2] * C4, C4, WC4_2, wgt[2], msk[2], pol_planes);
      f[g].x = __fadd_rn(__fadd_rn(a0.x, a1.x), a2.x);
      f[g].y = __fadd_rn(__fadd_rn(a0.y, a1.y), a2.y);
      f[g].z = __fadd_rn(__fadd_rn(a0.z, a1.z), a2.z);
      f[g].w = __fadd_rn(__fadd_rn(a0.w, a1.w), a2.w);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = h16 * 16 + g * 4;
    unsigned char* a = atile + (m & 3) * 4;
    *reinterpret_cast<uint32_t*>(a + mn_tile_off(c, m >> 2)) = rna_tf32(f[g].x);
    *reinterpret_cast<uint32_t*>(a + mn_tile_off(c + 1, m >> 2)) = rna_tf32(f[g].y);
    *reinterpret_cast<uint32_t*>(a + mn_tile_off(c + 2, m >> 2)) = rna_tf32(f[g].z);
    *reinterpret_cast<uint32_t*>(a + mn_tile_off(c + 3, m >> 2)) = rna_tf32(f[g].w);
  }
}

// Layer 2 and 3 issue (A operand in TMEM), shared by the lattice path and the per-query path
__device__ __forceinline__ void issue_layer2(uint32_t tmem, uint32_t sbase, uint32_t mbar) {
  constexpr uint32_t kI2 = umma_idesc_tf32(128, kMlpC);
#pragma unroll
  for (int k = 0; k < kMlpH / 8; ++k)
    umma_tf32_ts(tmem + 64, tmem + k * 8, umma_desc(sbase + kHOffW2 + (k >> 2) * 4096 + (k & 3) * 32), kI2, k > 0);
  umma_commit(mbar);
}
__device__ __forceinline__ void issue_layer3(uint32_t tmem, uint32_t sbase, uint32_t mbar) {
  constexpr uint32_t kI3 = umma_idesc_tf32(128, kMlpNOut);
#pragma unroll
  for (int k = 0; k < kMlpC / 8; ++k)
    umma_tf32_ts(tmem + 96, tmem + 64 + k * 8, umma_desc(sbase + kHOffW3 + k * 32), kI3, k > 0);
  umma_commit(mbar);
}
__device__ __forceinline__ void issue_layer1(uint32_t tmem, uint32_t sbase, uint32_t mbar) {
  // D1[128 x 64] = A[128 x 32] . W1^T, A MN-major: one 4096-byte pair of K atoms per step
  constexpr uint32_t kI1 = umma_idesc_tf32(128, kMlpH) | (1u << 15);
#pragma unroll
  for (int k = 0; k < kMlpC / 8; ++k)
    umma_tf32(tmem + 0, umma_desc_mn(sbase + kHOffA + k * 4096, 512, 2048), umma_desc(sbase + kHOffW1 + k * 32), kI1, k > 0);
  umma_commit(mbar);
}

// A block that is not a lattice: every row staged per query, the three layers one after the other (no pipelining:
// this is the rare path). Out of line so that its registers do not shape the lattice path. Returns the mbarrier phase.
template <int ARITH>
__device__ __noinline__ uint32_t fallback_block(const SampleParams& P, const float* q00, int wd, int d, int ni, int nj,
                                                int nk, const float4* pl0, const float4* pl1, const float4* pl2,
                                                unsigned char* smem, uint32_t tmem, uint32_t mbar1, uint32_t phase,
                                                float* lg, int ncls, bool e_ok) {
  const int warp = threadIdx.x >> 5;
  const uint32_t sbase = smem_u32(smem), mbar2 = mbar1 + 8, mbar3 = mbar1 + 16;
  const uint32_t t_quad = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int half = warp >> 2;
  for (int ii = 0; ii < ni; ++ii) {
    stage_per_query<ARITH>(P, q00 + (int64_t)ii * wd * 3, d, nj, nk, pl0, pl1, pl2, smem + kHOffA);
    fence_async_smem_mlp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) { tc_fence_after(); issue_layer1(tmem, sbase, mbar1); }
    mbar_wait(mbar1, phase);
    tc_fence_after();
    relu_tf32_inplace(t_quad + half * 32);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) { tc_fence_after(); issue_layer2(tmem, sbase, mbar2); }
    mbar_wait(mbar2, phase);
    tc_fence_after();
    relu_tf32_inplace16(t_quad + 64 + half * 16);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) { tc_fence_after(); issue_layer3(tmem, sbase, mbar3); }
    mbar_wait(mbar3, phase);
    phase ^= 1;
    tc_fence_after();
    if (warp < 4) {
      float v[16];
      tmem_ld16(t_quad + 96, v);
      if (e_ok) {
#pragma unroll
        for (int c = 0; c < kMlpNOut; ++c)
          if (c < ncls) st_cs_f1(lg + (int64_t)c * P.Q + ii * wd, v[c]);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  return phase;
}

template <int ARITH>
__global__ void __launch_bounds__(kGridThreads, kHeadCtasPerSm)
sample3_grid_head_kernel(const __grid_constant__ HeadParams HP) {
  using Cfg = HeadCfg;
  constexpr int BI = kHeadBI, BJ = kBJ, C4 = 8, C = 32;
  constexpr int QPT = BI * BJ * kBK / kGridThreads;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_mbar[3];
  __shared__ uint32_t s_tmem;
  __shared__ int s_vote[2];
  __shared__ int s_next;
  const GridParams& G = HP.G;
  const SampleParams& P = G.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // TMEM first (see tp_mlp.cu: the SM holds back the next CTA until this one has given up the allocation permit)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kHeadTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mbar1 = smem_u32(&s_mbar[0]), mbar2 = mbar1 + 8, mbar3 = mbar1 + 16;
  float* const T0 = reinterpret_cast<float*>(smem + kHOffTab);
  float* const T1 = T0 + 32 * Cfg::S0;
  float* const T2 = T1 + 32 * Cfg::S1;
  float4* const s_w = reinterpret_cast<float4*>(smem + kHOffTab + kHTabBytes);
  int2* const s_om = reinterpret_cast<int2*>(s_w + Cfg::E);

  // ---- one-time setup: weights (TF32, K-major SWIZZLE_128B), mbarriers ---------------------------------------
  for (int i = tid; i < kMlpH * kMlpC / 4; i += kGridThreads) {  // W1 [64][32]
    const float4 w = __ldg(reinterpret_cast<const float4*>(HP.w1) + i);
    *reinterpret_cast<float4*>(smem + kHOffW1 + swz128(i / (kMlpC / 4), i % (kMlpC / 4))) =
        make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
  }
  for (int i = tid; i < kMlpC * kMlpH / 4; i += kGridThreads) {  // W2 [32][64]: two K blocks of [32][32]
    const int n = i / (kMlpH / 4), c = i % (kMlpH / 4);
    const float4 w = __ldg(reinterpret_cast<const float4*>(HP.w2) + i);
    *reinterpret_cast<float4*>(smem + kHOffW2 + (c >> 3) * 4096 + swz128(n, c & 7)) =
        make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
  }
  for (int i = tid; i < kMlpNOut * kMlpC / 4; i += kGridThreads) {  // W3 [16][32], rows >= ncls are zero
    const int n = i / (kMlpC / 4);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < HP.ncls) w = __ldg(reinterpret_cast<const float4*>(HP.w3) + i);
    *reinterpret_cast<float4*>(smem + kHOffW3 + swz128(n, i % (kMlpC / 4))) =
        make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar1), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar2), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar3), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_vote[0] = 0;
  }
  fence_async_smem_mlp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t t_quad = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's lane quadrant
  const int half = warp >> 2;                                           // two warps per quadrant split the columns
  uint32_t phase = 0;

  const unsigned long long pol_planes = policy_evict_last();
  const int wd = G.w * G.d;
  const int nblocks = G.nblocks;
  const int l8 = tid & 7, ent = tid >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  float* const w0 = T0 + (4 * l8) * Cfg::S0 + ent;
  float* const w1t = T1 + (4 * l8) * Cfg::S1 + (ent ^ swz_bits(4 * l8));
  float* const w2t = T2 + (4 * l8) * Cfg::S2 + (ent ^ swz_bits(4 * l8));
  const int ak = tid & (kBK - 1), aj = (tid / kBK) % BJ, ia = tid / (kBK * BJ);
  const int kg = lane & 3, jj = lane >> 2;  // staging: lane -> (j, 4 consecutive k) = 16-byte chunk `lane` of a channel row
  // epilogue: TMEM lane m = (warp & 3) * 32 + lane <-> query (j = m >> 4, k = m & 15) of the tile's lattice row
  const int em = (warp & 3) * 32 + lane, ej = em >> 4, ek = em & 15;
  int zeroed = 0, nblk_done = 0;
  // staging bases of this thread (see `stage` below)
  unsigned char* const st_a = smem + kHOffA + mn_tile_off(warp, lane);
  const float* const st_t0 = T0 + warp * Cfg::S0 + jj;
  const float* const st_t1 = T1 + warp * Cfg::S1 + jj * kBK;
  const float* const st_t2 = T2 + warp * Cfg::S2;
  static_assert(kGridThreads / 32 == 8, "stage: 8 warps, channel = warp + 8 n");

  unsigned int ticket = 0;
  int blk = blockIdx.x;
  bool traced = false;
  HEAD_G(0);
  while (blk < nblocks) {
    if (!traced) HEAD_G(1);
    if (tid == 0) {  // consumed after the barrier that ends phase B
      ticket = atomicAdd(&g_head_ticket[HP.slot], 1u);
      s_next = (int)(gridDim.x + ticket);
    }
    const BlockPos bp = block_pos<BI>(G, blk);
    const int b = bp.b, i0 = bp.i0, j0 = bp.j0, k0 = bp.k0;
    const int ni = min(BI, G.h - i0), nj = min(BJ, G.w - j0), nk = min(kBK, G.d - k0);
    const float* q00 = P.queries + ((int64_t)b * P.Q + ((int64_t)i0 * G.w + j0) * G.d + k0) * 3;

    // ---- A / B: the block's queries, lattice check, one bilinear footprint per table entry (tp_sample_grid.cuh) ---
    const bool ok = grid_block_is_lattice<BI>(G, q00, ni, nj, nk, wd, tid);
    const int live = grid_build_records<ARITH, BI>(G, q00, ni, nj, nk, wd, C4, s_w, s_om, tid);
    grid_cast_vote(s_vote, nblk_done, live, ok, tid);
    __syncthreads();
    if (!traced) HEAD_G(2);
    const int vote = s_vote[nblk_done & 1];
    ++nblk_done;
    const bool separable = !(vote & 8);

    const int next = s_next;
    if (next < nblocks) {  // next block's queries: DRAM -> L2 behind this block's work
      const BlockPos np = block_pos<BI>(G, next);
      if (np.j0 + aj < G.w && np.k0 + ak < G.d && (ak & 1) == 0) {
        const float* n00 = P.queries + ((int64_t)np.b * P.Q + ((int64_t)np.i0 * G.w + np.j0) * G.d + np.k0) * 3;
#pragma unroll
        for (int t = 0; t < QPT; ++t) {
          const int ii = ia + t * (kGridThreads / (kBK * BJ));
          if (np.i0 + ii < G.h) prefetch_l2(n00 + (ii * wd + aj * G.d + ak) * 3);
        }
      }
    }

    const float4* const pl0 = reinterpret_cast<const float4*>(P.plane[0] + (int64_t)b * P.bstride[0]);
    const float4* const pl1 = reinterpret_cast<const float4*>(P.plane[1] + (int64_t)b * P.bstride[1]);
    const float4* const pl2 = reinterpret_cast<const float4*>(P.plane[2] + (int64_t)b * P.bstride[2]);
    float* const lg = HP.logits + (int64_t)b * HP.ncls * P.Q + ((int64_t)i0 * G.w + j0 + ej) * G.d + k0 + ek;
    const bool e_ok = ej < nj && ek < nk;  // this thread's epilogue query exists

    if (separable && (vote & 7) == 0) {
      // the whole block lies outside all three planes: features 0, and the bias-free head maps 0 to 0
      if (warp < 4 && e_ok)
        for (int ii = 0; ii < ni; ++ii)
          for (int c = 0; c < HP.ncls; ++c) st_cs_f1(lg + (int64_t)c * P.Q + ii * wd, 0.f);
      __syncthreads();  // s_next is rewritten at the top of the next block
      blk = next;
      continue;
    }
    pdl_wait();  // first read of the planes (tp_common.cuh); blocks outside all of them never get here
    if (!separable) {
      phase = fallback_block<ARITH>(P, q00, wd, G.d, ni, nj, nk, pl0, pl1, pl2, smem, tmem, mbar1, phase, lg, HP.ncls, e_ok);
      zeroed = 0;  // (the tables themselves are untouched, but keep the invariant simple)
      blk = next;
      continue;
    }
    {
      // ---- C: the three 2-D tables (32 channels = one chunk) ------------------------------------------------
      build_table<Cfg::E0 / 32>(vote & 1, zeroed & 1, pl0 + l8, C4, WC4_0, s_w + ent, s_om + ent, w0, Cfg::S0, true, pol_planes);
      build_table<Cfg::E1 / 32>(vote & 2, zeroed & 2, pl1 + l8, C4, WC4_1, s_w + Cfg::E0 + ent, s_om + Cfg::E0 + ent, w1t, Cfg::S1, true, pol_planes);
      build_table<Cfg::E2 / 32>(vote & 4, zeroed & 4, pl2 + l8, C4, WC4_2, s_w + Cfg::E0 + Cfg::E1 + ent, s_om + Cfg::E0 + Cfg::E1 + ent, w2t, Cfg::S2, true, pol_planes);
      zeroed = ~vote & 7;
      __syncthreads();
    }
    if (!traced) HEAD_G(3);

    // A block with one live plane has far fewer distinct queries than lattice points: with only yz in range the features
    // do not depend on i (one tile serves all rows of the block), with only xz they do not depend on j (the block's
    // BI x 16 (i, k) pairs fit one tile: tile row = i * 16 + k). Same features bit for bit, 4x fewer tile chains: on the
    // 640k occupancy lattice half of the blocks are of these two kinds.
    const bool only_yz = (vote & 7) == 2, only_xz = (vote & 7) == 4;
    const int ntiles = (only_yz || only_xz) ? 1 : ni;
    // features of lattice row ii -> A tile (MN-major: channel rows of 128 queries, 16-byte chunk = 4 consecutive k)
    // Warp w stages channels w, w + 8, w + 16, w + 24: everything that depends on the channel is (per-thread constant)
    // + n * (compile-time stride) — the A-tile chunk moves by two K atoms (4096 B), the table rows by 8 rows, and the
    // column swizzle of the yz / xz tables is ((n & 3) << 2) for warps 0-7.
    auto stage = [&](int ii) {
      if (only_xz) {
        // lane >> 2 plays i here: chunk `lane` of a channel row = tile rows (lane >> 2) * 16 + 4 kg .. + 3
        const bool ik_ok = (jj < ni) && (kg * 4 < nk);
        const float* t2p = st_t2 + (jj < BI ? jj : 0) * kBK;
#pragma unroll
        for (int n = 0; n < C / 8; ++n) {
          const int xk = (kg * 4) ^ ((n & 3) << 2);
          uint4 r = make_uint4(0u, 0u, 0u, 0u);
          if (ik_ok) {  // (0 + 0) + xz, as the general path computes it
            const float4 s2 = *reinterpret_cast<const float4*>(t2p + n * 8 * Cfg::S2 + xk);
            r.x = rna_tf32(__fadd_rn(0.f, s2.x));
            r.y = rna_tf32(__fadd_rn(0.f, s2.y));
            r.z = rna_tf32(__fadd_rn(0.f, s2.z));
            r.w = rna_tf32(__fadd_rn(0.f, s2.w));
          }
          *reinterpret_cast<uint4*>(st_a + n * 4096) = r;
        }
      } else {
        const bool jk_ok = (jj < nj) && (kg * 4 < nk);
        const float* t0p = st_t0 + ii * BJ;
        const float* t2p = st_t2 + ii * kBK;
#pragma unroll
        for (int n = 0; n < C / 8; ++n) {
          const int xk = (kg * 4) ^ ((n & 3) << 2);
          const float4 s1 = *reinterpret_cast<const float4*>(st_t1 + n * 8 * Cfg::S1 + xk);
          const float s0 = t0p[n * 8 * Cfg::S0];
          const float4 s2 = *reinterpret_cast<const float4*>(t2p + n * 8 * Cfg::S2 + xk);
          uint4 r = make_uint4(0u, 0u, 0u, 0u);
          if (jk_ok) {
            r.x = rna_tf32(__fadd_rn(__fadd_rn(s0, s1.x), s2.x));  // (xy + yz) + xz  (triplane_occ.py:345)
            r.y = rna_tf32(__fadd_rn(__fadd_rn(s0, s1.y), s2.y));
            r.z = rna_tf32(__fadd_rn(__fadd_rn(s0, s1.z), s2.z));
            r.w = rna_tf32(__fadd_rn(__fadd_rn(s0, s1.w), s2.w));
          }
          *reinterpret_cast<uint4*>(st_a + n * 4096) = r;
        }
      }
      fence_async_smem_mlp();
    };
    stage(0);
    tc_fence_before();
    __syncthreads();
    if (!traced) HEAD_G(4);
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_layer1(tmem, sbase, mbar1);
    }
    for (int ii = 0; ii < ntiles; ++ii) {
      const bool has_next = ii + 1 < ntiles;
      mbar_wait(mbar1, phase);
      tc_fence_after();
      relu_tf32_inplace16(t_quad + half * 32);  // D1 -> layer-2 A operand, in place in TMEM (16 columns at a time:
      relu_tf32_inplace16(t_quad + half * 32 + 16);  //  80 registers per thread at 3 CTAs per SM)
      tmem_wait_st();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {  // layer 2: D2[128 x 32] = relu(D1)[128 x 64] . W2^T
        tc_fence_after();
        issue_layer2(tmem, sbase, mbar2);
      }
      if (has_next) stage(ii + 1);  // layer 1 of this row has consumed the A tile
      mbar_wait(mbar2, phase);
      tc_fence_after();
      relu_tf32_inplace16(t_quad + 64 + half * 16);
      tmem_wait_st();
      tc_fence_before();
      __syncthreads();  // also publishes the staged A tile
      if (warp == 0 && elect_one()) {  // layer 3: D3[128 x 16] = relu(D2)[128 x 32] . W3^T, then layer 1 of the next row
        tc_fence_after();
        issue_layer3(tmem, sbase, mbar3);
        if (has_next) issue_layer1(tmem, sbase, mbar1);
      }
      mbar_wait(mbar3, phase);
      phase ^= 1;
      tc_fence_after();
      if (warp < 4) {
        float v[16];
        tmem_ld16(t_quad + 96, v);
        if (only_yz) {  // the same logits for every lattice row of the block
          if (e_ok)
            for (int r = 0; r < ni; ++r)
#pragma unroll
              for (int c = 0; c < kMlpNOut; ++c)
                if (c < HP.ncls) st_cs_f1(lg + (int64_t)c * P.Q + r * wd, v[c]);
        } else if (only_xz) {  // tile row em = i * 16 + k: the same logits for every j of the block
          if (ej < ni && ek < nk) {
            float* lx = HP.logits + (int64_t)b * HP.ncls * P.Q + ((int64_t)(i0 + ej) * G.w + j0) * G.d + k0 + ek;
            for (int r = 0; r < nj; ++r)
#pragma unroll
              for (int c = 0; c < kMlpNOut; ++c)
                if (c < HP.ncls) st_cs_f1(lx + (int64_t)c * P.Q + r * G.d, v[c]);
          }
        } else if (e_ok) {
#pragma unroll
          for (int c = 0; c < kMlpNOut; ++c)
            if (c < HP.ncls) st_cs_f1(lg + (int64_t)c * P.Q + ii * wd, v[c]);
        }
      }
      tc_fence_before();
      if (!traced && ntiles == BI) HEAD_G(5 + ii);
    }
    if (!traced && ntiles == BI) { HEAD_G(9); traced = true; }
    __syncthreads();  // tables and records are rewritten by the next block
    blk = next;
  }
  if (tid == 0 && ticket == (unsigned)(nblocks - 1)) g_head_ticket[HP.slot] = 0u;  // nobody draws after the last ticket
  HEAD_G(10);
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kHeadTmemCols) : "memory");
  }
}

}  // namespace tp

using namespace tp;

static int grid_head_entry(const tp_plane planes[3], const float* queries, const int32_t dims[3], int32_t batch,
                           const tp_sample_geom* sg, int32_t arith, const float* w1, const float* w2, const float* w3,
                           int32_t num_classes, float* logits, void* stream, bool pdl) {
  constexpr int C = kMlpC;
  if (!dims) return fail(TP_E_NULL, "tp_sample3_grid_head_tf32: null dims");
  const int h = dims[0], w = dims[1], d = dims[2];
  if (h < 0 || w < 0 || d < 0) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: bad dims %d %d %d", h, w, d);
  if (num_classes <= 0 || num_classes > kMlpNOut) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: num_classes=%d must be in 1..%d", num_classes, kMlpNOut);
  if (batch <= 0) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: bad B=%d", batch);
  const int64_t Q = (int64_t)h * w * d;
  if (Q == 0) return 0;
  if (d & 3) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: d=%d must be a multiple of 4 (use tp_sample3_nhwc_f32 + tp_mlp_head_tf32)", d);
  if (Q * 3 >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: too many queries per sample");
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_sample3_grid_head_tf32: unknown arith %d", arith);
  if (!planes || !queries || !sg || !w1 || !w2 || !w3 || !logits) return fail(TP_E_NULL, "tp_sample3_grid_head_tf32: null argument");
  if (((uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)w3) & 15) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: weights must be 16-byte aligned");
  HeadParams HP;
  GridParams& G = HP.G;
  SampleParams& P = G.S;
  for (int k = 0; k < 3; ++k) {
    if (!planes[k].data) return fail(TP_E_NULL, "tp_sample3_grid_head_tf32: plane %d is null", k);
    if (planes[k].H <= 0 || planes[k].W <= 0 || (int64_t)planes[k].H * planes[k].W * C >= (int64_t)1 << 31 ||
        planes[k].H >= (1 << 20) || planes[k].W >= (1 << 20))
      return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: plane %d H=%d W=%d unsupported", k, planes[k].H, planes[k].W);
    if ((uintptr_t)planes[k].data & 15 || (planes[k].batch_stride & 3))
      return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: plane %d not 16-byte aligned", k);
    P.plane[k] = planes[k].data;
    P.bstride[k] = planes[k].batch_stride;
    P.H[k] = planes[k].H;
    P.W[k] = planes[k].W;
    P.lo[k] = sg->lo[k];
    P.vs[k] = sg->vs[k];
    P.rcp_vs[k] = 1.0f / sg->vs[k];
    P.half[k] = sg->half[k];
    P.rcp_half[k] = 1.0f / sg->half[k];
  }
  P.queries = queries;
  P.out = nullptr;
  P.Q = Q;
  P.C = C;
  P.tiles_per_sample = 0;
  P.tiles = 0;
  G.h = h; G.w = w; G.d = d;
  G.vec_ok = 1;
  G.nkb = (d + kBK - 1) / kBK;
  G.nib = (h + kHeadBI - 1) / kHeadBI;
  G.njb = (w + kBJ - 1) / kBJ;
  const int64_t nb = (int64_t)batch * G.nib * G.njb * G.nkb;
  if (nb >= ((int64_t)1 << 30)) return fail(TP_E_SHAPE, "tp_sample3_grid_head_tf32: too many queries");
  G.nblocks = (int)nb;
  static std::atomic<unsigned> launch_seq{0};
  HP.slot = (int)(launch_seq.fetch_add(1, std::memory_order_relaxed) % kHeadTicketSlots);
  HP.w1 = w1; HP.w2 = w2; HP.w3 = w3; HP.logits = logits; HP.ncls = num_classes;
  const int64_t cap = (int64_t)kHeadCtasPerSm * kSMs;
  const unsigned grid = (unsigned)(nb < cap ? nb : cap);
  cudaStream_t s = (cudaStream_t)stream;
  const int ai = arith == TP_ARITH_TORCH_CUDA ? 0 : 1;
  if (ai == 0) TP_CUDA(opt_in_smem<sample3_grid_head_kernel<TP_ARITH_TORCH_CUDA>>(kHeadSmem));
  else TP_CUDA(opt_in_smem<sample3_grid_head_kernel<TP_ARITH_TORCH_CPU>>(kHeadSmem));
  if (ai == 0) launch_kernel(sample3_grid_head_kernel<TP_ARITH_TORCH_CUDA>, grid, kGridThreads, kHeadSmem, s, pdl, HP);
  else launch_kernel(sample3_grid_head_kernel<TP_ARITH_TORCH_CPU>, grid, kGridThreads, kHeadSmem, s, pdl, HP);
  TP_LAUNCH_CHECK("sample3_grid_head_kernel");
  return 0;
}

extern "C" int tp_sample3_grid_head_tf32(const tp_plane planes[3], const float* queries, const int32_t dims[3],
                                         int32_t batch, const tp_sample_geom* sg, int32_t arith, const float* w1,
                                         const float* w2, const float* w3, int32_t num_classes, float* logits,
                                         void* stream) {
  return grid_head_entry(planes, queries, dims, batch, sg, arith, w1, w2, w3, num_classes, logits, stream, false);
}

extern "C" int tp_sample3_grid_head_nchw_tf32(const tp_plane planes_nchw[3], const float* queries,
                                              const int32_t dims[3], int32_t batch, const tp_sample_geom* sg,
                                              int32_t arith, const float* w1, const float* w2, const float* w3,
                                              int32_t num_classes, float* logits, float* ws, int64_t ws_floats,
                                              void* stream) {
  if (dims && (int64_t)dims[0] * dims[1] * dims[2] == 0) return 0;
  tp_plane nhwc[3];
  if (int rc = planes3_to_workspace("tp_sample3_grid_head_nchw_tf32", planes_nchw, kMlpC, batch, ws, ws_floats, nhwc, stream)) return rc;
  return grid_head_entry(nhwc, queries, dims, batch, sg, arith, w1, w2, w3, num_classes, logits, stream, true);
}
