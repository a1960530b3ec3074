// Occupancy head on the decode output: the reference's `Mlp` (mmdet3d/models/dense_heads/mlp.py:25-70) is three
// bias-free 1x1x1 Conv3d with ReLU between them, C -> 2C -> C -> num_classes, applied to the sampled triplane
// features [B,C,X,Y,Z] (triplane_occ.py:182-186). Per query that is a 32 -> 64 -> 32 -> 5 dense contraction: the one
// GEMM-shaped piece next to the hot path. Here the three layers run back to back on the 5th-gen
// tensor cores for a tile of 128 queries: layer-1 A (activations) and the weights B in shared memory in the
// canonical K-major SWIZZLE_128B layout, tcgen05.mma kind::tf32 issued by one elected lane, accumulators in TMEM,
// read back with tcgen05.ld for the ReLU and written back in place with tcgen05.st: the next layer's MMA takes its A
// operand from TMEM. The 2C and C wide intermediates never leave the SM: 4C bytes in and 4*num_classes bytes out
// per query instead of three passes over [B,2C,Q] tensors.
// Precision: TF32 operands with fp32 accumulation — what cuDNN gives the reference's Conv3d by default
// (torch.backends.cudnn.allow_tf32 = True); operands are rounded to nearest TF32, not truncated.
#include "tp_umma.cuh"

namespace tp {

constexpr int kMlpThreads = 128;   // 4 warps = the 4 TMEM lane quadrants = 128 rows of a tile
constexpr int kTmemCols = 128;     // D1 [0,64) | D2 [64,96) | D3 [96,112)

// shared-memory map (bytes, every operand tile 1024-byte aligned for the 128-byte swizzle). Only the layer-1
// activations and the weights live here: the hidden activations stay in TMEM (below). 35 KB per CTA; the CTA count
// per SM is set by TMEM, 4 x 128 columns.
constexpr int kOffA1 = 0;                       // [128 x 32]  layer-1 A, 16 KB
constexpr int kOffW1 = 16384;                   // [64 x 32]   8 KB
constexpr int kOffW2 = kOffW1 + 8192;           // [32 x 64]   2 K-blocks of 4 KB
constexpr int kOffW3 = kOffW2 + 8192;           // [16 x 32]   2 KB
constexpr int kMlpSmem = kOffW3 + 2048 + 1024;  // + slack to align the base to 1024 B

#ifdef TP_MLP_TRACE  // tools/micro/mlp_trace.cu: phase timestamps of CTA 0, start / end of every CTA (not in the library build)
__device__ long long g_mlp_trace[16 * 12];
__device__ unsigned long long g_mlp_cta[2 * 1024];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MLP_T(n) do { if (tid == 0 && blockIdx.x == 0 && it < 12) g_mlp_trace[it * 16 + (n)] = clock64(); } while (0)
#define MLP_G(n) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_mlp_cta[blockIdx.x * 2 + (n)] = gtimer(); } while (0)
#define MLP_P(n) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_mlp_trace[11 * 16 + (n)] = clock64(); } while (0)
#else
#define MLP_P(n) do {} while (0)
#define MLP_G(n) do {} while (0)
#define MLP_T(n) do {} while (0)
#endif

struct MlpParams {
  const float* feats;  // [B, C, Q]
  const float* w1;     // [2C, C]
  const float* w2;     // [C, 2C]
  const float* w3;     // [ncls, C]
  float* logits;       // [B, ncls, Q]
  int Q;                       // < 2^27 (host-checked): channel offsets c * Q fit 32 bits
  int tiles_per_sample, tiles;
  int ncls;
};

template <bool VEC>
__global__ void __launch_bounds__(kMlpThreads, 4)
mlp_head_kernel(const MlpParams P) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_mbar[3];  // layer 1 / 2 / 3 complete
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  MLP_P(0);
  MLP_G(0);
  // TMEM first: the SM does not start the next CTA of this kernel until the resident one has allocated and given up
  // its allocation permit (measured: with the allocation after the weight staging the four CTAs of an SM started
  // 3.2, 5.2 and 7.7 us apart — 11 us of a 38 us kernel — instead of 0.6 us apart)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t mbar1 = smem_u32(&s_mbar[0]), mbar2 = mbar1 + 8, mbar3 = mbar1 + 16;

  const int lane = tid & 31;
  // Tile row (= TMEM lane = thread) = query inside the tile. VEC (Q % 4 == 0, 16-byte aligned feats): a lane loads, for 8
  // channels, the float4 of 4 consecutive queries 4g..4g+3 (g = lane & 7; channel 4i + (lane >> 3)) — 8 requests of 4
  // full lines per warp instead of 32 of one line — and stores it as one 16-byte chunk of an MN-major A tile (the layout
  // the features have in memory: queries contiguous per channel), so staging is 8 conflict-free 16-byte stores.
  const int g8 = lane & 7, cl = lane >> 3;
  const int qrow = tid;  // query (inside the tile) of this thread's TMEM lane

  // (sample, tile inside the sample) of this CTA's tiles, advanced without divisions
  auto advance = [&](int& b, int& t, int by) {
    t += by;
    while (t >= P.tiles_per_sample) { t -= P.tiles_per_sample; ++b; }
  };
  const uint32_t qs = (uint32_t)P.Q;
  // the 32 values this thread stages for one tile
  auto load_rows = [&](int b, int t, bool live, float* x) {
    if (VEC) {
      const int q = t * 128 + warp * 32 + 4 * g8;
      if (live && q < P.Q) {
        const float* f = P.feats + ((int64_t)b * kMlpC + cl) * P.Q + q;  // one 64-bit base, 32-bit channel offsets
#pragma unroll
        for (int i = 0; i < kMlpC / 4; ++i) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(f + (size_t)((uint32_t)(4 * i) * qs)));
          x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        return;
      }
    } else {
      const int q = t * 128 + tid;
      if (live && q < P.Q) {
        const float* f = P.feats + ((int64_t)b * kMlpC) * P.Q + q;
#pragma unroll
        for (int c = 0; c < kMlpC; ++c) x[c] = __ldg(f + (size_t)((uint32_t)c * qs));
        return;
      }
    }
#pragma unroll
    for (int c = 0; c < kMlpC; ++c) x[c] = 0.f;
  };
  float xn[kMlpC];  // next tile's rows: loaded while this tile's three MMAs run (the first: during the setup below)
  int b = 0, t = 0, bn = 0, tn = 0;  // current and next tile
  advance(b, t, blockIdx.x);
  bn = b; tn = t;
  load_rows(b, t, (int)blockIdx.x < P.tiles, xn);

  MLP_P(1);
  // ---- one-time setup: weights into their swizzled tiles, TMEM, mbarrier ---------------------------------
#ifndef TP_MLP_TRACE_NOW
  {
    // all nine 16-byte weight loads of a thread in flight before the first conversion
    constexpr int N1 = kMlpH * kMlpC / 4 / kMlpThreads, N2 = kMlpC * kMlpH / 4 / kMlpThreads;  // 4, 4
    static_assert(kMlpNOut * kMlpC / 4 == kMlpThreads, "W3 is one float4 per thread");
    float4 a1[N1], a2[N2], a3 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < N1; ++u) a1[u] = __ldg(reinterpret_cast<const float4*>(P.w1) + tid + u * kMlpThreads);
#pragma unroll
    for (int u = 0; u < N2; ++u) a2[u] = __ldg(reinterpret_cast<const float4*>(P.w2) + tid + u * kMlpThreads);
    if (tid / (kMlpC / 4) < P.ncls) a3 = __ldg(reinterpret_cast<const float4*>(P.w3) + tid);  // rows >= ncls stay zero
    auto put = [&](int off, float4 w) {
      *reinterpret_cast<float4*>(smem + off) = make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
    };
#pragma unroll
    for (int u = 0; u < N1; ++u) {  // W1 [64][32]: row n, 16-byte chunk k/4
      const int i = tid + u * kMlpThreads;
      put(kOffW1 + swz128(i / (kMlpC / 4), i % (kMlpC / 4)), a1[u]);
    }
#pragma unroll
    for (int u = 0; u < N2; ++u) {  // W2 [32][64]: two K blocks of [32][32]
      const int i = tid + u * kMlpThreads, n = i / (kMlpH / 4), c = i % (kMlpH / 4);
      put(kOffW2 + (c >> 3) * 4096 + swz128(n, c & 7), a2[u]);
    }
    put(kOffW3 + swz128(tid / (kMlpC / 4), tid % (kMlpC / 4)), a3);  // W3 [16][32]
  }
#endif
  MLP_P(2);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar1), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar2), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar3), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem_mlp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  MLP_P(3);
  const uint32_t tmem = s_tmem;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's lane quadrant
  constexpr uint32_t kI1 = umma_idesc_tf32(128, kMlpH), kI2 = umma_idesc_tf32(128, kMlpC), kI3 = umma_idesc_tf32(128, kMlpNOut);
  int it = -1;
  // Software pipeline over this CTA's tiles: layer 1 of tile n+1 is issued together with layer 3 of tile n, and the
  // A1 tile of n+1 is staged (and the rows of n+2 requested) while layer 2 of tile n runs, so a tile costs two CTA
  // barriers and one exposed MMA latency instead of four and three.
  auto stage_a1 = [&]() {  // xn -> swizzled K-major A1 tile, then request the following tile's rows
    if (VEC) {
      // MN-major A1: this lane's 4 queries are 16-byte chunk 8 warp + g8 of K-row (channel) c
#pragma unroll
      for (int i = 0; i < kMlpC / 4; ++i) {
        const int c = 4 * i + cl;
        *reinterpret_cast<uint4*>(smem + kOffA1 + mn_tile_off(c, warp * 8 + g8)) =
            make_uint4(rna_tf32(xn[4 * i]), rna_tf32(xn[4 * i + 1]), rna_tf32(xn[4 * i + 2]), rna_tf32(xn[4 * i + 3]));
      }
    } else {
#pragma unroll
      for (int c16 = 0; c16 < kMlpC / 4; ++c16)
        *reinterpret_cast<uint4*>(smem + kOffA1 + swz128(tid, c16)) =
            make_uint4(rna_tf32(xn[c16 * 4]), rna_tf32(xn[c16 * 4 + 1]), rna_tf32(xn[c16 * 4 + 2]), rna_tf32(xn[c16 * 4 + 3]));
    }
    fence_async_smem_mlp();
  };
  auto issue_layer1 = [&]() {  // D1[128 x 64] = A1[128 x 32] . W1^T
#pragma unroll
    for (int k = 0; k < kMlpC / 8; ++k)
      umma_tf32(tmem + 0, VEC ? umma_desc_mn(sbase + kOffA1 + k * 4096, 512, 2048) : umma_desc(sbase + kOffA1 + k * 32),
                umma_desc(sbase + kOffW1 + k * 32), VEC ? (kI1 | (1u << 15)) : kI1, k > 0);
    umma_commit(mbar1);
  };
  uint32_t phase = 0;  // all three mbarriers complete once per tile
#ifdef TP_MLP_TRACE_NOLOOP
  int tile = P.tiles;
#else
  int tile = blockIdx.x;
#endif
  if (tile < P.tiles) {
    stage_a1();
    advance(bn, tn, gridDim.x);
    load_rows(bn, tn, tile + (int)gridDim.x < P.tiles, xn);
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_layer1();
    }
  }
  for (; tile < P.tiles; tile += (int)gridDim.x) {
    ++it;
    MLP_T(0);
    const bool qv = t * 128 + qrow < P.Q;  // (b, t): the tile whose layer 1 is in flight
    const bool has_next = tile + (int)gridDim.x < P.tiles;
    mbar_wait(mbar1, phase);
    MLP_T(1);
    tc_fence_after();
    relu_tf32_inplace(t_lane + 0);  // ReLU in place: D1 becomes the layer-2 A operand, still in TMEM
    relu_tf32_inplace(t_lane + 32);
    tmem_wait_st();
    MLP_T(2);
    tc_fence_before();
    __syncthreads();
    MLP_T(3);
    // ---- layer 2: D2[128 x 32] = relu(D1)[128 x 64] . W2^T -------------------------------------------------
    if (warp == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kMlpH / 8; ++k)
        umma_tf32_ts(tmem + 64, tmem + k * 8, umma_desc(sbase + kOffW2 + (k >> 2) * 4096 + (k & 3) * 32), kI2, k > 0);
      umma_commit(mbar2);
    }
    MLP_T(4);
    // ---- under layer 2: the next tile's A1 (layer 1 of this tile has consumed the buffer) ------------------
    const int b_cur = b, t_cur = t;
    if (has_next) {
      stage_a1();
      b = bn; t = tn;
      advance(bn, tn, gridDim.x);
      load_rows(bn, tn, tile + 2 * (int)gridDim.x < P.tiles, xn);
    }
    MLP_T(5);
    mbar_wait(mbar2, phase);
    MLP_T(6);
    tc_fence_after();
    relu_tf32_inplace(t_lane + 64);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();  // also publishes the staged A1 tile
    MLP_T(7);
    // ---- layer 3: D3[128 x 16] = relu(D2)[128 x 32] . W3^T; then layer 1 of the next tile (D1 is free) -----
    if (warp == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kMlpC / 8; ++k)
        umma_tf32_ts(tmem + 96, tmem + 64 + k * 8, umma_desc(sbase + kOffW3 + k * 32), kI3, k > 0);
      umma_commit(mbar3);
      if (has_next) issue_layer1();
    }
    MLP_T(8);
    mbar_wait(mbar3, phase);
    MLP_T(9);
    phase ^= 1;
    tc_fence_after();
    {
      float v[16];
      tmem_ld16(t_lane + 96, v);
      if (qv) {
        float* o = P.logits + ((int64_t)b_cur * P.ncls) * P.Q + (t_cur * 128 + qrow);
#pragma unroll
        for (int c = 0; c < kMlpNOut; ++c)
          if (c < P.ncls) st_cs_f1(o + (size_t)((uint32_t)c * qs), v[c]);
      }
    }
    // D3 is overwritten two barriers from here (layer 3 of the next tile), D1 / D2 after this tile's waits: no
    // barrier needed at the end of a tile
    tc_fence_before();
    MLP_T(10);
  }
  __syncthreads();
  MLP_P(4);
  MLP_G(1);
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

}  // namespace tp

using namespace tp;

extern "C" int tp_mlp_head_tf32(const float* feats, int64_t Q, int32_t batch, int32_t C, const float* w1,
                                const float* w2, const float* w3, int32_t num_classes, float* logits, void* stream) {
  if (C != kMlpC) return fail(TP_E_SHAPE, "tp_mlp_head_tf32: input_dim=%d (this build: %d, configs/triplane_occ.py)", C, kMlpC);
  if (num_classes <= 0 || num_classes > kMlpNOut) return fail(TP_E_SHAPE, "tp_mlp_head_tf32: num_classes=%d must be in 1..%d", num_classes, kMlpNOut);
  if (batch <= 0 || Q < 0) return fail(TP_E_SHAPE, "tp_mlp_head_tf32: bad B=%d Q=%lld", batch, (long long)Q);
  if (Q >= ((int64_t)1 << 27) || ((Q + 127) / 128) * batch >= ((int64_t)1 << 31))
    return fail(TP_E_SHAPE, "tp_mlp_head_tf32: Q=%lld x B=%d too large for one call (Q < 2^27, B*Q < 2^38)", (long long)Q, batch);
  if (Q == 0) return 0;
  if (!feats || !w1 || !w2 || !w3 || !logits) return fail(TP_E_NULL, "tp_mlp_head_tf32: null argument");
  if (((uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)w3) & 15) return fail(TP_E_SHAPE, "tp_mlp_head_tf32: weights must be 16-byte aligned");
  MlpParams P;
  P.feats = feats; P.w1 = w1; P.w2 = w2; P.w3 = w3; P.logits = logits;
  P.Q = (int)Q; P.ncls = num_classes;
  P.tiles_per_sample = (int)((Q + 127) / 128);
  P.tiles = P.tiles_per_sample * batch;
  TP_CUDA(opt_in_smem<mlp_head_kernel<true>>(kMlpSmem));
  TP_CUDA(opt_in_smem<mlp_head_kernel<false>>(kMlpSmem));
  const int64_t cap = (int64_t)kSMs * 4;  // 128 of the 512 TMEM columns per CTA: 4 per SM
  const int grid = (int)(P.tiles < cap ? P.tiles : cap);
  const bool vec = (Q % 4 == 0) && (((uintptr_t)feats & 15) == 0);
  if (vec) mlp_head_kernel<true><<<grid, kMlpThreads, kMlpSmem, (cudaStream_t)stream>>>(P);
  else mlp_head_kernel<false><<<grid, kMlpThreads, kMlpSmem, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("mlp_head_kernel");
  return 0;
}
