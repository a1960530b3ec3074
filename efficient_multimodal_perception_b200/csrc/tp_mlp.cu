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
#include "tp_common.cuh"

namespace tp {

constexpr int kMlpThreads = 128;   // 4 warps = the 4 TMEM lane quadrants = 128 rows of a tile
constexpr int kMlpC = 32;          // input channels (configs/triplane_occ.py: Mlp(input_dim=32))
constexpr int kMlpH = 64;          // hidden = 2C
constexpr int kMlpNOut = 16;       // num_classes padded to the next multiple of 16 (UMMA N for M = 128)
constexpr int kTmemCols = 128;     // D1 [0,64) | D2 [64,96) | D3 [96,112)

// shared-memory map (bytes, every operand tile 1024-byte aligned for the 128-byte swizzle). Only the layer-1
// activations and the weights live here: the hidden activations stay in TMEM (below). 35 KB per CTA; the CTA count
// per SM is set by TMEM, 4 x 128 columns.
constexpr int kOffA1 = 0;                       // [128 x 32]  layer-1 A, 16 KB
constexpr int kOffW1 = 16384;                   // [64 x 32]   8 KB
constexpr int kOffW2 = kOffW1 + 8192;           // [32 x 64]   2 K-blocks of 4 KB
constexpr int kOffW3 = kOffW2 + 8192;           // [16 x 32]   2 KB
constexpr int kMlpSmem = kOffW3 + 2048 + 1024;  // + slack to align the base to 1024 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of (row r, 16-byte chunk c16 of the 128-byte row) inside a K-major SWIZZLE_128B tile:
// 8-row groups of 1024 B, the chunk index XORed with the row inside the group (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t swz128(int r, int c16) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // layout type SWIZZLE_128B
  return d;
}
// MN-major operand (rows of the tile contiguous in memory, cute's make_umma_desc<Major::MN>). For 32-bit types the only
// MN-major layout is SWIZZLE_128B_BASE32B: 512-byte atoms of 4 K-rows x 128 B (32 MN elements), the 32-byte chunk index
// XORed with the K-row (Swizzle<2,5,2>); `lbo` bytes between atoms along MN, `sbo` bytes between atoms along K
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of (K-row c, 16-byte chunk u of the 128 MN elements = 32 chunks) in a [32 x 128] tile of such atoms laid out
// with lbo = 512 (4 atoms side by side along MN) and sbo = 2048
__device__ __forceinline__ uint32_t mn_tile_off(int c, int u) {
  return (uint32_t)((c >> 2) * 2048 + (u >> 3) * 512 + (c & 3) * 128 + ((((u & 7) >> 1) ^ (c & 3)) << 5) + (u & 1) * 16);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, A operand read from TMEM (lane = row, one 32-bit column per K element): an accumulator tile that went
// through ReLU in place is the next layer's A without a trip through shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (cute::elect_one_sync): the tcgen05.mma / commit issuer
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem_mlp() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 16 consecutive columns of TMEM -> 16 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 columns at once, load and wait in one statement so that no use of the registers can be scheduled before the wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" ::"r"(r[0]),
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
      "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Round to nearest TF32, ties away from zero — what cvt.rna.tf32.f32 returns for every finite input — is "add half a
// TF32 ulp to the magnitude, drop the 13 low mantissa bits". The tensor core ignores those 13 bits of a kind::tf32
// operand (checked bit for bit on B200 with tools/micro/mlp_bits.cu: masked and unmasked operands give identical
// logits), so the rounding is ONE integer add instead of the ~8 SASS instructions cvt.rna expands to.
// Inf stays Inf, NaN stays NaN, the largest finite values round to Inf.
__device__ __forceinline__ uint32_t rna_tf32(float x) { return __float_as_uint(x) + 0x1000u; }
// ReLU and the rounding in one integer max: negative floats are negative integers. (NaN -> 0 like fmaxf(NaN, 0).)
__device__ __forceinline__ uint32_t relu_rna_tf32(uint32_t bits) { return (uint32_t)max((int)(bits + 0x1000u), 0); }
// ReLU + round to nearest TF32 of 32 accumulator columns, in place in TMEM
__device__ __forceinline__ void relu_tf32_inplace(uint32_t taddr) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = relu_rna_tf32(r[i]);
  tmem_st32(taddr, r);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

#ifdef TP_MLP_TRACE  // tools/micro/mlp_trace.cu: phase timestamps of CTA 0 (not part of the library build)
__device__ long long g_mlp_trace[16 * 12];
#define MLP_T(n) do { if (tid == 0 && blockIdx.x == 0 && it < 12) g_mlp_trace[it * 16 + (n)] = clock64(); } while (0)
__device__ unsigned long long g_mlp_cta[2 * 1024];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MLP_G(n) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_mlp_cta[blockIdx.x * 2 + (n)] = gtimer(); } while (0)
#define MLP_P(n) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_mlp_trace[11 * 16 + (n)] = clock64(); } while (0)
#else
__device__ unsigned long long g_mlp_cta[2 * 1024];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MLP_G(n) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_mlp_cta[blockIdx.x * 2 + (n)] = gtimer(); } while (0)
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
  static bool opted_in[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !opted_in[dev]) {
    TP_CUDA(cudaFuncSetAttribute(mlp_head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmem));
    TP_CUDA(cudaFuncSetAttribute(mlp_head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmem));
    if (dev >= 0 && dev < 64) opted_in[dev] = true;
  }
  const int64_t cap = (int64_t)kSMs * 4;  // 128 of the 512 TMEM columns per CTA: 4 per SM
  const int grid = (int)(P.tiles < cap ? P.tiles : cap);
  const bool vec = (Q % 4 == 0) && (((uintptr_t)feats & 15) == 0);
  if (vec) mlp_head_kernel<true><<<grid, kMlpThreads, kMlpSmem, (cudaStream_t)stream>>>(P);
  else mlp_head_kernel<false><<<grid, kMlpThreads, kMlpSmem, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("mlp_head_kernel");
  return 0;
}
