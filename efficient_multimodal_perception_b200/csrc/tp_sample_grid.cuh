// Building blocks of the lattice decode (tp_sample_grid.cu) shared with the decode + head kernel (tp_sample_head.cu):
// block geometry, table layout in shared memory, the per-plane table builder, block index -> lattice coordinates.
#pragma once
#include "tp_sample_dev.cuh"

namespace tp {

struct GridParams {
  SampleParams S;  // S.Q = h*w*d
  int h, w, d;
  int nib, njb, nkb;  // blocks along h, w, d
  int nblocks;        // batch * nib * njb * nkb
  int vec_ok;
  // generated lattice (tp_sample3_lattice_nhwc_f32): q_a(n) = (n + 0.5) * gen_step[a] + gen_org[a], S.queries == nullptr
  float gen_org[3], gen_step[3];
  int pdl;  // launched as the programmatic dependent of the layout conversion: wait before the first plane read
};

// roi()'s voxel centres, op by op (triplane_occ.py:311-316: three separate fp32 tensor ops, never contracted)
__device__ __forceinline__ float lattice_coord(const GridParams& G, int a, int n) {
  return __fadd_rn(__fmul_rn(__fadd_rn((float)n, 0.5f), G.gen_step[a]), G.gen_org[a]);
}

#ifdef TP_GRID_TRACE  // tools/micro/grid_trace.cu: per-CTA timeline (not part of the library build)
__device__ unsigned long long g_grid_cta[8 * 1024];
__device__ __forceinline__ unsigned long long grid_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define GRID_G(n) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_grid_cta[blockIdx.x * 8 + (n)] = grid_gtimer(); } while (0)
#else
#define GRID_G(n) do {} while (0)
#endif

constexpr int kGridThreads = 256;
// resident CTAs per SM and table entries in flight per thread: 3 CTAs (80 registers) with 3 entries (12 independent
// 16-byte loads) in flight. The table phase is a chain of dependent L2 round trips, (entries / 32) / in-flight of
// them per block; 4 CTAs (64 registers) only leave room for 2 in flight. Measured after the plane-skip change, same
// box: 640k lattice 21.4 vs 22.0 us, elev 26.4 vs 29.2, 8 x roi 42.9 vs 46.6, roi 9.9 vs 9.9 (3 CTAs with 2 in
// flight: 22.3 / 27.5 / 45.3 / 10.2).
#ifndef TP_GRID_CTAS_PER_SM
#define TP_GRID_CTAS_PER_SM 3
#endif
constexpr int kGridCtasPerSm = TP_GRID_CTAS_PER_SM;
constexpr int kBK = 16;  // lattice block extent along d
constexpr int kBJ = 8;   // ... along w: a warp-wide 16-byte store covers 8 (j) x 4 (k/4) = 512 contiguous bytes when d == 16
constexpr int kFallbackWarps = 4;

template <int BI>
struct GridCfg {
  static constexpr int E0 = BI * kBJ, E1 = kBJ * kBK, E2 = BI * kBK;  // table entries (xy, yz, xz)
  static constexpr int E = E0 + E1 + E2;
  // Table rows are channels. xy rows are read with 4-byte loads: stride == 1 (mod 32) makes the
  // channel-major writes conflict-free. yz / xz rows are read as 16-byte k-runs: stride == 4 (mod 32)
  // plus the column swizzle below does the same while keeping every run aligned.
  static constexpr int S0 = E0 + 1, S1 = E1 + 4, S2 = E2 + 4;
  static constexpr int kTableWords = 32 * (S0 + S1 + S2);
  static constexpr int kFallbackWords = kFallbackWarps * (kParamWords + kTileWords);
  static constexpr int kWords = kTableWords > kFallbackWords ? kTableWords : kFallbackWords;
  static constexpr int kSmemBytes = kWords * 4 + E * (16 + 8);
  static_assert(kWords % 4 == 0, "entry records are 16-byte aligned");
  static_assert(E0 % 32 == 0 && E1 % 32 == 0 && E2 % 32 == 0, "one plane per warp-round of 32 entries");
  static_assert(BI * kBJ * kBK % kGridThreads == 0, "whole queries per thread");
};

// column swizzle of the yz / xz tables: XOR bits 2-3 of the column with bits 3-4 of the channel.
// Keeps every aligned 4-column group (one float4 of 4 consecutive k) intact.
__device__ __forceinline__ int swz_bits(int c) { return ((c >> 3) & 3) << 2; }

// loads without side effects: not volatile, so the scheduler may batch the taps of several table
// entries ahead of the fmas that consume them (planes are read-only for the whole launch)
__device__ __forceinline__ float4 ld_plane_f4(const float4* p, unsigned long long pol) {
  float4 r;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "l"(p), "l"(pol));
  return r;
}
// predicated streaming store; no "memory" clobber: nothing in this kernel reads the output, and the
// clobber would pin every later shared-memory load behind the store
__device__ __forceinline__ void st_out_f4(float* p, float4 v, unsigned long long pol, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "@p st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;\n\t}" ::"l"(p),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol), "r"((int)pred));
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct Taps {
  float4 v00, v01, v10, v11;
};
// the four taps of one table entry, this lane's 4 channels; out-of-bounds taps are not touched
__device__ __forceinline__ Taps load_taps(const float4* __restrict__ pl, int off, int C4, int WC4, int mk,
                                          unsigned long long pol) {
  const float4* t0 = pl + off;
  const float4* t1 = t0 + WC4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  Taps t;
  t.v00 = (mk & 1) ? ld_plane_f4(t0, pol) : z;
  t.v01 = (mk & 2) ? ld_plane_f4(t0 + C4, pol) : z;
  t.v10 = (mk & 4) ? ld_plane_f4(t1, pol) : z;
  t.v11 = (mk & 8) ? ld_plane_f4(t1 + C4, pol) : z;
  return t;
}
// same accumulation as plane_taps<true>: nw, ne, sw, se in order, out-of-bounds taps skipped
__device__ __forceinline__ float4 accum_taps(const Taps& t, float4 w, int mk) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mk & 1) a = fma4(t.v00, w.x, a);
  if (mk & 2) a = fma4(t.v01, w.y, a);
  if (mk & 4) a = fma4(t.v10, w.z, a);
  if (mk & 8) a = fma4(t.v11, w.w, a);
  return a;
}

// One 2-D table (NR rounds of 32 entries) for this thread's 4 channels: two entries in flight, i.e. 8 independent
// 16-byte loads before the first fma. `live` and `zero` are block-uniform: a plane with no in-bounds tap in the whole
// block costs nothing (its table keeps, or is set to, the +0 the masked accumulation would have produced).
#ifndef TP_TAB_IN_FLIGHT
#define TP_TAB_IN_FLIGHT 3
#endif
constexpr int kTabInFlight = TP_TAB_IN_FLIGHT;  // table entries (x 4 taps x 16 B) a thread has in flight

template <int NR>
__device__ __forceinline__ void build_table(bool live, bool zero, const float4* __restrict__ pl, int C4, int WC4,
                                            const float4* s_w, const int2* s_om, float* t, int stride, bool cvalid,
                                            unsigned long long pol) {
  if (!live) {
    if (!zero) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        t[r * 32] = 0.f;
        t[r * 32 + stride] = 0.f;
        t[r * 32 + 2 * stride] = 0.f;
        t[r * 32 + 3 * stride] = 0.f;
      }
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < NR; r += kTabInFlight) {
    float4 wgt[kTabInFlight];
    int mk[kTabInFlight];
    Taps tp[kTabInFlight];
#pragma unroll
    for (int u = 0; u < kTabInFlight; ++u) {
      if (r + u < NR) {
        wgt[u] = s_w[(r + u) * 32];
        const int2 om = s_om[(r + u) * 32];
        mk[u] = cvalid ? om.y : 0;
        tp[u] = load_taps(pl, om.x, C4, WC4, mk[u], pol);
      }
    }
#pragma unroll
    for (int u = 0; u < kTabInFlight; ++u) {
      if (r + u < NR) {
        const float4 a = accum_taps(tp[u], wgt[u], mk[u]);
        float* tt = t + (r + u) * 32;
        tt[0] = a.x;
        tt[stride] = a.y;
        tt[2 * stride] = a.z;
        tt[3 * stride] = a.w;
      }
    }
  }
}

// lattice block -> coordinates. nblk = nib * njb * nkb blocks per sample, k fastest.
struct BlockPos {
  int b, i0, j0, k0;
};
template <int BI>
__device__ __forceinline__ BlockPos block_pos(const GridParams& G, int blk) {
  BlockPos p;
  const int kb = blk % G.nkb; blk /= G.nkb;
  const int jb = blk % G.njb; blk /= G.njb;
  p.i0 = (blk % G.nib) * BI;
  p.b = blk / G.nib;
  p.j0 = jb * kBJ;
  p.k0 = kb * kBK;
  return p;
}

// ---- phases A and B of a lattice block, shared by the forward, backward and decode + head kernels ---------------

// A: every thread reads its queries of the block (k and j fixed per thread, i strided) and compares them bit for bit
// with the block's representatives: x with the row's first query, y with the column's, z with the depth's.
template <int BI>
__device__ __forceinline__ bool grid_block_is_lattice(const GridParams& G, const float* q00, int ni, int nj, int nk,
                                                      int wd, int tid) {
  constexpr int QPT = BI * kBJ * kBK / kGridThreads;
  const int ak = tid & (kBK - 1), aj = (tid / kBK) % kBJ, ia = tid / (kBK * kBJ);
  bool ok = true;
  if (aj < nj && ak < nk) {
    const unsigned yr = __float_as_uint(__ldg(q00 + aj * G.d * 3 + 1));
    const unsigned zr = __float_as_uint(__ldg(q00 + ak * 3 + 2));
#pragma unroll
    for (int t = 0; t < QPT; ++t) {
      const int ii = ia + t * (kGridThreads / (kBK * kBJ));
      if (ii < ni) {
        const float* qi = q00 + ii * wd * 3;
        const float* qp = qi + (aj * G.d + ak) * 3;
        const unsigned x = __float_as_uint(__ldg(qp)), y = __float_as_uint(__ldg(qp + 1)),
                       z = __float_as_uint(__ldg(qp + 2));
        ok &= (x == __float_as_uint(__ldg(qi))) & (y == yr) & (z == zr);
      }
    }
  }
  return ok;
}

// B: one record per table entry (index pair) from the block's representative queries: 4 bilinear weights, nw tap
// offset in float4 units, in-bounds mask. Returns this thread's "plane p has an in-bounds tap" bits.
// (Measured: computing the 2 (BI + 8 + 16) distinct 1-D taps once and combining pairs saves ~half of this phase's
// instructions but needs a barrier in between — same time on the 640k lattice, 2 % slower on 8 x roi.)
// GEN: the representative coordinates are generated from the lattice indices (org = block origin {i0, j0, k0})
// instead of being read from the query tensor.
template <int ARITH, int BI, bool GEN = false>
__device__ __forceinline__ int grid_build_records(const GridParams& G, const float* q00, int ni, int nj, int nk, int wd,
                                                  int C4, float4* s_w, int2* s_om, int tid, int gi0 = 0, int gj0 = 0,
                                                  int gk0 = 0) {
  using Cfg = GridCfg<BI>;
  const SampleParams& P = G.S;
  int live = 0;
  for (int e = tid; e < Cfg::E; e += kGridThreads) {
    int pl, a0, a1, e0i, e1i, n0, n1, s0, s1;  // s: query stride of the two lattice indices
    if (e < Cfg::E0) {                      // xy(i,j): x -> W, y -> H of plane 0
      pl = 0; a0 = 0; a1 = 1; e0i = e / kBJ; e1i = e % kBJ; n0 = ni; n1 = nj; s0 = wd; s1 = G.d;
    } else if (e < Cfg::E0 + Cfg::E1) {     // yz(j,k): y -> W, z -> H of plane 1
      const int r = e - Cfg::E0;
      pl = 1; a0 = 1; a1 = 2; e0i = r / kBK; e1i = r % kBK; n0 = nj; n1 = nk; s0 = G.d; s1 = 1;
    } else {                                // xz(i,k): x -> W, z -> H of plane 2
      const int r = e - Cfg::E0 - Cfg::E1;
      pl = 2; a0 = 0; a1 = 2; e0i = r / kBK; e1i = r % kBK; n0 = ni; n1 = nk; s0 = wd; s1 = 1;
    }
    float4 wgt = make_float4(0.f, 0.f, 0.f, 0.f);
    int base = 0, mask = 0;
    if (e0i < n0 && e1i < n1) {
      float c0, c1;
      if (GEN) {  // axis 0 <-> i, 1 <-> j, 2 <-> k
        c0 = lattice_coord(G, a0, (a0 == 0 ? gi0 : gj0) + e0i);
        c1 = lattice_coord(G, a1, (a1 == 1 ? gj0 : gk0) + e1i);
      } else {
        c0 = __ldg(q00 + e0i * s0 * 3 + a0);
        c1 = __ldg(q00 + e1i * s1 * 3 + a1);
      }
      const float g0 = grid_coord<ARITH>(P, c0, a0);
      const float g1 = grid_coord<ARITH>(P, c1, a1);
      plane_setup<ARITH>(g0, g1, P.W[pl], P.H[pl], wgt, base, mask);
    }
    s_w[e] = wgt;
    s_om[e] = make_int2(base * C4, mask);
    if (mask) live |= 1 << pl;
  }
  return live;
}

// block-wide OR of (live planes | 8 if some query broke the lattice): warp redux + one shared atomic. s_vote[2] is a
// double buffer: the other slot is cleared here for the next block; the caller provides the barrier and reads
// s_vote[n & 1] after it.
__device__ __forceinline__ void grid_cast_vote(int* s_vote, int n, int live, bool ok, int tid) {
  const int bits = __reduce_or_sync(0xffffffffu, live | (ok ? 0 : 8));
  if ((tid & 31) == 0 && bits) atomicOr(&s_vote[n & 1], bits);
  if (tid == 0) s_vote[(n + 1) & 1] = 0;
}

}  // namespace tp
