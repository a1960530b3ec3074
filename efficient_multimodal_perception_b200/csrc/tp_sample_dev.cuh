// Device-side building blocks shared by the two decode kernels (tp_sample.cu: arbitrary query
// streams; tp_sample_grid.cu: [B,h,w,d,3] query lattices). Keeping ONE copy of the coordinate
// chain, the bilinear weights and the tap accumulation is what makes both kernels bit-identical to
// each other and to torch-CUDA's grid_sample (+ (xy + yz) + xz).
#pragma once
#include "tp_common.cuh"

namespace tp {

struct SampleParams {
  const float* plane[3];  // channels-last [B,H,W,C]
  int64_t bstride[3];
  int H[3], W[3];
  float lo[3], vs[3], rcp_vs[3], half[3], rcp_half[3];
  const float* queries;  // [B,Q,3]
  float* out;            // [B,C,Q]
  int64_t Q;
  int64_t tiles_per_sample;
  int64_t tiles;
  int C;
};

// host side, tp_sample.cu: the per-query launch (pdl: it directly follows our conversion kernel, see tp_common.cuh),
// the workspace layout of the channels-last copies and the conversion launch
int sample3_flat(const tp_plane planes[3], int32_t C, const float* queries, int64_t Q, int32_t batch,
                 const tp_sample_geom* sg, int32_t arith, float* out, void* stream, bool pdl);
int planes3_workspace(const char* who, const tp_plane planes_nchw[3], int32_t C, int32_t batch, float* ws,
                      int64_t ws_floats, tp_plane nhwc[3], float* dsts[3]);
int planes3_to_workspace(const char* who, const tp_plane planes_nchw[3], int32_t C, int32_t batch, float* ws,
                         int64_t ws_floats, tp_plane nhwc[3], void* stream);

// ATen grid_sampler_unnormalize, align_corners=False.
//  CUDA (GridSampler.cuh): ((g + 1) * size - 1) / 2   -- nvcc contracts the mul+sub into one fma
//  CPU  (GridSamplerKernel.cpp): (g + 1) * (size / 2) - 0.5
template <int ARITH>
__device__ __forceinline__ float unnormalize(float g, float size) {
  float t = __fadd_rn(g, 1.0f);
  if (ARITH == TP_ARITH_TORCH_CUDA) return __fmul_rn(__fmaf_rn(t, size, -1.0f), 0.5f);
  if (ARITH == TP_ARITH_TORCH_CPU) return __fsub_rn(__fmul_rn(t, __fmul_rn(size, 0.5f)), 0.5f);
  return __fmul_rn(__fsub_rn(__fmul_rn(t, size), 1.0f), 0.5f);  // 2: CUDA formula, no contraction
}

// normalised grid coordinate of one axis: g = ((p - lo) (/) vs) (/) (S/2) - 1
// (triplane_occ.py:333-337, point_triplane.py:451-458)
template <int ARITH>
__device__ __forceinline__ float grid_coord(const SampleParams& P, float p, int a) {
  float v = tp_voxel_coord<ARITH == TP_ARITH_TORCH_CPU ? TP_ARITH_TORCH_CPU : TP_ARITH_TORCH_CUDA>(
      p, P.lo[a], P.vs[a], P.rcp_vs[a]);
  float n = (ARITH == TP_ARITH_TORCH_CPU) ? __fdiv_rn(v, P.half[a]) : __fmul_rn(v, P.rcp_half[a]);
  return __fsub_rn(n, 1.0f);
}

// One axis of one bilinear footprint: the two 1-D weights, the low pixel index and which of the two
// pixels are inside [0,size). A plane's four taps are the outer product of two of these.
struct AxisTap {
  float a1;  // x1 - ix   (weight of the low pixel)
  float a0;  // ix - x0   (weight of the high pixel)
  int i0;    // floor(ix) as ATen's static_cast<int>(::floor(ix)) (cvt.rzi saturates, NaN -> 0)
  int inb;   // bit0: low pixel in bounds, bit1: high pixel in bounds
};

template <int ARITH>
__device__ __forceinline__ AxisTap axis_tap(float g, int size) {
  AxisTap t;
  const float ix = unnormalize<ARITH>(g, (float)size);
  const float f0 = floorf(ix);
  const float f1 = __fadd_rn(f0, 1.0f);
  t.a1 = __fsub_rn(f1, ix);
  t.a0 = __fsub_rn(ix, f0);
  t.i0 = (int)f0;
  const int i1 = t.i0 + 1;
  t.inb = (int)((t.i0 >= 0) & (t.i0 < size)) | ((int)((i1 >= 0) & (i1 < size)) << 1);
  return t;
}

// per-query setup for one plane from its two axes (tx -> W, ty -> H): 4 weights, nw pixel index,
// 4-bit in-bounds mask. nw=(x1-ix)(y1-iy) ne=(ix-x0)(y1-iy) sw=(x1-ix)(iy-y0) se=(ix-x0)(iy-y0)
__device__ __forceinline__ void plane_from_axes(const AxisTap& tx, const AxisTap& ty, int W, float4& w,
                                                int& base, int& mask) {
  w.x = __fmul_rn(tx.a1, ty.a1);
  w.y = __fmul_rn(tx.a0, ty.a1);
  w.z = __fmul_rn(tx.a1, ty.a0);
  w.w = __fmul_rn(tx.a0, ty.a0);
  const int bx0 = tx.inb & 1, bx1 = tx.inb >> 1, by0 = ty.inb & 1, by1 = ty.inb >> 1;
  mask = (bx0 & by0) | ((bx1 & by0) << 1) | ((bx0 & by1) << 2) | ((bx1 & by1) << 3);
  // keep the base small when nothing is in bounds so base*C cannot overflow
  base = mask ? (ty.i0 * W + tx.i0) : 0;
}

template <int ARITH>
__device__ __forceinline__ void plane_setup(float gx, float gy, int W, int H, float4& w, int& base,
                                            int& mask) {
  const AxisTap tx = axis_tap<ARITH>(gx, W), ty = axis_tap<ARITH>(gy, H);
  plane_from_axes(tx, ty, W, w, base, mask);
}

// ATen accumulates out_acc += val * w for nw, ne, sw, se in that order (fma-contracted by nvcc),
// starting from 0 and skipping out-of-bounds taps.
__device__ __forceinline__ float4 fma4(float4 v, float w, float4 a) {
  return make_float4(__fmaf_rn(v.x, w, a.x), __fmaf_rn(v.y, w, a.y), __fmaf_rn(v.z, w, a.z),
                     __fmaf_rn(v.w, w, a.w));
}
__device__ __forceinline__ float4 mul4(float4 v, float w) {  // == fma(v, w, +0) bit for bit
  return make_float4(__fmaf_rn(v.x, w, 0.f), __fmaf_rn(v.y, w, 0.f), __fmaf_rn(v.z, w, 0.f),
                     __fmaf_rn(v.w, w, 0.f));
}

// One plane, one query, this lane's 4 channels. C4 = C/4 is a compile-time constant in the
// specialised kernels so the x-neighbour is an immediate offset and only the row stride needs a
// second 64-bit pointer.
template <bool MASKED>
__device__ __forceinline__ float4 plane_taps(const float4* __restrict__ pl, int off, int C4, int WC4,
                                             float4 w, int mk, unsigned long long pol) {
  const float4* t0 = pl + off;
  const float4* t1 = t0 + WC4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!MASKED) {
    const float4 v00 = ld_keep_f4(t0, pol), v01 = ld_keep_f4(t0 + C4, pol), v10 = ld_keep_f4(t1, pol),
                 v11 = ld_keep_f4(t1 + C4, pol);
    return fma4(v11, w.w, fma4(v10, w.z, fma4(v01, w.y, mul4(v00, w.x))));
  }
  float4 a = z;
  if (mk & 1) a = fma4(ld_keep_f4(t0, pol), w.x, a);
  if (mk & 2) a = fma4(ld_keep_f4(t0 + C4, pol), w.y, a);
  if (mk & 4) a = fma4(ld_keep_f4(t1, pol), w.z, a);
  if (mk & 8) a = fma4(ld_keep_f4(t1 + C4, pol), w.w, a);
  return a;
}

// ---- backward: 16-byte vector reductions into channels-last gradient planes --------------------
__device__ __forceinline__ void red_add_f4(float4* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 scale4(float4 g, float w) {
  return make_float4(__fmul_rn(g.x, w), __fmul_rn(g.y, w), __fmul_rn(g.z, w), __fmul_rn(g.w, w));
}
// grad_out * weight into the in-bounds taps of one plane, this lane's 4 channels
__device__ __forceinline__ void scatter_taps(float4* __restrict__ gp, int off, int C4, int WC4, float4 w, int mk,
                                             float4 g) {
  float4* t0 = gp + off;
  float4* t1 = t0 + WC4;
  if (mk & 1) red_add_f4(t0, scale4(g, w.x));
  if (mk & 2) red_add_f4(t0 + C4, scale4(g, w.y));
  if (mk & 4) red_add_f4(t1, scale4(g, w.z));
  if (mk & 8) red_add_f4(t1 + C4, scale4(g, w.w));
}

// ---- one warp, one tile of 32 queries ----------------------------------------------------------
constexpr int kParamStride = 20;  // words per query: 16 used; 20 keeps 8-lane STS.128 phases conflict-free
constexpr int kParamWords = 32 * kParamStride + 16;  // + 4-word skew per 8-query group (LDS.128 side)
constexpr int kTileWords = 32 * 32;  // [32 channels][32 queries], column XOR-swizzled by (channel >> 2) & 7

__device__ __forceinline__ int param_base(int qi) { return qi * kParamStride + (qi >> 3) * 4; }

// The tile's 32 queries are two runs of 16 consecutive query indices: lanes 0-15 <-> run0 + lane
// (valid while lane < n0), lanes 16-31 <-> run1 + lane - 16 (valid while lane - 16 < n1). The flat
// kernel passes run1 = run0 + 16; the lattice kernel's fallback passes two (i,j) columns.
// sp: kParamWords floats, st: kTileWords floats of this warp's shared memory.
// vec_ok: runs, Q and out are 16-byte aligned and n0, n1 are multiples of 4 (zero fast path).
template <int ARITH, int C4T>
__device__ __forceinline__ void sample_tile(const SampleParams& P, const int b, const int64_t run0,
                                            const int64_t run1, const int n0, const int n1, float* sp,
                                            float* st, const unsigned long long pol_planes,
                                            const unsigned long long pol_out, const bool vec_ok) {
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3, l8 = lane & 7;
  const int C4 = C4T ? C4T : (P.C >> 2);
  const int C = C4 * 4;
  const int nchunk = (C4 + 7) >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  const int l16 = lane & 15;
  const bool hi = lane >= 16;
  const int64_t q = (hi ? run1 : run0) + l16;
  const bool qvalid = l16 < (hi ? n1 : n0);

  // ---- per-query coordinate chain (one lane per query) ---------------------------------
  int anymask = 0;
  {
    float4 w[3];
    int base[3], mask[3];
    if (qvalid) {
      const float* qp = P.queries + ((int64_t)b * P.Q + q) * 3;
      const float g0 = grid_coord<ARITH>(P, __ldg(qp), 0);
      const float g1 = grid_coord<ARITH>(P, __ldg(qp + 1), 1);
      const float g2 = grid_coord<ARITH>(P, __ldg(qp + 2), 2);
      plane_setup<ARITH>(g0, g1, P.W[0], P.H[0], w[0], base[0], mask[0]);  // (x,y)
      plane_setup<ARITH>(g1, g2, P.W[1], P.H[1], w[1], base[1], mask[1]);  // (y,z)
      plane_setup<ARITH>(g0, g2, P.W[2], P.H[2], w[2], base[2], mask[2]);  // (x,z)
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        w[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        base[k] = 0;
        mask[k] = 0;
      }
    }
    anymask = mask[0] | (mask[1] << 4) | (mask[2] << 8);
    float4* dst = reinterpret_cast<float4*>(sp + param_base(lane));
    dst[0] = w[0];
    dst[1] = w[1];
    dst[2] = w[2];
    // tap offsets in float4 units (base * C/4 < 2^29, checked on the host)
    dst[3] = make_float4(__int_as_float(base[0] * C4), __int_as_float(base[1] * C4),
                         __int_as_float(base[2] * C4), __int_as_float(anymask));
  }
  const bool tile_empty = __all_sync(0xffffffffu, anymask == 0);
  const bool tile_full = __all_sync(0xffffffffu, anymask == 0xfff);
  __syncwarp();
  if (!tile_empty) pdl_wait();  // first read of the (possibly still being converted) planes

  const float4* pl0 = reinterpret_cast<const float4*>(P.plane[0] + (int64_t)b * P.bstride[0]) + l8;
  const float4* pl1 = reinterpret_cast<const float4*>(P.plane[1] + (int64_t)b * P.bstride[1]) + l8;
  const float4* pl2 = reinterpret_cast<const float4*>(P.plane[2] + (int64_t)b * P.bstride[2]) + l8;

  for (int ch = 0; ch < nchunk; ++ch) {
    const int cmax = min(32, C - ch * 32);
    float* obase = P.out + ((int64_t)b * C + ch * 32) * P.Q;
    if (tile_empty) {
      // every query of the tile misses all three planes: zeros, no gathers, no staging
      if (vec_ok) {
        const int r0 = lane >> 3, pos = (lane & 7) * 4;
        const int off = pos & 15;
        if (off < (pos >= 16 ? n1 : n0)) {
          float* o = obase + (pos >= 16 ? run1 : run0) + off;
          for (int c = r0; c < cmax; c += 4)
            st_stream_f4(reinterpret_cast<float4*>(o + (int64_t)c * P.Q), make_float4(0.f, 0.f, 0.f, 0.f), pol_out);
        }
      } else if (qvalid) {
        for (int c = 0; c < cmax; ++c) st_stream_f1(obase + (int64_t)c * P.Q + q, 0.f, pol_out);
      }
      continue;
    }
    const int choff = ch * 8;
    const bool chunk_full = cmax == 32;  // warp-uniform; partial chunks only when C % 32 != 0
    float* tcol = st + (l8 * 4) * 32;
    if (tile_full && chunk_full) {
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        const int qi = sub * 8 + pass;  // 8 consecutive queries per 8-lane group
        const float4* prm = reinterpret_cast<const float4*>(sp + param_base(qi));
        const float4 w0 = prm[0], w1 = prm[1], w2 = prm[2], bm = prm[3];
        const float4 a0 = plane_taps<false>(pl0 + choff, __float_as_int(bm.x), C4, WC4_0, w0, 15, pol_planes);
        const float4 a1 = plane_taps<false>(pl1 + choff, __float_as_int(bm.y), C4, WC4_1, w1, 15, pol_planes);
        const float4 a2 = plane_taps<false>(pl2 + choff, __float_as_int(bm.z), C4, WC4_2, w2, 15, pol_planes);
        float* t = tcol + (qi ^ l8);
        t[0] = __fadd_rn(__fadd_rn(a0.x, a1.x), a2.x);  // (xy + yz) + xz  (triplane_occ.py:345)
        t[32] = __fadd_rn(__fadd_rn(a0.y, a1.y), a2.y);
        t[64] = __fadd_rn(__fadd_rn(a0.z, a1.z), a2.z);
        t[96] = __fadd_rn(__fadd_rn(a0.w, a1.w), a2.w);
      }
    } else {
      const bool cvalid = ch * 32 + l8 * 4 < C;  // C % 4 == 0
#pragma unroll 2
      for (int pass = 0; pass < 8; ++pass) {
        const int qi = sub * 8 + pass;
        const float4* prm = reinterpret_cast<const float4*>(sp + param_base(qi));
        const float4 w0 = prm[0], w1 = prm[1], w2 = prm[2], bm = prm[3];
        const int m = cvalid ? __float_as_int(bm.w) : 0;
        const float4 a0 = plane_taps<true>(pl0 + choff, __float_as_int(bm.x), C4, WC4_0, w0, m & 15, pol_planes);
        const float4 a1 = plane_taps<true>(pl1 + choff, __float_as_int(bm.y), C4, WC4_1, w1, (m >> 4) & 15, pol_planes);
        const float4 a2 = plane_taps<true>(pl2 + choff, __float_as_int(bm.z), C4, WC4_2, w2, (m >> 8) & 15, pol_planes);
        float* t = tcol + (qi ^ l8);
        t[0] = __fadd_rn(__fadd_rn(a0.x, a1.x), a2.x);
        t[32] = __fadd_rn(__fadd_rn(a0.y, a1.y), a2.y);
        t[64] = __fadd_rn(__fadd_rn(a0.z, a1.z), a2.z);
        t[96] = __fadd_rn(__fadd_rn(a0.w, a1.w), a2.w);
      }
    }
    __syncwarp();
    // ---- coalesced write of the [32 channels][32 queries] tile -------------------------
    if (qvalid) {
      float* o = obase + q;
      const int64_t Qs = P.Q;
#pragma unroll 8
      for (int c = 0; c < cmax; ++c, o += Qs) st_stream_f1(o, st[c * 32 + (lane ^ ((c >> 2) & 7))], pol_out);
    }
    __syncwarp();
  }
}

}  // namespace tp
