// Triplane decode: fused normalise + 3-plane bilinear gather + sum.
// Replaces the five sample_points_triplane variants of the reference
// (triplane.py:490-514, triplane_occ.py:321-348, triplane_elev.py:286-313,
//  point_triplane.py:439-466, point_triplane_occ.py:407-440): three F.grid_sample launches
// plus two adds and six temporaries become one gather kernel that reads each query once and
// writes each output element once.
//
// Layout: planes are read channels-last [B,H,W,C] so one tap of one query is C contiguous floats
// (8 lanes x 16 B per 32 channels => every warp-wide LDG.128 touches 4 fully used 128-B lines);
// the result is transposed through shared memory so the reference's channel-major [B,C,Q] output
// is written as 128-B coalesced rows.
#include <atomic>

#include "tp_common.cuh"

namespace tp {

struct SampleParams {
  const float* plane[3];
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

constexpr int kWarpsPerCta = 4;
constexpr int kParamStride = 20;  // words per query: 16 used; 20 keeps 8-lane STS.128 phases conflict-free
constexpr int kParamWords = 32 * kParamStride + 16;  // + 4-word skew per 8-query group (LDS.128 side)
constexpr int kTileWords = 32 * 32;  // [32 channels][32 queries], column XOR-swizzled by (channel >> 2) & 7

__device__ __forceinline__ int param_base(int qi) { return qi * kParamStride + (qi >> 3) * 4; }

// Dynamic tile scheduler state: a ring of (next tile, finished warps) pairs; the host picks a slot
// per launch, the last warp of a launch resets it. In-range and out-of-range tiles differ ~5x in
// cost, so a static split leaves SMs idle.
constexpr int kSchedSlots = 1024;
__device__ unsigned int g_sched[kSchedSlots][2];  // [next tile, finished warps]

// per-query setup for one plane: 4 weights, nw pixel index, 4-bit in-bounds mask
template <int ARITH>
__device__ __forceinline__ void plane_setup(float gx, float gy, int W, int H, float4& w, int& base,
                                            int& mask) {
  float ix = unnormalize<ARITH>(gx, (float)W);
  float iy = unnormalize<ARITH>(gy, (float)H);
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float fx1 = __fadd_rn(fx0, 1.0f), fy1 = __fadd_rn(fy0, 1.0f);
  // nw=(x1-ix)(y1-iy) ne=(ix-x0)(y1-iy) sw=(x1-ix)(iy-y0) se=(ix-x0)(iy-y0)
  float ax1 = __fsub_rn(fx1, ix), ax0 = __fsub_rn(ix, fx0);
  float ay1 = __fsub_rn(fy1, iy), ay0 = __fsub_rn(iy, fy0);
  w.x = __fmul_rn(ax1, ay1);
  w.y = __fmul_rn(ax0, ay1);
  w.z = __fmul_rn(ax1, ay0);
  w.w = __fmul_rn(ax0, ay0);
  // float->int like ATen's static_cast<int>(::floor(ix)) (cvt.rzi saturates, NaN -> 0)
  int x0 = (int)fx0, y0 = (int)fy0;
  int x1 = x0 + 1, y1 = y0 + 1;
  bool bx0 = (x0 >= 0) & (x0 < W), bx1 = (x1 >= 0) & (x1 < W);
  bool by0 = (y0 >= 0) & (y0 < H), by1 = (y1 >= 0) & (y1 < H);
  mask = (int)(bx0 & by0) | ((int)(bx1 & by0) << 1) | ((int)(bx0 & by1) << 2) |
         ((int)(bx1 & by1) << 3);
  // keep the base small when nothing is in bounds so base*C cannot overflow
  base = mask ? (y0 * W + x0) : 0;
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

// C4T: C/4 known at compile time (8, 24, 32 for the reference's C = 32, 96, 128) or 0 = runtime.
template <int ARITH, int C4T>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8)
sample3_kernel(const SampleParams P, const int sched_slot) {
  __shared__ __align__(16) float s_param[kWarpsPerCta][kParamWords];
  __shared__ __align__(16) float s_tile[kWarpsPerCta][kTileWords];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;
  float* sp = s_param[warp];
  float* st = s_tile[warp];
  const int C4 = C4T ? C4T : (P.C >> 2);
  const int C = C4 * 4;
  const int nchunk = (C4 + 7) >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  unsigned int* sched = g_sched[sched_slot];
  // planes (MBs) are re-read by every query while the output (10x larger) streams through L2 once
  const unsigned long long pol_planes = policy_evict_last(), pol_out = policy_evict_first();
  const bool q_vec4 = ((P.Q & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);

  for (;;) {
    // warp-granular dynamic scheduling: no CTA barrier, in-range and out-of-range tiles balance out
    unsigned int t32 = 0;
    if (lane == 0) t32 = atomicAdd(&sched[0], 1u);
    const int64_t tile = __shfl_sync(0xffffffffu, t32, 0);
    if (tile >= P.tiles) break;
    const int b = (int)(tile / P.tiles_per_sample);
    const int64_t q0 = (tile - (int64_t)b * P.tiles_per_sample) * 32;
    const int64_t q = q0 + lane;
    const bool qvalid = q < P.Q;

    // ---- per-query coordinate chain (one lane per query) ---------------------------------
    int anymask = 0;
    {
      float4 w[3];
      int base[3], mask[3];
      if (qvalid) {
        const float* qp = P.queries + ((int64_t)b * P.Q + q) * 3;
        float p[3] = {__ldg(qp), __ldg(qp + 1), __ldg(qp + 2)};
        float g[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float v = tp_voxel_coord<ARITH == TP_ARITH_TORCH_CPU ? TP_ARITH_TORCH_CPU
                                                               : TP_ARITH_TORCH_CUDA>(
              p[a], P.lo[a], P.vs[a], P.rcp_vs[a]);
          float n = (ARITH == TP_ARITH_TORCH_CPU) ? __fdiv_rn(v, P.half[a])
                                                  : __fmul_rn(v, P.rcp_half[a]);
          g[a] = __fsub_rn(n, 1.0f);
        }
        plane_setup<ARITH>(g[0], g[1], P.W[0], P.H[0], w[0], base[0], mask[0]);  // (x,y)
        plane_setup<ARITH>(g[1], g[2], P.W[1], P.H[1], w[1], base[1], mask[1]);  // (y,z)
        plane_setup<ARITH>(g[0], g[2], P.W[2], P.H[2], w[2], base[2], mask[2]);  // (x,z)
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

    const float4* pl0 = reinterpret_cast<const float4*>(P.plane[0] + (int64_t)b * P.bstride[0]) + l8;
    const float4* pl1 = reinterpret_cast<const float4*>(P.plane[1] + (int64_t)b * P.bstride[1]) + l8;
    const float4* pl2 = reinterpret_cast<const float4*>(P.plane[2] + (int64_t)b * P.bstride[2]) + l8;

    for (int ch = 0; ch < nchunk; ++ch) {
      const int cmax = min(32, C - ch * 32);
      float* orow = P.out + ((int64_t)b * C + ch * 32) * P.Q + q0;
      if (tile_empty) {
        // every query of the tile misses all three planes: zeros, no gathers, no staging
        if (q_vec4) {
          const int r0 = lane >> 3, c4q = (lane & 7) * 4;
          if (q0 + c4q < P.Q) {
            for (int c = r0; c < cmax; c += 4)
              st_stream_f4(reinterpret_cast<float4*>(orow + (int64_t)c * P.Q + c4q), make_float4(0.f, 0.f, 0.f, 0.f), pol_out);
          }
        } else if (qvalid) {
          for (int c = 0; c < cmax; ++c) st_stream_f1(orow + (int64_t)c * P.Q + lane, 0.f, pol_out);
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
        float* o = orow + lane;
        const int64_t Qs = P.Q;
#pragma unroll 8
        for (int c = 0; c < cmax; ++c, o += Qs) st_stream_f1(o, st[c * 32 + (lane ^ ((c >> 2) & 7))], pol_out);
      }
      __syncwarp();
    }
  }
  // last warp out resets the scheduler slot for the next launch that draws it
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1u) == gridDim.x * kWarpsPerCta - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

// NCHW -> NHWC through a padded 32x32 shared tile. grid = (ceil(HW/32), ceil(C/32), B)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, int64_t src_bstride, float* __restrict__ dst,
                    int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* s = src + (int64_t)b * src_bstride;
  float* d = dst + (int64_t)b * C * HW;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    int c = c0 + ty + j, p = p0 + tx;
    if (c < C && p < HW) tile[ty + j][tx] = __ldg(s + (int64_t)c * HW + p);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    int p = p0 + ty + j, c = c0 + tx;
    if (c < C && p < HW) d[(int64_t)p * C + c] = tile[tx][ty + j];
  }
}

struct Planes3 {
  const float* src[3];
  float* dst[3];
  int64_t src_bstride[3];
  int HW[3];
  int tiles_x[3];  // ceil(HW/32) per plane; blockIdx.x runs over the three planes back to back
};

// all three planes in one launch. grid = (sum_p ceil(HW_p/32), ceil(C/32), B)
__global__ void __launch_bounds__(256)
nchw_to_nhwc3_kernel(const Planes3 P, int C) {
  __shared__ float tile[32][33];
  int bx = blockIdx.x, k = 0;
  if (bx >= P.tiles_x[0]) { bx -= P.tiles_x[0]; k = 1; if (bx >= P.tiles_x[1]) { bx -= P.tiles_x[1]; k = 2; } }
  const int HW = P.HW[k];
  const int b = blockIdx.z;
  const int p0 = bx * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = P.src[k] + (int64_t)b * P.src_bstride[k];
  float* d = P.dst[k] + (int64_t)b * C * HW;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    int c = c0 + ty + j, p = p0 + tx;
    if (c < C && p < HW) tile[ty + j][tx] = __ldg(s + (int64_t)c * HW + p);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    int p = p0 + ty + j, c = c0 + tx;
    if (c < C && p < HW) d[(int64_t)p * C + c] = tile[tx][ty + j];
  }
}

}  // namespace tp

using namespace tp;

extern "C" int tp_planes3_nchw_to_nhwc_f32(const tp_plane planes_nchw[3], float* const dst[3], int32_t batch,
                                           int32_t C, void* stream) {
  if (!planes_nchw || !dst) return fail(TP_E_NULL, "tp_planes3_nchw_to_nhwc_f32: null argument");
  if (batch <= 0 || C <= 0) return fail(TP_E_SHAPE, "tp_planes3_nchw_to_nhwc_f32: bad shape B=%d C=%d", batch, C);
  Planes3 P;
  int tx = 0;
  for (int k = 0; k < 3; ++k) {
    if (!planes_nchw[k].data || !dst[k]) return fail(TP_E_NULL, "tp_planes3_nchw_to_nhwc_f32: plane %d is null", k);
    if (planes_nchw[k].H <= 0 || planes_nchw[k].W <= 0) return fail(TP_E_SHAPE, "tp_planes3_nchw_to_nhwc_f32: plane %d shape", k);
    P.src[k] = planes_nchw[k].data;
    P.dst[k] = dst[k];
    P.src_bstride[k] = planes_nchw[k].batch_stride;
    P.HW[k] = planes_nchw[k].H * planes_nchw[k].W;
    P.tiles_x[k] = (P.HW[k] + 31) / 32;
    tx += P.tiles_x[k];
  }
  dim3 grid(tx, (C + 31) / 32, batch);
  nchw_to_nhwc3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, C);
  TP_LAUNCH_CHECK("nchw_to_nhwc3_kernel");
  return 0;
}

extern "C" int tp_planes_nchw_to_nhwc_f32(const float* src, int64_t src_batch_stride, float* dst,
                                          int32_t batch, int32_t C, int32_t H, int32_t W,
                                          void* stream) {
  if (!src || !dst) return fail(TP_E_NULL, "tp_planes_nchw_to_nhwc_f32: null plane pointer");
  if (batch <= 0 || C <= 0 || H <= 0 || W <= 0)
    return fail(TP_E_SHAPE, "tp_planes_nchw_to_nhwc_f32: bad shape B=%d C=%d H=%d W=%d", batch, C, H, W);
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, batch);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_batch_stride, dst, C, HW);
  TP_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int tp_sample3_nhwc_f32(const tp_plane planes[3], int32_t C, const float* queries,
                                   int64_t Q, int32_t batch, const tp_sample_geom* sg,
                                   int32_t arith, float* out, void* stream) {
  if (C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: C=%d must be a positive multiple of 4", C);
  if (batch <= 0 || Q < 0) return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: bad B=%d Q=%lld", batch, (long long)Q);
  if (Q == 0) return 0;
  if (!planes || !queries || !out || !sg) return fail(TP_E_NULL, "tp_sample3_nhwc_f32: null argument");
  SampleParams P;
  for (int k = 0; k < 3; ++k) {
    if (!planes[k].data) return fail(TP_E_NULL, "tp_sample3_nhwc_f32: plane %d is null", k);
    if (planes[k].H <= 0 || planes[k].W <= 0 ||
        (int64_t)planes[k].H * planes[k].W * C >= (int64_t)1 << 31 || planes[k].H >= (1 << 20) ||
        planes[k].W >= (1 << 20))
      return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: plane %d H=%d W=%d unsupported", k, planes[k].H, planes[k].W);
    if ((uintptr_t)planes[k].data & 15 || (planes[k].batch_stride & 3))
      return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: plane %d not 16-byte aligned", k);
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
  P.out = out;
  P.Q = Q;
  P.C = C;
  P.tiles_per_sample = (Q + 31) / 32;
  P.tiles = P.tiles_per_sample * batch;
  if (P.tiles + kWarpsPerCta >= ((int64_t)1 << 32)) return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: too many queries");
  const int64_t ctas_needed = (P.tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = (int64_t)kSMs * 8;  // persistent: 8 resident CTAs per SM pull tiles dynamically
  const int grid = (int)(ctas_needed < cap ? ctas_needed : cap);
  static std::atomic<unsigned> next_slot{0};
  const int slot = (int)(next_slot.fetch_add(1) % kSchedSlots);
  cudaStream_t s = (cudaStream_t)stream;
#define TP_SAMPLE(A, C4T) sample3_kernel<A, C4T><<<grid, kWarpsPerCta * 32, 0, s>>>(P, slot)
#define TP_SAMPLE_C(A)                                                   \
  switch (C) { case 32: TP_SAMPLE(A, 8); break; case 96: TP_SAMPLE(A, 24); break; \
               case 128: TP_SAMPLE(A, 32); break; default: TP_SAMPLE(A, 0); break; }
  switch (arith) {
    case TP_ARITH_TORCH_CUDA: TP_SAMPLE_C(TP_ARITH_TORCH_CUDA) break;
    case TP_ARITH_TORCH_CPU:  TP_SAMPLE_C(TP_ARITH_TORCH_CPU) break;
    case 2:                   TP_SAMPLE(2, 0); break;
    default: return fail(TP_E_ENUM, "tp_sample3_nhwc_f32: unknown arith %d", arith);
  }
#undef TP_SAMPLE_C
#undef TP_SAMPLE
  TP_LAUNCH_CHECK("sample3_kernel");
  return 0;
}

extern "C" int tp_sample3_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const float* queries,
                                   int64_t Q, int32_t batch, const tp_sample_geom* sg,
                                   int32_t arith, float* out, float* ws, int64_t ws_floats,
                                   void* stream) {
  if (!planes_nchw || !ws) return fail(TP_E_NULL, "tp_sample3_nchw_f32: null argument");
  tp_plane nhwc[3];
  int64_t need = 0;
  for (int k = 0; k < 3; ++k) need += (int64_t)batch * C * planes_nchw[k].H * planes_nchw[k].W;
  if (ws_floats < need)
    return fail(TP_E_WORKSPACE, "tp_sample3_nchw_f32: workspace %lld < %lld floats", (long long)ws_floats, (long long)need);
  float* w = ws;
  float* dsts[3];
  for (int k = 0; k < 3; ++k) {
    dsts[k] = w;
    nhwc[k].data = w;
    nhwc[k].H = planes_nchw[k].H;
    nhwc[k].W = planes_nchw[k].W;
    nhwc[k].batch_stride = (int64_t)C * planes_nchw[k].H * planes_nchw[k].W;
    w += (int64_t)batch * nhwc[k].batch_stride;
  }
  if (int rc = tp_planes3_nchw_to_nhwc_f32(planes_nchw, dsts, batch, C, stream)) return rc;
  return tp_sample3_nhwc_f32(nhwc, C, queries, Q, batch, sg, arith, out, stream);
}
