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
constexpr int kParamWords = 32 * 16 + 16;  // 16 words per query + 4-word skew per 8-query group
constexpr int kTileStride = 33;            // [32 channels][32 queries] padded: conflict-free both ways
constexpr int kTileWords = 32 * kTileStride;

__device__ __forceinline__ int param_base(int qi) { return qi * 16 + (qi >> 3) * 4; }

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

template <int ARITH>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
sample3_kernel(const SampleParams P) {
  __shared__ __align__(16) float s_param[kWarpsPerCta][kParamWords];
  __shared__ float s_tile[kWarpsPerCta][kTileWords];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;
  float* sp = s_param[warp];
  float* st = s_tile[warp];
  const int C = P.C;
  const int nchunk = (C + 31) >> 5;
  const int64_t gwarp = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  const int64_t nwarp = (int64_t)gridDim.x * kWarpsPerCta;

  for (int64_t tile = gwarp; tile < P.tiles; tile += nwarp) {
    const int b = (int)(tile / P.tiles_per_sample);
    const int64_t q0 = (tile - (int64_t)b * P.tiles_per_sample) * 32;
    const int64_t q = q0 + lane;
    const bool qvalid = q < P.Q;

    // ---- per-query coordinate chain (one lane per query) ---------------------------------
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
      float4* dst = reinterpret_cast<float4*>(sp + param_base(lane));
      dst[0] = w[0];
      dst[1] = w[1];
      dst[2] = w[2];
      dst[3] = make_float4(__int_as_float(base[0]), __int_as_float(base[1]),
                           __int_as_float(base[2]),
                           __int_as_float(mask[0] | (mask[1] << 4) | (mask[2] << 8)));
    }
    __syncwarp();

    const float* pl0 = P.plane[0] + (int64_t)b * P.bstride[0];
    const float* pl1 = P.plane[1] + (int64_t)b * P.bstride[1];
    const float* pl2 = P.plane[2] + (int64_t)b * P.bstride[2];

    for (int ch = 0; ch < nchunk; ++ch) {
      const int c4 = ch * 32 + l8 * 4;  // first of this lane's 4 channels
      const bool cvalid = c4 < C;       // C % 4 == 0
#pragma unroll 2
      for (int pass = 0; pass < 8; ++pass) {
        const int qi = sub * 8 + pass;  // 8 consecutive queries per 8-lane group
        const float4* prm = reinterpret_cast<const float4*>(sp + param_base(qi));
        const float4 w0 = prm[0], w1 = prm[1], w2 = prm[2], bm = prm[3];
        const int m = cvalid ? __float_as_int(bm.w) : 0;
        float4 acc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float* pl = k == 0 ? pl0 : (k == 1 ? pl1 : pl2);
          const int Wk = P.W[k];
          const int base = __float_as_int(k == 0 ? bm.x : (k == 1 ? bm.y : bm.z));
          const float4 w = k == 0 ? w0 : (k == 1 ? w1 : w2);
          const int mk = (m >> (4 * k)) & 15;
          const float* t00 = pl + (int64_t)base * C + c4;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 v00 = z, v01 = z, v10 = z, v11 = z;
          if (mk & 1) v00 = __ldg(reinterpret_cast<const float4*>(t00));
          if (mk & 2) v01 = __ldg(reinterpret_cast<const float4*>(t00 + C));
          if (mk & 4) v10 = __ldg(reinterpret_cast<const float4*>(t00 + (int64_t)Wk * C));
          if (mk & 8) v11 = __ldg(reinterpret_cast<const float4*>(t00 + (int64_t)(Wk + 1) * C));
          // ATen accumulates out_acc += val * w for nw, ne, sw, se in that order, skipping
          // out-of-bounds taps (fma-contracted). Skipped taps must not contribute -0/NaN.
          float4 a = z;
          if (mk & 1) { a.x = __fmaf_rn(v00.x, w.x, a.x); a.y = __fmaf_rn(v00.y, w.x, a.y); a.z = __fmaf_rn(v00.z, w.x, a.z); a.w = __fmaf_rn(v00.w, w.x, a.w); }
          if (mk & 2) { a.x = __fmaf_rn(v01.x, w.y, a.x); a.y = __fmaf_rn(v01.y, w.y, a.y); a.z = __fmaf_rn(v01.z, w.y, a.z); a.w = __fmaf_rn(v01.w, w.y, a.w); }
          if (mk & 4) { a.x = __fmaf_rn(v10.x, w.z, a.x); a.y = __fmaf_rn(v10.y, w.z, a.y); a.z = __fmaf_rn(v10.z, w.z, a.z); a.w = __fmaf_rn(v10.w, w.z, a.w); }
          if (mk & 8) { a.x = __fmaf_rn(v11.x, w.w, a.x); a.y = __fmaf_rn(v11.y, w.w, a.y); a.z = __fmaf_rn(v11.z, w.w, a.z); a.w = __fmaf_rn(v11.w, w.w, a.w); }
          acc[k] = a;
        }
        // (xy + yz) + xz  (triplane_occ.py:345)
        float4 r;
        r.x = __fadd_rn(__fadd_rn(acc[0].x, acc[1].x), acc[2].x);
        r.y = __fadd_rn(__fadd_rn(acc[0].y, acc[1].y), acc[2].y);
        r.z = __fadd_rn(__fadd_rn(acc[0].z, acc[1].z), acc[2].z);
        r.w = __fadd_rn(__fadd_rn(acc[0].w, acc[1].w), acc[2].w);
        float* t = st + (l8 * 4) * kTileStride + qi;
        t[0] = r.x;
        t[kTileStride] = r.y;
        t[2 * kTileStride] = r.z;
        t[3 * kTileStride] = r.w;
      }
      __syncwarp();
      // ---- coalesced write of the [32 channels][32 queries] tile -------------------------
      if (qvalid) {
        float* o = P.out + ((int64_t)b * C + ch * 32) * P.Q + q;
        const int cmax = min(32, C - ch * 32);
#pragma unroll 8
        for (int c = 0; c < cmax; ++c) st_cs_f1(o + (int64_t)c * P.Q, st[c * kTileStride + lane]);
      }
      __syncwarp();
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

}  // namespace tp

using namespace tp;

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
  if (!planes || !queries || !out || !sg) return fail(TP_E_NULL, "tp_sample3_nhwc_f32: null argument");
  if (C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: C=%d must be a positive multiple of 4", C);
  if (batch <= 0 || Q < 0) return fail(TP_E_SHAPE, "tp_sample3_nhwc_f32: bad B=%d Q=%lld", batch, (long long)Q);
  if (Q == 0) return 0;
  SampleParams P;
  for (int k = 0; k < 3; ++k) {
    if (!planes[k].data) return fail(TP_E_NULL, "tp_sample3_nhwc_f32: plane %d is null", k);
    if (planes[k].H <= 0 || planes[k].W <= 0 ||
        (int64_t)planes[k].H * planes[k].W * C >= (int64_t)1 << 31)
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
  const int64_t ctas_needed = (P.tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = (int64_t)kSMs * 16;  // persistent beyond 16 CTAs/SM worth of tiles
  const int grid = (int)(ctas_needed < cap ? ctas_needed : cap);
  cudaStream_t s = (cudaStream_t)stream;
  switch (arith) {
    case TP_ARITH_TORCH_CUDA: sample3_kernel<TP_ARITH_TORCH_CUDA><<<grid, kWarpsPerCta * 32, 0, s>>>(P); break;
    case TP_ARITH_TORCH_CPU:  sample3_kernel<TP_ARITH_TORCH_CPU><<<grid, kWarpsPerCta * 32, 0, s>>>(P); break;
    case 2:                   sample3_kernel<2><<<grid, kWarpsPerCta * 32, 0, s>>>(P); break;
    default: return fail(TP_E_ENUM, "tp_sample3_nhwc_f32: unknown arith %d", arith);
  }
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
  for (int k = 0; k < 3; ++k) {
    if (!planes_nchw[k].data) return fail(TP_E_NULL, "tp_sample3_nchw_f32: plane %d is null", k);
    int rc = tp_planes_nchw_to_nhwc_f32(planes_nchw[k].data, planes_nchw[k].batch_stride, w, batch, C,
                                        planes_nchw[k].H, planes_nchw[k].W, stream);
    if (rc) return rc;
    nhwc[k].data = w;
    nhwc[k].H = planes_nchw[k].H;
    nhwc[k].W = planes_nchw[k].W;
    nhwc[k].batch_stride = (int64_t)C * planes_nchw[k].H * planes_nchw[k].W;
    w += (int64_t)batch * nhwc[k].batch_stride;
  }
  return tp_sample3_nhwc_f32(nhwc, C, queries, Q, batch, sg, arith, out, stream);
}
