// Triplane decode for ragged point subsets: ONE launch for a list of variable-length query segments, each bound to
// one sample's planes. Replaces the per-(sample, camera) Python loops around sample_points_triplane in the
// contrastive branches of the reference (triplane.py:438-458: B x 6 calls on SAM-labelled subsets of ~3-6 k points;
// point_triplane.py:365-372, 389-403: per-sample calls), each of which was three grid_sample launches plus the
// normalisation temporaries on a [1,C,1,N_s] problem that cannot fill the GPU.
//
// queries [T, 3] = the segments concatenated; seg_offsets [S+1] (device, int64); seg_batch [S] (device, int32; NULL:
// segment s reads sample s). The result is POINT-major, out [T, C]: every caller consumes the sampled features as
// [N_s, C] rows (`features.permute(1, 0)` before SupConLoss, triplane.py:453-455; `.squeeze().T` in
// point_triplane.py:372), and a query's C channels are contiguous in the channels-last planes, so a row leaves as
// 16-byte stores without the transpose through shared memory the channel-major kernels need.
// Same coordinate chain, weights and accumulation order as tp_sample.cu (tp_sample_dev.cuh): bit-identical values.
#include "tp_sample_dev.cuh"

namespace tp {

struct SegParams {
  SampleParams S;              // S.queries [T,3]; S.out [T,C] point-major (forward); S.Q = T; S.tiles = ceil(T/32)
  const int64_t* seg_offsets;  // [nseg+1]
  const int32_t* seg_batch;    // [nseg] or nullptr
  int nseg, batch;
  int bs4[3];                  // plane batch stride in float4 units
  float* gplane[3];            // backward: channels-last gradient planes (pre-zeroed)
  const float* gout;           // backward: [T, C] point-major
};

constexpr int kSegWarps = 4;

// segment of a concatenated query index: largest s with seg_offsets[s] <= q
__device__ __forceinline__ int seg_find(const int64_t* __restrict__ off, int nseg, int64_t q) {
  int lo = 0, hi = nseg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= q) lo = mid; else hi = mid;
  }
  return lo;
}

// per-query record: 3 x float4 weights, {tap offsets x3, 12-bit mask}, sample index (word 16 of the 20-word slot)
template <int ARITH>
__device__ __forceinline__ int seg_setup(const SegParams& G, int64_t q, bool qvalid, float* sp, int lane, int C4) {
  const SampleParams& P = G.S;
  float4 w[3] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  int base[3] = {0, 0, 0}, mask[3] = {0, 0, 0}, b = 0;
  if (qvalid) {
    const int s = seg_find(G.seg_offsets, G.nseg, q);
    b = G.seg_batch ? __ldg(G.seg_batch + s) : s;
    if (b >= 0 && b < G.batch) {
      const float* qp = P.queries + q * 3;
      const float g0 = grid_coord<ARITH>(P, __ldg(qp), 0);
      const float g1 = grid_coord<ARITH>(P, __ldg(qp + 1), 1);
      const float g2 = grid_coord<ARITH>(P, __ldg(qp + 2), 2);
      plane_setup<ARITH>(g0, g1, P.W[0], P.H[0], w[0], base[0], mask[0]);  // (x,y)
      plane_setup<ARITH>(g1, g2, P.W[1], P.H[1], w[1], base[1], mask[1]);  // (y,z)
      plane_setup<ARITH>(g0, g2, P.W[2], P.H[2], w[2], base[2], mask[2]);  // (x,z)
    } else {
      b = 0;  // a segment bound to a sample that does not exist reads nothing and yields zeros
    }
  }
  const int anymask = mask[0] | (mask[1] << 4) | (mask[2] << 8);
  float4* dst = reinterpret_cast<float4*>(sp + param_base(lane));
  dst[0] = w[0];
  dst[1] = w[1];
  dst[2] = w[2];
  // tap offsets in float4 units INCLUDING the sample's plane offset (< 2^31, host-checked): the gather loop below then
  // has no per-query 64-bit arithmetic (it was 4 M IMAD + 1.4 M LDC of 18 M warp instructions on configs[2])
  dst[3] = make_float4(__int_as_float(b * G.bs4[0] + base[0] * C4), __int_as_float(b * G.bs4[1] + base[1] * C4),
                       __int_as_float(b * G.bs4[2] + base[2] * C4), __int_as_float(anymask));
  return anymask;
}

template <int ARITH, bool BACKWARD>
__global__ void __launch_bounds__(kSegWarps * 32, 8)
sample3_seg_kernel(const SegParams G) {
  __shared__ __align__(16) float s_param[kSegWarps][kParamWords];
  const SampleParams& P = G.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;
  float* sp = s_param[warp];
  const int C4 = P.C >> 2, C = P.C;
  const int nchunk = (C4 + 7) >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  const unsigned long long pol_planes = policy_evict_last(), pol_out = policy_evict_first();

  const float4* const pl0 = reinterpret_cast<const float4*>(P.plane[0]) + l8;
  const float4* const pl1 = reinterpret_cast<const float4*>(P.plane[1]) + l8;
  const float4* const pl2 = reinterpret_cast<const float4*>(P.plane[2]) + l8;
  float4* const gp0 = reinterpret_cast<float4*>(G.gplane[0]) + l8;
  float4* const gp1 = reinterpret_cast<float4*>(G.gplane[1]) + l8;
  float4* const gp2 = reinterpret_cast<float4*>(G.gplane[2]) + l8;
  for (int64_t tile = (int64_t)blockIdx.x * kSegWarps + warp; tile < P.tiles; tile += (int64_t)gridDim.x * kSegWarps) {
    const int64_t q = tile * 32 + lane;
    const int anymask = seg_setup<ARITH>(G, q, q < P.Q, sp, lane, C4);
    const bool tile_empty = __all_sync(0xffffffffu, anymask == 0);
    __syncwarp();
    if (BACKWARD && tile_empty) continue;
    if (!BACKWARD && !tile_empty) pdl_wait();  // first read of the planes (tp_common.cuh)
    for (int ch = 0; ch < nchunk; ++ch) {
      const bool cvalid = ch * 32 + l8 * 4 < C;
#pragma unroll 2
      for (int pass = 0; pass < 8; ++pass) {
        const int qi = sub * 8 + pass;
        const int64_t qq = tile * 32 + qi;
        if (qq >= P.Q) continue;
        const float* rec = sp + param_base(qi);
        const float4* prm = reinterpret_cast<const float4*>(rec);
        const float4 w0 = prm[0], w1 = prm[1], w2 = prm[2], bm = prm[3];
        const int m = cvalid ? __float_as_int(bm.w) : 0;
        if (!BACKWARD) {
          const float4 a0 = plane_taps<true>(pl0 + ch * 8, __float_as_int(bm.x), C4, WC4_0, w0, m & 15, pol_planes);
          const float4 a1 = plane_taps<true>(pl1 + ch * 8, __float_as_int(bm.y), C4, WC4_1, w1, (m >> 4) & 15, pol_planes);
          const float4 a2 = plane_taps<true>(pl2 + ch * 8, __float_as_int(bm.z), C4, WC4_2, w2, (m >> 8) & 15, pol_planes);
          if (cvalid) {
            float4 r;  // (xy + yz) + xz  (triplane.py:512)
            r.x = __fadd_rn(__fadd_rn(a0.x, a1.x), a2.x);
            r.y = __fadd_rn(__fadd_rn(a0.y, a1.y), a2.y);
            r.z = __fadd_rn(__fadd_rn(a0.z, a1.z), a2.z);
            r.w = __fadd_rn(__fadd_rn(a0.w, a1.w), a2.w);
            st_stream_f4(reinterpret_cast<float4*>(P.out + qq * C) + ch * 8 + l8, r, pol_out);
          }
        } else {
          if (m == 0) continue;
          const float4 g = __ldg(reinterpret_cast<const float4*>(G.gout + qq * C) + ch * 8 + l8);
          scatter_taps(gp0 + ch * 8, __float_as_int(bm.x), C4, WC4_0, w0, m & 15, g);
          scatter_taps(gp1 + ch * 8, __float_as_int(bm.y), C4, WC4_1, w1, (m >> 4) & 15, g);
          scatter_taps(gp2 + ch * 8, __float_as_int(bm.z), C4, WC4_2, w2, (m >> 8) & 15, g);
        }
      }
    }
    __syncwarp();
  }
}

static int seg_fill(SegParams& G, const char* who, const tp_plane planes[3], int32_t C, const float* queries,
                    int64_t total, const int64_t* seg_offsets, const int32_t* seg_batch, int32_t nseg, int32_t batch,
                    const tp_sample_geom* sg, int32_t arith, const float* io) {
  if (C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "%s: C=%d must be a positive multiple of 4", who, C);
  if (batch <= 0 || nseg <= 0 || total < 0) return fail(TP_E_SHAPE, "%s: bad B=%d S=%d T=%lld", who, batch, nseg, (long long)total);
  if (!planes || !queries || !io || !sg || !seg_offsets) return fail(TP_E_NULL, "%s: null argument", who);
  if (!seg_batch && nseg > batch) return fail(TP_E_SHAPE, "%s: %d segments but %d samples and no seg_batch", who, nseg, batch);
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "%s: unknown arith %d", who, arith);
  if ((uintptr_t)io & 15) return fail(TP_E_SHAPE, "%s: the [T,C] tensor must be 16-byte aligned", who);
  SampleParams& P = G.S;
  for (int k = 0; k < 3; ++k) {
    if (!planes[k].data) return fail(TP_E_NULL, "%s: plane %d is null", who, k);
    if (planes[k].H <= 0 || planes[k].W <= 0 || (int64_t)planes[k].H * planes[k].W * C >= (int64_t)1 << 31 ||
        planes[k].H >= (1 << 20) || planes[k].W >= (1 << 20))
      return fail(TP_E_SHAPE, "%s: plane %d H=%d W=%d unsupported", who, k, planes[k].H, planes[k].W);
    if ((uintptr_t)planes[k].data & 15 || (planes[k].batch_stride & 3))
      return fail(TP_E_SHAPE, "%s: plane %d not 16-byte aligned", who, k);
    P.plane[k] = planes[k].data;
    G.gplane[k] = const_cast<float*>(planes[k].data);
    P.bstride[k] = planes[k].batch_stride;
    if (((int64_t)batch + 1) * (planes[k].batch_stride >> 2) >= ((int64_t)1 << 31))
      return fail(TP_E_SHAPE, "%s: plane %d: batch * batch_stride must stay below 2^33 floats", who, k);
    G.bs4[k] = (int)(planes[k].batch_stride >> 2);
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
  G.gout = nullptr;
  P.Q = total;
  P.C = C;
  P.tiles_per_sample = 0;
  P.tiles = (total + 31) / 32;
  G.seg_offsets = seg_offsets;
  G.seg_batch = seg_batch;
  G.nseg = nseg;
  G.batch = batch;
  return 0;
}

}  // namespace tp

using namespace tp;

static int seg_entry(const tp_plane planes[3], int32_t C, const float* queries, int64_t total,
                     const int64_t* seg_offsets, const int32_t* seg_batch, int32_t nseg, int32_t batch,
                     const tp_sample_geom* sg, int32_t arith, float* out, void* stream, bool pdl) {
  if (total == 0) return 0;
  SegParams G;
  if (int rc = seg_fill(G, "tp_sample3_seg_nhwc_f32", planes, C, queries, total, seg_offsets, seg_batch, nseg, batch, sg,
                        arith, out))
    return rc;
  G.S.out = out;
  const int64_t need = (G.S.tiles + kSegWarps - 1) / kSegWarps;
  const int64_t cap = (int64_t)kSMs * 8;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t s = (cudaStream_t)stream;
  if (arith == TP_ARITH_TORCH_CUDA) launch_kernel(sample3_seg_kernel<TP_ARITH_TORCH_CUDA, false>, grid, kSegWarps * 32, 0, s, pdl, G);
  else launch_kernel(sample3_seg_kernel<TP_ARITH_TORCH_CPU, false>, grid, kSegWarps * 32, 0, s, pdl, G);
  TP_LAUNCH_CHECK("sample3_seg_kernel");
  return 0;
}

extern "C" int tp_sample3_seg_nhwc_f32(const tp_plane planes[3], int32_t C, const float* queries, int64_t total,
                                       const int64_t* seg_offsets, const int32_t* seg_batch, int32_t nseg,
                                       int32_t batch, const tp_sample_geom* sg, int32_t arith, float* out,
                                       void* stream) {
  return seg_entry(planes, C, queries, total, seg_offsets, seg_batch, nseg, batch, sg, arith, out, stream, false);
}

extern "C" int tp_sample3_seg_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const float* queries, int64_t total,
                                       const int64_t* seg_offsets, const int32_t* seg_batch, int32_t nseg,
                                       int32_t batch, const tp_sample_geom* sg, int32_t arith, float* out, float* ws,
                                       int64_t ws_floats, void* stream) {
  if (total == 0) return 0;
  tp_plane nhwc[3];
  if (int rc = planes3_to_workspace("tp_sample3_seg_nchw_f32", planes_nchw, C, batch, ws, ws_floats, nhwc, stream)) return rc;
  return seg_entry(nhwc, C, queries, total, seg_offsets, seg_batch, nseg, batch, sg, arith, out, stream, true);
}

extern "C" int tp_sample3_seg_backward_nhwc_f32(const tp_plane gplanes_nhwc[3], int32_t C, const float* queries,
                                                int64_t total, const int64_t* seg_offsets,
                                                const int32_t* seg_batch, int32_t nseg, int32_t batch,
                                                const tp_sample_geom* sg, int32_t arith, const float* grad_out,
                                                void* stream) {
  if (total == 0) return 0;
  SegParams G;
  if (int rc = seg_fill(G, "tp_sample3_seg_backward_nhwc_f32", gplanes_nhwc, C, queries, total, seg_offsets, seg_batch,
                        nseg, batch, sg, arith, grad_out))
    return rc;
  G.gout = grad_out;
  const int64_t need = (G.S.tiles + kSegWarps - 1) / kSegWarps;
  const int64_t cap = (int64_t)kSMs * 8;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t s = (cudaStream_t)stream;
  if (arith == TP_ARITH_TORCH_CUDA) sample3_seg_kernel<TP_ARITH_TORCH_CUDA, true><<<grid, kSegWarps * 32, 0, s>>>(G);
  else sample3_seg_kernel<TP_ARITH_TORCH_CPU, true><<<grid, kSegWarps * 32, 0, s>>>(G);
  TP_LAUNCH_CHECK("sample3_seg_backward_kernel");
  return 0;
}
