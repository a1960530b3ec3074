// Backward kernels of the triplane hot path: the reference's configs are training
// configs, so a drop-in needs the gradients PyTorch's autograd would have produced for
//   * torch_scatter.scatter_max + SparseMaxPool3d       (point_triplane_projector.py:104,113-115)
//     -> the gradient of a pooled cell goes to the point(s) that attain the maximum;
//   * F.grid_sample w.r.t. its input                     (the five sample_points_triplane, point_to_cam)
//     -> scatter-add of grad_out * bilinear weight into the four taps (ATen grid_sampler_2d_backward).
// Gradients w.r.t. point / query coordinates are not produced (they are data in every config).
//
// Scatter-adds use 16-byte vector reductions (red.global.add.v4.f32, sm_90+) on channels-last
// gradient planes: one reduction covers 4 channels of a tap, 8 lanes cover 128 contiguous bytes.
#include "tp_sample_dev.cuh"

namespace tp {

// ---------------------------------------------------------------------------------------------
// decode backward: d(out[b,c,q]) -> d(planes), channels-last gradient planes (pre-zeroed)
// ---------------------------------------------------------------------------------------------
struct SampleBwdParams {
  SampleParams S;      // S.plane[k] are unused; S.out is unused
  float* gplane[3];    // [B,H,W,C] each, batch stride = S.bstride[k]
  const float* gout;   // [B,C,Q]
};

constexpr int kBwdWarps = 4;

template <int ARITH>
__global__ void __launch_bounds__(kBwdWarps * 32)
sample3_backward_kernel(const SampleBwdParams B) {
  __shared__ __align__(16) float s_param[kBwdWarps][kParamWords];
  __shared__ __align__(16) float s_tile[kBwdWarps][kTileWords];
  const SampleParams& P = B.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;
  float* sp = s_param[warp];
  float* st = s_tile[warp];
  const int C4 = P.C >> 2, C = P.C;
  const int nchunk = (C4 + 7) >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;

  for (int64_t tile = (int64_t)blockIdx.x * kBwdWarps + warp; tile < P.tiles; tile += (int64_t)gridDim.x * kBwdWarps) {
    const int b = (int)(tile / P.tiles_per_sample);
    const int64_t q0 = (tile - (int64_t)b * P.tiles_per_sample) * 32;
    const int64_t q = q0 + lane;
    const bool qvalid = q < P.Q;
    int anymask = 0;
    {
      float4 w[3] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
      int base[3] = {0, 0, 0}, mask[3] = {0, 0, 0};
      if (qvalid) {
        const float* qp = P.queries + ((int64_t)b * P.Q + q) * 3;
        const float g0 = grid_coord<ARITH>(P, __ldg(qp), 0);
        const float g1 = grid_coord<ARITH>(P, __ldg(qp + 1), 1);
        const float g2 = grid_coord<ARITH>(P, __ldg(qp + 2), 2);
        plane_setup<ARITH>(g0, g1, P.W[0], P.H[0], w[0], base[0], mask[0]);
        plane_setup<ARITH>(g1, g2, P.W[1], P.H[1], w[1], base[1], mask[1]);
        plane_setup<ARITH>(g0, g2, P.W[2], P.H[2], w[2], base[2], mask[2]);
      }
      anymask = mask[0] | (mask[1] << 4) | (mask[2] << 8);
      float4* dst = reinterpret_cast<float4*>(sp + param_base(lane));
      dst[0] = w[0];
      dst[1] = w[1];
      dst[2] = w[2];
      dst[3] = make_float4(__int_as_float(base[0] * C4), __int_as_float(base[1] * C4),
                           __int_as_float(base[2] * C4), __int_as_float(anymask));
    }
    const bool tile_empty = __all_sync(0xffffffffu, anymask == 0);
    __syncwarp();
    if (tile_empty) continue;
    float4* gp0 = reinterpret_cast<float4*>(B.gplane[0] + (int64_t)b * P.bstride[0]) + l8;
    float4* gp1 = reinterpret_cast<float4*>(B.gplane[1] + (int64_t)b * P.bstride[1]) + l8;
    float4* gp2 = reinterpret_cast<float4*>(B.gplane[2] + (int64_t)b * P.bstride[2]) + l8;
    for (int ch = 0; ch < nchunk; ++ch) {
      const int cmax = min(32, C - ch * 32);
      // coalesced read of the [32 channels][32 queries] slab of grad_out, swizzled like the forward tile
      const float* grow = B.gout + ((int64_t)b * C + ch * 32) * P.Q + q;
      for (int c = 0; c < 32; ++c)
        st[c * 32 + (lane ^ ((c >> 2) & 7))] = (qvalid && c < cmax) ? __ldg(grow + (int64_t)c * P.Q) : 0.f;
      __syncwarp();
      const bool cvalid = ch * 32 + l8 * 4 < C;
      const float* tcol = st + (l8 * 4) * 32;
      for (int pass = 0; pass < 8; ++pass) {
        const int qi = sub * 8 + pass;
        const float4* prm = reinterpret_cast<const float4*>(sp + param_base(qi));
        const float4 w0 = prm[0], w1 = prm[1], w2 = prm[2], bm = prm[3];
        const int m = cvalid ? __float_as_int(bm.w) : 0;
        if (m == 0) continue;
        const float* t = tcol + (qi ^ l8);
        const float4 g = make_float4(t[0], t[32], t[64], t[96]);
        scatter_taps(gp0 + ch * 8, __float_as_int(bm.x), C4, WC4_0, w0, m & 15, g);
        scatter_taps(gp1 + ch * 8, __float_as_int(bm.y), C4, WC4_1, w1, (m >> 4) & 15, g);
        scatter_taps(gp2 + ch * 8, __float_as_int(bm.z), C4, WC4_2, w2, (m >> 8) & 15, g);
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// encode backward: d(pooled planes) -> d(point features)
// ---------------------------------------------------------------------------------------------
struct EncodeBwdParams {
  const float* feats;      // [N, feat_stride]
  const int32_t* idx;      // [N,3]
  const int64_t* offsets;  // [B+1]
  const float* out[3];     // forward outputs (max mode) or nullptr
  const float* gout[3];    // gradients of the three outputs
  const int32_t* count;    // pooled-cell counts (mean mode), xy | yz | xz
  float* gfeats;           // [N, C]
  int64_t n_total, feat_stride;
  int64_t cells_before[3]; // offset of each plane's cells in `count`
  GeomDev g;
  int batch, C, mode, clamp_zero;  // mode: TP_REDUCE_MAX or TP_REDUCE_MEAN
};

__global__ void __launch_bounds__(128)
encode_backward_kernel(const EncodeBwdParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (n >= P.n_total) return;
  const int ix = __ldg(P.idx + n * 3), iy = __ldg(P.idx + n * 3 + 1), iz = __ldg(P.idx + n * 3 + 2);
  const GeomDev& g = P.g;
  const int C4 = P.C >> 2;
  float4* grow = reinterpret_cast<float4*>(P.gfeats + n * P.C);
  const bool inside = (ix >= 0) & (ix < g.grid[0]) & (iy >= 0) & (iy < g.grid[1]) & (iz >= 0) & (iz < g.grid[2]);
  int64_t cell[3] = {-1, -1, -1};
  if (inside) {
    const int b = tp_find_batch(P.offsets, P.batch, n);
    const int px = ix / g.pool[0], py = iy / g.pool[1], pz = iz / g.pool[2];
    // same cell order as the forward: xy [B,X,Y,Zp], yz [B,Y,Z,Xp], xz [B,X,Z,Yp]
    if (pz < g.pooled[2]) cell[0] = (((int64_t)b * g.grid[0] + ix) * g.grid[1] + iy) * g.pooled[2] + pz;
    if (px < g.pooled[0]) cell[1] = (((int64_t)b * g.grid[1] + iy) * g.grid[2] + iz) * g.pooled[0] + px;
    if (py < g.pooled[1]) cell[2] = (((int64_t)b * g.grid[0] + ix) * g.grid[2] + iz) * g.pooled[1] + py;
  }
  const float4* frow = reinterpret_cast<const float4*>(P.feats + n * P.feat_stride);
  for (int c4 = lane; c4 < C4; c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 f = (P.mode == TP_REDUCE_MAX) ? __ldg(frow + c4) : acc;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (cell[k] < 0 || !P.gout[k]) continue;
      const float4 go = __ldg(reinterpret_cast<const float4*>(P.gout[k]) + cell[k] * C4 + c4);
      if (P.mode == TP_REDUCE_MAX) {
        const float4 o = __ldg(reinterpret_cast<const float4*>(P.out[k]) + cell[k] * C4 + c4);
        // the maximiser gets the gradient; with clamp_zero a clamped (non-positive) maximum gets none
        const bool cz = P.clamp_zero != 0;
        if (f.x == o.x && !(cz && !(o.x > 0.f))) acc.x += go.x;
        if (f.y == o.y && !(cz && !(o.y > 0.f))) acc.y += go.y;
        if (f.z == o.z && !(cz && !(o.z > 0.f))) acc.z += go.z;
        if (f.w == o.w && !(cz && !(o.w > 0.f))) acc.w += go.w;
      } else {
        const float inv = 1.0f / (float)max(1, __ldg(P.count + P.cells_before[k] + cell[k]));
        acc.x += go.x * inv; acc.y += go.y * inv; acc.z += go.z * inv; acc.w += go.w * inv;
      }
    }
    grow[c4] = acc;
  }
}

}  // namespace tp

using namespace tp;

extern "C" int tp_sample3_backward_nhwc_f32(const tp_plane gplanes_nhwc[3], int32_t C, const float* queries,
                                            int64_t Q, int32_t batch, const tp_sample_geom* sg, int32_t arith,
                                            const float* grad_out, void* stream) {
  if (C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "tp_sample3_backward_nhwc_f32: C=%d must be a positive multiple of 4", C);
  if (batch <= 0 || Q < 0) return fail(TP_E_SHAPE, "tp_sample3_backward_nhwc_f32: bad B=%d Q=%lld", batch, (long long)Q);
  if (Q == 0) return 0;
  if (!gplanes_nhwc || !queries || !grad_out || !sg) return fail(TP_E_NULL, "tp_sample3_backward_nhwc_f32: null argument");
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_sample3_backward_nhwc_f32: unknown arith %d", arith);
  SampleBwdParams B;
  SampleParams& P = B.S;
  for (int k = 0; k < 3; ++k) {
    if (!gplanes_nhwc[k].data) return fail(TP_E_NULL, "tp_sample3_backward_nhwc_f32: plane %d is null", k);
    if (gplanes_nhwc[k].H <= 0 || gplanes_nhwc[k].W <= 0 ||
        (int64_t)gplanes_nhwc[k].H * gplanes_nhwc[k].W * C >= (int64_t)1 << 31)
      return fail(TP_E_SHAPE, "tp_sample3_backward_nhwc_f32: plane %d shape unsupported", k);
    if ((uintptr_t)gplanes_nhwc[k].data & 15 || (gplanes_nhwc[k].batch_stride & 3))
      return fail(TP_E_SHAPE, "tp_sample3_backward_nhwc_f32: plane %d not 16-byte aligned", k);
    P.plane[k] = nullptr;
    B.gplane[k] = const_cast<float*>(gplanes_nhwc[k].data);
    P.bstride[k] = gplanes_nhwc[k].batch_stride;
    P.H[k] = gplanes_nhwc[k].H;
    P.W[k] = gplanes_nhwc[k].W;
    P.lo[k] = sg->lo[k];
    P.vs[k] = sg->vs[k];
    P.rcp_vs[k] = 1.0f / sg->vs[k];
    P.half[k] = sg->half[k];
    P.rcp_half[k] = 1.0f / sg->half[k];
  }
  P.queries = queries;
  P.out = nullptr;
  B.gout = grad_out;
  P.Q = Q;
  P.C = C;
  P.tiles_per_sample = (Q + 31) / 32;
  P.tiles = P.tiles_per_sample * batch;
  const int64_t need = (P.tiles + kBwdWarps - 1) / kBwdWarps;
  const int64_t cap = (int64_t)kSMs * 12;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t s = (cudaStream_t)stream;
  if (arith == TP_ARITH_TORCH_CUDA) sample3_backward_kernel<TP_ARITH_TORCH_CUDA><<<grid, kBwdWarps * 32, 0, s>>>(B);
  else sample3_backward_kernel<TP_ARITH_TORCH_CPU><<<grid, kBwdWarps * 32, 0, s>>>(B);
  TP_LAUNCH_CHECK("sample3_backward_kernel");
  return 0;
}

extern "C" int tp_encode_backward_f32(const float* feats, int64_t feat_stride, int32_t C, const int32_t* idx,
                                      int64_t n_total, const int64_t* offsets, int32_t batch, const tp_geom* geom,
                                      int32_t reduce, int32_t clamp_zero, const float* out_xy, const float* out_yz,
                                      const float* out_xz, const int32_t* cell_count, const float* gout_xy,
                                      const float* gout_yz, const float* gout_xz, float* grad_feats, void* stream) {
  if (!geom) return fail(TP_E_NULL, "tp_encode_backward_f32: null geometry");
  if (C <= 0 || (C & 3) || batch <= 0 || n_total < 0) return fail(TP_E_SHAPE, "tp_encode_backward_f32: bad C=%d B=%d N=%lld", C, batch, (long long)n_total);
  if (reduce != TP_REDUCE_MAX && reduce != TP_REDUCE_MEAN) return fail(TP_E_ENUM, "tp_encode_backward_f32: reduce must be MAX or MEAN");
  if (n_total == 0) return 0;
  if (!idx || !offsets || !grad_feats) return fail(TP_E_NULL, "tp_encode_backward_f32: null argument");
  if (reduce == TP_REDUCE_MAX && (!feats || (feat_stride & 3) || ((uintptr_t)feats & 15)))
    return fail(TP_E_SHAPE, "tp_encode_backward_f32: feats must be 16-byte aligned rows");
  if (reduce == TP_REDUCE_MEAN && !cell_count) return fail(TP_E_NULL, "tp_encode_backward_f32: mean needs cell_count");
  int64_t cells[3];
  if (tp_encode_cells(geom, batch, cells) < 0) return fail(TP_E_SHAPE, "tp_encode_backward_f32: bad geometry");
  EncodeBwdParams P;
  P.feats = feats; P.idx = idx; P.offsets = offsets; P.count = cell_count; P.gfeats = grad_feats;
  P.out[0] = out_xy; P.out[1] = out_yz; P.out[2] = out_xz;
  P.gout[0] = gout_xy; P.gout[1] = gout_yz; P.gout[2] = gout_xz;
  for (int k = 0; k < 3; ++k)
    if (reduce == TP_REDUCE_MAX && P.gout[k] && !P.out[k]) return fail(TP_E_NULL, "tp_encode_backward_f32: plane %d has a gradient but no forward output", k);
  P.n_total = n_total; P.feat_stride = feat_stride;
  P.cells_before[0] = 0; P.cells_before[1] = cells[0]; P.cells_before[2] = cells[0] + cells[1];
  P.g = make_geom_dev(*geom);
  P.batch = batch; P.C = C; P.mode = reduce; P.clamp_zero = clamp_zero;
  const int64_t ctas = (n_total + 3) / 4;
  if (ctas >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_encode_backward_f32: too many points");
  encode_backward_kernel<<<(unsigned)ctas, 128, 0, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("encode_backward_kernel");
  return 0;
}
