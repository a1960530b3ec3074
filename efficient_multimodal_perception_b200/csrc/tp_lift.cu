// Camera -> point lift: PointTriplane.point_to_cam (point_triplane.py:164-241, twin
// point_triplane_occ.py:163-239). For every LiDAR point: project with each camera's lidar2image
// matrix, perspective divide, image augmentation (resize, crop, flip, the identity rotation),
// in-image test, then a bilinear sample of that camera's feature map (zeros padding,
// align_corners=False; the reference feeds the image ROW as grid-x, :230-235) summed over the
// cameras that see the point. The reference runs B x 6 Python iterations of ~15 small kernels plus
// a boolean-mask gather/scatter each; here it is one launch, one read of each point and one write
// of its feature row.
//
// Feature maps are read channels-last [B, ncam, Hf, Wf, Cf] (tp_planes_nchw_to_nhwc_f32 with
// batch = B * ncam converts the encoder's NCHW output): a tap is Cf contiguous floats, a warp reads
// 512 contiguous bytes per instruction.
#include "tp_cam.cuh"
#include "tp_sample_dev.cuh"

namespace tp {

constexpr int kLiftPts = 8;    // points per warp tile
constexpr int kLiftCams = 8;   // camera slots per point (2 per lane in the projection phase)
constexpr int kLiftWarps = 4;

struct LiftParams {
  const float* points;     // [N, point_stride], xyz first
  const int64_t* offsets;  // [B+1]
  const float* feats;      // [B, ncam, Hf, Wf, Cf]
  const float* cams;       // [B, ncam, 20]: lidar2image row-major 4x4, resize, crop_x, crop_y, flip
  float* out;              // [N, Cf]
  int64_t n_total;
  int point_stride, batch, ncam, Hf, Wf, Cf;
  float R0, R1;            // resize_dims = img_shape[::-1]  (point_triplane.py:176)
  float half0, half1;      // R0 / 2.0, R1 / 2.0
  float rcp0, rcp1;        // fp32(1 / R0), fp32(1 / R1)
};

// one (point, camera): 4 bilinear weights + nw tap offset (in float4 units) + in-bounds mask (0 = the
// camera does not see the point)
template <int ARITH>
__device__ __forceinline__ void lift_setup(const LiftParams& P, const float* __restrict__ cam, float px, float py,
                                           float pz, float4& w, int& off, int& mask) {
  const CamPixel cp = cam_project(cam, px, py, pz, P.R0, P.R1, P.half0, P.half1);
  const float x = cp.x, y = cp.y;
  const bool valid = cp.valid;
  // swap to (row, col); normalise 2*row/H - 1, 2*col/W - 1 with H = R0, W = R1 (:229-233)
  float g0 = __fmul_rn(2.0f, y), g1 = __fmul_rn(2.0f, x);
  g0 = (ARITH == TP_ARITH_TORCH_CPU) ? __fdiv_rn(g0, P.R0) : __fmul_rn(g0, P.rcp0);
  g1 = (ARITH == TP_ARITH_TORCH_CPU) ? __fdiv_rn(g1, P.R1) : __fmul_rn(g1, P.rcp1);
  g0 = __fsub_rn(g0, 1.0f);
  g1 = __fsub_rn(g1, 1.0f);
  // grid[..., 0] = row -> feature-map x (width Wf); grid[..., 1] = col -> feature-map y (height Hf)
  int base;
  plane_setup<ARITH>(g0, g1, P.Wf, P.Hf, w, base, mask);
  if (!valid) mask = 0;
  off = mask ? base * (P.Cf >> 2) : 0;
  // a camera that sees the point but whose 4 taps are all outside still contributes +0 (no effect)
}

template <int ARITH>
__global__ void __launch_bounds__(kLiftWarps * 32)
lift_kernel(const LiftParams P) {
  __shared__ __align__(16) float4 s_w[kLiftWarps][kLiftPts * kLiftCams];
  __shared__ __align__(8) int2 s_om[kLiftWarps][kLiftPts * kLiftCams];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * kLiftWarps + warp;
  const int64_t p0 = tile * kLiftPts;
  if (p0 >= P.n_total) return;
  const unsigned long long pol_feat = policy_evict_last(), pol_out = policy_evict_first();
  const int C4 = P.Cf >> 2;

  // ---- projection: lane -> (point lane & 7, cameras lane >> 3 and (lane >> 3) + 4) -------------
  {
    const int pi = lane & 7;
    const int64_t p = p0 + pi;
    const bool pv = p < P.n_total;
    float px = 0.f, py = 0.f, pz = 0.f;
    int b = 0;
    if (pv) {
      const float* pp = P.points + p * P.point_stride;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      b = tp_find_batch(P.offsets, P.batch, p);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cam = (lane >> 3) + 4 * h;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      int off = 0, mask = 0, seen = 0;
      if (pv && cam < P.ncam) {
        lift_setup<ARITH>(P, P.cams + ((int64_t)b * P.ncam + cam) * 20, px, py, pz, w, off, mask);
        seen = mask != 0;
      }
      s_w[warp][pi * kLiftCams + cam] = w;
      // bits 0-3: tap mask; the map index rides in the offset: + (b * ncam + cam) * Hf * Wf * C4
      s_om[warp][pi * kLiftCams + cam] =
          make_int2(off, seen ? (mask | ((b * P.ncam + cam) << 4)) : 0);
    }
  }
  __syncwarp();

  // ---- gather: the whole warp per point, 512 contiguous bytes per load instruction ---------------
  const int64_t map4 = (int64_t)P.Hf * P.Wf * C4;  // float4 per feature map
  const int WC4 = P.Wf * C4;
  const float4* feats = reinterpret_cast<const float4*>(P.feats);
  for (int pi = 0; pi < kLiftPts; ++pi) {
    const int64_t p = p0 + pi;
    if (p >= P.n_total) break;
    float4* orow = reinterpret_cast<float4*>(P.out + p * P.Cf);
    // which cameras see this point (warp-uniform bit mask): most points are seen by none or by one
    const unsigned seen = __ballot_sync(0xffffffffu, lane < P.ncam && s_om[warp][pi * kLiftCams + (lane & 7)].y != 0);
    if (seen == 0) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 6
      for (int c4 = lane; c4 < C4; c4 += 32) st_stream_f4(orow + c4, z, pol_out);
      continue;
    }
    if ((seen & (seen - 1)) == 0) {
      // exactly one camera: 0 + f == f bit for bit except f == -0 (-> +0), which the add below reproduces
      const int cam = __ffs(seen) - 1;
      const int2 om = s_om[warp][pi * kLiftCams + cam];
      const float4 w = s_w[warp][pi * kLiftCams + cam];
      const float4* pl = feats + (int64_t)(om.y >> 4) * map4;
#pragma unroll 3
      for (int c4 = lane; c4 < C4; c4 += 32) {
        const float4 f = plane_taps<true>(pl + c4, om.x, C4, WC4, w, om.y & 15, pol_feat);
        st_stream_f4(orow + c4, make_float4(__fadd_rn(0.f, f.x), __fadd_rn(0.f, f.y), __fadd_rn(0.f, f.z),
                                            __fadd_rn(0.f, f.w)), pol_out);
      }
      continue;
    }
    for (int c4 = lane; c4 < C4; c4 += 32) {
      float4 total = make_float4(0.f, 0.f, 0.f, 0.f);
      for (unsigned m = seen; m; m &= m - 1) {  // cameras in ascending order
        const int cam = __ffs(m) - 1;
        const int2 om = s_om[warp][pi * kLiftCams + cam];
        const float4 w = s_w[warp][pi * kLiftCams + cam];
        const float4* pl = feats + (int64_t)(om.y >> 4) * map4 + c4;
        const float4 f = plane_taps<true>(pl, om.x, C4, WC4, w, om.y & 15, pol_feat);
        // point_feature[valid] += features (:236), cameras in ascending order from zeros
        total.x = __fadd_rn(total.x, f.x);
        total.y = __fadd_rn(total.y, f.y);
        total.z = __fadd_rn(total.z, f.z);
        total.w = __fadd_rn(total.w, f.w);
      }
      st_stream_f4(orow + c4, total, pol_out);
    }
  }
}

// backward w.r.t. the feature maps: grad_out[p, :] * weight into the four taps of every camera that
// sees the point (channels-last gradient maps, pre-zeroed by the caller)
template <int ARITH>
__global__ void __launch_bounds__(kLiftWarps * 32)
lift_backward_kernel(const LiftParams P, const float* __restrict__ gout, float* __restrict__ gfeats) {
  __shared__ __align__(16) float4 s_w[kLiftWarps][kLiftPts * kLiftCams];
  __shared__ __align__(8) int2 s_om[kLiftWarps][kLiftPts * kLiftCams];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p0 = ((int64_t)blockIdx.x * kLiftWarps + warp) * kLiftPts;
  if (p0 >= P.n_total) return;
  const int C4 = P.Cf >> 2;
  {
    const int pi = lane & 7;
    const int64_t p = p0 + pi;
    const bool pv = p < P.n_total;
    float px = 0.f, py = 0.f, pz = 0.f;
    int b = 0;
    if (pv) {
      const float* pp = P.points + p * P.point_stride;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      b = tp_find_batch(P.offsets, P.batch, p);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cam = (lane >> 3) + 4 * h;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      int off = 0, mask = 0;
      if (pv && cam < P.ncam) lift_setup<ARITH>(P, P.cams + ((int64_t)b * P.ncam + cam) * 20, px, py, pz, w, off, mask);
      s_w[warp][pi * kLiftCams + cam] = w;
      s_om[warp][pi * kLiftCams + cam] = make_int2(off, mask ? (mask | ((b * P.ncam + cam) << 4)) : 0);
    }
  }
  __syncwarp();
  const int64_t map4 = (int64_t)P.Hf * P.Wf * C4;
  const int WC4 = P.Wf * C4;
  float4* gf = reinterpret_cast<float4*>(gfeats);
  for (int pi = 0; pi < kLiftPts; ++pi) {
    const int64_t p = p0 + pi;
    if (p >= P.n_total) break;
    const float4* grow = reinterpret_cast<const float4*>(gout + p * P.Cf);
    for (int cam = 0; cam < P.ncam; ++cam) {
      const int2 om = s_om[warp][pi * kLiftCams + cam];
      if (om.y == 0) continue;  // warp-uniform
      const float4 w = s_w[warp][pi * kLiftCams + cam];
      for (int c4 = lane; c4 < C4; c4 += 32)
        scatter_taps(gf + (int64_t)(om.y >> 4) * map4 + c4, om.x, C4, WC4, w, om.y & 15, __ldg(grow + c4));
    }
  }
}

}  // namespace tp

using namespace tp;

static int lift_params(LiftParams& P, const char* who, const float* points, int32_t point_stride, int64_t n_total,
                       const int64_t* offsets, int32_t batch, int32_t ncam, int32_t Hf, int32_t Wf, int32_t Cf,
                       const float* cams, float resize_dim0, float resize_dim1) {
  if (n_total < 0 || batch <= 0 || point_stride < 3) return fail(TP_E_SHAPE, "%s: bad N=%lld B=%d stride=%d", who, (long long)n_total, batch, point_stride);
  if (ncam <= 0 || ncam > kLiftCams) return fail(TP_E_SHAPE, "%s: ncam=%d must be in 1..%d", who, ncam, kLiftCams);
  if (Hf <= 0 || Wf <= 0 || Cf <= 0 || (Cf & 3)) return fail(TP_E_SHAPE, "%s: bad feature map %dx%dx%d (Cf %% 4 == 0)", who, Hf, Wf, Cf);
  if ((int64_t)batch * ncam >= (1 << 27) || (int64_t)Hf * Wf * Cf >= ((int64_t)1 << 31))
    return fail(TP_E_SHAPE, "%s: feature maps too large", who);
  P.points = points; P.offsets = offsets; P.cams = cams;
  P.n_total = n_total; P.point_stride = point_stride; P.batch = batch; P.ncam = ncam;
  P.Hf = Hf; P.Wf = Wf; P.Cf = Cf;
  P.R0 = resize_dim0; P.R1 = resize_dim1;
  P.half0 = (float)((double)resize_dim0 / 2.0); P.half1 = (float)((double)resize_dim1 / 2.0);
  P.rcp0 = 1.0f / resize_dim0; P.rcp1 = 1.0f / resize_dim1;
  return 0;
}

extern "C" int tp_lift_cam_backward_f32(const float* points, int32_t point_stride, int64_t n_total,
                                        const int64_t* offsets, int32_t batch, int32_t ncam, int32_t Hf,
                                        int32_t Wf, int32_t Cf, const float* cams, float resize_dim0,
                                        float resize_dim1, int32_t arith, const float* grad_out,
                                        float* grad_feats_nhwc, void* stream) {
  LiftParams P;
  if (int rc = lift_params(P, "tp_lift_cam_backward_f32", points, point_stride, n_total, offsets, batch, ncam, Hf, Wf,
                           Cf, cams, resize_dim0, resize_dim1))
    return rc;
  if (n_total == 0) return 0;
  if (!points || !offsets || !cams || !grad_out || !grad_feats_nhwc) return fail(TP_E_NULL, "tp_lift_cam_backward_f32: null argument");
  if (((uintptr_t)grad_feats_nhwc & 15) || ((uintptr_t)grad_out & 15)) return fail(TP_E_SHAPE, "tp_lift_cam_backward_f32: buffers not 16-byte aligned");
  P.feats = nullptr; P.out = nullptr;
  const int64_t tiles = (n_total + kLiftPts - 1) / kLiftPts;
  const int64_t ctas = (tiles + kLiftWarps - 1) / kLiftWarps;
  if (ctas >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_lift_cam_backward_f32: too many points");
  cudaStream_t s = (cudaStream_t)stream;
  switch (arith) {
    case TP_ARITH_TORCH_CUDA: lift_backward_kernel<TP_ARITH_TORCH_CUDA><<<(unsigned)ctas, kLiftWarps * 32, 0, s>>>(P, grad_out, grad_feats_nhwc); break;
    case TP_ARITH_TORCH_CPU:  lift_backward_kernel<TP_ARITH_TORCH_CPU><<<(unsigned)ctas, kLiftWarps * 32, 0, s>>>(P, grad_out, grad_feats_nhwc); break;
    default: return fail(TP_E_ENUM, "tp_lift_cam_backward_f32: unknown arith %d", arith);
  }
  TP_LAUNCH_CHECK("lift_backward_kernel");
  return 0;
}

extern "C" int tp_lift_cam_f32(const float* points, int32_t point_stride, int64_t n_total,
                               const int64_t* offsets, int32_t batch, const float* feats_nhwc,
                               int32_t ncam, int32_t Hf, int32_t Wf, int32_t Cf, const float* cams,
                               float resize_dim0, float resize_dim1, int32_t arith, float* out,
                               void* stream) {
  LiftParams P;
  if (int rc = lift_params(P, "tp_lift_cam_f32", points, point_stride, n_total, offsets, batch, ncam, Hf, Wf, Cf, cams,
                           resize_dim0, resize_dim1))
    return rc;
  if (n_total == 0) return 0;
  if (!points || !offsets || !feats_nhwc || !cams || !out) return fail(TP_E_NULL, "tp_lift_cam_f32: null argument");
  if (((uintptr_t)feats_nhwc & 15) || ((uintptr_t)out & 15)) return fail(TP_E_SHAPE, "tp_lift_cam_f32: feats / out not 16-byte aligned");
  P.feats = feats_nhwc; P.out = out;
  const int64_t tiles = (n_total + kLiftPts - 1) / kLiftPts;
  const int64_t ctas = (tiles + kLiftWarps - 1) / kLiftWarps;
  if (ctas >= ((int64_t)1 << 31)) return fail(TP_E_SHAPE, "tp_lift_cam_f32: too many points");
  cudaStream_t s = (cudaStream_t)stream;
  switch (arith) {
    case TP_ARITH_TORCH_CUDA: lift_kernel<TP_ARITH_TORCH_CUDA><<<(unsigned)ctas, kLiftWarps * 32, 0, s>>>(P); break;
    case TP_ARITH_TORCH_CPU:  lift_kernel<TP_ARITH_TORCH_CPU><<<(unsigned)ctas, kLiftWarps * 32, 0, s>>>(P); break;
    default: return fail(TP_E_ENUM, "tp_lift_cam_f32: unknown arith %d", arith);
  }
  TP_LAUNCH_CHECK("lift_kernel");
  return 0;
}
