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

#include "tp_sample_dev.cuh"

namespace tp {

constexpr int kWarpsPerCta = 4;

// Dynamic tile scheduler state: a ring of (next tile, finished warps) pairs; the host picks a slot
// per launch, the last warp of a launch resets it. In-range and out-of-range tiles differ ~5x in
// cost, so a static split leaves SMs idle.
constexpr int kSchedSlots = 1024;
__device__ unsigned int g_sched[kSchedSlots][2];  // [next tile, finished warps]

// C4T: C/4 known at compile time (8, 24, 32 for the reference's C = 32, 96, 128) or 0 = runtime.
template <int ARITH, int C4T>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8)
sample3_kernel(const SampleParams P, const int sched_slot) {
  __shared__ __align__(16) float s_param[kWarpsPerCta][kParamWords];
  __shared__ __align__(16) float s_tile[kWarpsPerCta][kTileWords];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
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
    const int64_t left = P.Q - q0;
    const int n0 = (int)(left < 16 ? left : 16);
    const int n1 = (int)(left < 16 ? 0 : (left < 32 ? left - 16 : 16));
    sample_tile<ARITH, C4T>(P, b, q0, q0 + 16, n0, n1, s_param[warp], s_tile[warp], pol_planes, pol_out,
                            q_vec4);
  }
  // Last warp out resets the scheduler slot for the next launch that draws it. No fence: the draw that ended this
  // warp's loop has returned (its value was consumed), so every draw precedes its warp's count-out, and the reset only
  // has to be visible to the next launch (ncu: the fence, an ERRBAR behind every outstanding store of the warp, and the
  // wait on it were 8 % of this kernel's stall samples on the range-image points).
  if (lane == 0 && atomicAdd(&sched[1], 1u) == gridDim.x * kWarpsPerCta - 1) {
    sched[0] = 0;
    sched[1] = 0;
  }
}

// NCHW -> NHWC through a padded 32x32 shared tile. grid = (ceil(HW/32), ceil(C/32), B)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, int64_t src_bstride, float* __restrict__ dst,
                    int C, int HW, int vec) {
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
  if (vec) {  // C % 4 == 0, 16-byte aligned destination: one 16-byte store per thread
    const int pl = threadIdx.x >> 3, l8 = threadIdx.x & 7, p = p0 + pl, c = c0 + 4 * l8;
    if (c < C && p < HW)
      *reinterpret_cast<float4*>(d + (int64_t)p * C + c) =
          make_float4(tile[4 * l8][pl], tile[4 * l8 + 1][pl], tile[4 * l8 + 2][pl], tile[4 * l8 + 3][pl]);
    return;
  }
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
  int vec;         // 16-byte stores possible
};

// all three planes in one launch. grid = (sum_p ceil(HW_p/32), ceil(C/32), B)
// The first instruction releases a dependent decode launch (tp_common.cuh). A round-2 rewrite on slabs of 32 channels x
// 128 pixels (16 loads in flight per thread, 16-byte stores, one wave of 384 CTAs) was SLOWER on the same box — 5.5 vs
// 4.5 us for one sample's planes, 24.4 vs 22.4 us at bs = 8: with every CTA resident at once the launch is a read
// phase followed by a write phase, while the 1 536 small CTAs of this kernel (1.3 waves) overlap the two.
__global__ void __launch_bounds__(256)
nchw_to_nhwc3_kernel(const Planes3 P, int C) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
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
  if (P.vec) {  // C % 4 == 0, 16-byte aligned destination: one 16-byte store per thread (pixel tid >> 3, channels 4 (tid & 7) ..)
    const int pl = threadIdx.x >> 3, l8 = threadIdx.x & 7, p = p0 + pl, c = c0 + 4 * l8;
    if (c < C && p < HW)
      *reinterpret_cast<float4*>(d + (int64_t)p * C + c) =
          make_float4(tile[4 * l8][pl], tile[4 * l8 + 1][pl], tile[4 * l8 + 2][pl], tile[4 * l8 + 3][pl]);
    return;
  }
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
  P.vec = (C & 3) == 0;
  for (int k = 0; k < 3; ++k) P.vec = P.vec && ((uintptr_t)dst[k] & 15) == 0;
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
  const int vec = ((C & 3) == 0 && ((uintptr_t)dst & 15) == 0) ? 1 : 0;
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_batch_stride, dst, C, HW, vec);
  TP_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

int tp::sample3_flat(const tp_plane planes[3], int32_t C, const float* queries, int64_t Q, int32_t batch,
                     const tp_sample_geom* sg, int32_t arith, float* out, void* stream, bool pdl) {
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
#define TP_SAMPLE(A, C4T) launch_kernel(sample3_kernel<A, C4T>, grid, kWarpsPerCta * 32, 0, s, pdl, P, slot)
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

extern "C" int tp_sample3_nhwc_f32(const tp_plane planes[3], int32_t C, const float* queries,
                                   int64_t Q, int32_t batch, const tp_sample_geom* sg,
                                   int32_t arith, float* out, void* stream) {
  return sample3_flat(planes, C, queries, Q, batch, sg, arith, out, stream, false);
}

// channels-last copies of reference-layout planes live in the caller's workspace, plane after plane
int tp::planes3_workspace(const char* who, const tp_plane planes_nchw[3], int32_t C, int32_t batch, float* ws,
                          int64_t ws_floats, tp_plane nhwc[3], float* dsts[3]) {
  if (!planes_nchw || !ws) return fail(TP_E_NULL, "%s: null argument", who);
  if (batch <= 0 || C <= 0) return fail(TP_E_SHAPE, "%s: bad shape B=%d C=%d", who, batch, C);
  int64_t need = 0;
  for (int k = 0; k < 3; ++k) need += (int64_t)batch * C * planes_nchw[k].H * planes_nchw[k].W;
  if (ws_floats < need)
    return fail(TP_E_WORKSPACE, "%s: workspace %lld < %lld floats", who, (long long)ws_floats, (long long)need);
  float* w = ws;
  for (int k = 0; k < 3; ++k) {
    dsts[k] = w;
    nhwc[k].data = w;
    nhwc[k].H = planes_nchw[k].H;
    nhwc[k].W = planes_nchw[k].W;
    nhwc[k].batch_stride = (int64_t)C * planes_nchw[k].H * planes_nchw[k].W;
    w += (int64_t)batch * nhwc[k].batch_stride;
  }
  return 0;
}

// ... converted by one launch that releases its dependent (the decode launch that follows)
int tp::planes3_to_workspace(const char* who, const tp_plane planes_nchw[3], int32_t C, int32_t batch, float* ws,
                             int64_t ws_floats, tp_plane nhwc[3], void* stream) {
  float* dsts[3];
  if (int rc = planes3_workspace(who, planes_nchw, C, batch, ws, ws_floats, nhwc, dsts)) return rc;
  return tp_planes3_nchw_to_nhwc_f32(planes_nchw, dsts, batch, C, stream);
}

extern "C" int tp_sample3_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const float* queries,
                                   int64_t Q, int32_t batch, const tp_sample_geom* sg,
                                   int32_t arith, float* out, float* ws, int64_t ws_floats,
                                   void* stream) {
  tp_plane nhwc[3];
  if (int rc = planes3_to_workspace("tp_sample3_nchw_f32", planes_nchw, C, batch, ws, ws_floats, nhwc, stream)) return rc;
  return sample3_flat(nhwc, C, queries, Q, batch, sg, arith, out, stream, true);
}
