// Gathers / scatters either side of the triplane decode. None of them is arithmetic-heavy; all of
// them are Python loops over (sample, camera) with boolean-mask compaction and ~15 small kernels per iteration in
// the reference, and all of them use nearest-pixel (truncating) indices, i.e. integer work:
//
//  * point / range-pixel features -> camera-image pixels
//        TriplaneMAE.forward            triplane.py:381-390      (coordinates from JointEncoder.interact)
//        PointTriplane.cam_rec_feat     point_triplane.py:243-309 (coordinates projected here)
//    `img[:, rows, cols] = feat[:, valid]` is an index_put with duplicate targets: undefined order on torch-CUDA,
//    last-in-list-order on torch-CPU. Here the order is DEFINED: the source with the highest index wins (what
//    torch-CPU does), via an int32 winner image (atomicMax) followed by a write-once gather of the dense output.
//  * JointEncoder.interact              joint_encoder.py:97-215: projection of the range image into every
//    camera, range_cam_coors, nearest-pixel gather of image features into the range image (sum over cameras),
//    and the index-put of the position embedding into the image features (same winner rule).
//  * InterpNet's neighbourhood search   interpnet.py:65 (torch_geometric.nn.radius -> torch_cluster.radius,
//    r = 1.0, max_num_neighbors = 32; un-vendored third-party op: first 32 sources in index order whose squared
//    distance is < r^2, per query, same sample only — restated from torch_cluster's CUDA kernel).
#include "tp_cam.cuh"

namespace tp {

__device__ __forceinline__ int64_t f2long(float v) { return (int64_t)v; }  // Tensor.long(): truncation toward zero

// ---------------------------------------------------------------------------------------------------------------
// winner images
// ---------------------------------------------------------------------------------------------------------------
// coors [M, npix, 2] fp32 (row, col), -1 = none  ->  winner [M, H, W] = max source pixel index that lands there.
// valid = long(row) > 0 (triplane.py:386: `cam_coors[..., 0] > 0` AFTER .long(); row 0 is dropped by the
// reference and therefore here). Targets outside [0,H) x [0,W) would be an index error in the reference: dropped.
__global__ void __launch_bounds__(256)
winner_from_coors_kernel(const float2* __restrict__ coors, int64_t total, int npix, int H, int W,
                         int* __restrict__ winner) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const float2 rc = __ldg(coors + t);
  const int64_t r = f2long(rc.x), c = f2long(rc.y);
  if (r > 0 && r < H && c >= 0 && c < W) {
    const int64_t m = t / npix;
    atomicMax(winner + (m * H + r) * W + c, (int)(t - m * npix));
  }
}

struct PointPixelParams {
  const float* points;     // [N, point_stride]
  const int64_t* offsets;  // [B+1]
  const float* cams;       // [B, ncam, 20]
  int* winner;             // [B, ncam, H, W], H = R0, W = R1
  int64_t n_total;
  int point_stride, batch, ncam, H, W;
  float R0, R1, half0, half1;
};

// cam_rec_feat (point_triplane.py:263-307): project every point into every camera; pixel = (long(y), long(x)).
// winner holds the point's index INSIDE its sample (points_feat is per sample in the reference).
__global__ void __launch_bounds__(256)
winner_from_points_kernel(const PointPixelParams P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n_total * P.ncam) return;
  const int64_t p = t / P.ncam;
  const int cam = (int)(t - p * P.ncam);
  const float* pp = P.points + p * P.point_stride;
  const int b = tp_find_batch(P.offsets, P.batch, p);
  const CamPixel cp = cam_project(P.cams + ((int64_t)b * P.ncam + cam) * 20, __ldg(pp), __ldg(pp + 1), __ldg(pp + 2),
                                  P.R0, P.R1, P.half0, P.half1);
  if (!cp.valid) return;
  const int64_t r = f2long(cp.y), c = f2long(cp.x);
  if (r >= 0 && r < P.H && c >= 0 && c < P.W)
    atomicMax(P.winner + (((int64_t)b * P.ncam + cam) * P.H + r) * P.W + c, (int)(p - __ldg(P.offsets + b)));
}

// out [M, C, H, W] = feat[m / imgs_per_feat][c][winner] or 0. feat addressed with explicit strides so that both the
// channel-major decode output [B, C, N] and point-major rows [N, C] (+ per-sample row offsets) can be the source.
struct WinnerGatherParams {
  const int* winner;  // [M, H*W]
  const float* feat;
  float* out;         // [M, C, H*W]
  const int64_t* feat_row0;  // optional [M / imgs_per_feat + 1]: first source row of each feature batch (point-major)
  int64_t feat_bstride, feat_cstride, feat_nstride;
  int64_t HW;
  int M, C, imgs_per_feat;
};

__global__ void __launch_bounds__(256)
winner_gather_kernel(const WinnerGatherParams P) {
  // one thread: 4 consecutive pixels of one image, all channels; 16-byte streaming stores
  const int64_t quads = (P.HW + 3) >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= quads * P.M) return;
  const int m = (int)(t / quads);
  const int64_t p0 = (t - (int64_t)m * quads) * 4;
  const int fb = m / P.imgs_per_feat;
  const float* f = P.feat + (int64_t)fb * P.feat_bstride + (P.feat_row0 ? __ldg(P.feat_row0 + fb) * P.feat_nstride : 0);
  int w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = (p0 + k < P.HW) ? __ldg(P.winner + (int64_t)m * P.HW + p0 + k) : -1;
  float* o = P.out + (int64_t)m * P.C * P.HW + p0;
  const bool vec = ((P.HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
  const bool any = (w[0] & w[1] & w[2] & w[3]) >= 0;  // some winner is non-negative
  for (int c = 0; c < P.C; ++c, o += P.HW) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (any) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (w[k] >= 0) v[k] = __ldg(f + (int64_t)c * P.feat_cstride + (int64_t)w[k] * P.feat_nstride);
    }
    if (vec) {
      st_cs_f4(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (p0 + k < P.HW) st_cs_f1(o + k, v[k]);
    }
  }
}

// backward of the winner gather w.r.t. feat: gfeat[fb][c][winner] += gout[m][c][pixel] (a source can win one pixel
// per image, so at most imgs_per_feat addends meet: atomics)
__global__ void __launch_bounds__(256)
winner_gather_backward_kernel(const WinnerGatherParams P, const float* __restrict__ gout, float* __restrict__ gfeat) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.HW * P.M) return;
  const int m = (int)(t / P.HW);
  const int64_t p = t - (int64_t)m * P.HW;
  const int w = __ldg(P.winner + t);
  if (w < 0) return;
  const int fb = m / P.imgs_per_feat;
  float* g = gfeat + (int64_t)fb * P.feat_bstride + (P.feat_row0 ? __ldg(P.feat_row0 + fb) * P.feat_nstride : 0) +
             (int64_t)w * P.feat_nstride;
  const float* go = gout + (int64_t)m * P.C * P.HW + p;
  for (int c = 0; c < P.C; ++c) atomicAdd(g + (int64_t)c * P.feat_cstride, __ldg(go + (int64_t)c * P.HW));
}

// ---------------------------------------------------------------------------------------------------------------
// JointEncoder.interact (joint_encoder.py:97-215)
// ---------------------------------------------------------------------------------------------------------------
struct RangeProjectParams {
  const float* range_points;  // [B, npix, 3]
  const float* range_image;   // [B, npix]  (the masked range image; > 0 = unmasked pixel)
  const float* cams;          // [B, ncam, 20]
  float* coors;               // [B, ncam, npix, 2]  (row, col) or -1
  int* fidx;                  // [B, ncam, npix]     feature-map pixel row * Wf + col, or -1
  int* winner;                // [B, ncam, Hf * Wf]  highest range pixel that lands on a feature pixel, pre-set to -1
  int64_t npix;
  int batch, ncam, Hf, Wf;
  float R0, R1, half0, half1, rcp0, rcp1;
};

template <int ARITH>
__global__ void __launch_bounds__(256)
range_project_kernel(const RangeProjectParams P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)P.batch * P.npix) return;
  const int b = (int)(t / P.npix);
  const int64_t pix = t - (int64_t)b * P.npix;
  const float* pp = P.range_points + t * 3;
  const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  // "pixels that contain no point": all three coordinates == 0 (:139-141)
  const bool has_point = !((px == 0.f) & (py == 0.f) & (pz == 0.f));
  const bool unmasked = __ldg(P.range_image + t) > 0.f;  // batch_range_mask (:135)
  for (int cam = 0; cam < P.ncam; ++cam) {
    const int64_t o = ((int64_t)b * P.ncam + cam) * P.npix + pix;
    float2 rc = make_float2(-1.f, -1.f);
    int fi = -1;
    if (has_point) {
      const CamPixel cp = cam_project(P.cams + ((int64_t)b * P.ncam + cam) * 20, px, py, pz, P.R0, P.R1, P.half0, P.half1);
      if (cp.valid) {
        rc = make_float2(cp.y, cp.x);  // swapped to (row, col) (:186-187)
        if (unmasked) {
          // valid_coor[:, 0] * Hf / R0, valid_coor[:, 1] * Wf / R1, .long()  (:203-205)
          const float fr = tp_div<ARITH>(__fmul_rn(cp.y, (float)P.Hf), P.R0, P.rcp0);
          const float fc = tp_div<ARITH>(__fmul_rn(cp.x, (float)P.Wf), P.R1, P.rcp1);
          const int64_t r = f2long(fr), c = f2long(fc);
          if (r >= 0 && r < P.Hf && c >= 0 && c < P.Wf) {  // rounding up to Hf / Wf would be an index error there
            fi = (int)(r * P.Wf + c);
            atomicMax(P.winner + ((int64_t)b * P.ncam + cam) * P.Hf * P.Wf + fi, (int)pix);
          }
        }
      }
    }
    reinterpret_cast<float2*>(P.coors)[o] = rc;
    P.fidx[o] = fi;
  }
}

// cam_range_features[b, :, pix] = sum over cameras (ascending, from zero) of img_features[b, cam, :, fidx] (:208)
struct RangeGatherParams {
  const int* fidx;   // [B, ncam, npix]
  const float* img;  // [B, ncam, C, Hf*Wf]
  float* out;        // [B, C, npix]
  int64_t npix;
  int batch, ncam, C, HWf;
};

__global__ void __launch_bounds__(256)
range_gather_kernel(const RangeGatherParams P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)P.batch * P.npix) return;
  const int b = (int)(t / P.npix);
  const int64_t pix = t - (int64_t)b * P.npix;
  int fi[8];
  bool any = false;
#pragma unroll
  for (int cam = 0; cam < 8; ++cam) {
    fi[cam] = cam < P.ncam ? __ldg(P.fidx + ((int64_t)b * P.ncam + cam) * P.npix + pix) : -1;
    any |= fi[cam] >= 0;
  }
  float* o = P.out + (int64_t)b * P.C * P.npix + pix;
  const float* img = P.img + (int64_t)b * P.ncam * P.C * P.HWf;
  for (int c = 0; c < P.C; ++c, o += P.npix) {
    float acc = 0.f;
    if (any) {
#pragma unroll
      for (int cam = 0; cam < 8; ++cam)
        if (fi[cam] >= 0) acc = __fadd_rn(acc, __ldg(img + ((int64_t)cam * P.C + c) * P.HWf + fi[cam]));
    }
    st_cs_f1(o, acc);
  }
}

__global__ void __launch_bounds__(256)
range_gather_backward_kernel(const RangeGatherParams P, const float* __restrict__ gout, float* __restrict__ gimg) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)P.batch * P.npix) return;
  const int b = (int)(t / P.npix);
  const int64_t pix = t - (int64_t)b * P.npix;
  const float* go = gout + (int64_t)b * P.C * P.npix + pix;
  float* gi = gimg + (int64_t)b * P.ncam * P.C * P.HWf;
  for (int cam = 0; cam < P.ncam; ++cam) {
    const int fi = __ldg(P.fidx + ((int64_t)b * P.ncam + cam) * P.npix + pix);
    if (fi < 0) continue;
    for (int c = 0; c < P.C; ++c) atomicAdd(gi + ((int64_t)cam * P.C + c) * P.HWf + fi, __ldg(go + (int64_t)c * P.npix));
  }
}

// img[m, c, p] (+)= pe[m, p, c] where winner[m, p] >= 0 (:212-213; m = b * ncam + cam, p = feature pixel).
// FWD: in place on img. !FWD (backward w.r.t. pe): gpe[m, p, c] = winner >= 0 ? gimg[m, c, p] : 0.
template <bool FWD>
__global__ void __launch_bounds__(256)
posembed_kernel(const int* __restrict__ winner, float* __restrict__ img, float* __restrict__ pe, int64_t M, int C, int HWf) {
  __shared__ float tile[32][33];
  // 32 pixels x 32 channels through shared memory: both sides coalesced
  const int64_t m = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float* im = img + m * C * HWf;
  float* pm = pe + m * (int64_t)HWf * C;
  const int* wm = winner + m * HWf;
  if (FWD) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int p = p0 + ty + j, c = c0 + tx;
      tile[ty + j][tx] = (p < HWf && c < C) ? pm[(int64_t)p * C + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int c = c0 + ty + j, p = p0 + tx;
      if (p < HWf && c < C && __ldg(wm + p) >= 0) im[(int64_t)c * HWf + p] = __fadd_rn(im[(int64_t)c * HWf + p], tile[tx][ty + j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int c = c0 + ty + j, p = p0 + tx;
      tile[ty + j][tx] = (p < HWf && c < C && __ldg(wm + p) >= 0) ? im[(int64_t)c * HWf + p] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int p = p0 + ty + j, c = c0 + tx;
      if (p < HWf && c < C) pm[(int64_t)p * C + c] = tile[tx][ty + j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// radius search
// ---------------------------------------------------------------------------------------------------------------
struct RadiusParams {
  const float* x;            // [Nx, 3] sources
  const int64_t* x_offsets;  // [B+1]
  const float* y;            // [Ny, 3] queries
  const int64_t* y_offsets;  // [B+1]
  int* col;                  // [Ny, max_nbr] source indices (global), -1 padded
  int* cnt;                  // [Ny]
  int64_t ny;
  int batch, max_nbr;
  float r2;
};

// one warp per query: 32 sources per step in index order, ballot, first (max_nbr - count) hits in lane order —
// exactly the sequence a single thread walking the sources in order would emit
__global__ void __launch_bounds__(128)
radius_kernel(const RadiusParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= P.ny) return;
  const int b = tp_find_batch(P.y_offsets, P.batch, q);
  const float qx = __ldg(P.y + q * 3), qy = __ldg(P.y + q * 3 + 1), qz = __ldg(P.y + q * 3 + 2);
  const int64_t x0 = __ldg(P.x_offsets + b), x1 = __ldg(P.x_offsets + b + 1);
  int count = 0;
  int* crow = P.col + q * P.max_nbr;
  for (int64_t s = x0; s < x1 && count < P.max_nbr; s += 32) {
    const int64_t i = s + lane;
    bool hit = false;
    if (i < x1) {
      const float dx = __fsub_rn(__ldg(P.x + i * 3), qx), dy = __fsub_rn(__ldg(P.x + i * 3 + 1), qy),
                  dz = __fsub_rn(__ldg(P.x + i * 3 + 2), qz);
      const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      hit = d2 < P.r2;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int rank = __popc(bal & ((1u << lane) - 1));
    if (hit && count + rank < P.max_nbr) crow[count + rank] = (int)i;
    count += __popc(bal);
  }
  count = min(count, P.max_nbr);
  for (int k = count + lane; k < P.max_nbr; k += 32) crow[k] = -1;
  if (lane == 0) P.cnt[q] = count;
}

}  // namespace tp

using namespace tp;

static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

extern "C" int tp_pixel_winner_coors_i32(const float* coors, int64_t n_images, int64_t npix, int32_t H, int32_t W,
                                         int32_t* winner, void* stream) {
  if (n_images < 0 || npix < 0 || H <= 0 || W <= 0 || npix >= ((int64_t)1 << 31))
    return fail(TP_E_SHAPE, "tp_pixel_winner_coors_i32: bad shape M=%lld npix=%lld H=%d W=%d", (long long)n_images, (long long)npix, H, W);
  if (n_images == 0) return 0;
  if (!winner || (npix && !coors)) return fail(TP_E_NULL, "tp_pixel_winner_coors_i32: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  TP_CUDA(cudaMemsetAsync(winner, 0xFF, (size_t)n_images * H * W * 4, s));
  const int64_t total = n_images * npix;
  if (total == 0) return 0;
  if (total >= ((int64_t)1 << 39)) return fail(TP_E_SHAPE, "tp_pixel_winner_coors_i32: too many coordinates");
  winner_from_coors_kernel<<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float2*>(coors), total, (int)npix, H, W, winner);
  TP_LAUNCH_CHECK("winner_from_coors_kernel");
  return 0;
}

extern "C" int tp_pixel_winner_points_i32(const float* points, int32_t point_stride, int64_t n_total,
                                          const int64_t* offsets, int32_t batch, const float* cams, int32_t ncam,
                                          float resize_dim0, float resize_dim1, int32_t* winner, void* stream) {
  if (batch <= 0 || ncam <= 0 || ncam > 8 || n_total < 0 || point_stride < 3 || !(resize_dim0 >= 1.f) || !(resize_dim1 >= 1.f))
    return fail(TP_E_SHAPE, "tp_pixel_winner_points_i32: bad shape B=%d ncam=%d N=%lld", batch, ncam, (long long)n_total);
  if (!winner || !offsets || !cams || (n_total && !points)) return fail(TP_E_NULL, "tp_pixel_winner_points_i32: null argument");
  PointPixelParams P;
  P.points = points; P.offsets = offsets; P.cams = cams; P.winner = winner;
  P.n_total = n_total; P.point_stride = point_stride; P.batch = batch; P.ncam = ncam;
  P.H = (int)resize_dim0; P.W = (int)resize_dim1;
  P.R0 = resize_dim0; P.R1 = resize_dim1; P.half0 = resize_dim0 / 2.0f; P.half1 = resize_dim1 / 2.0f;
  cudaStream_t s = (cudaStream_t)stream;
  TP_CUDA(cudaMemsetAsync(winner, 0xFF, (size_t)batch * ncam * P.H * P.W * 4, s));
  if (n_total == 0) return 0;
  winner_from_points_kernel<<<blocks_for(n_total * ncam, 256), 256, 0, s>>>(P);
  TP_LAUNCH_CHECK("winner_from_points_kernel");
  return 0;
}

static int winner_params(WinnerGatherParams& P, const char* who, const int32_t* winner, int64_t n_images, int64_t HW,
                         int32_t C, int32_t imgs_per_feat, const float* feat, int64_t feat_bstride,
                         int64_t feat_cstride, int64_t feat_nstride, const int64_t* feat_row0, const float* io) {
  if (n_images < 0 || HW <= 0 || C <= 0 || imgs_per_feat <= 0 || n_images >= ((int64_t)1 << 31))
    return fail(TP_E_SHAPE, "%s: bad shape M=%lld HW=%lld C=%d", who, (long long)n_images, (long long)HW, C);
  if (n_images && (!winner || !feat || !io)) return fail(TP_E_NULL, "%s: null argument", who);
  P.winner = winner; P.feat = feat; P.out = const_cast<float*>(io); P.feat_row0 = feat_row0;
  P.feat_bstride = feat_bstride; P.feat_cstride = feat_cstride; P.feat_nstride = feat_nstride;
  P.HW = HW; P.M = (int)n_images; P.C = C; P.imgs_per_feat = imgs_per_feat;
  return 0;
}

extern "C" int tp_winner_gather_f32(const int32_t* winner, int64_t n_images, int64_t HW, int32_t C,
                                    int32_t imgs_per_feat, const float* feat, int64_t feat_bstride,
                                    int64_t feat_cstride, int64_t feat_nstride, const int64_t* feat_row0, float* out,
                                    void* stream) {
  WinnerGatherParams P;
  if (int rc = winner_params(P, "tp_winner_gather_f32", winner, n_images, HW, C, imgs_per_feat, feat, feat_bstride,
                             feat_cstride, feat_nstride, feat_row0, out))
    return rc;
  if (n_images == 0) return 0;
  winner_gather_kernel<<<blocks_for(((HW + 3) / 4) * n_images, 256), 256, 0, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("winner_gather_kernel");
  return 0;
}

extern "C" int tp_winner_gather_backward_f32(const int32_t* winner, int64_t n_images, int64_t HW, int32_t C,
                                             int32_t imgs_per_feat, const float* grad_out, float* grad_feat,
                                             int64_t feat_bstride, int64_t feat_cstride, int64_t feat_nstride,
                                             const int64_t* feat_row0, void* stream) {
  WinnerGatherParams P;
  if (int rc = winner_params(P, "tp_winner_gather_backward_f32", winner, n_images, HW, C, imgs_per_feat, grad_feat,
                             feat_bstride, feat_cstride, feat_nstride, feat_row0, grad_out))
    return rc;
  if (n_images == 0) return 0;
  winner_gather_backward_kernel<<<blocks_for(HW * n_images, 256), 256, 0, (cudaStream_t)stream>>>(P, grad_out, grad_feat);
  TP_LAUNCH_CHECK("winner_gather_backward_kernel");
  return 0;
}

extern "C" int tp_range_project_f32(const float* range_points, const float* range_image, int64_t npix, int32_t batch,
                                    const float* cams, int32_t ncam, float resize_dim0, float resize_dim1, int32_t Hf,
                                    int32_t Wf, int32_t arith, float* coors, int32_t* fidx, int32_t* winner,
                                    void* stream) {
  if (batch <= 0 || ncam <= 0 || ncam > 8 || npix < 0 || npix >= ((int64_t)1 << 31) || Hf <= 0 || Wf <= 0 ||
      !(resize_dim0 >= 1.f) || !(resize_dim1 >= 1.f))
    return fail(TP_E_SHAPE, "tp_range_project_f32: bad shape B=%d ncam=%d npix=%lld Hf=%d Wf=%d", batch, ncam, (long long)npix, Hf, Wf);
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_range_project_f32: unknown arith %d", arith);
  if (!cams || !winner || (npix && (!range_points || !range_image || !coors || !fidx)))
    return fail(TP_E_NULL, "tp_range_project_f32: null argument");
  RangeProjectParams P;
  P.range_points = range_points; P.range_image = range_image; P.cams = cams; P.coors = coors; P.fidx = fidx;
  P.winner = winner; P.npix = npix; P.batch = batch; P.ncam = ncam; P.Hf = Hf; P.Wf = Wf;
  P.R0 = resize_dim0; P.R1 = resize_dim1; P.half0 = resize_dim0 / 2.0f; P.half1 = resize_dim1 / 2.0f;
  P.rcp0 = 1.0f / resize_dim0; P.rcp1 = 1.0f / resize_dim1;
  cudaStream_t s = (cudaStream_t)stream;
  TP_CUDA(cudaMemsetAsync(winner, 0xFF, (size_t)batch * ncam * Hf * Wf * 4, s));
  if (npix == 0) return 0;
  if (arith == TP_ARITH_TORCH_CUDA) range_project_kernel<TP_ARITH_TORCH_CUDA><<<blocks_for((int64_t)batch * npix, 256), 256, 0, s>>>(P);
  else range_project_kernel<TP_ARITH_TORCH_CPU><<<blocks_for((int64_t)batch * npix, 256), 256, 0, s>>>(P);
  TP_LAUNCH_CHECK("range_project_kernel");
  return 0;
}

static int range_gather_params(RangeGatherParams& P, const char* who, const int32_t* fidx, int64_t npix, int32_t batch,
                               int32_t ncam, int32_t C, int32_t HWf, const float* img, const float* io) {
  if (batch <= 0 || ncam <= 0 || ncam > 8 || C <= 0 || HWf <= 0 || npix < 0)
    return fail(TP_E_SHAPE, "%s: bad shape B=%d ncam=%d C=%d", who, batch, ncam, C);
  if (npix && (!fidx || !img || !io)) return fail(TP_E_NULL, "%s: null argument", who);
  P.fidx = fidx; P.img = img; P.out = const_cast<float*>(io); P.npix = npix; P.batch = batch; P.ncam = ncam; P.C = C; P.HWf = HWf;
  return 0;
}

extern "C" int tp_range_gather_f32(const int32_t* fidx, int64_t npix, int32_t batch, int32_t ncam, const float* img_features,
                                   int32_t C, int32_t HWf, float* out, void* stream) {
  RangeGatherParams P;
  if (int rc = range_gather_params(P, "tp_range_gather_f32", fidx, npix, batch, ncam, C, HWf, img_features, out)) return rc;
  if (npix == 0) return 0;
  range_gather_kernel<<<blocks_for((int64_t)batch * npix, 256), 256, 0, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("range_gather_kernel");
  return 0;
}

extern "C" int tp_range_gather_backward_f32(const int32_t* fidx, int64_t npix, int32_t batch, int32_t ncam,
                                            const float* grad_out, int32_t C, int32_t HWf, float* grad_img, void* stream) {
  RangeGatherParams P;
  if (int rc = range_gather_params(P, "tp_range_gather_backward_f32", fidx, npix, batch, ncam, C, HWf, grad_img, grad_out)) return rc;
  if (npix == 0) return 0;
  range_gather_backward_kernel<<<blocks_for((int64_t)batch * npix, 256), 256, 0, (cudaStream_t)stream>>>(P, grad_out, grad_img);
  TP_LAUNCH_CHECK("range_gather_backward_kernel");
  return 0;
}

extern "C" int tp_posembed_scatter_f32(const int32_t* winner, int64_t n_images, int32_t C, int32_t HWf, float* img_features,
                                       const float* pos_embed, void* stream) {
  if (n_images < 0 || n_images > 65535 || C <= 0 || HWf <= 0) return fail(TP_E_SHAPE, "tp_posembed_scatter_f32: bad shape M=%lld C=%d HWf=%d", (long long)n_images, C, HWf);
  if (n_images == 0) return 0;
  if (!winner || !img_features || !pos_embed) return fail(TP_E_NULL, "tp_posembed_scatter_f32: null argument");
  dim3 grid((HWf + 31) / 32, (C + 31) / 32, (unsigned)n_images);
  posembed_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(winner, img_features, const_cast<float*>(pos_embed), n_images, C, HWf);
  TP_LAUNCH_CHECK("posembed_kernel");
  return 0;
}

extern "C" int tp_posembed_scatter_backward_f32(const int32_t* winner, int64_t n_images, int32_t C, int32_t HWf,
                                                const float* grad_img, float* grad_pos_embed, void* stream) {
  if (n_images < 0 || n_images > 65535 || C <= 0 || HWf <= 0) return fail(TP_E_SHAPE, "tp_posembed_scatter_backward_f32: bad shape");
  if (n_images == 0) return 0;
  if (!winner || !grad_img || !grad_pos_embed) return fail(TP_E_NULL, "tp_posembed_scatter_backward_f32: null argument");
  dim3 grid((HWf + 31) / 32, (C + 31) / 32, (unsigned)n_images);
  posembed_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(winner, const_cast<float*>(grad_img), grad_pos_embed, n_images, C, HWf);
  TP_LAUNCH_CHECK("posembed_backward_kernel");
  return 0;
}

extern "C" int tp_radius_i32(const float* x, const int64_t* x_offsets, const float* y, const int64_t* y_offsets,
                             int64_t ny, int32_t batch, float r, int32_t max_num_neighbors, int32_t* col,
                             int32_t* count, void* stream) {
  if (batch <= 0 || ny < 0 || max_num_neighbors <= 0 || !(r >= 0.f)) return fail(TP_E_SHAPE, "tp_radius_i32: bad arguments");
  if (ny == 0) return 0;
  if (!x_offsets || !y || !y_offsets || !col || !count) return fail(TP_E_NULL, "tp_radius_i32: null argument");
  RadiusParams P;
  P.x = x; P.x_offsets = x_offsets; P.y = y; P.y_offsets = y_offsets; P.col = col; P.cnt = count; P.ny = ny;
  P.batch = batch; P.max_nbr = max_num_neighbors; P.r2 = r * r;
  radius_kernel<<<blocks_for(ny, 4), 128, 0, (cudaStream_t)stream>>>(P);
  TP_LAUNCH_CHECK("radius_kernel");
  return 0;
}
