// The reference's LiDAR -> camera-pixel projection chain, op by op. Shared by the camera -> point lift (tp_lift.cu:
// point_triplane.py:190-227), the range-image / camera interaction (tp_project.cu: joint_encoder.py:125-185) and the
// point -> pixel scatter (tp_project.cu: point_triplane.py:263-301): all three write the same Python lines.
#pragma once
#include "tp_common.cuh"

namespace tp {

struct CamPixel {
  float x, y;  // augmented pixel coordinates: x = column (width axis), y = row
  bool valid;  // inside the resized image: (y < R0) & (x < R1) & (y >= 0) & (x >= 0)
};

// cam: 20 floats = lidar2image 4x4 row-major, resize, crop[0], crop[1], flip (0/1).
// R0, R1 = resize_dims = img_shape[::-1]; half0 = R0 / 2.0, half1 = R1 / 2.0.
__device__ __forceinline__ CamPixel cam_project(const float* __restrict__ cam, float px, float py, float pz, float R0,
                                                float R1, float half0, float half1) {
  // einsum("cij,hj->chi"): sum_j M[i][j] * hom[j], fp32 fma chain in j order from 0
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = __fmaf_rn(__ldg(cam + 4 * i), px, 0.f);
    a = __fmaf_rn(__ldg(cam + 4 * i + 1), py, a);
    a = __fmaf_rn(__ldg(cam + 4 * i + 2), pz, a);
    c[i] = __fmaf_rn(__ldg(cam + 4 * i + 3), 1.0f, a);
  }
  const float z = fmaxf(c[2], 1e-5f);  // torch.maximum(z, 1e-5) (NaN propagates below through the compares)
  float x = __fdiv_rn(c[0], z), y = __fdiv_rn(c[1], z);
  const float resize = __ldg(cam + 16), crop_x = __ldg(cam + 17), crop_y = __ldg(cam + 18);
  const bool flip = __ldg(cam + 19) != 0.f;
  x = __fsub_rn(__fmul_rn(x, resize), crop_x);
  y = __fsub_rn(__fmul_rn(y, resize), crop_y);
  if (flip) x = __fsub_rn(R1, x);
  // "-= W/2, rotate by 0, += W/2": the rotation is the identity on finite values, the subtract/add pair is not
  // (it rounds twice) and is replayed
  x = __fadd_rn(__fsub_rn(x, half1), half1);
  y = __fadd_rn(__fsub_rn(y, half0), half0);
  CamPixel r;
  r.x = x;
  r.y = y;
  r.valid = (y < R0) & (x < R1) & (y >= 0.f) & (x >= 0.f) & !isnan(c[2]);
  return r;
}

}  // namespace tp
