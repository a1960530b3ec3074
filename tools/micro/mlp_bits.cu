// Writes the logits of the occupancy-head kernel for a fixed pseudo-random input to argv[1] (raw fp32), to compare
// builds bit for bit (e.g. -DTP_MLP_NOMASK: operands with their 13 low mantissa bits left in place).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 [-DTP_MLP_NOMASK] -I efficient_multimodal_perception_b200/csrc \
//        -I include -o build/micro/mlp_bits tools/micro/mlp_bits.cu
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../efficient_multimodal_perception_b200/csrc/tp_mlp.cu"
namespace tp {
int fail(int code, const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); return code; }
int check_cuda(cudaError_t e, const char* what) { if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -9; } return 0; }
}
static unsigned long long rng = 88172645463325252ull;
static float rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (float)((rng >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0); }
int main(int argc, char** argv) {
  const int64_t Q = 100003;
  std::vector<float> hx(Q * 32), h1(64 * 32), h2(64 * 32), h3(5 * 32), ho(Q * 5);
  for (auto& v : hx) v = rnd() * 3.f;
  for (auto& v : h1) v = rnd() * 0.2f;
  for (auto& v : h2) v = rnd() * 0.15f;
  for (auto& v : h3) v = rnd() * 0.2f;
  float *x, *w1, *w2, *w3, *o;
  cudaMalloc(&x, Q * 32 * 4); cudaMalloc(&o, Q * 5 * 4);
  cudaMalloc(&w1, 8192); cudaMalloc(&w2, 8192); cudaMalloc(&w3, 640);
  cudaMemcpy(x, hx.data(), Q * 32 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(w1, h1.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(w2, h2.data(), 8192, cudaMemcpyHostToDevice);
  cudaMemcpy(w3, h3.data(), 640, cudaMemcpyHostToDevice);
  if (tp_mlp_head_tf32(x, Q, 1, 32, w1, w2, w3, 5, o, nullptr)) return 1;
  if (cudaMemcpy(ho.data(), o, Q * 5 * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("failed\n"); return 1; }
  FILE* f = fopen(argv[1], "wb"); fwrite(ho.data(), 4, ho.size(), f); fclose(f);
  double s = 0; for (float v : ho) s += v;
  printf("sum %.9g first %.9g %.9g\n", s, ho[0], ho[1]);
  return 0;
}
