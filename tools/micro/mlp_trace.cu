// Phase timestamps (clock64) of CTA 0 of the occupancy-head kernel, first 12 tiles, 640k queries.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DTP_MLP_TRACE -I efficient_multimodal_perception_b200/csrc \
//        -I include -o gpurun_out/mlp_trace tools/micro/mlp_trace.cu
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../efficient_multimodal_perception_b200/csrc/tp_mlp.cu"
namespace tp {
int fail(int code, const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); return code; }
int check_cuda(cudaError_t e, const char* what) { if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -9; } return 0; }
}
int main() {
  const int64_t Q = 640000;
  float *x, *w1, *w2, *w3, *o;
  cudaMalloc(&x, Q * 32 * 4); cudaMalloc(&o, Q * 5 * 4);
  cudaMalloc(&w1, 64 * 32 * 4); cudaMalloc(&w2, 64 * 32 * 4); cudaMalloc(&w3, 5 * 32 * 4);
  cudaMemset(x, 0, Q * 32 * 4); cudaMemset(w1, 0, 8192); cudaMemset(w2, 0, 8192); cudaMemset(w3, 0, 640);
  for (int r = 0; r < 3; ++r) {
    int rc = tp_mlp_head_tf32(x, Q, 1, 32, w1, w2, w3, 5, o, nullptr);
    if (rc) return 1;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  long long t[16 * 12];
  cudaMemcpyFromSymbol(t, tp::g_mlp_trace, sizeof(t));
  const char* names[11] = {"top", "layer1 done", "relu1 stored", "sync", "layer2 issued", "next A1 staged + rows requested",
                           "layer2 done", "relu2 + sync", "layer3 + next layer1 issued", "layer3 done", "logits stored"};
  for (int it = 0; it < 10; ++it) {
    printf("tile %2d:", it);
    for (int n = 1; n < 11; ++n) printf(" %5lld", t[it * 16 + n] - t[it * 16 + n - 1]);
    printf("  | total %lld\n", t[it * 16 + 10] - t[it * 16]);
  }
  printf("prologue: first loads issued %lld, weights staged %lld, tmem+sync %lld | loop %lld | whole CTA %lld cycles\n",
         t[177] - t[176], t[178] - t[177], t[179] - t[178], t[180] - t[179], t[180] - t[176]);
  {
    static unsigned long long g[2048];
    cudaMemcpyFromSymbol(g, tp::g_mlp_cta, sizeof(g));
    unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
    for (int i = 0; i < 592; ++i) { s0 = g[2*i] < s0 ? g[2*i] : s0; s1 = g[2*i] > s1 ? g[2*i] : s1; e0 = g[2*i+1] < e0 ? g[2*i+1] : e0; e1 = g[2*i+1] > e1 ? g[2*i+1] : e1; }
    printf("CTA starts span %llu ns, first end +%llu ns, last end +%llu ns\n", s1 - s0, e0 - s0, e1 - s0);
    for (int i = 0; i < 592; i += 37) printf("  cta %3d: start +%llu, dur %llu ns\n", i, g[2*i] - s0, g[2*i+1] - g[2*i]);
  }
  printf("columns:"); for (int n = 1; n < 11; ++n) printf(" [%s]", names[n]); printf("\n");
  return 0;
}
