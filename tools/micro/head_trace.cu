// Phase timeline (globaltimer) of the decode + head kernel on the 640k-query occupancy lattice: for every CTA the first
// block that runs tile chains: A+B, C, first stage, the four tiles; plus CTA start / end.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTP_HEAD_TRACE -I efficient_multimodal_perception_b200/csrc \
//        -I include -o build/micro/head_trace tools/micro/head_trace.cu build/csrc/tp_{sample,sample_grid,api,voxelize,encode,lift,backward,backward_grid,mlp}.o
#include <algorithm>
#include <cstdio>
#include <vector>
#include "../../efficient_multimodal_perception_b200/csrc/tp_sample_head.cu"
int main() {
  const int h = 200, w = 200, d = 16, C = 32, S = 128;
  const int64_t Q = (int64_t)h * w * d;
  std::vector<float> hq(Q * 3), hp(3 * S * S * C), hw(64 * 32 * 2 + 5 * 32);
  for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) for (int k = 0; k < d; ++k) {
    float* q = &hq[((int64_t)(i * w + j) * d + k) * 3];
    q[0] = (i + 0.5f) * 0.5f - 50.f; q[1] = (j + 0.5f) * 0.5f - 50.f; q[2] = (k + 0.5f) * 0.5f - 5.f;
  }
  unsigned s = 12345;
  for (auto& v : hp) { s = s * 1664525u + 1013904223u; v = (s >> 8) * (1.f / 16777216.f) - 0.5f; }
  for (auto& v : hw) { s = s * 1664525u + 1013904223u; v = ((s >> 8) * (1.f / 16777216.f) - 0.5f) * 0.3f; }
  float *q, *p, *o, *wt;
  cudaMalloc(&q, Q * 12); cudaMalloc(&p, hp.size() * 4); cudaMalloc(&o, Q * 5 * 4); cudaMalloc(&wt, hw.size() * 4);
  cudaMemcpy(q, hq.data(), Q * 12, cudaMemcpyHostToDevice); cudaMemcpy(p, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(wt, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
  tp_plane pl[3]; for (int k = 0; k < 3; ++k) { pl[k].data = p + (size_t)k * S * S * C; pl[k].batch_stride = 3 * S * S * C; pl[k].H = S; pl[k].W = S; }
  tp_sample_geom sg = {{-25.f, -25.f, -5.f}, {0.4f, 0.4f, 0.1f}, {64.f, 64.f, 64.f}};
  int dims[3] = {h, w, d};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 5; ++r) {
    if (r == 4) cudaEventRecord(e0);
    if (tp_sample3_grid_head_tf32(pl, q, dims, 1, &sg, 0, wt, wt + 2048, wt + 4096, 5, o, nullptr)) { printf("%s\n", tp_last_error()); return 1; }
  }
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("last launch (events): %.1f us\n", ms * 1e3);
  static unsigned long long g[16 * 1024];
  cudaMemcpyFromSymbol(g, tp::g_head_cta, sizeof(g));
  const int n = 444;
  unsigned long long s0 = ~0ull;
  for (int i = 0; i < n; ++i) s0 = std::min(s0, g[16 * i]);
  auto pr = [](const char* nm, std::vector<unsigned long long>& v) { if (v.empty()) return; std::sort(v.begin(), v.end());
    printf("%-34s n=%4zu  min %6llu  p10 %6llu  median %6llu  p90 %6llu  max %6llu ns\n", nm, v.size(), v[0], v[v.size()/10], v[v.size()/2], v[v.size()*9/10], v.back()); };
  std::vector<unsigned long long> st, en, setup, ab, c, stg, tile[4], blk;
  int nblk_hist[8] = {};
  for (int i = 0; i < n; ++i) {
    const unsigned long long* t = g + 16 * i;
    st.push_back(t[0] - s0); en.push_back(t[10] - s0);
    if (t[9] > t[4] && t[4] > t[3] && t[3] >= t[2]) {  // the traced block of this CTA
      c.push_back(t[3] - t[2]); stg.push_back(t[4] - t[3]);
      tile[0].push_back(t[5] - t[4]); for (int k = 1; k < 4; ++k) tile[k].push_back(t[5 + k] - t[4 + k]);
    }
  }
  pr("CTA start", st); pr("CTA end", en);
  for (int i = 0; i < n; ++i) { setup.push_back(g[16 * i + 1] - g[16 * i]); ab.push_back(g[16 * i + 2] - g[16 * i + 1]); }
  pr("setup (TMEM, weights)", setup); pr("A+B of the first traced block", ab);
  pr("C (tables) of the traced block", c); pr("first stage + barrier", stg);
  for (int k = 0; k < 4; ++k) { char nm[32]; snprintf(nm, sizeof nm, "tile %d (chain)", k); pr(nm, tile[k]); }
  return 0;
}
