// Per-CTA timeline (globaltimer at entry / start of the second block / exit) of the lattice decode kernel on the
// 640k-query occupancy lattice (200 x 200 x 16, C = 32, three 128 x 128 planes).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTP_GRID_TRACE -I efficient_multimodal_perception_b200/csrc \
//        -I include -o build/micro/grid_trace tools/micro/grid_trace.cu build/csrc/tp_{sample,sample_head,api,voxelize,encode,lift,backward,backward_grid,mlp}.o
#include <cstdarg>
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../efficient_multimodal_perception_b200/csrc/tp_sample_grid.cu"
int main() {
  const int h = 200, w = 200, d = 16, C = 32, S = 128;
  const int64_t Q = (int64_t)h * w * d;
  std::vector<float> hq(Q * 3), hp(3 * S * S * C);
  for (int i = 0; i < h; ++i) for (int j = 0; j < w; ++j) for (int k = 0; k < d; ++k) {
    float* q = &hq[((int64_t)(i * w + j) * d + k) * 3];
    q[0] = (i + 0.5f) * 0.5f - 50.f; q[1] = (j + 0.5f) * 0.5f - 50.f; q[2] = (k + 0.5f) * 0.5f - 5.f;
  }
  unsigned s = 12345; for (auto& v : hp) { s = s * 1664525u + 1013904223u; v = (s >> 8) * (1.f / 16777216.f) - 0.5f; }
  float *q, *p, *o;
  cudaMalloc(&q, Q * 12); cudaMalloc(&p, hp.size() * 4); cudaMalloc(&o, Q * C * 4);
  cudaMemcpy(q, hq.data(), Q * 12, cudaMemcpyHostToDevice); cudaMemcpy(p, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice);
  tp_plane pl[3]; for (int k = 0; k < 3; ++k) { pl[k].data = p + (size_t)k * S * S * C; pl[k].batch_stride = 3 * S * S * C; pl[k].H = S; pl[k].W = S; }
  tp_sample_geom sg = {{-25.f, -25.f, -5.f}, {0.4f, 0.4f, 0.1f}, {64.f, 64.f, 64.f}};  // z plane axis: 80 cells of 0.1 in a 128 grid
  int dims[3] = {h, w, d};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 5; ++r) {
    if (r == 4) cudaEventRecord(e0);
    if (tp_sample3_grid_nhwc_f32(pl, C, q, dims, 1, &sg, 0, o, nullptr)) { printf("%s\n", tp_last_error()); return 1; }
  }
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("last launch (events): %.1f us\n", ms * 1e3);
  static unsigned long long g[8192];
  cudaMemcpyFromSymbol(g, tp::g_grid_cta, sizeof(g));
  const int n = 592;
  unsigned long long s0 = ~0ull;
  for (int i = 0; i < n; ++i) s0 = std::min(s0, g[8 * i]);
  std::vector<unsigned long long> st, e1v, e2v;
  for (int i = 0; i < n; ++i) { st.push_back(g[8*i] - s0); if (g[8*i+1] >= g[8*i]) e2v.push_back(g[8*i+2] - s0); else e1v.push_back(g[8*i+2] - s0); }
  auto pr = [](const char* nm, std::vector<unsigned long long>& v) { if (v.empty()) return; std::sort(v.begin(), v.end());
    printf("%-28s n=%4zu  min %6llu  p10 %6llu  median %6llu  p90 %6llu  max %6llu ns\n", nm, v.size(), v[0], v[v.size()/10], v[v.size()/2], v[v.size()*9/10], v.back()); };
  {  // phases of every CTA's first block, split by whether the table phase gathered anything
    std::vector<unsigned long long> ab[3], bar[3], c[3], dd[3], tot[3];
    for (int i = 0; i < n; ++i) {
      const unsigned long long* t = g + 8 * i;
      const int ib = i / 25, jb = i % 25;  // BI = 8 on the 200 x 200 x 16 lattice: planes cover lattice indices [50, 150)
      const bool xin = ib * 8 + 7 >= 50 && ib * 8 < 150, yin = jb * 8 + 7 >= 50 && jb * 8 < 150;
      const int cls = (xin ? 1 : 0) + (yin ? 1 : 0);
      ab[cls].push_back(t[3] - t[0]); bar[cls].push_back(t[4] - t[3]); c[cls].push_back(t[5] - t[4]); dd[cls].push_back(t[6] - t[5]);
      tot[cls].push_back(t[6] - t[0]);
    }
    const char* cname[3] = {"no plane in range (store-only)", "one plane in range (xz or yz)", "all three planes in range"};
    for (int cls = 0; cls < 3; ++cls) {
      printf("-- first block, %s (BI = 8 only)\n", cname[cls]);
      pr("  A+B (queries, footprints)", ab[cls]); pr("  barrier", bar[cls]); pr("  C (tables)", c[cls]); pr("  D (stores)", dd[cls]); pr("  whole block", tot[cls]);
    }
  }
  pr("CTA start", st); pr("end (CTAs with 1 block)", e1v); pr("end (CTAs with 2 blocks)", e2v);
  return 0;
}
