// Probe: how fast can one B200 gather 512-byte rows in random order, as a function of the number of
// independent 16-byte loads each lane keeps in flight (U) and of resident warps per SM?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather_probe tools/micro/gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ld_hint(const float4* p, unsigned long long pol) {
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
// MODE 0: plain __ldg; 1: L2::cache_hint evict_last; 2: L2::cache_hint evict_normal; 3: evict_first
template <int U, int MODE>
__global__ void gather_m(const float4* __restrict__ rows, const int* __restrict__ ids, int n_ids, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  unsigned long long pol;
  if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = warp * U; i + U <= n_ids; i += nwarp * U) {
    float4 x[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float4* p = rows + (size_t)__ldg(ids + i + j) * 32 + lane;
      x[j] = MODE == 0 ? __ldg(p) : ld_hint(p, pol);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) { acc.x += x[j].x; acc.y += x[j].y; acc.z += x[j].z; acc.w += x[j].w; }
  }
  out[(size_t)warp * 32 + lane] = acc;
}

template <int U>
__global__ void gather(const float4* __restrict__ rows, const int* __restrict__ ids, int n_ids, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = warp * U; i + U <= n_ids; i += nwarp * U) {
    float4 x[U];
#pragma unroll
    for (int j = 0; j < U; ++j) x[j] = __ldg(rows + (size_t)__ldg(ids + i + j) * 32 + lane);
#pragma unroll
    for (int j = 0; j < U; ++j) { acc.x += x[j].x; acc.y += x[j].y; acc.z += x[j].z; acc.w += x[j].w; }
  }
  out[(size_t)warp * 32 + lane] = acc;
}

int main() {
  const int n_rows = 350000, n_ids = 3 * n_rows;
  std::vector<int> h(n_ids);
  srand(1);
  for (auto& v : h) v = (int)((((long long)rand() << 15) ^ rand()) % n_rows);
  float4 *rows, *out; int* ids;
  cudaMalloc(&rows, (size_t)n_rows * 512); cudaMalloc(&out, (size_t)148 * 64 * 32 * 16 * 4); cudaMalloc(&ids, n_ids * 4);
  cudaMemset(rows, 0, (size_t)n_rows * 512);
  cudaMemcpy(ids, h.data(), n_ids * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int wps : {8, 16, 32, 64}) {
    const int ctas = 148 * wps / 4;
#define RUN(U) { for (int r = 0; r < 3; ++r) gather<U><<<ctas, 128>>>(rows, ids, n_ids, out); \
    cudaEventRecord(a); for (int r = 0; r < 10; ++r) gather<U><<<ctas, 128>>>(rows, ids, n_ids, out); cudaEventRecord(b); \
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); \
    printf("warps/SM %2d  U=%2d  %7.1f us  %6.0f GB/s\n", wps, U, ms * 100, (double)n_ids * 512 / (ms / 10 * 1e-3) / 1e9); }
    RUN(1) RUN(2) RUN(4) RUN(8) RUN(16)
  }
  {
    const int ctas = 148 * 10;
#define RUNM(M) { for (int r = 0; r < 3; ++r) gather_m<4, M><<<ctas, 128>>>(rows, ids, n_ids, out); \
    cudaEventRecord(a); for (int r = 0; r < 10; ++r) gather_m<4, M><<<ctas, 128>>>(rows, ids, n_ids, out); cudaEventRecord(b); \
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); \
    printf("40 warps/SM U=4 mode %d (0 ldg, 1 hint evict_last, 2 hint evict_normal, 3 hint evict_first) %7.1f us %6.0f GB/s\n", M, ms * 100, (double)n_ids * 512 / (ms / 10 * 1e-3) / 1e9); }
    RUNM(0) RUNM(1) RUNM(2) RUNM(3)
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
