// Probe: write bandwidth of an [C=32][h=200][w=200][d=16] fp32 tensor (82 MB) when every CTA writes, for all 32
// channels and BI lattice rows, a run of RUN bytes (the D phase of tp_sample_grid.cu writes RUN = 512), vs a
// linear fill. build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_probe tools/micro/store_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st4(float4* p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol));
}
__global__ void linear(float4* out, size_t n4) {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    st4(out + i, make_float4(1, 2, 3, 4), pol);
}
// block = BI rows x (RUN bytes of the w*d axis); warp w handles channels w, w+8, ...; per (c, i): RUN bytes
template <int BI, int RUN>
__global__ void blocked(float4* out, int nblk_j, int nblk_i, int rowbytes, size_t chbytes) {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int blk = blockIdx.x; blk < nblk_j * nblk_i; blk += gridDim.x) {
    const int jb = blk % nblk_j, ib = blk / nblk_j;
    char* base = (char*)out + (size_t)ib * BI * rowbytes + (size_t)jb * RUN;
    for (int c = warp; c < 32; c += 8) {
#pragma unroll
      for (int ii = 0; ii < BI; ++ii)
#pragma unroll
        for (int o = 0; o < RUN; o += 512)
          st4((float4*)(base + c * chbytes + (size_t)ii * rowbytes + o) + lane, make_float4(1, 2, 3, 4), pol);
    }
  }
}
int main() {
  const int h = 200, w = 200, d = 16, C = 32;
  const int rowbytes = w * d * 4;  // 12800
  const size_t chbytes = (size_t)h * rowbytes, total = chbytes * C;
  float4* out[6];
  for (auto& o : out) cudaMalloc(&o, total);  // rotate over 6 x 82 MB > 3 x L2
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
#define TIME(name, launch) { for (int r = 0; r < 6; ++r) { float4* o = out[r % 6]; launch; } cudaEventRecord(a); \
    for (int r = 0; r < 60; ++r) { float4* o = out[r % 6]; launch; } cudaEventRecord(b); cudaEventSynchronize(b); \
    cudaEventElapsedTime(&ms, a, b); printf("%-28s %6.1f us  %6.0f GB/s\n", name, ms / 60 * 1e3, total / (ms / 60 * 1e-3) / 1e9); }
  TIME("linear 592x256", (linear<<<592, 256>>>(o, total / 16)))
  TIME("linear 2368x256", (linear<<<2368, 256>>>(o, total / 16)))
  TIME("blocked BI=8 RUN=512", (blocked<8, 512><<<592, 256>>>(o, rowbytes / 512, h / 8, rowbytes, chbytes)))
  TIME("blocked BI=4 RUN=512", (blocked<4, 512><<<592, 256>>>(o, rowbytes / 512, h / 4, rowbytes, chbytes)))
  TIME("blocked BI=4 RUN=2560", (blocked<4, 2560><<<592, 256>>>(o, rowbytes / 2560, h / 4, rowbytes, chbytes)))
  TIME("blocked BI=2 RUN=2560", (blocked<2, 2560><<<592, 256>>>(o, rowbytes / 2560, h / 2, rowbytes, chbytes)))
  TIME("blocked BI=1 RUN=12800", (blocked<1, 12800><<<592, 256>>>(o, 1, h, rowbytes, chbytes)))
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
