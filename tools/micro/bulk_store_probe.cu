// Probe: write bandwidth of cp.async.bulk.global.shared::cta stores of RUN bytes each (the D phase of the lattice
// decode kernel would issue one per (channel, lattice row): 512 B), one store per thread and round, from a
// 32 KB shared staging buffer, same [32][200][200][16] fp32 destination pattern as store_probe.cu.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_store_probe tools/micro/bulk_store_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(void* g, const void* s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
               "r"((unsigned)__cvta_generic_to_shared(s)), "r"(bytes) : "memory");
}
template <int BI, int RUN>
__global__ void blocked(char* out, int nblk_j, int nblk_i, int rowbytes, size_t chbytes) {
  extern __shared__ __align__(128) char stage[];  // 32 KB
  for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) ((float4*)stage)[i] = make_float4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int per_block = 32 * BI;  // (channel, row) pairs
  for (int blk = blockIdx.x; blk < nblk_j * nblk_i; blk += gridDim.x) {
    const int jb = blk % nblk_j, ib = blk / nblk_j;
    char* base = out + (size_t)ib * BI * rowbytes + (size_t)jb * RUN;
    for (int p = threadIdx.x; p < per_block; p += blockDim.x) {
      const int c = p / BI, ii = p % BI;
      bulk_store(base + c * chbytes + (size_t)ii * rowbytes, stage + (p * RUN) % 32768, RUN);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  const int h = 200, w = 200, d = 16, C = 32;
  const int rowbytes = w * d * 4;
  const size_t chbytes = (size_t)h * rowbytes, total = chbytes * C;
  char* out[6];
  for (auto& o : out) cudaMalloc(&o, total);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
#define TIME(name, launch) { for (int r = 0; r < 6; ++r) { char* o = out[r % 6]; launch; } cudaEventRecord(a); \
    for (int r = 0; r < 60; ++r) { char* o = out[r % 6]; launch; } cudaEventRecord(b); cudaEventSynchronize(b); \
    cudaEventElapsedTime(&ms, a, b); printf("%-34s %6.1f us  %6.0f GB/s\n", name, ms / 60 * 1e3, total / (ms / 60 * 1e-3) / 1e9); }
  TIME("bulk BI=8 RUN=512  296x256", (blocked<8, 512><<<296, 256, 32768>>>(o, rowbytes / 512, h / 8, rowbytes, chbytes)))
  TIME("bulk BI=8 RUN=512  592x256", (blocked<8, 512><<<592, 256, 32768>>>(o, rowbytes / 512, h / 8, rowbytes, chbytes)))
  TIME("bulk BI=8 RUN=512  592x64", (blocked<8, 512><<<592, 64, 32768>>>(o, rowbytes / 512, h / 8, rowbytes, chbytes)))
  TIME("bulk BI=4 RUN=2560 592x128", (blocked<4, 2560><<<592, 128, 32768>>>(o, rowbytes / 2560, h / 4, rowbytes, chbytes)))
  TIME("bulk BI=1 RUN=12800 592x32", (blocked<1, 12800><<<592, 32, 32768>>>(o, 1, h, rowbytes, chbytes)))
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
