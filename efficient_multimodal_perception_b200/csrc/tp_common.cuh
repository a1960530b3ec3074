// Shared device/host helpers for libtriplane (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "triplane.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtriplane is written for sm_100a (B200) only"
#endif

namespace tp {

constexpr int kSMs = 148;  // B200; grids are sized in multiples of this

// ---- error plumbing (tp_api.cu owns the storage) -------------------------------------------
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define TP_CUDA(expr)                                  \
  do {                                                 \
    cudaError_t _e = (expr);                           \
    if (_e != cudaSuccess) return ::tp::check_cuda(_e, #expr); \
  } while (0)
#define TP_LAUNCH_CHECK(name)                          \
  do {                                                 \
    cudaError_t _e = cudaGetLastError();               \
    if (_e != cudaSuccess) return ::tp::check_cuda(_e, name); \
  } while (0)

// > 48 KB of dynamic shared memory needs a per-device opt-in. One atomic device bitmask per kernel
// (the kernel is a template argument, so two kernels of the same signature do not share it);
// racing threads at worst repeat the idempotent cudaFuncSetAttribute.
template <auto Kern>
inline cudaError_t opt_in_smem(int bytes) {
  static std::atomic<unsigned long long> done{0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < 64;
  if (tracked && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && tracked) done.fetch_or(1ull << dev, std::memory_order_release);
  return e;
}

// ---- programmatic dependent launch: layout conversion -> decode inside one C-ABI call ----------
// The *_nchw_* entry points launch the NCHW -> channels-last conversion and then the decode kernel with
// programmaticStreamSerializationAllowed: the conversion calls pdl_launch_dependents() first thing, so the decode
// grid is scheduled while the conversion is still running and does everything that does not need the planes (query
// load, lattice check, tap records, all-zero blocks) before its first pdl_wait(). The attribute is only ever set on
// a launch that directly follows our own conversion kernel (itself launched with full stream ordering), so nothing
// but that conversion can overlap the decode. pdl_wait() is a no-op in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- the reference's coordinate chain, op by op, never contracted -----------------------------
// (p - lo) (/) vs : torch-CUDA multiplies by the fp32 reciprocal of the Python-float divisor,
// torch-CPU divides (SURVEY §7 "bit-exact voxel indices").
struct AxisMap {
  float lo, d, rcp;  // d = divisor, rcp = 1.0f / d computed in fp32 on the host
};

template <int ARITH>
__device__ __forceinline__ float tp_div(float a, float d, float rcp) {
  if (ARITH == TP_ARITH_TORCH_CUDA) return __fmul_rn(a, rcp);
  return __fdiv_rn(a, d);
}

// voxel coordinate v = (p - lo) / vs  (point_triplane.py:153-156)
template <int ARITH>
__device__ __forceinline__ float tp_voxel_coord(float p, float lo, float vs, float rcp_vs) {
  return tp_div<ARITH>(__fsub_rn(p, lo), vs, rcp_vs);
}

struct GeomDev {
  float lo[3], hi[3], vs[3], rcp_vs[3];
  int grid[3], pool[3], pooled[3];  // pooled = (grid - pool) / pool + 1
};

inline GeomDev make_geom_dev(const tp_geom& g) {
  GeomDev d;
  for (int a = 0; a < 3; ++a) {
    d.lo[a] = g.lo[a];
    d.hi[a] = g.hi[a];
    d.vs[a] = g.vs[a];
    d.rcp_vs[a] = 1.0f / g.vs[a];
    d.grid[a] = g.grid[a];
    d.pool[a] = g.pool[a];
    d.pooled[a] = (g.grid[a] - g.pool[a]) / g.pool[a] + 1;
  }
  return d;
}

// strict crop of point_triplane.py:148-150 + int32 truncation of :159
template <int ARITH>
__device__ __forceinline__ bool tp_crop_index(const GeomDev& g, float x, float y, float z, int& ix,
                                              int& iy, int& iz) {
  bool keep = (x > g.lo[0]) & (x < g.hi[0]) & (y > g.lo[1]) & (y < g.hi[1]) & (z > g.lo[2]) &
              (z < g.hi[2]);
  ix = __float2int_rz(tp_voxel_coord<ARITH>(x, g.lo[0], g.vs[0], g.rcp_vs[0]));
  iy = __float2int_rz(tp_voxel_coord<ARITH>(y, g.lo[1], g.vs[1], g.rcp_vs[1]));
  iz = __float2int_rz(tp_voxel_coord<ARITH>(z, g.lo[2], g.vs[2], g.rcp_vs[2]));
  return keep;
}

// sample index of a concatenated row: largest b with offsets[b] <= i  (B is small)
__device__ __forceinline__ int tp_find_batch(const int64_t* __restrict__ offsets, int batch,
                                             int64_t i) {
  int lo = 0, hi = batch;  // offsets[lo] <= i < offsets[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// streaming store: written once, never re-read by this kernel
__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cs_f1(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// L2 eviction policies (createpolicy): streamed-once data must not evict the small hot working set.
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_keep_f4(const float4* p, unsigned long long pol) {
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_stream_f1(float* p, float v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

}  // namespace tp
