// Error plumbing, device queries and the host-buffer entry points of libtriplane.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "tp_common.cuh"

namespace tp {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

// Per-thread device arena for the host-buffer entry points: grows, never shrinks, freed by
// tp_host_arena_release(). One stream per thread; H2D -> kernels -> D2H are ordered on it.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_d2h = nullptr;  // result copies of chunk i overlap the kernel and the H2D of chunk i+1
  cudaEvent_t ev[8] = {};
  void* enc_ws = nullptr;  // encode workspace keeps its "clean head table" invariant across calls
  size_t enc_ws_bytes = 0;
  int reserve(size_t bytes) {
    if (!stream) {
      TP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      TP_CUDA(cudaStreamCreateWithFlags(&stream_d2h, cudaStreamNonBlocking));
      for (auto& e : ev) TP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (bytes > cap) {
      if (base) { TP_CUDA(cudaStreamSynchronize(stream)); TP_CUDA(cudaStreamSynchronize(stream_d2h)); TP_CUDA(cudaFree(base)); base = nullptr; cap = 0; }
      TP_CUDA(cudaMalloc((void**)&base, bytes));
      cap = bytes;
    }
    used = 0;
    return 0;
  }
  template <class T>
  T* take(size_t count) {
    size_t off = (used + 255) / 256 * 256;
    used = off + count * sizeof(T);
    return reinterpret_cast<T*>(base + off);
  }
  static size_t pad(size_t bytes) { return (bytes + 255) / 256 * 256 + 256; }
};
static thread_local Arena g_arena;

}  // namespace tp

using namespace tp;

extern "C" const char* tp_last_error(void) { return g_err; }
extern "C" int tp_version(void) { return TP_VERSION; }
extern "C" int tp_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

extern "C" void tp_host_arena_release(void) {
  Arena& A = g_arena;
  if (A.stream) cudaStreamSynchronize(A.stream);
  if (A.base) cudaFree(A.base);
  if (A.enc_ws) cudaFree(A.enc_ws);
  if (A.stream) cudaStreamDestroy(A.stream);
  if (A.stream_d2h) { cudaStreamSynchronize(A.stream_d2h); cudaStreamDestroy(A.stream_d2h); }
  for (auto& e : A.ev) if (e) cudaEventDestroy(e);
  A = Arena();
}

static int sample3_host_impl(const float* planes_host[3], const int32_t HW[6],
                             const int64_t plane_batch_stride[3], int32_t C,
                             const float* queries_host, int64_t Q, const int32_t* dims, int32_t batch,
                             const tp_sample_geom* sg, int32_t arith, float* out_host) {
  if (!planes_host || !HW || !plane_batch_stride || !queries_host || !out_host || !sg)
    return fail(TP_E_NULL, "tp_sample3_host: null argument");
  if (C <= 0 || (C & 3) || batch <= 0 || Q < 0) return fail(TP_E_SHAPE, "tp_sample3_host: C=%d B=%d Q=%lld", C, batch, (long long)Q);
  if (Q == 0) return 0;
  Arena& A = g_arena;
  size_t plane_elems[3], total = 0;
  for (int k = 0; k < 3; ++k) {
    if (!planes_host[k]) return fail(TP_E_NULL, "tp_sample3_host: plane %d null", k);
    plane_elems[k] = (size_t)C * HW[2 * k] * HW[2 * k + 1];
    total += 2 * Arena::pad(plane_elems[k] * batch * 4);  // NCHW copy + NHWC copy
  }
  const size_t qbytes = (size_t)batch * Q * 3 * 4, obytes = (size_t)batch * C * Q * 4;
  total += Arena::pad(qbytes) + Arena::pad(obytes);
  if (int rc = A.reserve(total)) return rc;
  cudaStream_t s = A.stream;
  tp_plane nchw[3], nhwc[3];
  for (int k = 0; k < 3; ++k) {
    float* d = A.take<float>(plane_elems[k] * batch);
    if (plane_batch_stride[k] == (int64_t)plane_elems[k]) {
      TP_CUDA(cudaMemcpyAsync(d, planes_host[k], plane_elems[k] * batch * 4, cudaMemcpyHostToDevice, s));
    } else {
      for (int b = 0; b < batch; ++b)
        TP_CUDA(cudaMemcpyAsync(d + (size_t)b * plane_elems[k], planes_host[k] + (size_t)b * plane_batch_stride[k],
                                plane_elems[k] * 4, cudaMemcpyHostToDevice, s));
    }
    nchw[k].data = d;
    nchw[k].batch_stride = (int64_t)plane_elems[k];
    nchw[k].H = HW[2 * k];
    nchw[k].W = HW[2 * k + 1];
    float* t = A.take<float>(plane_elems[k] * batch);
    if (int rc = tp_planes_nchw_to_nhwc_f32(d, nchw[k].batch_stride, t, batch, C, nchw[k].H, nchw[k].W, s)) return rc;
    nhwc[k] = nchw[k];
    nhwc[k].data = t;
  }
  float* dq = A.take<float>((size_t)batch * Q * 3);
  float* dout = A.take<float>((size_t)batch * C * Q);
  // Pipeline over chunks of the query range (rows of the lattice for the 5-D entry): the device->host copy of
  // chunk i (the 4C bytes per query that dominate the PCIe time) runs on its own stream while chunk i+1's
  // queries go up and its kernel runs. Each chunk's result is a compact [C, q_chunk] block on the device and
  // lands in the caller's [B, C, Q] tensor with a strided copy.
  const int64_t unit = dims ? (int64_t)dims[1] * dims[2] : 32;  // queries per indivisible slice
  const int64_t units = dims ? dims[0] : (Q + 31) / 32;
  int nchunk = (obytes / batch >= ((size_t)8 << 20) && units >= 4) ? 4 : 1;
  const int64_t upc = (units + nchunk - 1) / nchunk;
  size_t off_out = 0;
  int evi = 0;
  for (int b = 0; b < batch; ++b) {
    tp_plane pb[3];
    for (int k = 0; k < 3; ++k) {
      pb[k] = nhwc[k];
      pb[k].data = nhwc[k].data + (size_t)b * nhwc[k].batch_stride;
    }
    for (int64_t u0 = 0; u0 < units; u0 += upc) {
      const int64_t q0 = u0 * unit;
      const int64_t qc = (u0 + upc < units ? upc * unit : Q - q0);
      const float* hq = queries_host + ((size_t)b * Q + q0) * 3;
      float* dqc = dq + ((size_t)b * Q + q0) * 3;
      float* doc = dout + off_out;
      off_out += (size_t)C * qc;
      TP_CUDA(cudaMemcpyAsync(dqc, hq, (size_t)qc * 12, cudaMemcpyHostToDevice, s));
      int rc;
      if (dims) {
        const int32_t cd[3] = {(int32_t)(qc / unit), dims[1], dims[2]};
        rc = tp_sample3_grid_nhwc_f32(pb, C, dqc, cd, 1, sg, arith, doc, s);
      } else {
        rc = tp_sample3_nhwc_f32(pb, C, dqc, qc, 1, sg, arith, doc, s);
      }
      if (rc) return rc;
      cudaEvent_t e = A.ev[evi++ & 7];
      TP_CUDA(cudaEventRecord(e, s));
      TP_CUDA(cudaStreamWaitEvent(A.stream_d2h, e, 0));
      TP_CUDA(cudaMemcpy2DAsync(out_host + (size_t)b * C * Q + q0, (size_t)Q * 4, doc, (size_t)qc * 4, (size_t)qc * 4,
                                (size_t)C, cudaMemcpyDeviceToHost, A.stream_d2h));
    }
  }
  TP_CUDA(cudaStreamSynchronize(A.stream_d2h));
  TP_CUDA(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int tp_sample3_host_f32(const float* planes_host[3], const int32_t HW[6],
                                   const int64_t plane_batch_stride[3], int32_t C,
                                   const float* queries_host, int64_t Q, int32_t batch,
                                   const tp_sample_geom* sg, int32_t arith, float* out_host) {
  return sample3_host_impl(planes_host, HW, plane_batch_stride, C, queries_host, Q, nullptr, batch, sg, arith,
                           out_host);
}

extern "C" int tp_sample3_grid_host_f32(const float* planes_host[3], const int32_t HW[6],
                                        const int64_t plane_batch_stride[3], int32_t C,
                                        const float* queries_host, const int32_t dims[3], int32_t batch,
                                        const tp_sample_geom* sg, int32_t arith, float* out_host) {
  if (!dims || dims[0] < 0 || dims[1] < 0 || dims[2] < 0) return fail(TP_E_SHAPE, "tp_sample3_grid_host_f32: bad dims");
  return sample3_host_impl(planes_host, HW, plane_batch_stride, C, queries_host,
                           (int64_t)dims[0] * dims[1] * dims[2], dims, batch, sg, arith, out_host);
}

extern "C" int tp_encode_host_f32(const float* feats_host, int32_t C, const float* points_host,
                                  int32_t point_stride, int64_t n, const int64_t* offsets_host,
                                  int32_t batch, const tp_geom* geom, int32_t arith, int32_t reduce,
                                  int32_t clamp_zero, float* out_xy_host, float* out_yz_host,
                                  float* out_xz_host) {
  if (!geom || !offsets_host) return fail(TP_E_NULL, "tp_encode_host_f32: null argument");
  if (n > 0 && (!feats_host || !points_host)) return fail(TP_E_NULL, "tp_encode_host_f32: null input");
  if (batch <= 0 || n < 0 || C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "tp_encode_host_f32: bad shape");
  Arena& A = g_arena;
  int64_t cells[3];
  if (tp_encode_cells(geom, batch, cells) < 0) return fail(TP_E_SHAPE, "tp_encode_host_f32: bad geometry");
  float* outs_host[3] = {out_xy_host, out_yz_host, out_xz_host};
  size_t total = Arena::pad((size_t)n * C * 4) + Arena::pad((size_t)n * point_stride * 4) +
                 Arena::pad((size_t)(batch + 1) * 8);
  for (int k = 0; k < 3; ++k)
    if (outs_host[k]) total += Arena::pad((size_t)cells[k] * C * 4);
  if (int rc = A.reserve(total)) return rc;
  cudaStream_t s = A.stream;
  const int64_t wsb = tp_encode_workspace_bytes(geom, batch, n);
  if (wsb < 0) return fail(TP_E_SHAPE, "tp_encode_host_f32: bad geometry");
  if ((size_t)wsb > A.enc_ws_bytes) {
    if (A.enc_ws) { TP_CUDA(cudaStreamSynchronize(s)); TP_CUDA(cudaFree(A.enc_ws)); A.enc_ws = nullptr; }
    TP_CUDA(cudaMalloc(&A.enc_ws, (size_t)wsb));
    A.enc_ws_bytes = (size_t)wsb;
    if (int rc = tp_encode_workspace_init(A.enc_ws, wsb, s)) return rc;
  }
  float* dfe = A.take<float>((size_t)n * C);
  float* dpt = A.take<float>((size_t)n * point_stride);
  int64_t* doff = A.take<int64_t>((size_t)batch + 1);
  float* douts[3] = {nullptr, nullptr, nullptr};
  for (int k = 0; k < 3; ++k)
    if (outs_host[k]) douts[k] = A.take<float>((size_t)cells[k] * C);
  if (n > 0) {
    TP_CUDA(cudaMemcpyAsync(dfe, feats_host, (size_t)n * C * 4, cudaMemcpyHostToDevice, s));
    TP_CUDA(cudaMemcpyAsync(dpt, points_host, (size_t)n * point_stride * 4, cudaMemcpyHostToDevice, s));
  }
  TP_CUDA(cudaMemcpyAsync(doff, offsets_host, (size_t)(batch + 1) * 8, cudaMemcpyHostToDevice, s));
  if (int rc = tp_encode_f32(dfe, C, C, nullptr, dpt, point_stride, n, doff, batch, geom, arith, reduce,
                             clamp_zero, douts[0], douts[1], douts[2], nullptr, A.enc_ws,
                             (int64_t)A.enc_ws_bytes, s))
    return rc;
  for (int k = 0; k < 3; ++k)
    if (outs_host[k])
      TP_CUDA(cudaMemcpyAsync(outs_host[k], douts[k], (size_t)cells[k] * C * 4, cudaMemcpyDeviceToHost, s));
  TP_CUDA(cudaStreamSynchronize(s));
  return 0;
}
