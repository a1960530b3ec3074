// Decode backward for [B,h,w,d,3] query tensors (the 5-D callers train too: TriplaneOcc's occupancy loss flows back
// through sample_points_triplane, triplane_occ.py:182-186). On a voxel-centre lattice the bilinear footprint of a
// query in plane xy depends on (i, j) only, so
//     d(plane xy)[tap(i,j)] += w_tap(i,j) * sum_k grad_out[c,i,j,k]
// and likewise yz sums over i and xz over j. A block of BI x 8 x 16 queries therefore needs BI*8 + 8*16 + BI*16
// scatter footprints instead of 3 * BI*8*16: ~10x fewer vector reductions into the gradient planes than the
// per-query kernel (tp_backward.cu), which is bound by exactly those.
//
// Same structure as the forward lattice kernel (tp_sample_grid.cu): A) the block's queries are checked bit for bit
// for the lattice structure, B) one record per table entry; then per chunk of 32 channels R) grad_out of the block
// is reduced into three tables in shared memory (sum over k / i / j: registers and warp shuffles), S) every entry
// scatters w_tap * table value into its four taps with red.global.add.v4.f32. A block that is not a lattice scatters
// per query inside the same launch. Sums are associated differently from the per-query kernel (and from ATen, whose
// atomics are unordered anyway): results agree to fp32 rounding.
#include "tp_sample_grid.cuh"

namespace tp {

struct GridBwdParams {
  GridParams G;      // G.S.plane[k] / G.S.out unused
  float* gplane[3];  // [B,H,W,C] gradient planes (pre-zeroed), batch stride = G.S.bstride[k]
  const float* gout; // [B,C,Q]
};

constexpr int kBwdGridCtasPerSm = 3;

// table entries of one plane: 4 channels (rows 4 l8 .. 4 l8 + 3) of entry column `col` -> the four taps
template <int NR>
__device__ __forceinline__ void scatter_table(bool live, float4* __restrict__ gp, int C4, int WC4, const float4* s_w,
                                              const int2* s_om, const float* t, int stride, bool cvalid) {
  if (!live) return;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int2 om = s_om[r * 32];
    const int mk = cvalid ? om.y : 0;
    if (mk == 0) continue;
    const float* tt = t + r * 32;
    const float4 g = make_float4(tt[0], tt[stride], tt[2 * stride], tt[3 * stride]);
    scatter_taps(gp, om.x, C4, WC4, s_w[r * 32], mk, g);
  }
}

template <int ARITH, int BI>
__global__ void __launch_bounds__(kGridThreads, kBwdGridCtasPerSm)
sample3_grid_backward_kernel(const __grid_constant__ GridBwdParams BP) {
  using Cfg = GridCfg<BI>;
  constexpr int BJ = kBJ;
  extern __shared__ __align__(16) float smem[];
  float4* const s_w = reinterpret_cast<float4*>(smem + Cfg::kWords);
  int2* const s_om = reinterpret_cast<int2*>(s_w + Cfg::E);
  __shared__ int s_vote[2];

  const GridParams& G = BP.G;
  const SampleParams& P = G.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C4 = P.C >> 2, C = P.C;
  const int wd = G.w * G.d;
  const int nblocks = G.nblocks;
  float* const T0 = smem;
  float* const T1 = T0 + 32 * Cfg::S0;
  float* const T2 = T1 + 32 * Cfg::S1;
  const int l8 = tid & 7, ent = tid >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  const int nchunk = (C4 + 7) >> 3;
  // scatter phase: this thread's table read positions (rows 4 l8 .., column = entry), as the forward kernel writes them
  const float* const r0 = T0 + (4 * l8) * Cfg::S0 + ent;
  const float* const r1 = T1 + (4 * l8) * Cfg::S1 + (ent ^ swz_bits(4 * l8));
  const float* const r2 = T2 + (4 * l8) * Cfg::S2 + (ent ^ swz_bits(4 * l8));
  const int kg = lane & 3, jj = lane >> 2;
  int nblk_done = 0;
  if (tid == 0) s_vote[0] = 0;
  __syncthreads();

  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const BlockPos bp = block_pos<BI>(G, blk);
    const int b = bp.b, i0 = bp.i0, j0 = bp.j0, k0 = bp.k0;
    const int ni = min(BI, G.h - i0), nj = min(BJ, G.w - j0), nk = min(kBK, G.d - k0);
    const int64_t qblk = ((int64_t)i0 * G.w + j0) * G.d + k0;  // first query of the block inside its sample
    const float* q00 = P.queries + ((int64_t)b * P.Q + qblk) * 3;

    // ---- A / B: as in the forward kernel (tp_sample_grid.cuh) ---------------------------------------------------
    const bool ok = grid_block_is_lattice<BI>(G, q00, ni, nj, nk, wd, tid);
    const int live = grid_build_records<ARITH, BI>(G, q00, ni, nj, nk, wd, C4, s_w, s_om, tid);
    grid_cast_vote(s_vote, nblk_done, live, ok, tid);
    __syncthreads();
    const int vote = s_vote[nblk_done & 1];
    ++nblk_done;

    float4* const gp0 = reinterpret_cast<float4*>(BP.gplane[0] + (int64_t)b * P.bstride[0]) + l8;
    float4* const gp1 = reinterpret_cast<float4*>(BP.gplane[1] + (int64_t)b * P.bstride[1]) + l8;
    float4* const gp2 = reinterpret_cast<float4*>(BP.gplane[2] + (int64_t)b * P.bstride[2]) + l8;

    if (vote & 8) {
      // ---- not a lattice: per-query scatter, 32 queries per pass (8 lanes x 4 channels per query) --------------
      for (int m = ent; m < BI * BJ * kBK; m += kGridThreads / 8) {
        const int mi = m / (BJ * kBK), mj = (m / kBK) % BJ, mk = m % kBK;
        if (mi >= ni || mj >= nj || mk >= nk) continue;
        const int64_t q = qblk + (int64_t)mi * wd + mj * G.d + mk;
        const float* qp = P.queries + ((int64_t)b * P.Q + q) * 3;
        const float gx = grid_coord<ARITH>(P, __ldg(qp), 0), gy = grid_coord<ARITH>(P, __ldg(qp + 1), 1),
                    gz = grid_coord<ARITH>(P, __ldg(qp + 2), 2);
        float4 w[3];
        int base[3], msk[3];
        plane_setup<ARITH>(gx, gy, P.W[0], P.H[0], w[0], base[0], msk[0]);
        plane_setup<ARITH>(gy, gz, P.W[1], P.H[1], w[1], base[1], msk[1]);
        plane_setup<ARITH>(gx, gz, P.W[2], P.H[2], w[2], base[2], msk[2]);
        if ((msk[0] | msk[1] | msk[2]) == 0) continue;
        for (int ch = 0; ch < nchunk; ++ch) {
          const int c = ch * 32 + l8 * 4;
          if (c >= C) continue;
          const float* go = BP.gout + ((int64_t)b * C + c) * P.Q + q;
          const float4 g = make_float4(__ldg(go), __ldg(go + P.Q), __ldg(go + 2 * P.Q), __ldg(go + 3 * P.Q));
          scatter_taps(gp0 + ch * 8, base[0] * C4, C4, WC4_0, w[0], msk[0], g);
          scatter_taps(gp1 + ch * 8, base[1] * C4, C4, WC4_1, w[1], msk[1], g);
          scatter_taps(gp2 + ch * 8, base[2] * C4, C4, WC4_2, w[2], msk[2], g);
        }
      }
      __syncthreads();
      continue;
    }
    if ((vote & 7) == 0) continue;  // nothing of the block is inside any plane: no gradient (records are rewritten
                                    // after the next block's barrier-free phase A only; see the barrier below)

    const bool jk_ok = (jj < nj) && (kg * 4 < nk);
    for (int ch = 0; ch < nchunk; ++ch) {
      // ---- R: grad_out of the block -> sums over k (xy), i (yz), j (xz), 32 channels ---------------------------
      const int cmax = min(32, C - ch * 32);
      for (int c = warp; c < 32; c += kGridThreads / 32) {
        const int xk = (kg * 4) ^ swz_bits(c);
        const float* go = BP.gout + ((int64_t)b * C + ch * 32 + c) * P.Q + qblk + jj * G.d + kg * 4;
        // all BI rows of this channel in flight before the first reduction (the loop below is a chain of shuffles)
        float4 gr[BI];
#pragma unroll
        for (int ii = 0; ii < BI; ++ii) {
          gr[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (jk_ok && ii < ni && c < cmax) gr[ii] = __ldcs(reinterpret_cast<const float4*>(go + (int64_t)ii * wd));
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ii = 0; ii < BI; ++ii) {
          const float4 g = gr[ii];
          acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
          // xz(i, k): sum over the 8 j of the block = lanes that differ in bits 2-4
          float4 s = g;
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
            s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
            s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
          }
          if (jj == 0) *reinterpret_cast<float4*>(T2 + c * Cfg::S2 + ii * kBK + xk) = s;
          // xy(i, j): sum over the 16 k = 4 components x the 4 lanes that differ in bits 0-1
          float t = (g.x + g.y) + (g.z + g.w);
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          if (kg == 0) T0[c * Cfg::S0 + ii * BJ + jj] = t;
        }
        *reinterpret_cast<float4*>(T1 + c * Cfg::S1 + jj * kBK + xk) = acc;  // yz(j, k): sum over i
      }
      __syncthreads();
      // ---- S: one scatter per table entry and tap ---------------------------------------------------------------
      const bool cvalid = ch * 32 + l8 * 4 < C;
      scatter_table<Cfg::E0 / 32>(vote & 1, gp0 + ch * 8, C4, WC4_0, s_w + ent, s_om + ent, r0, Cfg::S0, cvalid);
      scatter_table<Cfg::E1 / 32>(vote & 2, gp1 + ch * 8, C4, WC4_1, s_w + Cfg::E0 + ent, s_om + Cfg::E0 + ent, r1, Cfg::S1, cvalid);
      scatter_table<Cfg::E2 / 32>(vote & 4, gp2 + ch * 8, C4, WC4_2, s_w + Cfg::E0 + Cfg::E1 + ent, s_om + Cfg::E0 + Cfg::E1 + ent, r2, Cfg::S2, cvalid);
      __syncthreads();  // tables (next chunk) and records (next block) are rewritten
    }
  }
}

template <int ARITH, int BI>
static void launch_grid_bwd(const GridBwdParams& BP, cudaStream_t s) {
  const int nb = BP.G.nblocks;
  const int cap = kBwdGridCtasPerSm * kSMs;
  auto kern = sample3_grid_backward_kernel<ARITH, BI>;
  opt_in_smem<sample3_grid_backward_kernel<ARITH, BI>>(GridCfg<BI>::kSmemBytes);  // a failure surfaces at the launch check
  kern<<<(unsigned)(nb < cap ? nb : cap), kGridThreads, GridCfg<BI>::kSmemBytes, s>>>(BP);
}

}  // namespace tp

using namespace tp;

extern "C" int tp_sample3_grid_backward_nhwc_f32(const tp_plane gplanes_nhwc[3], int32_t C, const float* queries,
                                                 const int32_t dims[3], int32_t batch, const tp_sample_geom* sg,
                                                 int32_t arith, const float* grad_out, void* stream) {
  if (!dims) return fail(TP_E_NULL, "tp_sample3_grid_backward_nhwc_f32: null dims");
  const int h = dims[0], w = dims[1], d = dims[2];
  if (h < 0 || w < 0 || d < 0) return fail(TP_E_SHAPE, "tp_sample3_grid_backward_nhwc_f32: bad dims %d %d %d", h, w, d);
  const int64_t Q = (int64_t)h * w * d;
  // the lattice path reads grad_out in 16-byte k-runs; anything else goes through the per-query kernel
  if ((d & 3) || (reinterpret_cast<uintptr_t>(grad_out) & 15) || Q == 0 || Q * 3 >= ((int64_t)1 << 31))
    return tp_sample3_backward_nhwc_f32(gplanes_nhwc, C, queries, Q, batch, sg, arith, grad_out, stream);
  if (C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "tp_sample3_grid_backward_nhwc_f32: C=%d must be a positive multiple of 4", C);
  if (batch <= 0) return fail(TP_E_SHAPE, "tp_sample3_grid_backward_nhwc_f32: bad B=%d", batch);
  if (!gplanes_nhwc || !queries || !grad_out || !sg) return fail(TP_E_NULL, "tp_sample3_grid_backward_nhwc_f32: null argument");
  if (arith != TP_ARITH_TORCH_CUDA && arith != TP_ARITH_TORCH_CPU) return fail(TP_E_ENUM, "tp_sample3_grid_backward_nhwc_f32: unknown arith %d", arith);
  GridBwdParams BP;
  GridParams& G = BP.G;
  SampleParams& P = G.S;
  for (int k = 0; k < 3; ++k) {
    if (!gplanes_nhwc[k].data) return fail(TP_E_NULL, "tp_sample3_grid_backward_nhwc_f32: plane %d is null", k);
    if (gplanes_nhwc[k].H <= 0 || gplanes_nhwc[k].W <= 0 ||
        (int64_t)gplanes_nhwc[k].H * gplanes_nhwc[k].W * C >= (int64_t)1 << 31)
      return fail(TP_E_SHAPE, "tp_sample3_grid_backward_nhwc_f32: plane %d shape unsupported", k);
    if ((uintptr_t)gplanes_nhwc[k].data & 15 || (gplanes_nhwc[k].batch_stride & 3))
      return fail(TP_E_SHAPE, "tp_sample3_grid_backward_nhwc_f32: plane %d not 16-byte aligned", k);
    P.plane[k] = nullptr;
    BP.gplane[k] = const_cast<float*>(gplanes_nhwc[k].data);
    P.bstride[k] = gplanes_nhwc[k].batch_stride;
    P.H[k] = gplanes_nhwc[k].H;
    P.W[k] = gplanes_nhwc[k].W;
    P.lo[k] = sg->lo[k];
    P.vs[k] = sg->vs[k];
    P.rcp_vs[k] = 1.0f / sg->vs[k];
    P.half[k] = sg->half[k];
    P.rcp_half[k] = 1.0f / sg->half[k];
  }
  P.queries = queries;
  P.out = nullptr;
  BP.gout = grad_out;
  P.Q = Q;
  P.C = C;
  P.tiles_per_sample = 0;
  P.tiles = 0;
  G.h = h; G.w = w; G.d = d;
  G.vec_ok = 1;
  G.nkb = (d + kBK - 1) / kBK;
  auto nblocks = [&](int bi) { return (int64_t)batch * ((h + bi - 1) / bi) * ((w + kBJ - 1) / kBJ) * G.nkb; };
  const int bi = nblocks(8) >= kBwdGridCtasPerSm * kSMs ? 8 : 4;
  G.nib = (h + bi - 1) / bi;
  G.njb = (w + kBJ - 1) / kBJ;
  if (nblocks(bi) >= ((int64_t)1 << 30)) return fail(TP_E_SHAPE, "tp_sample3_grid_backward_nhwc_f32: too many queries");
  G.nblocks = (int)nblocks(bi);
  cudaStream_t s = (cudaStream_t)stream;
  if (arith == TP_ARITH_TORCH_CUDA) {
    if (bi == 8) launch_grid_bwd<TP_ARITH_TORCH_CUDA, 8>(BP, s); else launch_grid_bwd<TP_ARITH_TORCH_CUDA, 4>(BP, s);
  } else {
    if (bi == 8) launch_grid_bwd<TP_ARITH_TORCH_CPU, 8>(BP, s); else launch_grid_bwd<TP_ARITH_TORCH_CPU, 4>(BP, s);
  }
  TP_LAUNCH_CHECK("sample3_grid_backward_kernel");
  return 0;
}
