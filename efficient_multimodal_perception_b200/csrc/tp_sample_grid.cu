// Triplane decode for [B,h,w,d,3] query tensors (triplane_occ.py:321-348, triplane_elev.py:286-313,
// point_triplane_occ.py:407-440): the reference's 5-D callers always pass a voxel-centre lattice
// (roi(), triplane_occ.py:291-318; get_reference_points(), triplane_elev.py:113-133) in which x
// depends only on the h index, y only on w and z only on d. Then
//     xy(i,j)  yz(j,k)  xz(i,k)
// are functions of two lattice indices each, and out[c,i,j,k] = (xy[c,i,j] + yz[c,j,k]) + xz[c,i,k]
// with exactly the per-plane values and the summation order of the per-query kernel: bit-identical
// output, 0.2-0.6 bilinear footprints per query instead of 3, and the kernel becomes a streaming
// write of the result.
//
// Nothing is assumed: every CTA takes a BI x BJ x 16 block of the lattice, loads its queries (they
// are read once, as the algorithmic-bytes model says) and checks bit-for-bit that the block is
// separable. A block that is not (jittered points, a permuted tensor, ...) falls back, inside the
// same launch, to the per-query tile routine of tp_sample.cu.
#include "tp_sample_grid.cuh"

namespace tp {

// The per-query path for a block that is not a lattice. Kept out of line: its register needs (the flat kernel's
// tile routine) must not shape the register allocation of the table path.
template <int ARITH, int C4T, int BI>
__device__ __noinline__ void grid_fallback(const GridParams& G, int b, int i0, int j0, int k0, float* smem) {
  constexpr int BJ = kBJ;
  const SampleParams& P = G.S;
  const int warp = threadIdx.x >> 5;
  if (warp >= kFallbackWarps) return;
  const unsigned long long pol_planes = policy_evict_last(), pol_out = policy_evict_first();
  const int ni = min(BI, G.h - i0), nj = min(BJ, G.w - j0), nk = min(kBK, G.d - k0);
  float* sp = smem + warp * (kParamWords + kTileWords);
  float* st = sp + kParamWords;
  for (int t = warp; t < BI * BJ / 2; t += kFallbackWarps) {
    const int r0 = 2 * t, r1 = 2 * t + 1;
    const int ia0 = r0 / BJ, ja0 = r0 % BJ, ia1 = r1 / BJ, ja1 = r1 % BJ;
    const int n0 = (ia0 < ni && ja0 < nj) ? nk : 0, n1 = (ia1 < ni && ja1 < nj) ? nk : 0;
    if (n0 == 0 && n1 == 0) continue;
    const int64_t run0 = ((int64_t)(i0 + ia0) * G.w + j0 + ja0) * G.d + k0;
    const int64_t run1 = ((int64_t)(i0 + ia1) * G.w + j0 + ja1) * G.d + k0;
    sample_tile<ARITH, C4T>(P, b, run0, run1, n0, n1, sp, st, pol_planes, pol_out, G.vec_ok != 0);
  }
}

// Persistent CTAs: block n+1's queries are prefetched into L2 while block n is gathered and written,
// so the only DRAM-latency-bound step of a block (reading its 12 B/query) is off the critical path.
template <int ARITH, int C4T, int BI, bool GEN>
__global__ void __launch_bounds__(kGridThreads, kGridCtasPerSm)
sample3_grid_kernel(const __grid_constant__ GridParams G) {
  using Cfg = GridCfg<BI>;
  constexpr int BJ = kBJ;
  constexpr int QPT = BI * BJ * kBK / kGridThreads;  // queries per thread
  extern __shared__ __align__(16) float smem[];  // Cfg::kSmemBytes: tables (or fallback tiles) + entry records
  float4* const s_w = reinterpret_cast<float4*>(smem + Cfg::kWords);  // bilinear weights (nw, ne, sw, se) per entry
  int2* const s_om = reinterpret_cast<int2*>(s_w + Cfg::E);  // {nw tap offset in float4 units, in-bounds mask}

  const SampleParams& P = G.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C4 = C4T ? C4T : (P.C >> 2);
  const int C = C4 * 4;
  const unsigned long long pol_planes = policy_evict_last(), pol_out = policy_evict_first();
  const int wd = G.w * G.d;  // h*w*d*3 < 2^31 (host-checked): in-sample query offsets fit 32 bits
  const int nblocks = G.nblocks;

  float* const T0 = smem;
  float* const T1 = T0 + 32 * Cfg::S0;
  float* const T2 = T1 + 32 * Cfg::S1;
  const int l8 = tid & 7, ent = tid >> 3;
  const int WC4_0 = P.W[0] * C4, WC4_1 = P.W[1] * C4, WC4_2 = P.W[2] * C4;
  const int nchunk = (C4 + 7) >> 3;
  // table write positions of this thread: rows 4*l8 .. 4*l8+3, column = entry (swizzled for yz / xz)
  float* const w0 = T0 + (4 * l8) * Cfg::S0 + ent;
  float* const w1 = T1 + (4 * l8) * Cfg::S1 + (ent ^ swz_bits(4 * l8));
  float* const w2 = T2 + (4 * l8) * Cfg::S2 + (ent ^ swz_bits(4 * l8));
  // query ownership in phase A: k and j fixed per thread, i = ia + t * (256 / 128)
  const int ak = tid & (kBK - 1), aj = (tid / kBK) % BJ, ia = tid / (kBK * BJ);
  // output phase: lane -> (j, 4 consecutive k)
  const int kg = lane & 3, jj = lane >> 2;

  int zeroed = 0;  // bit p: table p holds zeros from an earlier block / chunk
  __shared__ int s_vote[2];
  int nblk_done = 0;
  if (tid == 0) s_vote[0] = 0;
  __syncthreads();
  GRID_G(0);
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    if (blk != (int)blockIdx.x) GRID_G(1);
    const BlockPos bp = block_pos<BI>(G, blk);  // recomputed per block: carrying it across the loop spills
    const int b = bp.b, i0 = bp.i0, j0 = bp.j0, k0 = bp.k0;
    const int ni = min(BI, G.h - i0), nj = min(BJ, G.w - j0), nk = min(kBK, G.d - k0);
    const float* q00 = P.queries + ((int64_t)b * P.Q + ((int64_t)i0 * G.w + j0) * G.d + k0) * 3;  // block origin

    // ---- A: read the block's queries once; is x = x(i), y = y(j), z = z(k) bit for bit? --------
    const bool ok = GEN ? true : grid_block_is_lattice<BI>(G, q00, ni, nj, nk, wd, tid);
    // ---- B: one bilinear footprint per table entry (index pair), from the representative queries -
    const int live = grid_build_records<ARITH, BI, GEN>(G, q00, ni, nj, nk, wd, C4, s_w, s_om, tid, i0, j0, k0);
    grid_cast_vote(s_vote, nblk_done, live, ok, tid);
    __syncthreads();
    const int vote = s_vote[nblk_done & 1];
    ++nblk_done;
    const bool separable = !(vote & 8);
    if (blk == (int)blockIdx.x) GRID_G(4);

    // ---- prefetch the next block's queries (DRAM -> L2) behind this block's gathers and stores ---
    const int next = blk + gridDim.x;
    if (!GEN && next < nblocks) {
      const BlockPos np = block_pos<BI>(G, next);
      if (np.j0 + aj < G.w && np.k0 + ak < G.d) {
        const float* n00 = P.queries + ((int64_t)np.b * P.Q + ((int64_t)np.i0 * G.w + np.j0) * G.d + np.k0) * 3;
        // one lane per 32 bytes of the (j,k) run is plenty: 12 B per query
        if ((ak & 1) == 0) {
#pragma unroll
          for (int t = 0; t < QPT; ++t) {
            const int ii = ia + t * (kGridThreads / (kBK * BJ));
            if (np.i0 + ii < G.h) prefetch_l2(n00 + (ii * wd + aj * G.d + ak) * 3);
          }
        }
      }
    }

    if (!separable) {
      // ---- per-query fallback: the flat kernel's tile routine on pairs of (i,j) columns ---------
      grid_fallback<ARITH, C4T, BI>(G, b, i0, j0, k0, smem);
      zeroed = 0;
      __syncthreads();  // the fallback tiles alias the next block's tables
    } else {
      const float4* const pl0 = reinterpret_cast<const float4*>(P.plane[0] + (int64_t)b * P.bstride[0]) + l8;
      const float4* const pl1 = reinterpret_cast<const float4*>(P.plane[1] + (int64_t)b * P.bstride[1]) + l8;
      const float4* const pl2 = reinterpret_cast<const float4*>(P.plane[2] + (int64_t)b * P.bstride[2]) + l8;
      const bool jk_ok = (jj < nj) && (kg * 4 < nk);
      if (G.pdl && (vote & 7)) pdl_wait();  // first read of the planes; a block outside all of them never waits
      float* const o_jk = P.out + (int64_t)b * C * P.Q + ((int64_t)i0 * G.w + j0 + jj) * G.d + k0 + kg * 4;

      for (int ch = 0; ch < nchunk; ++ch) {
        // ---- C: the three 2-D tables for 32 channels: 8 lanes x 16 B per bilinear tap -----------
        // two entries per thread in flight: 8 independent 16-byte loads before the first fma
        const bool cvalid = ch * 32 + l8 * 4 < C;
        // block-uniform: a plane nothing of which is in range is not touched at all
        build_table<Cfg::E0 / 32>(vote & 1, zeroed & 1, pl0 + ch * 8, C4, WC4_0, s_w + ent, s_om + ent, w0, Cfg::S0, cvalid, pol_planes);
        build_table<Cfg::E1 / 32>(vote & 2, zeroed & 2, pl1 + ch * 8, C4, WC4_1, s_w + Cfg::E0 + ent, s_om + Cfg::E0 + ent, w1, Cfg::S1, cvalid, pol_planes);
        build_table<Cfg::E2 / 32>(vote & 4, zeroed & 4, pl2 + ch * 8, C4, WC4_2, s_w + Cfg::E0 + Cfg::E1 + ent, s_om + Cfg::E0 + Cfg::E1 + ent, w2, Cfg::S2, cvalid, pol_planes);
        zeroed = ~vote & 7;  // tables that hold zeros now (and keep them while their plane stays out of range)
        __syncthreads();
        if (blk == (int)blockIdx.x) GRID_G(5);

        // ---- D: out[c, i, j, k..k+3] = (xy[c,i,j] + yz[c,j,k..]) + xz[c,i,k..], 16-byte stores ---
        const int cmax = min(32, C - ch * 32);
        if ((vote & 7) == 0) {  // (0 + 0) + 0: the whole block lies outside all three planes
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int c = warp; c < cmax; c += kGridThreads / 32) {
            float* o = o_jk + (int64_t)(ch * 32 + c) * P.Q;
#pragma unroll
            for (int ii = 0; ii < BI; ++ii, o += wd) st_out_f4(o, z4, pol_out, jk_ok && ii < ni);
          }
        } else
        for (int c = warp; c < cmax; c += kGridThreads / 32) {
          const int xk = (kg * 4) ^ swz_bits(c);
          const float4 s1 = *reinterpret_cast<const float4*>(T1 + c * Cfg::S1 + jj * kBK + xk);
          const float* t0 = T0 + c * Cfg::S0 + jj;
          const float* t2 = T2 + c * Cfg::S2 + xk;
          float* o = o_jk + (int64_t)(ch * 32 + c) * P.Q;
#pragma unroll
          for (int ii = 0; ii < BI; ++ii, o += wd) {
            const float s0 = t0[ii * BJ];
            const float4 s2 = *reinterpret_cast<const float4*>(t2 + ii * kBK);
            float4 r;
            r.x = __fadd_rn(__fadd_rn(s0, s1.x), s2.x);  // (xy + yz) + xz  (triplane_occ.py:345)
            r.y = __fadd_rn(__fadd_rn(s0, s1.y), s2.y);
            r.z = __fadd_rn(__fadd_rn(s0, s1.z), s2.z);
            r.w = __fadd_rn(__fadd_rn(s0, s1.w), s2.w);
            st_out_f4(o, r, pol_out, jk_ok && ii < ni);
          }
        }
        if (ch + 1 < nchunk) __syncthreads();
        if (blk == (int)blockIdx.x) GRID_G(6);
      }
    }
  }
  GRID_G(2);
}

template <int ARITH, int C4T, int BI, bool GEN>
static void launch_grid(const GridParams& G, int batch, cudaStream_t s, bool pdl) {
  // persistent CTAs (kGridCtasPerSm per SM)
  const int64_t blocks = G.nblocks < kGridCtasPerSm * kSMs ? G.nblocks : kGridCtasPerSm * kSMs;
  auto kern = sample3_grid_kernel<ARITH, C4T, BI, GEN>;
  opt_in_smem<sample3_grid_kernel<ARITH, C4T, BI, GEN>>(GridCfg<BI>::kSmemBytes);  // a failure surfaces at the launch check
  launch_kernel(kern, (unsigned)blocks, kGridThreads, GridCfg<BI>::kSmemBytes, s, pdl, G);
}

}  // namespace tp

using namespace tp;

// queries == nullptr: generated lattice (lat_org / lat_step), no fallback to the per-query kernel
static int grid_entry(const tp_plane planes[3], int32_t C, const float* queries, const float* lat_org,
                      const float* lat_step, const int32_t dims[3], int32_t batch, const tp_sample_geom* sg,
                      int32_t arith, float* out, void* stream, const tp_plane* planes_nchw = nullptr,
                      float* const* nhwc_dst = nullptr) {
  // planes_nchw != nullptr: `planes` describes the (not yet written) channels-last workspace copies nhwc_dst[3]
  const bool gen = queries == nullptr;
  bool pdl = false;
  if (!dims) return fail(TP_E_NULL, "tp_sample3_grid_nhwc_f32: null dims");
  const int h = dims[0], w = dims[1], d = dims[2];
  if (h < 0 || w < 0 || d < 0) return fail(TP_E_SHAPE, "tp_sample3_grid_nhwc_f32: bad dims %d %d %d", h, w, d);
  const int64_t Q = (int64_t)h * w * d;
  // lattice path needs 16-byte aligned k-runs; anything else goes through the per-query kernel
  if ((d & 3) || (reinterpret_cast<uintptr_t>(out) & 15) || arith == 2 || Q == 0 || Q * 3 >= ((int64_t)1 << 31)) {
    if (gen) return Q == 0 ? 0 : fail(TP_E_SHAPE, "tp_sample3_lattice_nhwc_f32: needs d %% 4 == 0, a 16-byte aligned out and h*w*d*3 < 2^31 (got %d %d %d)", h, w, d);
    if (Q == 0) return 0;
    if (planes_nchw)
      if (int rc = tp_planes3_nchw_to_nhwc_f32(planes_nchw, nhwc_dst, batch, C, stream)) return rc;
    return sample3_flat(planes, C, queries, Q, batch, sg, arith, out, stream, planes_nchw != nullptr);
  }
  if (C <= 0 || (C & 3)) return fail(TP_E_SHAPE, "tp_sample3_grid_nhwc_f32: C=%d must be a positive multiple of 4", C);
  if (batch <= 0) return fail(TP_E_SHAPE, "tp_sample3_grid_nhwc_f32: bad B=%d", batch);
  if (!planes || !out || !sg || (gen && (!lat_org || !lat_step))) return fail(TP_E_NULL, "tp_sample3_grid_nhwc_f32: null argument");
  GridParams G;
  for (int a = 0; a < 3; ++a) {
    G.gen_org[a] = gen ? lat_org[a] : 0.f;
    G.gen_step[a] = gen ? lat_step[a] : 0.f;
  }
  SampleParams& P = G.S;
  for (int k = 0; k < 3; ++k) {
    if (!planes[k].data) return fail(TP_E_NULL, "tp_sample3_grid_nhwc_f32: plane %d is null", k);
    if (planes[k].H <= 0 || planes[k].W <= 0 ||
        (int64_t)planes[k].H * planes[k].W * C >= (int64_t)1 << 31 || planes[k].H >= (1 << 20) ||
        planes[k].W >= (1 << 20))
      return fail(TP_E_SHAPE, "tp_sample3_grid_nhwc_f32: plane %d H=%d W=%d unsupported", k, planes[k].H, planes[k].W);
    if ((uintptr_t)planes[k].data & 15 || (planes[k].batch_stride & 3))
      return fail(TP_E_SHAPE, "tp_sample3_grid_nhwc_f32: plane %d not 16-byte aligned", k);
    P.plane[k] = planes[k].data;
    P.bstride[k] = planes[k].batch_stride;
    P.H[k] = planes[k].H;
    P.W[k] = planes[k].W;
    P.lo[k] = sg->lo[k];
    P.vs[k] = sg->vs[k];
    P.rcp_vs[k] = 1.0f / sg->vs[k];
    P.half[k] = sg->half[k];
    P.rcp_half[k] = 1.0f / sg->half[k];
  }
  P.queries = queries;
  P.out = out;
  P.Q = Q;
  P.C = C;
  P.tiles_per_sample = 0;
  P.tiles = 0;
  G.h = h; G.w = w; G.d = d;
  G.vec_ok = 1;
  G.nkb = (d + kBK - 1) / kBK;
  // block shape BI x 8 x 16: the larger one shares each yz footprint between 8 lattice rows instead
  // of 4; the smaller one is for grids that would not give every resident CTA a block
  auto nblocks = [&](int bi) { return (int64_t)batch * ((h + bi - 1) / bi) * ((w + kBJ - 1) / kBJ) * G.nkb; };
  const int cfg = nblocks(8) >= kGridCtasPerSm * kSMs ? 0 : 1;
  const int bi = cfg == 0 ? 8 : 4;
  G.nib = (h + bi - 1) / bi;
  G.njb = (w + kBJ - 1) / kBJ;
  if (nblocks(bi) >= ((int64_t)1 << 30)) return fail(TP_E_SHAPE, "tp_sample3_grid_nhwc_f32: too many queries");
  G.nblocks = (int)nblocks(bi);
  cudaStream_t s = (cudaStream_t)stream;
  if (planes_nchw) {
    // Conversion launch first. The decode becomes its programmatic dependent only when every CTA gets a single block
    // (roi: 13.4 vs 15.1 us): dependents are placed on the SMs in the order the conversion CTAs retire, and the
    // persistent grid's static block -> CTA map then no longer spreads the expensive blocks (all three planes in
    // range) evenly over the SMs — 640k lattice 28.5 us as a dependent vs 27.0 us in plain stream order, wherever the
    // wait is placed (top of the kernel, after the first block's footprints, before the first gather).
    if (int rc = tp_planes3_nchw_to_nhwc_f32(planes_nchw, nhwc_dst, batch, C, stream)) return rc;
    pdl = G.nblocks <= kGridCtasPerSm * kSMs;
  }
  G.pdl = pdl ? 1 : 0;
#define TP_GRID_G(A, C4T, BI) \
  if (gen) launch_grid<A, C4T, BI, true>(G, batch, s, pdl); else launch_grid<A, C4T, BI, false>(G, batch, s, pdl);
#define TP_GRID_B(A, C4T) \
  switch (cfg) { case 0: TP_GRID_G(A, C4T, 8) break; default: TP_GRID_G(A, C4T, 4) break; }
#define TP_GRID_C(A)                                                      \
  switch (C) { case 32: TP_GRID_B(A, 8) break; case 96: TP_GRID_B(A, 24) break; \
               case 128: TP_GRID_B(A, 32) break; default: TP_GRID_B(A, 0) break; }
  switch (arith) {
    case TP_ARITH_TORCH_CUDA: TP_GRID_C(TP_ARITH_TORCH_CUDA) break;
    case TP_ARITH_TORCH_CPU:  TP_GRID_C(TP_ARITH_TORCH_CPU) break;
    default: return fail(TP_E_ENUM, "tp_sample3_grid_nhwc_f32: unknown arith %d", arith);
  }
#undef TP_GRID_C
#undef TP_GRID_B
#undef TP_GRID_G
  TP_LAUNCH_CHECK("sample3_grid_kernel");
  return 0;
}

extern "C" int tp_sample3_grid_nhwc_f32(const tp_plane planes[3], int32_t C, const float* queries,
                                        const int32_t dims[3], int32_t batch, const tp_sample_geom* sg,
                                        int32_t arith, float* out, void* stream) {
  if (!queries) return fail(TP_E_NULL, "tp_sample3_grid_nhwc_f32: null queries");
  return grid_entry(planes, C, queries, nullptr, nullptr, dims, batch, sg, arith, out, stream);
}

extern "C" int tp_sample3_lattice_nhwc_f32(const tp_plane planes[3], int32_t C, const int32_t dims[3],
                                           const float origin[3], const float step[3], int32_t batch,
                                           const tp_sample_geom* sg, int32_t arith, float* out, void* stream) {
  if (!origin || !step) return fail(TP_E_NULL, "tp_sample3_lattice_nhwc_f32: null lattice");
  return grid_entry(planes, C, nullptr, origin, step, dims, batch, sg, arith, out, stream);
}

extern "C" int tp_sample3_grid_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const float* queries,
                                        const int32_t dims[3], int32_t batch, const tp_sample_geom* sg,
                                        int32_t arith, float* out, float* ws, int64_t ws_floats,
                                        void* stream) {
  if (!queries) return fail(TP_E_NULL, "tp_sample3_grid_nchw_f32: null queries");
  tp_plane nhwc[3];
  float* dsts[3];
  if (int rc = planes3_workspace("tp_sample3_grid_nchw_f32", planes_nchw, C, batch, ws, ws_floats, nhwc, dsts)) return rc;
  return grid_entry(nhwc, C, queries, nullptr, nullptr, dims, batch, sg, arith, out, stream, planes_nchw, dsts);
}

extern "C" int tp_sample3_lattice_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const int32_t dims[3],
                                           const float origin[3], const float step[3], int32_t batch,
                                           const tp_sample_geom* sg, int32_t arith, float* out, float* ws,
                                           int64_t ws_floats, void* stream) {
  if (!origin || !step) return fail(TP_E_NULL, "tp_sample3_lattice_nchw_f32: null lattice");
  tp_plane nhwc[3];
  float* dsts[3];
  if (int rc = planes3_workspace("tp_sample3_lattice_nchw_f32", planes_nchw, C, batch, ws, ws_floats, nhwc, dsts)) return rc;
  return grid_entry(nhwc, C, nullptr, origin, step, dims, batch, sg, arith, out, stream, planes_nchw, dsts);
}
