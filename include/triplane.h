/*
 * triplane.h — C ABI of libtriplane.so: the B200 (sm_100a) triplane hot path.
 *
 * The reference (charyyev/efficient_multimodal_perception) has NO FFI for this path: its boundary
 * is a set of Python methods that call torch / torch_scatter / spconv ops. Each entry point below
 * names the reference call site (file:line under /root/reference) whose arithmetic it replaces.
 * The Python modules in efficient_multimodal_perception_b200/ bind these with ctypes and keep the
 * reference's method names and forward() signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host` or the comment says host;
 *   - the caller allocates every output and workspace buffer; nothing here allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never
 *     synchronise the device (the *_host entry points are the exception and say so);
 *   - return value: 0 = ok, <0 = argument error (TP_E_*), >0 = a cudaError_t; tp_last_error()
 *     returns a thread-local message for the last non-zero return;
 *   - all arithmetic is fp32 / int32, compiled WITHOUT --use_fast_math; the coordinate chain
 *     replays the reference's op sequence (see tp_arith).
 */
#ifndef TRIPLANE_H_
#define TRIPLANE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TP_VERSION 1

#if defined(__GNUC__)
#define TP_API __attribute__((visibility("default")))
#else
#define TP_API
#endif

/* argument errors */
#define TP_E_NULL      (-1)  /* required pointer is NULL */
#define TP_E_SHAPE     (-2)  /* size / stride / channel constraint violated */
#define TP_E_ENUM      (-3)  /* unknown enum value */
#define TP_E_WORKSPACE (-4)  /* workspace too small */

/*
 * Which PyTorch device's op sequence the coordinate chain replays. The reference writes
 *   v = (p - lo) / vs ;  g = v / (S/2) - 1            (triplane_occ.py:333-337, point_triplane.py:451-458)
 * with Python-float divisors. torch-CUDA evaluates `tensor / python_float` as a multiply by the fp32
 * reciprocal, torch-CPU as a true division; ATen's grid_sampler un-normalises as
 * ((g+1)*W-1)/2 on CUDA and (g+1)*(W/2)-0.5 on CPU. Indices differ in rare ulp cases, so the
 * caller says which one it wants to be bit-compatible with.
 */
typedef enum tp_arith {
  TP_ARITH_TORCH_CUDA = 0, /* x * fp32(1/d); unnormalise ((g+1)*W-1)/2 (fma-contracted as nvcc does) */
  TP_ARITH_TORCH_CPU  = 1  /* x / d (IEEE);  unnormalise (g+1)*(W/2) - 0.5                           */
} tp_arith;

typedef enum tp_reduce {
  TP_REDUCE_MAX  = 0, /* reference: torch_scatter.scatter_max + SparseMaxPool3d (projector.py:104,113-115) */
  TP_REDUCE_MEAN = 1, /* north-star extension: sum / count per pooled cell                               */
  TP_REDUCE_SUM  = 2, /* partial sums for the point-sharded multi-GPU mean (divide after the all-reduce)  */
  TP_REDUCE_MAX_PARTIAL = 3 /* point-sharded multi-GPU max: empty cells are -inf so that an all-reduce(max)
                               over GPUs is exact; tp_encode_finalize_max_f32 then maps -inf -> 0          */
} tp_reduce;

/* Voxel geometry: pc_range / voxel_size / grid_size of configs/point_triplane.py:8-10. Host struct. */
typedef struct tp_geom {
  float lo[3];       /* pc_range[0:3]                                   */
  float hi[3];       /* pc_range[3:6]                                   */
  float vs[3];       /* voxel_size                                      */
  int32_t grid[3];   /* grid_size (X, Y, Z)                             */
  int32_t pool[3];   /* pooling kernel (kx, ky, kz) = int(grid/split), projector.py:53-58 */
} tp_geom;

/* One feature plane for decode. */
typedef struct tp_plane {
  const float* data;     /* NCHW [B, C, H, W] or NHWC [B, H, W, C] depending on the entry point */
  int64_t batch_stride;  /* elements between consecutive samples (3*C*H*W for the stacked [B,3,C,H,W]) */
  int32_t H, W;          /* grid_sample: first coordinate of the pair -> W (last dim), second -> H */
} tp_plane;

/* Per-axis affine for decode: v_a = (p_a - lo_a) / vs_a ; g_a = v_a / half_a - 1  (half_a = S_a / 2) */
typedef struct tp_sample_geom {
  float lo[3];
  float vs[3];
  float half[3];
} tp_sample_geom;

TP_API const char* tp_last_error(void);
TP_API int tp_version(void);
/* number of SMs of the current device (grid sizing); <0 on error */
TP_API int tp_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * a1  voxelize_points — point_triplane.py:133-161 (twin point_triplane_occ.py:132-160)
 *
 * points  [n_total, point_stride] fp32, samples concatenated; in_offsets [B+1] int64 (device).
 * Stable compaction of the points with lo < p < hi (strict, all three axes), copying `ncols`
 * columns, and idx = int32(trunc((p - lo) (/) vs)) per tp_arith. Order inside a sample is kept
 * (the reference's boolean-mask indexing). out_offsets [B+1] int64 (device) receives the
 * compacted sample boundaries; out_offsets[B] = N'.
 * workspace: tp_voxelize_workspace_bytes(n_total) bytes.
 * ------------------------------------------------------------------------------------------- */
TP_API int64_t tp_voxelize_workspace_bytes(int64_t n_total);
TP_API int tp_voxelize_f32(const float* points, int64_t n_total, int32_t point_stride, int32_t ncols,
                    const int64_t* in_offsets, int32_t batch,
                    const tp_geom* geom, int32_t arith,
                    float* out_points /* [>=N', ncols] */, int32_t* out_idx /* [>=N', 3] */,
                    int64_t* out_offsets /* [B+1] */,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* uncompacted variant: keep[n] (uint8) and idx[n,3] for every raw point (idx undefined where !keep) */
TP_API int tp_voxel_index_f32(const float* points, int64_t n_total, int32_t point_stride,
                       const tp_geom* geom, int32_t arith,
                       uint8_t* keep, int32_t* idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3  triplane encode — point_triplane_projector.py:99-115
 *     torch.unique -> torch_scatter.scatter_max -> 3 x SparseMaxPool3d -> .dense()
 *     -> permute(...).flatten(3), fused: max over all points of a pooled cell, written once.
 *
 * feats   [n_total, C] fp32 (row stride feat_stride elements), C % 4 == 0
 * Either idx [n_total,3] int32 (the reference's grid_ind, batch by offsets) or, when idx == NULL,
 * xyz points [n_total, point_stride] from which crop + index are computed in-kernel (fused a1).
 * offsets [B+1] int64 device: sample boundaries.
 * Outputs, channels-last, exactly the post-permute/flatten layouts of projector.py:113-115:
 *   out_xy [B, X, Y, Zp*C]   out_yz [B, Y, Z, Xp*C]   out_xz [B, X, Z, Yp*C]
 * with Xp = (X-kx)/kx+1 etc. Points whose pooled index falls outside the pooled extent are dropped
 * for that plane (spconv stride/kernel semantics); points with idx outside [0,grid) are dropped.
 * Empty cells are 0. clamp_zero != 0 -> max(0, .) (spconv native-path ambiguity).
 * cell_count (optional, may be NULL): int32 [cells] number of points per pooled cell, same cell order
 * as the three outputs concatenated (xy, yz, xz).
 * Any of out_xy/out_yz/out_xz may be NULL to skip that plane.
 * workspace: tp_encode_workspace_bytes() bytes; it must be zeroed once before the FIRST call
 * (tp_encode_workspace_init) and every call leaves its tile counters zero again. One workspace per
 * stream: calls sharing a workspace must be stream-ordered.
 * ------------------------------------------------------------------------------------------- */
TP_API int64_t tp_encode_cells(const tp_geom* geom, int32_t batch, int64_t cells_per_plane[3]);
TP_API int64_t tp_encode_workspace_bytes(const tp_geom* geom, int32_t batch, int64_t n_total);
TP_API int tp_encode_workspace_init(void* workspace, int64_t workspace_bytes, void* stream);
TP_API int tp_encode_f32(const float* feats, int64_t feat_stride, int32_t C,
                  const int32_t* idx, const float* points, int32_t point_stride,
                  int64_t n_total, const int64_t* offsets, int32_t batch,
                  const tp_geom* geom, int32_t arith, int32_t reduce, int32_t clamp_zero,
                  float* out_xy, float* out_yz, float* out_xz, int32_t* cell_count,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* 8f#3 (i)  PointTriplaneProjector.forward up to the first per-plane Linear WITHOUT the dense pooled tensors —
 * point_triplane_projector.py:99-115 + the first layer of mlp_xy / mlp_yz / mlp_xz (:60-64, 113-115):
 *     hidden_p[b, row, :] = act( b1_p + sum over the occupied pooled cells (row, g) of W1_p[:, g*C:(g+1)*C] . max-pooled cell )
 * which equals Linear(k*C -> C) applied to the dense flattened tensor of tp_encode_f32 (empty cells contribute 0).
 * One C x C product per OCCUPIED cell (fp32 FMA), nothing of the 430 MB per sample is written or read.
 * feats / idx / points / offsets / geom / arith / clamp_zero as tp_encode_f32 (max reduction). C % 4 == 0, C <= 128,
 * at most 64 pooled cells per axis. w1t[p]: the Linear weight W1_p [C, G_p*C] re-laid as [G_p][C (in)][C (out)]
 * (G_0 = Zp, G_1 = Xp, G_2 = Yp); b1[p] [C]; relu != 0 applies the ReLU that follows the Linear.
 * hidden[0] [B, X*Y, C], hidden[1] [B, Y*Z, C], hidden[2] [B, X*Z, C]. Deterministic (fixed summation order).
 * workspace: tp_projector_sparse_workspace_bytes() bytes (address space for the cell slots; only occupied slots are
 * touched), no initialisation needed. */
TP_API int64_t tp_projector_sparse_workspace_bytes(const tp_geom* geom, int32_t batch, int32_t C);
TP_API int tp_projector_sparse_f32(const float* feats, int64_t feat_stride, int32_t C, const int32_t* idx,
                            const float* points, int32_t point_stride, int64_t n_total, const int64_t* offsets,
                            int32_t batch, const tp_geom* geom, int32_t arith, int32_t clamp_zero,
                            const float* const w1t[3], const float* const b1[3], int32_t relu,
                            float* const hidden[3], void* workspace, int64_t workspace_bytes, void* stream);

/* Divide partial SUM planes by counts (after a point-sharded all-reduce): out[cell,:] /= max(cnt,1). */
TP_API int tp_encode_finalize_mean_f32(float* planes, const int32_t* cell_count, int64_t cells, int32_t C,
                                void* stream);

/* After all-reduce(max) of TP_REDUCE_MAX_PARTIAL planes: -inf -> 0 (cell empty on every GPU), optional
 * clamp at 0. In place over n_floats contiguous floats. */
TP_API int tp_encode_finalize_max_f32(float* planes, int64_t n_floats, int32_t clamp_zero, void* stream);

/* unq_cnt of projector.py:99 as a dense grid: counts [B, X, Y, Z] int32 += 1 per in-range point.
 * counts must be zeroed by the caller. */
TP_API int tp_voxel_counts_i32(const int32_t* idx, int64_t n_total, const int64_t* offsets, int32_t batch,
                        const tp_geom* geom, int32_t* counts, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4  sample_points_triplane — triplane.py:490-514, triplane_occ.py:321-348, triplane_elev.py:286-313,
 *     point_triplane.py:439-466, point_triplane_occ.py:407-440:
 *     normalise + 3 x F.grid_sample(bilinear, zeros, align_corners=False) + (xy + yz) + xz, fused.
 *
 * planes[3]: (x,y)->plane 0, (y,z)->plane 1, (x,z)->plane 2; first coordinate of each pair -> W.
 * queries [B, Q, 3] fp32; out [B, C, Q] fp32 (the reference's channel-major result).
 * tp_planes_nchw_to_nhwc_f32 converts a reference-layout plane [B,C,H,W] to the channels-last
 * [B,H,W,C] copy the gather kernel reads with 16-byte loads. tp_sample3_nchw_f32 does both
 * (nhwc_workspace >= sum_p B*C*H_p*W_p floats).
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_planes_nchw_to_nhwc_f32(const float* src, int64_t src_batch_stride, float* dst,
                               int32_t batch, int32_t C, int32_t H, int32_t W, void* stream);
/* the three planes in ONE launch: dst[k] receives [B, H_k, W_k, C] */
TP_API int tp_planes3_nchw_to_nhwc_f32(const tp_plane planes_nchw[3], float* const dst[3], int32_t batch,
                                int32_t C, void* stream);
TP_API int tp_sample3_nhwc_f32(const tp_plane planes_nhwc[3], int32_t C,
                        const float* queries, int64_t Q, int32_t batch,
                        const tp_sample_geom* sg, int32_t arith,
                        float* out, void* stream);
TP_API int tp_sample3_nchw_f32(const tp_plane planes_nchw[3], int32_t C,
                        const float* queries, int64_t Q, int32_t batch,
                        const tp_sample_geom* sg, int32_t arith,
                        float* out, float* nhwc_workspace, int64_t nhwc_workspace_floats,
                        void* stream);

/* [B,h,w,d,3] query tensors (the 5-D callers: triplane_occ.py:338-346, triplane_elev.py:303-311,
 * point_triplane_occ.py:427-438). Same result, bit for bit, as tp_sample3_nhwc_f32 on the flattened
 * [B, h*w*d, 3] queries; out [B, C, h*w*d]. Every 3-D block of the tensor is tested on the device for
 * the voxel-centre-lattice structure of roi() (triplane_occ.py:291-318) / get_reference_points()
 * (triplane_elev.py:113-133) -- x a function of the h index only, y of w, z of d -- and, where it holds,
 * the three per-plane sums are evaluated once per index pair instead of once per query. Blocks where
 * it does not hold take the per-query path inside the same launch. dims = {h, w, d}. */
TP_API int tp_sample3_grid_nhwc_f32(const tp_plane planes_nhwc[3], int32_t C,
                             const float* queries, const int32_t dims[3], int32_t batch,
                             const tp_sample_geom* sg, int32_t arith,
                             float* out, void* stream);
TP_API int tp_sample3_grid_nchw_f32(const tp_plane planes_nchw[3], int32_t C,
                             const float* queries, const int32_t dims[3], int32_t batch,
                             const tp_sample_geom* sg, int32_t arith,
                             float* out, float* nhwc_workspace, int64_t nhwc_workspace_floats,
                             void* stream);

/* The same decode for the reference's regular query grids WITHOUT a query tensor: roi() (triplane_occ.py:291-318,
 * point_triplane_occ.py:378-405) builds ref_3d[i,j,k] = ((i, j, k) + 0.5) * voxel_size + occ_range[0:3] once and
 * `.repeat`s it per batch every step (triplane_occ.py:153,249). Here the coordinates are generated in the kernel by the
 * same three fp32 operations (add 0.5, multiply, add; never contracted), so the 12 bytes per query are not read:
 * bit-identical to tp_sample3_grid_nhwc_f32 on the materialised tensor. dims = {h, w, d}, d % 4 == 0, out 16-byte aligned
 * (else TP_E_SHAPE: materialise the points and use the entry point above). */
TP_API int tp_sample3_lattice_nhwc_f32(const tp_plane planes_nhwc[3], int32_t C, const int32_t dims[3],
                                const float origin[3], const float step[3], int32_t batch,
                                const tp_sample_geom* sg, int32_t arith, float* out, void* stream);

/* Reference-layout (NCHW) planes in, for every decode entry point: ONE C-ABI call = the NCHW -> channels-last conversion
 * launch + the decode launch, chained by programmatic dependent launch (the decode grid is scheduled while the
 * conversion runs and does everything that does not need the planes -- query load, lattice check, tap records, blocks
 * outside all planes -- before it waits for the converted copy). nhwc_workspace >= sum_p B*C*H_p*W_p floats. Results
 * are those of the *_nhwc_* entry point on the converted planes, bit for bit. */
TP_API int tp_sample3_lattice_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const int32_t dims[3],
                                const float origin[3], const float step[3], int32_t batch,
                                const tp_sample_geom* sg, int32_t arith, float* out,
                                float* nhwc_workspace, int64_t nhwc_workspace_floats, void* stream);
TP_API int tp_sample3_seg_nchw_f32(const tp_plane planes_nchw[3], int32_t C, const float* queries, int64_t total,
                            const int64_t* seg_offsets, const int32_t* seg_batch, int32_t nseg, int32_t batch,
                            const tp_sample_geom* sg, int32_t arith, float* out,
                            float* nhwc_workspace, int64_t nhwc_workspace_floats, void* stream);
TP_API int tp_sample3_grid_head_nchw_tf32(const tp_plane planes_nchw[3], const float* queries, const int32_t dims[3],
                                   int32_t batch, const tp_sample_geom* sg, int32_t arith, const float* w1,
                                   const float* w2, const float* w3, int32_t num_classes, float* logits,
                                   float* nhwc_workspace, int64_t nhwc_workspace_floats, void* stream);

/* Ragged point subsets in ONE launch: the per-(sample, camera) loops around sample_points_triplane in the contrastive
 * branches (triplane.py:438-458: B x 6 SAM-labelled subsets; point_triplane.py:365-372, 389-403). queries [T, 3] = the
 * segments concatenated; seg_offsets [S+1] int64 (device); seg_batch [S] int32 (device; NULL: segment s reads sample
 * s) names the sample whose planes a segment reads (out of range: the segment yields zeros). out [T, C] is POINT-major:
 * every caller turns the reference's [1,C,1,N] result into [N, C] rows (`features.permute(1, 0)`, triplane.py:453-455).
 * Values are bit-identical to tp_sample3_nhwc_f32. The backward scatters grad_out [T, C] into ZEROED channels-last
 * gradient planes, like tp_sample3_backward_nhwc_f32. */
TP_API int tp_sample3_seg_nhwc_f32(const tp_plane planes_nhwc[3], int32_t C, const float* queries, int64_t total,
                            const int64_t* seg_offsets, const int32_t* seg_batch, int32_t num_segments,
                            int32_t batch, const tp_sample_geom* sg, int32_t arith, float* out, void* stream);
TP_API int tp_sample3_seg_backward_nhwc_f32(const tp_plane gplanes_nhwc[3], int32_t C, const float* queries,
                                     int64_t total, const int64_t* seg_offsets, const int32_t* seg_batch,
                                     int32_t num_segments, int32_t batch, const tp_sample_geom* sg, int32_t arith,
                                     const float* grad_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a2  point_to_cam — point_triplane.py:164-241 (twin point_triplane_occ.py:163-239)
 *     einsum(lidar2image, hom_points) -> / max(z, 1e-5) -> * resize - crop (-> flip) -> the -W/2, rotate
 *     by 0, +W/2 round trip -> in-image mask -> (row, col) normalised by resize_dims -> F.grid_sample
 *     of that camera's feature map (row fed as grid-x) -> sum over the cameras that see the point.
 *
 * points      [n_total, point_stride] fp32 (xyz first), samples concatenated; offsets [B+1] int64 (device)
 * feats_nhwc  [B, ncam, Hf, Wf, Cf] fp32 channels-last copy of img_features [B, ncam, Cf, Hf, Wf]
 *             (tp_planes_nchw_to_nhwc_f32 with batch = B * ncam), Cf % 4 == 0, ncam <= 8
 * cams        [B, ncam, 20] fp32 (device): lidar2image 4x4 row-major, resize, crop[0], crop[1], flip (0/1)
 *             (img_metas[b]['lidar2image'][cam], ['imgs_aug'][cam], point_triplane.py:177-199)
 * resize_dim0/1 = img_metas[0]['img_shape'][::-1]  (point_triplane.py:176)
 * out         [n_total, Cf]: row n = sum over cameras of the bilinear sample; 0 where no camera sees it.
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_lift_cam_f32(const float* points, int32_t point_stride, int64_t n_total,
                    const int64_t* offsets, int32_t batch, const float* feats_nhwc,
                    int32_t ncam, int32_t Hf, int32_t Wf, int32_t Cf, const float* cams,
                    float resize_dim0, float resize_dim1, int32_t arith, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward: what autograd produces in the reference for the three ops above.
 * No gradients w.r.t. point / query coordinates (they are data in every config).
 *
 * tp_sample3_backward_nhwc_f32: ATen grid_sampler_2d_backward w.r.t. the input of the three grid_sample
 *   calls: gplanes_nhwc[k].data (channels-last [B,H,W,C], ZEROED by the caller, written here) +=
 *   grad_out[b,c,q] * bilinear weight at every in-bounds tap. grad_out [B,C,Q]. Scatter-add order is not
 *   deterministic (as in ATen): sums agree to fp32 rounding.
 * tp_encode_backward_f32: reduce = TP_REDUCE_MAX: grad_feats[n,c] = sum over the three planes of
 *   gout_p[cell_p(n), c] where feats[n,c] attains out_p[cell_p(n), c] (the scatter_max argmax routing followed by
 *   spconv's max-pool backward; exact ties, measure-zero for real features, give every maximiser the gradient);
 *   TP_REDUCE_MEAN: gout / count. idx [N,3] int32 as in tp_encode_f32; any gout_p may be NULL.
 * tp_lift_cam_backward_f32: gradient w.r.t. the channels-last feature maps [B,ncam,Hf,Wf,Cf] (ZEROED by the caller).
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_sample3_backward_nhwc_f32(const tp_plane gplanes_nhwc[3], int32_t C, const float* queries, int64_t Q,
                                 int32_t batch, const tp_sample_geom* sg, int32_t arith, const float* grad_out,
                                 void* stream);
/* the same for [B,h,w,d,3] query tensors (dims = {h, w, d}): lattice blocks (checked on the device, as in
 * tp_sample3_grid_nhwc_f32) first sum grad_out over the lattice index a plane does not depend on, then scatter one
 * footprint per index pair: ~10x fewer reductions into the gradient planes. Other blocks scatter per query. */
TP_API int tp_sample3_grid_backward_nhwc_f32(const tp_plane gplanes_nhwc[3], int32_t C, const float* queries,
                                      const int32_t dims[3], int32_t batch, const tp_sample_geom* sg, int32_t arith,
                                      const float* grad_out, void* stream);
TP_API int tp_encode_backward_f32(const float* feats, int64_t feat_stride, int32_t C, const int32_t* idx,
                           int64_t n_total, const int64_t* offsets, int32_t batch, const tp_geom* geom,
                           int32_t reduce, int32_t clamp_zero, const float* out_xy, const float* out_yz,
                           const float* out_xz, const int32_t* cell_count, const float* gout_xy,
                           const float* gout_yz, const float* gout_xz, float* grad_feats, void* stream);
TP_API int tp_lift_cam_backward_f32(const float* points, int32_t point_stride, int64_t n_total,
                             const int64_t* offsets, int32_t batch, int32_t ncam, int32_t Hf, int32_t Wf,
                             int32_t Cf, const float* cams, float resize_dim0, float resize_dim1, int32_t arith,
                             const float* grad_out, float* grad_feats_nhwc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a6 / 8f#3  occupancy head on the decode output — mmdet3d/models/dense_heads/mlp.py:25-70 (Mlp.forward):
 *     conv1 (C -> 2C, 1x1x1, no bias) + ReLU, conv2 (2C -> C) + ReLU, conv3 (C -> num_classes), fused for tiles
 *     of 128 queries on the tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM). TF32 inputs, fp32
 *     accumulation: the precision cuDNN gives the reference's Conv3d by default.
 * feats [B, C, Q] (the decode output), w1 [2C, C], w2 [C, 2C], w3 [num_classes, C] (the Conv3d weights with the
 * 1x1x1 kernel dims dropped), logits [B, num_classes, Q]. C == 32, num_classes <= 16 in this build.
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_mlp_head_tf32(const float* feats, int64_t Q, int32_t batch, int32_t C, const float* w1,
                     const float* w2, const float* w3, int32_t num_classes, float* logits, void* stream);

/* 8f#3 (ii)  occupancy decode + head in one kernel: tp_sample3_grid_nhwc_f32 (C = 32) followed by tp_mlp_head_tf32
 *     without the [B,32,Q] feature tensor in between (TriplaneOcc: sample_points_triplane, triplane_occ.py:321-348,
 *     then occ_head, triplane_occ.py:182-186). planes_nhwc: three channels-last [B,H,W,32] planes; queries
 *     [B,h,w,d,3] (d % 4 == 0); logits [B, num_classes, h*w*d]. The features fed to the head are bit for bit the
 *     decode kernel's, the logits equal tp_mlp_head_tf32 on them. Blocks of the query tensor that are not a
 *     lattice are evaluated per query inside the same launch. */
TP_API int tp_sample3_grid_head_tf32(const tp_plane planes_nhwc[3], const float* queries, const int32_t dims[3],
                              int32_t batch, const tp_sample_geom* sg, int32_t arith, const float* w1,
                              const float* w2, const float* w3, int32_t num_classes, float* logits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 8f#4  gathers / scatters either side of the decode (nearest-pixel, i.e. integer index work)
 *
 * Feature -> camera-pixel scatter: TriplaneMAE.forward triplane.py:381-390 and PointTriplane.cam_rec_feat
 * point_triplane.py:243-309: `img[:, rows, cols] = feat[:, valid]` with duplicate targets. torch-CUDA leaves the order
 * undefined; torch-CPU keeps the LAST source in list order. This library defines the order as torch-CPU's: the source
 * with the highest index wins. Two steps: a winner image (int32, -1 = empty; written by the call, atomicMax), then a
 * write-once gather of the dense output.
 *   tp_pixel_winner_coors_i32: coors [M, npix, 2] fp32 (row, col) as range_cam_coors of JointEncoder.interact; a source
 *     is used iff long(row) > 0 (triplane.py:386 tests AFTER .long(): row 0 is dropped there too) and inside H x W.
 *   tp_pixel_winner_points_i32: points [N, stride] (samples concatenated, offsets [B+1]), cams as tp_lift_cam_f32;
 *     projects every point into every camera with the chain of point_triplane.py:263-301, pixel = (long(y), long(x));
 *     winner [B, ncam, R0, R1] holds the point index inside its sample.
 *   tp_winner_gather_f32: out [M, C, HW] = feat[m / imgs_per_feat][c][winner[m, pix]] (0 where winner < 0). feat is
 *     addressed as feat + fb * feat_bstride + (feat_row0 ? feat_row0[fb] * feat_nstride : 0) + c * feat_cstride +
 *     n * feat_nstride: channel-major decode output [B, C, N] or point-major rows [N, C].
 *   tp_winner_gather_backward_f32: grad_feat (ZEROED by the caller, same addressing) += grad_out at the winners.
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_pixel_winner_coors_i32(const float* coors, int64_t n_images, int64_t npix, int32_t H, int32_t W,
                              int32_t* winner, void* stream);
TP_API int tp_pixel_winner_points_i32(const float* points, int32_t point_stride, int64_t n_total,
                               const int64_t* offsets, int32_t batch, const float* cams, int32_t ncam,
                               float resize_dim0, float resize_dim1, int32_t* winner, void* stream);
TP_API int tp_winner_gather_f32(const int32_t* winner, int64_t n_images, int64_t HW, int32_t C, int32_t imgs_per_feat,
                         const float* feat, int64_t feat_bstride, int64_t feat_cstride, int64_t feat_nstride,
                         const int64_t* feat_row0, float* out, void* stream);
TP_API int tp_winner_gather_backward_f32(const int32_t* winner, int64_t n_images, int64_t HW, int32_t C,
                                  int32_t imgs_per_feat, const float* grad_out, float* grad_feat,
                                  int64_t feat_bstride, int64_t feat_cstride, int64_t feat_nstride,
                                  const int64_t* feat_row0, void* stream);

/* JointEncoder.interact — joint_encoder.py:97-215, everything but the position_encoder MLP:
 *   tp_range_project_f32: range_points [B, npix, 3], range_image [B, npix] (masked input; > 0 = unmasked), cams
 *     [B, ncam, 20] -> coors [B, ncam, npix, 2] = range_cam_coors (row, col; -1 where the pixel holds no point
 *     (x = y = z = 0) or the camera does not see it, :180-187), fidx [B, ncam, npix] = nearest feature-map pixel
 *     long(row * Hf / R0) * Wf + long(col * Wf / R1) of the unmasked visible points (-1 otherwise, :190-205), winner
 *     [B, ncam, Hf*Wf] = highest range pixel landing on each feature pixel (for the position-embedding index-put).
 *   tp_range_gather_f32: out [B, C, npix] = sum over cameras, ascending from zero, of img_features[b, cam, :, fidx]
 *     (img_features [B, ncam, C, Hf*Wf], the reference's NCHW maps; :208). Backward: grad_img (ZEROED) += grad_out.
 *   tp_posembed_scatter_f32: img_features[m, :, p] += pos_embed[m, p, :] where winner[m, p] >= 0, in place
 *     (m = b * ncam + cam; pos_embed [M, Hf*Wf, C] = position_encoder of the winners' points; :211-213, duplicates
 *     resolved as above). Backward: grad_pos_embed[m, p, :] = winner >= 0 ? grad_img[m, :, p] : 0.
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_range_project_f32(const float* range_points, const float* range_image, int64_t npix, int32_t batch,
                         const float* cams, int32_t ncam, float resize_dim0, float resize_dim1, int32_t Hf,
                         int32_t Wf, int32_t arith, float* coors, int32_t* fidx, int32_t* winner, void* stream);
TP_API int tp_range_gather_f32(const int32_t* fidx, int64_t npix, int32_t batch, int32_t ncam,
                        const float* img_features, int32_t C, int32_t HWf, float* out, void* stream);
TP_API int tp_range_gather_backward_f32(const int32_t* fidx, int64_t npix, int32_t batch, int32_t ncam,
                                 const float* grad_out, int32_t C, int32_t HWf, float* grad_img, void* stream);
TP_API int tp_posembed_scatter_f32(const int32_t* winner, int64_t n_images, int32_t C, int32_t HWf,
                            float* img_features, const float* pos_embed, void* stream);
TP_API int tp_posembed_scatter_backward_f32(const int32_t* winner, int64_t n_images, int32_t C, int32_t HWf,
                                     const float* grad_img, float* grad_pos_embed, void* stream);

/* InterpNet's neighbourhood search — interpnet.py:44,65: torch_geometric.nn.radius(x = sources, y = queries, r,
 * batch_x, batch_y) -> torch_cluster.radius with max_num_neighbors = 32 (un-vendored third-party op, restated from
 * its CUDA kernel): for every query the FIRST max_num_neighbors sources of the same sample, in index order, with
 * squared distance < r^2. x [Nx, 3] / y [Ny, 3] samples concatenated, x_offsets / y_offsets [B+1] int64 (device).
 * col [Ny, max_num_neighbors] int32 global source indices, -1 padded; count [Ny]. (row, col) pairs of the reference =
 * (q, col[q, k]) for k < count[q], queries ascending. */
TP_API int tp_radius_i32(const float* x, const int64_t* x_offsets, const float* y, const int64_t* y_offsets,
                  int64_t ny, int32_t batch, float r, int32_t max_num_neighbors, int32_t* col, int32_t* count,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU, point-sharded encode (the reference is data-parallel only and has no counterpart).
 *
 * Partial planes: tp_encode_f32(reduce = TP_REDUCE_MAX_PARTIAL / TP_REDUCE_SUM) on every rank, an all-reduce of the
 * planes by the caller's communication library (the Python host uses torch.distributed / NCCL, dist.py), then
 * tp_encode_finalize_max_f32 / tp_encode_finalize_mean_f32.
 *
 * Owner exchange, tp_route_points_f32: every rank pushes its points into the owners' receive buffers over peer memory.
 * Rank r owns rows [x_bounds[r], x_bounds[r+1]) of the xy and xz planes and rows [y_bounds[r], y_bounds[r+1]) of the
 * yz plane. peer_*[r] are THIS process's device pointers to rank r's buffers (CUDA IPC / VMM mappings, e.g. torch
 * symmetric memory): peer_cnt[r] int32[2] zeroed, peer_idx_x/y[r] int32 [capacity, 3] filled with -1, peer_feat_x/y[r]
 * float [capacity, C]; capacity >= the global point count. The caller separates reset, route and the owners' encodes
 * with cross-rank barriers. Afterwards every owner runs tp_encode_f32 twice on its buffers with n_total = capacity:
 * idx_x (x relative to the slab) with grid (x_bounds[r+1]-x_bounds[r], Y, Z) for xy and xz, idx_y with grid
 * (X, y_bounds[r+1]-y_bounds[r], Z) for yz, same pool: the results are the slabs of the single-GPU planes, bit for bit.
 * world <= 8, one sample per call (arrival order is arbitrary, so rows of several samples cannot be told apart). */
TP_API int tp_route_points_f32(const float* points, int32_t point_stride, const float* feats, int64_t feat_stride,
                        int32_t C, int64_t n, const tp_geom* geom, int32_t arith, int32_t world,
                        const int32_t* x_bounds, const int32_t* y_bounds, void* const* peer_cnt,
                        void* const* peer_idx_x, void* const* peer_feat_x, void* const* peer_idx_y,
                        void* const* peer_feat_y, int64_t capacity, void* stream);

/* Collective hooks for callers without torch.distributed (the Python host uses torch.distributed / NCCL for the same
 * exchange). NCCL is opened at run time (dlopen "libnccl.so.2"); errors: 1000 + ncclResult_t.
 *   tp_comm_unique_id: rank 0 creates the 128-byte id and ships it to the other ranks by its own means;
 *   tp_comm_init: collective over `world` processes (one GPU each, the current device); tp_comm_destroy frees it;
 *   tp_allreduce_planes: in place over n_floats contiguous floats of partial planes (the three planes may live in one
 *     buffer) with max (reduce = TP_REDUCE_MAX / _MAX_PARTIAL) or sum (TP_REDUCE_SUM / _MEAN), plus, when n_counts > 0,
 *     a sum of the int32 cell counts in the same NCCL group; follow with tp_encode_finalize_max_f32 / _mean_f32. */
TP_API int tp_comm_unique_id(void* id_out_128_bytes);
TP_API int tp_comm_init(void** comm_out, int32_t world, int32_t rank, const void* unique_id_128_bytes);
TP_API int tp_comm_destroy(void* comm);
TP_API int tp_allreduce_planes(void* comm, float* planes, int64_t n_floats, int32_t reduce, int32_t* cell_count,
                        int64_t n_counts, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry points (what a non-PyTorch caller binds; bench.py's `e2e` leg).
 * All pointers are HOST memory (pinned recommended). They allocate a per-thread cached device
 * arena, copy in, run the kernels above, copy out and synchronise the internal stream.
 * ------------------------------------------------------------------------------------------- */
TP_API int tp_sample3_host_f32(const float* planes_nchw_host[3], const int32_t HW[6] /* H0,W0,H1,W1,H2,W2 */,
                        const int64_t plane_batch_stride[3], int32_t C,
                        const float* queries_host, int64_t Q, int32_t batch,
                        const tp_sample_geom* sg, int32_t arith, float* out_host);
/* same, for a [B,h,w,d,3] query tensor (tp_sample3_grid_nhwc_f32 on the device); dims = {h, w, d} */
TP_API int tp_sample3_grid_host_f32(const float* planes_nchw_host[3], const int32_t HW[6],
                             const int64_t plane_batch_stride[3], int32_t C,
                             const float* queries_host, const int32_t dims[3], int32_t batch,
                             const tp_sample_geom* sg, int32_t arith, float* out_host);
TP_API int tp_encode_host_f32(const float* feats_host, int32_t C, const float* points_host,
                       int32_t point_stride, int64_t n_total, const int64_t* offsets_host,
                       int32_t batch, const tp_geom* geom, int32_t arith, int32_t reduce,
                       int32_t clamp_zero, float* out_xy_host, float* out_yz_host, float* out_xz_host);
/* free the cached arena of the calling thread */
TP_API void tp_host_arena_release(void);

#ifdef __cplusplus
}
#endif
#endif /* TRIPLANE_H_ */
