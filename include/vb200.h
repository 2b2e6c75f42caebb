/*
 * vb200.h -- C ABI of libvb200.so: the B200-native (sm_100a) 2D->3D feature path of Vampire.
 *
 * This is the drop-in boundary.  The reference (cskkxjk/Vampire) has no FFI of its own: its
 * setup.py registers a BuildExtension with NO ext_modules (/root/reference/setup.py:10-26) and the
 * path is four methods of the backbone nn.Module plus one inline expression
 * (/root/reference/src/layers/backbones/base_vampire2.py, "BV2" below).  The entry points here
 * are what a CUDA extension in that setup.py would bind, one per reference call site:
 *
 *   vb200_get_pixel          <- BaseVAMPIRE2.get_pixel                         BV2:351-388
 *   vb200_get_geometry       <- BaseVAMPIRE2.get_geometry (+ nan_to_num)       BV2:314-349, 612
 *   vb200_lift_indices       <- the integers behind F.grid_sample in get_voxel_feats   BV2:493-507
 *   vb200_render_indices     <- the integers behind F.grid_sample in the render        BV2:397-419
 *   vb200_lift_pool_fwd/bwd  <- depth (x) ctx outer product + get_voxel_feats  BV2:553, 483-516
 *   vb200_render_fwd/bwd     <- volume_rendering_from_multiple_views           BV2:391-467
 *                               with ModifyLaplaceDensity        src/utils/render_utils.py:30-46
 *
 * Conventions
 *   - plain C: pointers, sizes, ints and floats only; no C++/torch types; no exceptions.
 *   - every pointer named d_* is DEVICE memory owned by the caller (the PyTorch caching
 *     allocator in the Python host); the library allocates no device memory.  Its only state is the
 *     instrumentation counters/events of vb200_trace_* and, per device, one pair of internal side streams with their
 *     events (vb200_render_fwd forks the BEV branch onto it; a mutex serialises the fork..join bookkeeping of
 *     concurrent callers, and the caller's stream is re-joined on every exit path).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant and
 *     thread-safe.
 *   - compiled for C = 16 context channels, K = 18 classes and at most 8 cameras (every reference experiment);
 *     other channel counts return VB200_ERR_ARG.
 *   - return 0 on success, a negative VB200_ERR_* otherwise; vb200_strerror() names it.
 *   - there is no CPU fallback: on a device that is not sm_100 the compute calls return
 *     VB200_ERR_ARCH.
 *   - feature tensors come in `dtype` VB200_F32 / VB200_BF16 / VB200_F16; all arithmetic and all
 *     geometry is fp32 (the reference runs the geometry under autocast(enabled=False), BV2:485).
 */
#ifndef VB200_H_
#define VB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_VERSION 203 /* 0.2.0 */

enum vb200_status {
  VB200_OK = 0,
  VB200_ERR_ARG = -1,       /* null pointer / bad size / unsupported channel count */
  VB200_ERR_DTYPE = -2,     /* unknown dtype code */
  VB200_ERR_ARCH = -3,      /* current device is not sm_100 */
  VB200_ERR_WORKSPACE = -4, /* workspace too small */
  VB200_ERR_CUDA = -5,      /* a CUDA runtime call or launch failed (see cudaGetLastError) */
  VB200_ERR_ALIGN = -6      /* a pointer is not 16-byte aligned */
};

enum vb200_dtype { VB200_F32 = 0, VB200_BF16 = 1, VB200_F16 = 2 };

/* memory layout of a (B, C, Z, Y, X) volume */
enum vb200_layout {
  VB200_NCDHW = 0, /* the reference's contiguous layout, x fastest           */
  VB200_NDHWC = 1  /* channels-last (torch.channels_last_3d), channel fastest */
};

/* backbone_conf['density_mode'] (BV2:191-194): what `self.density` is */
enum vb200_density {
  VB200_DENSITY_SDF = 0,  /* 'sdf':   ModifyLaplaceDensity(beta, bias) of an SDF feature (render_utils.py:30-46)        */
  VB200_DENSITY_NAIVE = 1 /* 'naive': nn.Sigmoid(); `beta` is not read (pass any valid device float), d beta = 0       */
};

/* Static description of the path: sizes + the fp32 constants the reference derives from its
 * backbone_conf (base_exp.py:40-92).  Floats are the Python doubles rounded to fp32 exactly as
 * torch rounds a Python scalar operand. */
typedef struct VbGrid {
  int32_t B, N;           /* samples, cameras per sample (6)                                    */
  int32_t D, fH, fW;      /* depth planes (86), feature rows (64), feature cols (176).  D == 1 selects the 2-D
                           * lift of the BaseBiLinear ablation (base_bilinear.py:471-517) in
                           * the lift entry points: no depth distribution, depth test z > 0 (pass d_lo = 0,
                           * d_hi = +inf), the single plane sampled at iz = 0; pass a depth tensor of ones     */
  int32_t vZ, vY, vX;     /* seg voxel grid (20, 256, 256)                                      */
  int32_t oZ, oY, oX;     /* det / BEV grid (10, 256, 256)                                      */
  int32_t C, K;           /* context channels (16), semantic classes (18)                       */
  int16_t has_bda;        /* 0: mats_dict has no 'bda_mat' (BV2:370,343 skip the bda products)  */
  int16_t density_mode;   /* enum vb200_density: how the density feature becomes sigma (BV2:191-194).  Shares a word
                           * with has_bda on purpose: the struct is a by-value kernel parameter, and growing it past
                           * 128 bytes cost the recomputing camera march 5-11 % (measured: ptxas spills differently) */
  float img_w_m1, img_h_m1; /* float(W-1), float(H-1)          BV2:499-500                      */
  float x_hi, y_hi;       /* float(W-0.5), float(H-0.5)        BV2:494-495                      */
  float d_lo, d_hi;       /* d_bound[0], d_bound[1]            BV2:496                          */
  float d_ext;            /* d_bound[1]-d_bound[0] in double, then fp32   BV2:501               */
  float seg_lo[3];        /* x/y/z_bound_seg[0]                BV2:397-399                      */
  float seg_ext[3];       /* x/y/z_bound_seg[1]-[0]            BV2:400-402                      */
  float bg_depth;         /* d_bound[1]                        BV2:436                          */
  float bev_delta;        /* z_bound_det[2]                    BV2:451                          */
  float sdf_bias;         /* density bias (-1)                 render_utils.py:35               */
  float beta_min;         /* 1e-4                              render_utils.py:31               */
  float term_eps;         /* early-termination threshold on transmittance; 0 disables          */
} VbGrid;

/* Lattice tables, DEVICE pointers to fp32 arrays built on the host with the reference's own torch
 * calls (BV2:243-293); recomputing them in-kernel is not bit-identical. */
typedef struct VbTables {
  const float* us;       /* [fW]  linspace(0, W-1, fW)                       */
  const float* vs;       /* [fH]  linspace(0, H-1, fH)                       */
  const float* ds;       /* [D]   arange(*d_bound)                           */
  const float* xs;       /* [vX]  seg voxel centres                          */
  const float* ys;       /* [vY]                                             */
  const float* zs;       /* [vZ]                                             */
  const float* oxs;      /* [oX]  det voxel centres                          */
  const float* oys;      /* [oY]                                             */
  const float* ozs;      /* [oZ]                                             */
  const float* mids;     /* [D-1] interval mid depths                        */
  const float* bev_mids; /* [oZ]  BEV level heights, top level first         */
} VbTables;

/* d_mats: (B, N, 6, 4, 4) fp32 row-major, prepared on the host with the reference's torch calls:
 *   slot 0 bda^-1 | 1 K.E^-1 | 2 ida | 3 ida^-1 | 4 E.K^-1 | 5 bda      (vampire_b200/matrices.py) */
#define VB200_MAT_SLOTS 6

int vb200_version(void);
const char* vb200_strerror(int status);
/* 0 if the current CUDA device is sm_100, VB200_ERR_ARCH otherwise */
int vb200_device_check(void);

/* ---- instrumentation (bench.py: gpu_launches + live per-kernel CUDA-event timing) ----------- */
int vb200_trace_num_kernels(void);
const char* vb200_trace_kernel_name(int kernel_id);
/* launches of kernel family `kernel_id` since load (any id outside the range: all families) */
long long vb200_launch_count(int kernel_id);
/* start (1) / stop (0) recording a CUDA-event pair around every kernel launch; clears old records */
int vb200_trace_enable(int on);
/* waits for the recorded events; fills per-family device milliseconds and launch counts
 * (arrays of vb200_trace_num_kernels() entries) and clears the records */
int vb200_trace_collect(double* total_ms, long long* launches);

/* ---- geometry ------------------------------------------------- */

/* G1: d_pix (B, N, vZ, vY, vX, 3) fp32 = get_pixel(...)                          BV2:351-388 */
int vb200_get_pixel(const VbGrid* g, const VbTables* t, const float* d_mats, float* d_pix, void* stream);

/* G2: d_geom (B, N, D, fH, fW, 3) fp32 = get_geometry(...); nan_to_num != 0 additionally applies
 * torch.nan_to_num(geom, -1e3)                                                  BV2:314-349, 612 */
int vb200_get_geometry(const VbGrid* g, const VbTables* t, const float* d_mats, float* d_geom,
                       int nan_to_num, void* stream);

/* L2: per (b, n, voxel): valid (uint8), base corner (x0, y0, z0) int16 x3, fractions fp32 x3 of the
 * align_corners=False trilinear lookup into the (D, fH, fW) frustum volume.  Any output may be NULL. */
int vb200_lift_indices(const VbGrid* g, const VbTables* t, const float* d_mats, uint8_t* d_valid,
                       int16_t* d_i0, float* d_frac, void* stream);

/* R2: per (b, n, s<D-1, h, w): geom_valid_mask (uint8), base voxel (x0, y0, z0) int16 x3 and fractions
 * of the align_corners=True lookup into the (vZ, vY, vX) volume.  d_geom may be NULL (geometry is
 * then recomputed from d_mats + nan_to_num) or a (B, N, D, fH, fW, 3) tensor as passed to the
 * reference's render.  Any output may be NULL. */
int vb200_render_indices(const VbGrid* g, const VbTables* t, const float* d_mats, const float* d_geom,
                         uint8_t* d_mask, int16_t* d_i0, float* d_frac, void* stream);

/* ---- lift + pool (Bk) --------------------------------------------------- */

/* Feature dtypes of the lift: `dtype` is the dtype of the depth distribution AND of the pooled volume (and of their
 * gradients), `ctx_dtype` the dtype of the context features (and of d_gctx).  Supported: dtype == ctx_dtype, or
 * dtype = VB200_F32 with a 16-bit ctx_dtype -- the reference under AMP, where softmax is autocast to fp32 and
 * depth.unsqueeze(2) * ctx.unsqueeze(3) (BV2:553) promotes the frustum, the grid_sample and the pooled volume to
 * fp32. */

/* bytes of scratch vb200_lift_pool_fwd needs (a channels-last copy of ctx in ctx_dtype) / _bwd needs */
size_t vb200_lift_pool_fwd_workspace(const VbGrid* g, int ctx_dtype);
size_t vb200_lift_pool_bwd_workspace(const VbGrid* g, int dtype);

/* Forward: out[b,c,z,y,x] = sum_n f[n,c] / (sum_n [|f[n,c]|>0] + 1e-6), f = valid * trilinear sample
 * of depth[b,n] (x) ctx[b,n] at the projected voxel centre -- the frustum tensor is never formed.
 *   d_depth (B,N,D,fH,fW) in `dtype`, d_ctx (B,N,C,fH,fW) in `ctx_dtype` (contiguous, reference layout)
 *   d_out   (B,C,vZ,vY,vX) in `dtype`, memory layout `out_layout`
 *   d_cnt   optional (B, vZ*vY*vX) uint64: per-channel non-zero camera count, 4 bits per channel
 *           (saved for the backward); may be NULL for inference. */
int vb200_lift_pool_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_depth,
                        const void* d_ctx, int dtype, int ctx_dtype, void* d_out, int out_layout, uint64_t* d_cnt,
                        void* d_workspace, size_t workspace_bytes, void* stream);

/* Backward: deterministic (sorted-segment gather, no float atomics).
 *   d_gout (B,C,vZ,vY,vX) in `dtype`, layout gout_layout; d_cnt from the forward
 *   d_gdepth (B,N,D,fH,fW) in `dtype`, d_gctx (B,N,C,fH,fW) in `ctx_dtype` (fully overwritten) */
int vb200_lift_pool_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_depth,
                        const void* d_ctx, int dtype, int ctx_dtype, const void* d_gout, int gout_layout,
                        const uint64_t* d_cnt, void* d_gdepth, void* d_gctx, void* d_workspace,
                        size_t workspace_bytes, void* stream);

/* ---- cached projection / sort plan (north-star kernel (a); SURVEY §7.1 K_proj) --------------------------
 * get_pixel (BV2:351-388) and the sampling coordinates of get_voxel_feats (BV2:493-507) depend on the camera
 * matrices only, and in validation / test those never change (deterministic ida,
 * src/datasets/nusc_det_seg_dataset.py:489-498; identity bda, src/exps/nuscenes/base_exp.py:113-120).  A plan holds,
 * for ONE sample's matrices, the compacted valid (voxel, camera) pairs twice: voxel-major for the forward gather
 * and sorted per destination pixel cell for the deterministic backward.  It stores only integers and exact
 * fractions produced by the strict fp32 chain, so the *_planned entry points are bit-identical to the plain
 * ones; they need neither the lattice tables nor the matrices. */
typedef struct VbLiftPlan {  /* all DEVICE pointers; 32 bytes */
  const uint32_t* head;      /* [vZ*vY*vX]  (first pair << 4) | number of valid cameras of the voxel              */
  const void* pairs;         /* [P] 16-byte records, voxel-major, cameras ascending:
                                {cam << 29 | (z0+1) << 20 | (y0+1) << 10 | (x0+1), fx, fy, fz}                     */
  const int32_t* cell_off;   /* [N*(fH+1)*(fW+1) + 1]  CSR over destination pixel cells (n, y0+1, x0+1)            */
  const void* cell_recs;     /* [P] 16-byte records, cell-major, sorted by (z0, voxel):
                                {(z0+1) << 21 | voxel, fx, fy, fz}                                                 */
} VbLiftPlan;

/* Build the plans of g->B samples into caller-owned arrays with `capacity` records per sample:
 *   d_head (B, nvox) uint32 | d_pairs, d_cell_recs (B, capacity) x 16 B | d_cell_off (B, nc + 1) int32
 *   d_num_pairs (B) int32: valid pairs P of each sample.  P > capacity means that sample's records were truncated:
 *   rebuild with capacity >= P (N * nvox always suffices).  Limits: fW, fH < 1022, D < 510, N <= 8, nvox <= 2^21. */
size_t vb200_lift_plan_workspace(const VbGrid* g, long long capacity);
int vb200_lift_plan_build(const VbGrid* g, const VbTables* t, const float* d_mats, uint32_t* d_head, void* d_pairs,
                          int32_t* d_cell_off, void* d_cell_recs, long long capacity, int32_t* d_num_pairs,
                          void* d_workspace, size_t workspace_bytes, void* stream);

/* vb200_lift_pool_fwd / _bwd driven by cached plans: d_plans = DEVICE array of g->B VbLiftPlan (16-byte aligned).
 * Forward workspace = vb200_lift_pool_fwd_workspace; backward = vb200_lift_pool_bwd_planned_workspace. */
size_t vb200_lift_pool_bwd_planned_workspace(const VbGrid* g, int dtype);
int vb200_lift_pool_fwd_planned(const VbGrid* g, const VbLiftPlan* d_plans, const void* d_depth, const void* d_ctx,
                                int dtype, int ctx_dtype, void* d_out, int out_layout, uint64_t* d_cnt,
                                void* d_workspace, size_t workspace_bytes, void* stream);
int vb200_lift_pool_bwd_planned(const VbGrid* g, const VbLiftPlan* d_plans, const void* d_depth, const void* d_ctx,
                                int dtype, int ctx_dtype, const void* d_gout, int gout_layout, const uint64_t* d_cnt,
                                void* d_gdepth, void* d_gctx, void* d_workspace, size_t workspace_bytes, void* stream);

/* Compatibility path for the reference's get_voxel_feats SIGNATURE (BV2:483), which is handed the
 * already materialised frustum tensor d_frustum (B,N,C,D,fH,fW) in `dtype`: same semantics, generic
 * gather (forward) / atomicAdd scatter (backward, like ATen).  Not the benchmarked path. */
size_t vb200_gather_pool_bwd_workspace(const VbGrid* g, int dtype);
int vb200_gather_pool_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_frustum,
                          int dtype, void* d_out, uint64_t* d_cnt, void* stream);
int vb200_gather_pool_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_gout,
                          const uint64_t* d_cnt, int dtype, void* d_gfrustum, void* d_workspace,
                          size_t workspace_bytes, void* stream);

/* ---- volume rendering (Bk) ------------------------------------------ */

/* Cached render plan of ONE sample (north-star kernel (a), camera side): for every (camera, ray, sample) the mask,
 * the base voxel and the three fractions behind F.grid_sample (BV2:397-419) and the step length of BV2:426, as the
 * strict fp32 chain get_geometry -> nan_to_num -> normalise -> unnormalise produces them.  Like the lift plan it
 * depends on the matrices only (constant in validation / test), and the march that reads it composites the same
 * samples with the same weights as the march that recomputes the geometry.  Layout: rays are grouped in the
 * 8 x 4 pixel patches a warp marches; record index = ((n * npatch + patch) * (D-1) + i) * 32 + lane. */
typedef struct VbRenderPlan {  /* all DEVICE pointers; 32 bytes */
  const void* steps;     /* [N * npatch * (D-1) * 32] x 16 B: {valid << 31 | base voxel, fx, fy, fz}             */
  const float* delta;    /* [same]  |p_{i+1} - p_i|                                                              */
  const int16_t* last;   /* [N * npatch * 32]  last valid sample of the ray, -1 if none                          */
  const void* box;       /* [N * npatch * (D-1)] x 8 B: per (warp, sample) the box of voxels covering all trilinear
                            corners of the warp's rays: {staged << 31 | first voxel, nx | ny << 8 | nz << 16}; bits 21..27
                            of a record's key = its base corner's index inside the box (shared-memory staged march)  */
} VbRenderPlan;

/* per-sample element counts of the three arrays: rays = N * npatch * 32, steps = rays * (D-1) */
size_t vb200_render_plan_rays(const VbGrid* g);
/* Build the plans of g->B samples: d_steps (B, steps) x 16 B | d_delta (B, steps) fp32 | d_last (B, rays) int16 |
 * d_box (B, steps / 32) x 8 B */
int vb200_render_plan_build(const VbGrid* g, const VbTables* t, const float* d_mats, void* d_steps, float* d_delta,
                            int16_t* d_last, void* d_box, void* stream);

typedef struct VbRenderIn {
  const void* density;  /* (B, 1, vZ, vY, vX) density_feature, `dtype`, NCDHW */
  const void* sem;      /* (B, K, vZ, vY, vX) semantic_logits                 */
  const void* rgb;      /* (B, 3, vZ, vY, vX)                                 */
  const void* feat;     /* (B, C, vZ, vY, vX) base features                   */
  const float* beta;    /* device pointer to the learnable density.beta scalar */
  const float* geom;    /* optional (B, N, D, fH, fW, 3) fp32 geometry as passed to the reference's
                           render; NULL = recompute from d_mats (+ nan_to_num), no 70 MB/sample read */
  const VbRenderPlan* plans; /* optional DEVICE array of B cached plans (forward only; ignored when geom is given):
                           the camera march reads its geometry from them instead of recomputing it */
  int32_t flags;        /* VB200_RENDER_* bits (forward only) */
  const void* packed;   /* optional, BACKWARD only: the channels-last copy of density | sem | rgb that the forward built
                           at the start of its workspace (B consecutive copies of vb200_render_packed_bytes() each,
                           present when the forward's workspace held all B samples).  Given, vb200_render_bwd does
                           not pack the volumes a second time; NULL = pack again */
} VbRenderIn;

/* voxel_output leaves vb200_render_fwd already multiplied by tanh(voxel_density) -- by voxel_density itself when
 * density_mode is VB200_DENSITY_NAIVE: the BEV epilogue of BV2:627-630 folded into the kernel that resamples the
 * features (both operands are in registers there).  Inference only: vb200_render_bwd differentiates the unfused
 * outputs. */
#define VB200_RENDER_TANH_EPILOGUE 1

typedef struct VbRenderOut {
  float* rgb;           /* (B, N, 3, fH, fW)                                  */
  float* seg;           /* (B, N, K, fH, fW)                                  */
  float* depth;         /* (B, N, 1, fH, fW)                                  */
  float* bev_rgb;       /* (B, 3, oY, oX)                                     */
  float* bev_seg;       /* (B, K, oY, oX)                                     */
  float* bev_height;    /* (B, 1, oY, oX)                                     */
  float* voxel_density; /* (B, 1, oZ, oY, oX)  sigma, top level first         */
  void* voxel_output;   /* (B, C, oZ, oY, oX)  `dtype`, resampled features    */
} VbRenderOut;

/* vb200_render_fwd_workspace = the minimum: BEV compositing weights + ONE packed camera volume (samples
 * are then packed and marched one at a time, the packed copy staying L2-resident).  Passing
 * minimum + (n-1) * vb200_render_packed_bytes lets n samples share a pack/march round. */
size_t vb200_render_fwd_workspace(const VbGrid* g, int dtype);
size_t vb200_render_packed_bytes(const VbGrid* g, int dtype);
size_t vb200_render_bwd_workspace(const VbGrid* g, int dtype);

/* branches: bit 0 = camera branch, bit 1 = BEV branch */
#define VB200_BRANCH_CAM 1
#define VB200_BRANCH_BEV 2

int vb200_render_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbRenderIn* in,
                     int dtype, const VbRenderOut* out, int branches, void* d_workspace,
                     size_t workspace_bytes, void* stream);

/* When both branches are requested, vb200_render_fwd forks the BEV kernels onto an internal per-device
 * side stream and joins before returning (the two branches are independent and both issue-bound).
 * enable = 0 serialises them on the caller's stream (used by bench.py to time kernels in isolation).
 * Concurrent render calls on one device (several host threads / streams) are safe; they share the side stream. */
int vb200_render_set_fork(int enable);

/* Depth-split camera march (opt-in).  When a render call has too few rays to fill the GPU (batch 1 at the R50 config),
 * the march can be launched as thread-block clusters of `segments` CTAs: each CTA composites one contiguous part of the
 * ray samples and the parts are folded front to back through distributed shared memory (out = out_0 + T_0 out_1 ...).
 * Same samples, same weights; the fold changes the summation order (within the forward tolerance), which is why it is
 * off by default: with it a sample's last bits would depend on the size of the batch it rides in.
 * segments: 0 = default (never, unless the environment variable VB200_MARCH_SPLIT says otherwise), 1 = never,
 * -1 = automatic (two segments while the launch is under one wave: -11 % / -16 % march time at batch 1, bf16 / fp32),
 * 2 / 4 / 8 = always that many.  Process-wide. */
int vb200_render_set_march_split(int segments);

typedef struct VbRenderGrad {
  /* cotangents of the eight outputs (same shapes/dtypes as VbRenderOut; NULL = zero) */
  const float* g_rgb;
  const float* g_seg;
  const float* g_depth;
  const float* g_bev_rgb;
  const float* g_bev_seg;
  const float* g_bev_height;
  const float* g_voxel_density;
  const void* g_voxel_output;
  /* gradients of the inputs (fully overwritten), `dtype`; g_beta is 1 fp32 */
  void* g_density;
  void* g_sem;
  void* g_rgb_in;
  void* g_feat;
  float* g_beta;
} VbRenderGrad;

/* `out` = the forward's outputs (the compositing backward reuses them: suffix sums are formed
 * as total - prefix). */
int vb200_render_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbRenderIn* in,
                     int dtype, const VbRenderOut* out, const VbRenderGrad* grad, int branches,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- the callers right after the path (SURVEY §8f "next" rows 2-3) ------------------------------ */

/* nn.UpsamplingBilinear2d(scale_factor=factor), align_corners=True, on `planes` fp32 maps of H x W
 * (the rendered rgb / semantic / depth maps, BV2:210, 616-626).  Backward is a deterministic gather. */
int vb200_upsample_bilinear_fwd(const float* d_in, float* d_out, int planes, int H, int W, int factor, void* stream);
int vb200_upsample_bilinear_bwd(const float* d_gout, float* d_gin, int planes, int H, int W, int factor, void* stream);

/* F.grid_sample(volume, normalise(points), align_corners=True) for point / occupancy queries (BV2:576-609).
 *   d_vol   (B, channels, vZ, vY, vX) in `dtype`;  d_pts (B or 1, P, 3) fp32 EGO coordinates
 *   d_rot3x3 optional (B, 9): points are first rotated by bda[:3,:3] (the Occ3D grid, BV2:598-601)
 *   border         1 = padding_mode='border' (semantic logits), 0 = zeros
 *   apply_density  1 = sample sigma(volume) instead of the volume (occ_density, BV2:609); needs d_beta
 *   mask_invalid   1 = multiply by the in-range mask (pts_sdf, BV2:595)
 *   d_out (B, channels, P) fp32;  d_valid optional (B, P) uint8 in-range mask (BV2:587-589)
 * Backward scatters with atomicAdd (as ATen does); d_gvol (B, channels, vZ, vY, vX) in `dtype`. */
size_t vb200_query_points_bwd_workspace(const VbGrid* g, int channels, int P, int dtype);
int vb200_query_points_fwd(const VbGrid* g, const void* d_vol, int dtype, int channels, const float* d_pts, int P,
                           int pts_batched, const float* d_rot3x3, int border, int apply_density, int mask_invalid,
                           const float* d_beta, float* d_out, uint8_t* d_valid, void* stream);
int vb200_query_points_bwd(const VbGrid* g, const void* d_vol, int dtype, int channels, const float* d_pts, int P,
                           int pts_batched, const float* d_rot3x3, int border, int apply_density, int mask_invalid,
                           const float* d_beta, const float* d_gout, void* d_gvol, float* d_gbeta,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- producer of the lift's depth input (SURVEY §8f "next" row 1) ---------------------------------
 * Replaces `.softmax(dim=1)` of BV2:551 on the (B*N, D, fH, fW) depth logits: softmax over the D planes
 * (stride `inner` = fH*fW elements), `outer` = B*N.  fp32 arithmetic; logits fp32 / bf16 / fp16; probabilities in
 * the logits' dtype or fp32 (the reference's autocast behaviour).  One DRAM read + one write (cp.async staging).
 * Backward: d_glogits = probs * (d_gprobs - sum_D probs * d_gprobs); probs and d_gprobs share `dtype`;
 * out_dtype == dtype, or any dtype when dtype is fp32 (fp32 softmax output, fp16 conv gradient under AMP). */
int vb200_depth_softmax_fwd(const void* d_logits, int in_dtype, void* d_probs, int out_dtype, long long outer, int D,
                            int inner, void* stream);
int vb200_depth_softmax_bwd(const void* d_probs, const void* d_gprobs, int dtype, void* d_glogits, int out_dtype,
                            long long outer, int D, int inner, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VB200_H_ */
