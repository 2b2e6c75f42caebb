// vb_lift_plan.cu -- north-star kernel (a): frustum geometry -> voxel-index computation fused with the
// per-cell sort/segment, as a CACHEABLE plan, and the lift + pool forward that consumes it.
//
// get_pixel (BV2:351-388) and the sampling coordinates of get_voxel_feats (BV2:493-507) depend on the camera
// matrices only.  In validation / test they never change (deterministic ida, nusc_det_seg_dataset.py:489-498;
// identity bda, base_exp.py:113-120), so the strict fp32 projection -- ~360 instructions per (voxel, camera)
// pair in a forward that is 77 % issue-bound, plus a count/scan/fill/sort per backward -- is computed ONCE per
// distinct set of matrices and kept in HBM (46 MB per sample at the R50 config; 180 GB of HBM3e make hundreds of
// distinct rigs cacheable).  The plan holds only integers the strict chain produced and the exact fractions
// `ix - x0`, so the kernels that consume it are bit-identical to the ones that recompute the projection.
//
//   plan (one per sample; VbLiftPlan in vb200.h)
//     head[vox]       (first << 4) | count            valid pairs of the voxel, cameras ascending
//     pairs[P]        {cam << 29 | (z0+1) << 20 | (y0+1) << 10 | (x0+1), fx, fy, fz}     voxel-major
//     cell_off[nc+1]  CSR over destination pixel cells (n, y0+1, x0+1)
//     cell_recs[P]    {(z0+1) << 21 | voxel, fx, fy, fz}   cell-major, sorted by (z0, voxel)  -> deterministic bwd
//
//   build:  plan_count (valid pairs per voxel + per cell) -> 2 exclusive scans -> plan_fill -> per-cell rank sort
#include "vb_lift_common.cuh"
#include "vb_lift_pairs.cuh"
#include "vb_scan.cuh"
#include "vb_trace.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kXBits = 10, kYBits = 10, kZBits = 9;   // key fields: x0+1 <= fW, y0+1 <= fH, z0+1 <= D

struct PlanProlog {
  bool has_bda, affine;
};
__device__ __forceinline__ PlanProlog plan_prolog(const VbGrid& g, const float* d_mats, float* s_m, float* s_q, int b) {
  stage_mats(s_m, d_mats, b, g.N);
  __syncthreads();
  PlanProlog p;
  p.has_bda = (g.has_bda != 0) && !block_is_identity(s_m);
  p.affine = block_pixel_affine(s_m, g.N, p.has_bda);
  stage_cull(s_q, s_m, g.N, p.has_bda);
  __syncthreads();
  return p;
}

__global__ void __launch_bounds__(kThreads) plan_count_kernel(VbGrid g, VbTables t, VbLiftDiv dv,
                                                              const float* __restrict__ d_mats, int* __restrict__ vcount,
                                                              int* __restrict__ ccount) {
  __shared__ float s_m[VB_MAX_CAMS * VB200_MAT_SLOTS * 16];
  __shared__ float s_q[VB_MAX_CAMS * 16];
  const int b = blockIdx.y;
  const PlanProlog pp = plan_prolog(g, d_mats, s_m, s_q, b);
  const int nvox = g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * kThreads + threadIdx.x;
  if (vox >= nvox) return;
  const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
  const float px = __ldg(t.xs + x), py = __ldg(t.ys + y), pz = __ldg(t.zs + z);
  const CellDims cd = cell_dims(g);
  int cnt = 0;
  for (int n = 0; n < g.N; ++n) {
    LiftCoord lc;
    if (!pair_coord(g, s_m, s_q, pp.has_bda, pp.affine, dv, n, px, py, pz, lc)) continue;
    ++cnt;
    atomicAdd(ccount + (size_t)b * cd.nc + (n * cd.ncy + (lc.y0 + 1)) * cd.ncx + (lc.x0 + 1), 1);
  }
  vcount[(size_t)b * nvox + vox] = cnt;
}

__global__ void __launch_bounds__(kThreads) plan_fill_kernel(VbGrid g, VbTables t, VbLiftDiv dv,
                                                             const float* __restrict__ d_mats,
                                                             const int* __restrict__ vfirst,
                                                             const int* __restrict__ cell_off, int* __restrict__ cursor,
                                                             uint32_t* __restrict__ head, uint4* __restrict__ pairs,
                                                             uint4* __restrict__ cell_tmp, long long capacity) {
  __shared__ float s_m[VB_MAX_CAMS * VB200_MAT_SLOTS * 16];
  __shared__ float s_q[VB_MAX_CAMS * 16];
  const int b = blockIdx.y;
  const PlanProlog pp = plan_prolog(g, d_mats, s_m, s_q, b);
  const int nvox = g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * kThreads + threadIdx.x;
  if (vox >= nvox) return;
  const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
  const float px = __ldg(t.xs + x), py = __ldg(t.ys + y), pz = __ldg(t.zs + z);
  const CellDims cd = cell_dims(g);
  const int first = vfirst[(size_t)b * (nvox + 1) + vox];
  int cnt = 0;
  for (int n = 0; n < g.N; ++n) {
    LiftCoord lc;
    if (!pair_coord(g, s_m, s_q, pp.has_bda, pp.affine, dv, n, px, py, pz, lc)) continue;
    // exactly the far weights tri_weights() forms: ix - (float)x0 (the near weights are 1 - that, bit for bit,
    // whenever the near corner is inside the grid; see vb_lift_plan.cu header of lift_fwd_planned_kernel)
    const uint32_t fx = __float_as_uint(lc.ix - (float)lc.x0), fy = __float_as_uint(lc.iy - (float)lc.y0),
                   fz = __float_as_uint(lc.iz - (float)lc.z0);
    const long long slot = (long long)first + cnt;
    if (slot < capacity) {
      const uint32_t key = ((uint32_t)n << (kXBits + kYBits + kZBits)) | ((uint32_t)(lc.z0 + 1) << (kXBits + kYBits)) |
                           ((uint32_t)(lc.y0 + 1) << kXBits) | (uint32_t)(lc.x0 + 1);
      pairs[(size_t)b * capacity + slot] = make_uint4(key, fx, fy, fz);
    }
    const int cell = (n * cd.ncy + (lc.y0 + 1)) * cd.ncx + (lc.x0 + 1);
    const long long cslot =
        (long long)cell_off[(size_t)b * (cd.nc + 1) + cell] + atomicAdd(cursor + (size_t)b * cd.nc + cell, 1);
    if (cslot < capacity)
      cell_tmp[(size_t)b * capacity + cslot] = make_uint4(((uint32_t)(lc.z0 + 1) << kVoxBits) | (uint32_t)vox, fx, fy, fz);
    ++cnt;
  }
  head[(size_t)b * nvox + vox] = ((uint32_t)first << 4) | (uint32_t)cnt;
}

// warp per cell: rank-sort the segment by its key (distinct: distinct voxels) => order fixed by (z0, voxel)
__global__ void __launch_bounds__(kThreads) plan_sort_cells_kernel(const int* __restrict__ cell_off,
                                                                   const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                                   int nc, long long capacity, int total_cells) {
  const int cell_g = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (cell_g >= total_cells) return;
  const int lane = threadIdx.x & 31;
  const int b = cell_g / nc, cell = cell_g % nc;
  const int* off = cell_off + (size_t)b * (nc + 1);
  const long long lo = off[cell];
  long long hi = off[cell + 1];
  if (hi > capacity) hi = capacity;       // overflowing build: the caller retries with a larger capacity
  const int L = (int)(hi - lo);
  if (L <= 0) return;
  const uint4* s = src + (size_t)b * capacity + lo;
  uint4* d = dst + (size_t)b * capacity + lo;
  for (int i = lane; i < L; i += 32) {
    const uint4 r = s[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) rank += (__ldg(&s[j].x) < r.x) ? 1 : 0;
    d[rank] = r;
  }
}

__global__ void plan_totals_kernel(const int* __restrict__ vfirst, int nvox, int32_t* __restrict__ num_pairs, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) num_pairs[b] = vfirst[(size_t)b * (nvox + 1) + nvox];
}

size_t a256(size_t n) { return (n + 255) & ~(size_t)255; }

struct PlanWs {
  size_t vcount, vfirst, ccount, cursor, chunk_sums, cell_tmp, total;
};
PlanWs plan_ws(const VbGrid* g, long long capacity) {
  const CellDims cd = cell_dims(*g);
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX;
  PlanWs w;
  size_t o = 0;
  w.vcount = o;     o += a256((size_t)g->B * nvox * 4);
  w.vfirst = o;     o += a256((size_t)g->B * (nvox + 1) * 4);
  w.ccount = o;     o += a256((size_t)g->B * cd.nc * 4);
  w.cursor = o;     o += a256((size_t)g->B * cd.nc * 4);
  w.chunk_sums = o; o += a256((size_t)g->B * 1024 * 4);
  w.cell_tmp = o;   o += a256((size_t)g->B * (size_t)capacity * 16);
  w.total = o;
  return w;
}

// ---- forward from a plan: one thread per voxel, loop over its valid pairs ----------------------------------------
// Same arithmetic as lift_pool_fwd_kernel (vb_lift.cu) on the same integers and fractions, cameras in the same
// (ascending) order => bit-identical output.  (Near weights: tri_weights() forms (float)(x0+1) - ix, here 1 - fx
// with fx = ix - (float)x0.  For x0 >= 0 the subtraction ix - x0 is exact, so both are the rounding of the same
// real number; for x0 = -1 the near corner is outside the grid and its weight is zeroed either way.)
#ifndef VB_LIFT_PLANNED_MINB
#define VB_LIFT_PLANNED_MINB 8   // measured (R50, B=8, bf16): 0.270 ms at 8 blocks/SM, 0.279 at 7, 0.339 at 5
#endif
constexpr int kFwdThreads = 128;

template <typename TD, typename TC, int C, int OUT_LAYOUT>
__global__ void __launch_bounds__(kFwdThreads, VB_LIFT_PLANNED_MINB) lift_fwd_planned_kernel(
    VbGrid g, const VbLiftPlan* __restrict__ plans, const TD* __restrict__ depth, const TC* __restrict__ ctx_nhwc,
    TD* __restrict__ out, uint64_t* __restrict__ cnt_out) {
  const int b = blockIdx.y;
  const int nvox = g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * kFwdThreads + threadIdx.x;
  if (vox >= nvox) return;
  const uint32_t* __restrict__ head = plans[b].head;
  const uint4* __restrict__ pairs = reinterpret_cast<const uint4*>(plans[b].pairs);
  const uint32_t h = __ldg(head + vox);
  const int cnt = (int)(h & 15u);
  const uint4* pr = pairs + (h >> 4);
  const int HW = g.fH * g.fW;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0f;
  int cams_seen = 0;
  uint64_t zero_cnt = 0;
  for (int j = 0; j < cnt; ++j) {
    const uint4 r = __ldg(pr + j);
    const int x0 = (int)(r.x & ((1u << kXBits) - 1)) - 1;
    const int y0 = (int)((r.x >> kXBits) & ((1u << kYBits) - 1)) - 1;
    const int z0 = (int)((r.x >> (kXBits + kYBits)) & ((1u << kZBits) - 1)) - 1;
    const int n = (int)(r.x >> (kXBits + kYBits + kZBits));
    const float fx = __uint_as_float(r.y), fy = __uint_as_float(r.z), fz = __uint_as_float(r.w);
    const TD* dcam = depth + (size_t)(b * g.N + n) * g.D * HW;
    const TC* ccam = ctx_nhwc + (size_t)(b * g.N + n) * HW * C;
    float f[C];
    lift_pair_gather<TD, TC, C>(g, dcam, ccam, HW, x0, y0, z0, 1.0f - fx, fx, 1.0f - fy, fy, 1.0f - fz, fz, f);
    lift_accumulate<C>(f, acc, cams_seen, zero_cnt);
  }
  lift_store<TD, C, OUT_LAYOUT>(out, cnt_out, b, nvox, vox, acc, cams_seen, zero_cnt);
}

template <typename TD, typename TC>
int launch_fwd_planned(const VbGrid* g, const VbLiftPlan* d_plans, const void* d_depth, const void* d_ctx, void* d_out,
                       int out_layout, uint64_t* d_cnt, void* ws, cudaStream_t st) {
  constexpr int C = 16;
  if (g->C != C) return VB200_ERR_ARG;
  TC* ctx_nhwc = reinterpret_cast<TC*>(ws);
  {
    dim3 grid(g->fH, g->B * g->N);
    const size_t smem = (size_t)C * (g->fW + 1) * sizeof(TC);
    VbTraceScope tr(VB_K_CTX_NHWC, st);
    ctx_to_nhwc_kernel<TC, C><<<grid, 256, smem, st>>>(reinterpret_cast<const TC*>(d_ctx), ctx_nhwc, g->fH, g->fW);
    VB_LAUNCH_CHECK();
  }
  const int nvox = g->vZ * g->vY * g->vX;
  dim3 grid(vb_ceil_div(nvox, kFwdThreads), g->B);
  VbTraceScope tr(VB_K_LIFT_FWD, st);
  if (out_layout == VB200_NCDHW)
    lift_fwd_planned_kernel<TD, TC, C, VB200_NCDHW><<<grid, kFwdThreads, 0, st>>>(
        *g, d_plans, reinterpret_cast<const TD*>(d_depth), ctx_nhwc, reinterpret_cast<TD*>(d_out), d_cnt);
  else
    lift_fwd_planned_kernel<TD, TC, C, VB200_NDHWC><<<grid, kFwdThreads, 0, st>>>(
        *g, d_plans, reinterpret_cast<const TD*>(d_depth), ctx_nhwc, reinterpret_cast<TD*>(d_out), d_cnt);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

bool plan_dims_ok(const VbGrid* g) {
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX;
  return g->N >= 1 && g->N <= VB_MAX_CAMS && g->fW + 1 < (1 << kXBits) && g->fH + 1 < (1 << kYBits) &&
         g->D + 1 < (1 << kZBits) && nvox <= (1u << kVoxBits) && (size_t)g->N * nvox < (1u << 28);
}

}  // namespace

extern "C" size_t vb200_lift_plan_workspace(const VbGrid* g, long long capacity) {
  if (!g || capacity < 0) return 0;
  return plan_ws(g, capacity).total;
}

extern "C" int vb200_lift_plan_build(const VbGrid* g, const VbTables* t, const float* d_mats, uint32_t* d_head,
                                     void* d_pairs, int32_t* d_cell_off, void* d_cell_recs, long long capacity,
                                     int32_t* d_num_pairs, void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && d_head && d_pairs && d_cell_off && d_cell_recs && d_num_pairs && d_workspace);
  VB_CHECK_ARG(g->B > 0 && capacity > 0 && g->D >= 1);
  VB_CHECK_ARG(plan_dims_ok(g));
  if (workspace_bytes < vb200_lift_plan_workspace(g, capacity)) return VB200_ERR_WORKSPACE;
  if (((uintptr_t)d_workspace | (uintptr_t)d_pairs | (uintptr_t)d_cell_recs) & 15) return VB200_ERR_ALIGN;
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(d_workspace);
  const PlanWs w = plan_ws(g, capacity);
  const CellDims cd = cell_dims(*g);
  const int nvox = g->vZ * g->vY * g->vX;
  int* vcount = reinterpret_cast<int*>(ws + w.vcount);
  int* vfirst = reinterpret_cast<int*>(ws + w.vfirst);
  int* ccount = reinterpret_cast<int*>(ws + w.ccount);
  int* cursor = reinterpret_cast<int*>(ws + w.cursor);
  int* chunk_sums = reinterpret_cast<int*>(ws + w.chunk_sums);
  uint4* cell_tmp = reinterpret_cast<uint4*>(ws + w.cell_tmp);
  const VbLiftDiv dv = vb_lift_div(g);
  if (cudaMemsetAsync(ccount, 0, (size_t)g->B * cd.nc * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
  if (cudaMemsetAsync(cursor, 0, (size_t)g->B * cd.nc * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
  VbTraceScope tr(VB_K_LIFT_PLAN, st, 10);
  dim3 vgrid(vb_ceil_div(nvox, kThreads), g->B);
  plan_count_kernel<<<vgrid, kThreads, 0, st>>>(*g, *t, dv, d_mats, vcount, ccount);
  VB_LAUNCH_CHECK();
  rc = vb_exclusive_scan(vcount, vfirst, chunk_sums, nvox, g->B, st);
  if (rc) return rc;
  rc = vb_exclusive_scan(ccount, d_cell_off, chunk_sums, cd.nc, g->B, st);
  if (rc) return rc;
  plan_fill_kernel<<<vgrid, kThreads, 0, st>>>(*g, *t, dv, d_mats, vfirst, d_cell_off, cursor, d_head,
                                              reinterpret_cast<uint4*>(d_pairs), cell_tmp, capacity);
  VB_LAUNCH_CHECK();
  const int total_cells = g->B * cd.nc;
  plan_sort_cells_kernel<<<vb_ceil_div(total_cells, kThreads / 32), kThreads, 0, st>>>(
      d_cell_off, cell_tmp, reinterpret_cast<uint4*>(d_cell_recs), cd.nc, capacity, total_cells);
  VB_LAUNCH_CHECK();
  plan_totals_kernel<<<vb_ceil_div(g->B, 64), 64, 0, st>>>(vfirst, nvox, d_num_pairs, g->B);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" int vb200_lift_pool_fwd_planned(const VbGrid* g, const VbLiftPlan* d_plans, const void* d_depth,
                                           const void* d_ctx, int dtype, int ctx_dtype, void* d_out, int out_layout,
                                           uint64_t* d_cnt, void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && d_plans && d_depth && d_ctx && d_out && d_workspace);
  VB_CHECK_ARG(g->B > 0 && g->D >= 1);
  VB_CHECK_ARG(plan_dims_ok(g));
  VB_CHECK_ARG(out_layout == VB200_NCDHW || out_layout == VB200_NDHWC);
  if (workspace_bytes < vb200_lift_pool_fwd_workspace(g, ctx_dtype)) return VB200_ERR_WORKSPACE;
  if (((uintptr_t)d_out | (uintptr_t)d_ctx | (uintptr_t)d_depth | (uintptr_t)d_plans) & 15) return VB200_ERR_ALIGN;
  if ((uintptr_t)d_workspace & 31) return VB200_ERR_ALIGN;   // the channels-last ctx copy is read with 256-bit loads
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
#define VB_CALL(TD, TC) launch_fwd_planned<TD, TC>(g, d_plans, d_depth, d_ctx, d_out, out_layout, d_cnt, d_workspace, st)
  VB_LIFT_DISPATCH(dtype, ctx_dtype, VB_CALL);
#undef VB_CALL
}
