// vb_lift_bwd.cu -- Bk for L1-L4: deterministic backward of the fused lift + pool.
//
// The forward is a gather (voxel <- 8 frustum cells per seeing camera), so its backward is a
// scatter into d_depth / d_ctx.  ATen's grid_sampler_3d_backward does that scatter with float
// atomicAdd into a 372 MB grad-frustum (order-dependent bits); here the scatter is turned back
// into a gather by SORTING the valid (voxel, camera) pairs by destination pixel cell
// (SURVEY §7.1 a/d, A.5.6):
//
//   plan_pairs<count>   histogram of valid pairs per cell (n, y0, x0)            [int atomics]
//   scan_chunks/sums/add exclusive scan -> CSR offsets (chunked 3-pass scan)
//   plan_pairs<fill>    records ((z0+1) << 21 | voxel) dropped into their cell's segment
//   sort_cells          warp per cell: rank-sort the segment => order fixed by (z0, voxel)
//   lift_bwd            warp per cell, ONE launch over all cells: per-warp shared-memory depth bins and register
//                       d_ctx accumulators, written out as the cell's own partials (4 corner pixels x D depth
//                       bins, 4 x C channels) with coalesced stores -- no read-modify-write of the gradients.
//                       Phase A (lane = pair): bit-exact re-projection (or the cached plan record), 8 depth loads,
//                       16-channel dot products, warp-shuffle segmented reduction over equal depth bins.
//                       Phase B (lane = channel): serial accumulation of d_ctx in sorted order.
//                       (Round 1 ran 4 launches over a 2x2 cell colouring and flushed with `gd[..] += v`: ncu put
//                        41 % of all stall samples on that dependent global load -> add -> store.)
//   lift_bwd_reduce     per pixel: the sum of the <= 4 cells it is a corner of, in fixed order, cast to the caller's
//                       dtype and layout (d_depth (B,N,D,fH,fW), d_ctx (B,N,C,fH,fW)).
//
// No float atomics anywhere => bit-reproducible gradients; no grad-frustum is ever formed.
#include "vb_lift_common.cuh"
#include "vb_lift_pairs.cuh"
#include "vb_scan.cuh"
#include "vb_trace.cuh"

namespace {

constexpr int kC = 16;
constexpr int kThreads = 256;

// ---- plan: count / fill ---------------------------------------------------------------------------
// (measured alternative: the count pass parking (cell, z0) per pair so that the fill pass need not project again --
//  B=1 0.105 -> 0.100 ms, B=8 0.61 -> 0.64 ms: the fill is bound by its atomics and scattered stores, not by the
//  projection; the two symmetric passes stay)
template <int MODE>
__global__ void __launch_bounds__(kThreads) plan_pairs_kernel(VbGrid g, VbTables t, VbLiftDiv dv,
                                                              const float* __restrict__ d_mats,
                                                              int* __restrict__ counts, const int* __restrict__ offsets,
                                                              int* __restrict__ cursor, uint32_t* __restrict__ recs) {
  __shared__ float s_m[VB_MAX_CAMS * VB200_MAT_SLOTS * 16];
  __shared__ float s_q[VB_MAX_CAMS * 16];
  const int b = blockIdx.y;
  stage_mats(s_m, d_mats, b, g.N);
  __syncthreads();
  const bool has_bda = (g.has_bda != 0) && !block_is_identity(s_m);
  const bool affine = block_pixel_affine(s_m, g.N, has_bda);
  stage_cull(s_q, s_m, g.N, has_bda);
  __syncthreads();
  const int nvox = g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * kThreads + threadIdx.x;
  if (vox >= nvox) return;
  const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
  const float px = __ldg(t.xs + x), py = __ldg(t.ys + y), pz = __ldg(t.zs + z);
  const CellDims cd = cell_dims(g);
  for (int n = 0; n < g.N; ++n) {
    LiftCoord lc;
    if (!pair_coord(g, s_m, s_q, has_bda, affine, dv, n, px, py, pz, lc)) continue;
    const int cell = (n * cd.ncy + (lc.y0 + 1)) * cd.ncx + (lc.x0 + 1);
    if (MODE == 0) {
      atomicAdd(counts + (size_t)b * cd.nc + cell, 1);
    } else {
      const int slot = offsets[(size_t)b * (cd.nc + 1) + cell] + atomicAdd(cursor + (size_t)b * cd.nc + cell, 1);
      recs[(size_t)b * g.N * nvox + slot] = ((uint32_t)(lc.z0 + 1) << kVoxBits) | (uint32_t)vox;
    }
  }
}

// ---- per-cell rank sort: warp per cell -----------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) sort_cells_kernel(const int* __restrict__ offsets,
                                                              const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                              int nc, size_t pairs_per_sample, int total_cells) {
  const int cell_g = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (cell_g >= total_cells) return;
  const int lane = threadIdx.x & 31;
  const int b = cell_g / nc, cell = cell_g % nc;
  const int* off = offsets + (size_t)b * (nc + 1);
  const int lo = off[cell], L = off[cell + 1] - lo;
  const uint32_t* s = src + (size_t)b * pairs_per_sample + lo;
  uint32_t* d = dst + (size_t)b * pairs_per_sample + lo;
  for (int i = lane; i < L; i += 32) {
    const uint32_t r = s[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) rank += (__ldg(s + j) < r) ? 1 : 0;   // records are distinct (distinct voxels)
    d[rank] = r;
  }
}

// ---- ctx (B,N,C,fH,fW) T -> (B,N,fH,fW,C) fp32 (same pre-pass as the forward) ---------------------------
template <typename T>
__global__ void __launch_bounds__(256) ctx_to_nhwc_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int fH,
                                                              int fW) {
  extern __shared__ float s_t[];  // [C][fW + 1]
  const int h = blockIdx.x, bn = blockIdx.y, ld = fW + 1;
  for (int i = threadIdx.x; i < kC * fW; i += blockDim.x) {
    const int c = i / fW, w = i % fW;
    s_t[c * ld + w] = VbType<T>::ld(src + (((size_t)bn * kC + c) * fH + h) * fW + w);
  }
  __syncthreads();
  float* out = dst + ((size_t)bn * fH + h) * fW * kC;
  for (int i = threadIdx.x; i < kC * fW; i += blockDim.x) out[i] = s_t[(i % kC) * ld + i / kC];
}

// ---- g' = d_out / (count + 1e-6), once per voxel, channels-last fp32 ---------------------------------------
// The gather below visits the voxels of a pixel cell, which lie along a camera ray and are scattered through
// the volume: read straight from an NCDHW cotangent every one of a pair's 16 channel values is its own 32-byte
// sector (ncu: 110 MB of DRAM reads per colour launch for 21 MB of useful values) and the 16 IEEE divisions by
// the camera counts are repeated for every camera that sees the voxel.  One streaming pass writes g' as 64
// contiguous bytes per voxel instead (the non-zero mask is not differentiated).
template <typename T, int GOUT_LAYOUT>
__global__ void __launch_bounds__(256) gout_prep_kernel(const T* __restrict__ gout, const uint64_t* __restrict__ cnt,
                                                        float* __restrict__ gprep, int nvox) {
  const int vox = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (vox >= nvox) return;
  const uint64_t cw = cnt[(size_t)b * nvox + vox];
  float gp[kC];
  if (GOUT_LAYOUT == VB200_NCDHW) {
#pragma unroll
    for (int c = 0; c < kC; ++c) gp[c] = VbType<T>::ld(gout + ((size_t)b * kC + c) * nvox + vox);
  } else {
    const T* gv = gout + ((size_t)b * nvox + vox) * kC;
    constexpr int Ln = VbLanes<T>::n;
#pragma unroll
    for (int q4 = 0; q4 < kC / Ln; ++q4) VbVec<T, Ln>::ld(gv + q4 * Ln, &gp[q4 * Ln]);
  }
#pragma unroll
  for (int c = 0; c < kC; ++c) gp[c] = gp[c] / ((float)((cw >> (4 * c)) & 0xf) + 1e-6f);
  float4* o = reinterpret_cast<float4*>(gprep + ((size_t)b * nvox + vox) * kC);
#pragma unroll
  for (int q4 = 0; q4 < kC / 4; ++q4) o[q4] = make_float4(gp[4 * q4], gp[4 * q4 + 1], gp[4 * q4 + 2], gp[4 * q4 + 3]);
}

// ---- the backward proper -------------------------------------------------------------------------------
#ifndef VB_LIFT_BWD_MINB
#define VB_LIFT_BWD_MINB 4
#endif
// PLANNED: the cell segments come from a cached VbLiftPlan (one per sample): 16-byte records
// {(z0+1) << 21 | voxel, fx, fy, fz} already sorted by (z0, voxel) -- no re-projection, no per-call plan.
template <typename T, bool PLANNED>
__global__ void __launch_bounds__(kThreads, VB_LIFT_BWD_MINB) lift_bwd_kernel(VbGrid g, VbTables t, VbLiftDiv dv,
                                                            const float* __restrict__ d_mats,
                                                            const VbLiftPlan* __restrict__ plans,
                                                            const T* __restrict__ depth,
                                                            const float* __restrict__ ctx_nhwc,
                                                            const float* __restrict__ gprep,
                                                            const int* __restrict__ offsets,
                                                            const uint32_t* __restrict__ recs,
                                                            float* __restrict__ part_depth,
                                                            float* __restrict__ part_ctx) {
  extern __shared__ float s_dyn[];
  __shared__ float s_m[VB200_MAT_SLOTS * 16];
  const CellDims cd = cell_dims(g);
  // blockIdx.y = b * N + n so that a block shares one camera's matrices
  const int bn = blockIdx.y, b = bn / g.N, n = bn % g.N;
  bool has_bda = false, affine = false;
  if (!PLANNED) {
    for (int i = threadIdx.x; i < VB200_MAT_SLOTS * 16; i += blockDim.x)
      s_m[i] = __ldg(d_mats + (size_t)bn * VB200_MAT_SLOTS * 16 + i);
    __syncthreads();
    has_bda = (g.has_bda != 0) && !block_is_identity(s_m);
    affine = block_pixel_affine(s_m, 1, has_bda);      // this block's one camera
  }

  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci = blockIdx.x * (kThreads / 32) + wid;
  if (ci >= cd.ncy * cd.ncx) return;
  const int cy = ci / cd.ncx, cx = ci % cd.ncx;
  const int y0 = cy - 1, x0 = cx - 1;
  const int cell = (n * cd.ncy + cy) * cd.ncx + cx;
  const int* off = PLANNED ? plans[b].cell_off : offsets + (size_t)b * (cd.nc + 1);
  const int seg_lo = off[cell], L = off[cell + 1] - seg_lo;
  if (L == 0) return;

  const int D = g.D, HW = g.fH * g.fW;
  const int nvox = g.vZ * g.vY * g.vX;
  // per-warp shared memory: bins[4][D] | pa[32][4] | pg[32][17] | cv[4][16]
  const int per_warp = 4 * D + 32 * 4 + 32 * 17 + 64;
  float* bins = s_dyn + wid * per_warp;
  float* pa = bins + 4 * D;
  float* pg = pa + 32 * 4;
  float* cv = pg + 32 * 17;

  // the cell's four corner pixels (zeros padding: clamp the address, remember in-bounds)
  const int xa = max(x0, 0), xb = min(x0 + 1, g.fW - 1);
  const int ya = max(y0, 0), yb = min(y0 + 1, g.fH - 1);
  const bool inx0 = x0 >= 0, inx1 = x0 + 1 < g.fW, iny0 = y0 >= 0, iny1 = y0 + 1 < g.fH;
  const int pxl[4] = {ya * g.fW + xa, ya * g.fW + xb, yb * g.fW + xa, yb * g.fW + xb};
  const float* ccam = ctx_nhwc + (size_t)bn * HW * kC;
  for (int i = lane; i < 64; i += 32) cv[i] = __ldg(ccam + (size_t)pxl[i >> 4] * kC + (i & 15));
  for (int i = lane; i < 4 * D; i += 32) bins[i] = 0.0f;
  __syncwarp();

  const T* dcam = depth + (size_t)bn * D * HW;
  const uint32_t* seg = PLANNED ? nullptr : recs + (size_t)b * g.N * nvox + seg_lo;
  const uint4* pseg = PLANNED ? reinterpret_cast<const uint4*>(plans[b].cell_recs) + seg_lo : nullptr;
  const int c_ = lane & 15, half = lane >> 4;
  float accC[4] = {0.f, 0.f, 0.f, 0.f};

  for (int r0 = 0; r0 < L; r0 += 32) {
    const int np = min(32, L - r0);
    const bool act = lane < np;
    // ---- phase A: lane = pair ----
    int z0 = 100000 + lane;   // inactive lanes: unique keys => single-lane runs of zeros
    float vA[4] = {0.f, 0.f, 0.f, 0.f}, vB[4] = {0.f, 0.f, 0.f, 0.f};
    if (act) {
      int vox;
      float wx0, wx1, wy0, wy1, wza, wzb;
      if (PLANNED) {
        const uint4 pr = __ldg(pseg + r0 + lane);
        vox = (int)(pr.x & ((1u << kVoxBits) - 1));
        z0 = (int)(pr.x >> kVoxBits) - 1;
        const float fx = __uint_as_float(pr.y), fy = __uint_as_float(pr.z), fz = __uint_as_float(pr.w);
        wx0 = inx0 ? 1.0f - fx : 0.0f; wx1 = inx1 ? fx : 0.0f;
        wy0 = iny0 ? 1.0f - fy : 0.0f; wy1 = iny1 ? fy : 0.0f;
        wza = z0 >= 0 ? 1.0f - fz : 0.0f;
        wzb = z0 + 1 < D ? fz : 0.0f;
      } else {
        const uint32_t rec = seg[r0 + lane];
        vox = (int)(rec & ((1u << kVoxBits) - 1));
        const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
        // same code as the plan => same (x0, y0, z0)
        const LiftCoord lc = pair_strict(g, s_m, has_bda, affine, dv, __ldg(t.xs + x), __ldg(t.ys + y), __ldg(t.zs + z));
        z0 = lc.z0;
        wx0 = inx0 ? (float)(lc.x0 + 1) - lc.ix : 0.0f; wx1 = inx1 ? lc.ix - (float)lc.x0 : 0.0f;
        wy0 = iny0 ? (float)(lc.y0 + 1) - lc.iy : 0.0f; wy1 = iny1 ? lc.iy - (float)lc.y0 : 0.0f;
        wza = lc.z0 >= 0 ? (float)(lc.z0 + 1) - lc.iz : 0.0f;
        wzb = lc.z0 + 1 < D ? lc.iz - (float)lc.z0 : 0.0f;
      }
      const int za = max(z0, 0), zb = min(z0 + 1, D - 1);
      const float wxy[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
      const T* d0 = dcam + (size_t)za * HW;
      const T* d1 = dcam + (size_t)zb * HW;
      // g' = d_out * valid / (count + 1e-6), prepared per voxel by gout_prep_kernel: 4 x 128-bit loads
      float gp[kC];
      {
        const float4* gv = reinterpret_cast<const float4*>(gprep + ((size_t)b * nvox + vox) * kC);
#pragma unroll
        for (int q4 = 0; q4 < kC / 4; ++q4) {
          const float4 v = __ldg(gv + q4);
          gp[4 * q4] = v.x; gp[4 * q4 + 1] = v.y; gp[4 * q4 + 2] = v.z; gp[4 * q4 + 3] = v.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float s = fmaf(wzb, VbType<T>::ld(d1 + pxl[k]), wza * VbType<T>::ld(d0 + pxl[k]));
        pa[lane * 4 + k] = wxy[k] * s;                       // d_ctx[c] += g'[c] * w_jk * S_jk
        float dot = 0.0f;
#pragma unroll
        for (int c = 0; c < kC; ++c) dot = fmaf(gp[c], cv[k * 16 + c], dot);
        vA[k] = wza * wxy[k] * dot;                          // d_depth[z0]   += w_z0 * w_jk * sum_c g' ctx
        vB[k] = wzb * wxy[k] * dot;                          // d_depth[z0+1] += w_z1 * ...
      }
#pragma unroll
      for (int c = 0; c < kC; ++c) pg[lane * 17 + c] = gp[c];
    }
    // segmented reduction over runs of equal z0 (contiguous: the segment is sorted by (z0, voxel))
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int zo = __shfl_down_sync(0xffffffffu, z0, o);
      const bool take = (lane + o < 32) && (zo == z0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = __shfl_down_sync(0xffffffffu, vA[k], o);
        const float bb = __shfl_down_sync(0xffffffffu, vB[k], o);
        if (take) { vA[k] += a; vB[k] += bb; }
      }
    }
    const int zprev = __shfl_up_sync(0xffffffffu, z0, 1);
    const bool head = act && (lane == 0 || zprev != z0);
    // heads of one round have distinct z0 => distinct bins inside each sub-phase
    if (head && z0 >= 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) bins[k * D + z0] += vA[k];
    }
    __syncwarp();
    if (head && z0 + 1 < D) {
#pragma unroll
      for (int k = 0; k < 4; ++k) bins[k * D + z0 + 1] += vB[k];
    }
    __syncwarp();
    // ---- phase B: lane = (channel, half); serial over the round's pairs in sorted order ----
    for (int p = half; p < np; p += 2) {
      const float gpc = pg[p * 17 + c_];
#pragma unroll
      for (int k = 0; k < 4; ++k) accC[k] = fmaf(gpc, pa[p * 4 + k], accC[k]);
    }
    __syncwarp();
  }

  // ---- flush: the cell's own partials, coalesced, no read-modify-write (empty cells return above and are skipped
  // by the reduce, which reads the same CSR offsets)
  float* pd = part_depth + ((size_t)b * cd.nc + cell) * 4 * D;
  for (int i = lane; i < 4 * D; i += 32) pd[i] = bins[i];
#pragma unroll
  for (int k = 0; k < 4; ++k) accC[k] += __shfl_down_sync(0xffffffffu, accC[k], 16);
  if (half == 0) {
    float* pc = part_ctx + ((size_t)b * cd.nc + cell) * 4 * kC;
#pragma unroll
    for (int k = 0; k < 4; ++k) pc[k * kC + c_] = accC[k];
  }
}

// ---- per pixel: sum of the cells it is a corner of -------------------------------------------------------------------
// Pixel (y, x) of camera n is corner k of cell (cy, cx) = (y + 1 - (k >> 1), x + 1 - (k & 1)): k = 0 (ya, xa) of the cell
// whose base is the pixel itself ... k = 3 (yb, xb) of the cell one up-left.  Fixed order k = 0..3 => deterministic.
// The partials are [cell][corner][bin] (bins contiguous), the gradients [bin][pixel] (pixels contiguous): a block
// transposes a tile of kReduceTile pixels of one image row through shared memory -- a warp reads one pixel's four
// rows of D bins with the lanes along the bins (coalesced), then the tile is written with the lanes along the pixels
// (coalesced).  The first version had the lanes along the pixels for both and re-fetched every 32-byte sector of the
// partials eight times from L2 (ncu: 525 us at 8 % issue, 29 % DRAM for 1.2 GB).
constexpr int kReduceTile = 32;
template <typename TD, typename TC>
__global__ void __launch_bounds__(256) lift_bwd_reduce_kernel(VbGrid g, const VbLiftPlan* __restrict__ plans,
                                                              const int* __restrict__ offsets,
                                                              const float* __restrict__ part_depth,
                                                              const float* __restrict__ part_ctx, TD* __restrict__ gdepth,
                                                              TC* __restrict__ gctx) {
  extern __shared__ float s_tile[];                  // [D + kC][kReduceTile + 1]
  constexpr int LD = kReduceTile + 1;
  const CellDims cd = cell_dims(g);
  const int tiles_x = (g.fW + kReduceTile - 1) / kReduceTile;
  const int y = blockIdx.x / tiles_x, x0 = (blockIdx.x % tiles_x) * kReduceTile;
  const int bn = blockIdx.y, b = bn / g.N, n = bn % g.N;
  const int D = g.D, HW = g.fH * g.fW;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int* off = plans ? plans[b].cell_off : offsets + (size_t)b * (cd.nc + 1);
  const int npx = min(kReduceTile, g.fW - x0);
  float* s_ctx = s_tile + (size_t)D * LD;
  for (int px = wid; px < npx; px += nw) {           // a warp per pixel, lanes along the bins / channels
    const int x = x0 + px;
    const float* pd[4];
    const float* pc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cell = (n * cd.ncy + (y + 1 - (k >> 1))) * cd.ncx + (x + 1 - (k & 1));
      const bool any = off[cell + 1] != off[cell];
      pd[k] = any ? part_depth + (((size_t)b * cd.nc + cell) * 4 + k) * D : nullptr;
      pc[k] = any ? part_ctx + (((size_t)b * cd.nc + cell) * 4 + k) * kC : nullptr;
    }
    for (int z = lane; z < D; z += 32) {
      float v = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (pd[k]) v += __ldg(pd[k] + z);
      s_tile[z * LD + px] = v;
    }
    if (lane < kC) {
      float v = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (pc[k]) v += __ldg(pc[k] + lane);
      s_ctx[lane * LD + px] = v;
    }
  }
  __syncthreads();
  TD* gd = gdepth + (size_t)bn * D * HW + (size_t)y * g.fW + x0;
  for (int i = threadIdx.x; i < D * kReduceTile; i += blockDim.x) {       // lanes along the pixels
    const int z = i / kReduceTile, px = i % kReduceTile;
    if (px < npx) gd[(size_t)z * HW + px] = VbType<TD>::cvt(s_tile[z * LD + px]);
  }
  TC* gc = gctx + (size_t)bn * kC * HW + (size_t)y * g.fW + x0;
  for (int i = threadIdx.x; i < kC * kReduceTile; i += blockDim.x) {
    const int c = i / kReduceTile, px = i % kReduceTile;
    if (px < npx) gc[(size_t)c * HW + px] = VbType<TC>::cvt(s_ctx[c * LD + px]);
  }
}

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

struct BwdLayout {
  size_t ctx, counts, offsets, cursor, chunk_sums, recs_a, recs_b, part_ctx, gprep, part_depth, total;
};
// planned: the per-call plan buffers (counts / offsets / cursor / records) are not needed
BwdLayout bwd_layout(const VbGrid* g, bool need_gdepth_ws, bool planned) {
  const CellDims cd = cell_dims(*g);
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX, HW = (size_t)g->fH * g->fW;
  BwdLayout l;
  size_t o = 0;
  l.ctx = o;     o += align256((size_t)g->B * g->N * HW * kC * 4);
  l.counts = o;  o += planned ? 0 : align256((size_t)g->B * cd.nc * 4);
  l.offsets = o; o += planned ? 0 : align256((size_t)g->B * (cd.nc + 1) * 4);
  l.cursor = o;  o += planned ? 0 : align256((size_t)g->B * cd.nc * 4);
  l.chunk_sums = o; o += planned ? 0 : align256((size_t)g->B * 1024 * 4);
  l.recs_a = o;  o += planned ? 0 : align256((size_t)g->B * g->N * nvox * 4);
  l.recs_b = o;  o += planned ? 0 : align256((size_t)g->B * g->N * nvox * 4);
  // per-cell partials: 4 corner pixels x (C channels | D depth bins) fp32 (R50: 95 MB per sample)
  l.part_ctx = o;   o += align256((size_t)g->B * cd.nc * 4 * kC * 4);
  l.gprep = o;      o += align256((size_t)g->B * nvox * kC * 4);
  l.part_depth = o; o += align256((size_t)g->B * cd.nc * 4 * g->D * 4);
  (void)need_gdepth_ws;
  l.total = o;
  return l;
}

// TD: dtype of depth / d_gout / d_gdepth; TC: dtype of ctx / d_gctx (see vb_lift_common.cuh).
// d_plans != nullptr: cached plans (device array of B VbLiftPlan), no per-call plan.
template <typename TD, typename TC>
int launch_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbLiftPlan* d_plans, const void* d_depth,
               const void* d_ctx, const void* d_gout, int gout_layout, const uint64_t* d_cnt, void* d_gdepth,
               void* d_gctx, char* ws, cudaStream_t st) {
  if (g->C != kC) return VB200_ERR_ARG;
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX, HW = (size_t)g->fH * g->fW;
  if (nvox > (1u << kVoxBits) || g->D + 1 >= (1 << (32 - kVoxBits))) return VB200_ERR_ARG;
  constexpr bool kF32 = sizeof(TD) == 4;
  (void)HW;
  const bool planned = d_plans != nullptr;
  const BwdLayout l = bwd_layout(g, !kF32, planned);
  const CellDims cd = cell_dims(*g);
  float* ctx_nhwc = reinterpret_cast<float*>(ws + l.ctx);
  int* counts = reinterpret_cast<int*>(ws + l.counts);
  int* offsets = reinterpret_cast<int*>(ws + l.offsets);
  int* cursor = reinterpret_cast<int*>(ws + l.cursor);
  int* chunk_sums = reinterpret_cast<int*>(ws + l.chunk_sums);
  uint32_t* recs_a = reinterpret_cast<uint32_t*>(ws + l.recs_a);
  uint32_t* recs_b = reinterpret_cast<uint32_t*>(ws + l.recs_b);
  float* part_ctx = reinterpret_cast<float*>(ws + l.part_ctx);
  float* part_depth = reinterpret_cast<float*>(ws + l.part_depth);

  VbLiftDiv dv = {};
  VbTables no_tables = {};
  if (!planned) dv = vb_lift_div(g);
  {
    VbTraceScope tr(VB_K_CTX_NHWC, st);
    ctx_to_nhwc_f32_kernel<TC><<<dim3(g->fH, g->B * g->N), 256, (size_t)kC * (g->fW + 1) * 4, st>>>(
        reinterpret_cast<const TC*>(d_ctx), ctx_nhwc, g->fH, g->fW);
    VB_LAUNCH_CHECK();
  }
  if (!planned) {
    if (cudaMemsetAsync(counts, 0, (size_t)g->B * cd.nc * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
    if (cudaMemsetAsync(cursor, 0, (size_t)g->B * cd.nc * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
    VbTraceScope tr(VB_K_LIFT_PLAN, st, 6);
    dim3 vgrid(vb_ceil_div(nvox, kThreads), g->B);
    plan_pairs_kernel<0><<<vgrid, kThreads, 0, st>>>(*g, *t, dv, d_mats, counts, nullptr, nullptr, nullptr);
    VB_LAUNCH_CHECK();
    {
      const int rc = vb_exclusive_scan(counts, offsets, chunk_sums, cd.nc, g->B, st);   // > 4 M pixel cells per sample: not a camera feature map
      if (rc) return rc;
    }
    plan_pairs_kernel<1><<<vgrid, kThreads, 0, st>>>(*g, *t, dv, d_mats, nullptr, offsets, cursor, recs_a);
    VB_LAUNCH_CHECK();
    const int total_cells = g->B * cd.nc;
    sort_cells_kernel<<<vb_ceil_div(total_cells, kThreads / 32), kThreads, 0, st>>>(
        offsets, recs_a, recs_b, cd.nc, (size_t)g->N * nvox, total_cells);
    VB_LAUNCH_CHECK();
  }
  {
    VbTraceScope tr(VB_K_LIFT_BWD, st, 3);
    const size_t smem = (size_t)(kThreads / 32) * (4 * g->D + 32 * 4 + 32 * 17 + 64) * sizeof(float);
    float* gprep = reinterpret_cast<float*>(ws + l.gprep);
    {
      dim3 pgrid(vb_ceil_div(nvox, 256), g->B);
      if (gout_layout == VB200_NCDHW)
        gout_prep_kernel<TD, VB200_NCDHW><<<pgrid, 256, 0, st>>>(reinterpret_cast<const TD*>(d_gout), d_cnt, gprep, (int)nvox);
      else
        gout_prep_kernel<TD, VB200_NDHWC><<<pgrid, 256, 0, st>>>(reinterpret_cast<const TD*>(d_gout), d_cnt, gprep, (int)nvox);
      VB_LAUNCH_CHECK();
    }
    auto kern = planned ? lift_bwd_kernel<TD, true> : lift_bwd_kernel<TD, false>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return VB200_ERR_CUDA;
    dim3 grid(vb_ceil_div((long long)cd.ncy * cd.ncx, kThreads / 32), g->B * g->N);
    kern<<<grid, kThreads, smem, st>>>(*g, planned ? no_tables : *t, dv, d_mats, d_plans,
                                       reinterpret_cast<const TD*>(d_depth), ctx_nhwc, gprep, offsets, recs_b,
                                       part_depth, part_ctx);
    VB_LAUNCH_CHECK();
    const size_t rsmem = (size_t)(g->D + kC) * (kReduceTile + 1) * sizeof(float);
    if (rsmem > 48 * 1024) return VB200_ERR_ARG;       // D > ~350 depth bins: not a configuration of the reference
    lift_bwd_reduce_kernel<TD, TC><<<dim3(g->fH * vb_ceil_div(g->fW, kReduceTile), g->B * g->N), 256, rsmem, st>>>(
        *g, d_plans, offsets, part_depth, part_ctx, reinterpret_cast<TD*>(d_gdepth), reinterpret_cast<TC*>(d_gctx));
    VB_LAUNCH_CHECK();
  }
  return VB200_OK;
}

int bwd_entry(const VbGrid* g, const VbTables* t, const float* d_mats, const VbLiftPlan* d_plans, const void* d_depth,
              const void* d_ctx, int dtype, int ctx_dtype, const void* d_gout, int gout_layout, const uint64_t* d_cnt,
              void* d_gdepth, void* d_gctx, void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && d_depth && d_ctx && d_gout && d_cnt && d_gdepth && d_gctx && d_workspace);
  VB_CHECK_ARG(d_plans || (t && d_mats));
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS);
  VB_CHECK_ARG(g->D >= 1);
  VB_CHECK_ARG(gout_layout == VB200_NCDHW || gout_layout == VB200_NDHWC);
  if (workspace_bytes < bwd_layout(g, dtype != VB200_F32, d_plans != nullptr).total) return VB200_ERR_WORKSPACE;
  if (((uintptr_t)d_workspace | (uintptr_t)d_gout | (uintptr_t)d_gctx | (uintptr_t)d_gdepth) & 15) return VB200_ERR_ALIGN;
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(d_workspace);
#define VB_CALL(TD, TC) \
  launch_bwd<TD, TC>(g, t, d_mats, d_plans, d_depth, d_ctx, d_gout, gout_layout, d_cnt, d_gdepth, d_gctx, ws, st)
  VB_LIFT_DISPATCH(dtype, ctx_dtype, VB_CALL);
#undef VB_CALL
}

}  // namespace

extern "C" size_t vb200_lift_pool_bwd_workspace(const VbGrid* g, int dtype) {
  if (!g) return 0;
  return bwd_layout(g, dtype != VB200_F32, false).total;
}

extern "C" size_t vb200_lift_pool_bwd_planned_workspace(const VbGrid* g, int dtype) {
  if (!g) return 0;
  return bwd_layout(g, dtype != VB200_F32, true).total;
}

extern "C" int vb200_lift_pool_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_depth,
                                   const void* d_ctx, int dtype, int ctx_dtype, const void* d_gout, int gout_layout,
                                   const uint64_t* d_cnt, void* d_gdepth, void* d_gctx, void* d_workspace,
                                   size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(t && d_mats);
  return bwd_entry(g, t, d_mats, nullptr, d_depth, d_ctx, dtype, ctx_dtype, d_gout, gout_layout, d_cnt, d_gdepth,
                   d_gctx, d_workspace, workspace_bytes, stream);
}

extern "C" int vb200_lift_pool_bwd_planned(const VbGrid* g, const VbLiftPlan* d_plans, const void* d_depth,
                                           const void* d_ctx, int dtype, int ctx_dtype, const void* d_gout,
                                           int gout_layout, const uint64_t* d_cnt, void* d_gdepth, void* d_gctx,
                                           void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(d_plans);
  return bwd_entry(g, nullptr, nullptr, d_plans, d_depth, d_ctx, dtype, ctx_dtype, d_gout, gout_layout, d_cnt,
                   d_gdepth, d_gctx, d_workspace, workspace_bytes, stream);
}
