// vb_march_planned.cuh -- camera ray march (R2-R4) driven by a cached render plan.
//
// get_geometry (BV2:314-349), its nan_to_num (BV2:612), the normalisation / mask / unnormalise of BV2:397-419 and
// the step lengths of BV2:426 depend on the camera matrices only.  In validation / test those never change
// (nusc_det_seg_dataset.py:489-498, base_exp.py:113-120), so the strict fp32 chain that march_fwd_kernel runs per
// (ray, sample) -- ~170 of its ~440 instructions per executed warp-step -- is computed ONCE per distinct matrices
// by render_plan_build_kernel and read back from HBM (115 MB per sample at the R50 config, read only up to each
// warp's termination).  The records hold exactly the integers / fractions / step lengths the strict chain produced,
// so this kernel composites the same samples with the same weights as march_fwd_kernel<.., NANSAFE = false>.
//
//   steps[((n * npatch + patch) * S + i) * 32 + lane] = {valid << 31 | v0, fx, fy, fz}   (patch-major: a warp reads
//   delta[same index]                                  = |p_{i+1} - p_i|                   512 + 128 contiguous bytes
//   last[(n * npatch + patch) * 32 + lane]             = last valid sample of the ray (-1: none)        per step)
//   v0 / f* are the inward-shifted base corner and far weights of the fast variant (see march_fwd_kernel).
#pragma once
#include "vb_render_common.cuh"

namespace {

constexpr uint32_t kPlanValid = 0x80000000u;
constexpr uint32_t kPlanVoxMask = (1u << 21) - 1;
// Staging boxes (the TMA-class variant, vb_march_staged.cuh): per (warp, sample) the axis-aligned box of voxels that
// covers all 8 trilinear corners of the warp's in-volume rays.  box[(n * npatch + patch) * S + i] = {staged << 31 | first
// voxel of the box, nx | ny << 8 | nz << 16}; a record's key carries the base corner's index INSIDE the box in bits
// 21..27 (rel = ((z - z_lo) * ny + (y - y_lo)) * nx + (x - x_lo)).  A box is staged when it holds at most kBoxCap
// voxels in at most 8 y-rows x 4 z-rows (one bulk copy per lane, lane = dz * 8 + dy).
constexpr int kBoxCap = 96;
constexpr int kPlanRelShift = 21;
constexpr uint32_t kPlanRelMask = 0x7fu;
constexpr uint32_t kBoxStaged = 0x80000000u;

__host__ __device__ inline int march_patches(const VbGrid& g) {
  return ((g.fW + kPatchW - 1) / kPatchW) * ((g.fH + kPatchH - 1) / kPatchH);
}

// ---- plan build: one thread per ray, the strict chain of march_fwd_kernel<FROM_MATS = true> -----------------
template <bool FASTDIV>
__global__ void __launch_bounds__(kMarchThreads) render_plan_build_kernel(VbGrid g, VbTables t, VbRenderDiv dv,
                                                                          const float* __restrict__ d_mats,
                                                                          uint4* __restrict__ steps,
                                                                          float* __restrict__ delta,
                                                                          int16_t* __restrict__ last,
                                                                          uint2* __restrict__ box,
                                                                          size_t rays_per_sample) {
  __shared__ float s_m[VB200_MAT_SLOTS * 16];
  const int b = blockIdx.z, n = blockIdx.y;
  for (int i = threadIdx.x; i < VB200_MAT_SLOTS * 16; i += blockDim.x)
    s_m[i] = __ldg(d_mats + (size_t)(b * g.N + n) * VB200_MAT_SLOTS * 16 + i);
  __syncthreads();
  const bool has_bda = (g.has_bda != 0) && !block_is_identity(s_m + 5 * 16);
  const bool affine = block_ida_inv_affine(s_m + 3 * 16);
  const int patches_x = (g.fW + kPatchW - 1) / kPatchW;
  const int npatch = march_patches(g);
  const int patch = blockIdx.x * (kMarchThreads / 32) + (threadIdx.x >> 5);
  if (patch >= npatch) return;
  const int lane = threadIdx.x & 31;
  const int w = (patch % patches_x) * kPatchW + (lane % kPatchW);
  const int h = (patch / patches_x) * kPatchH + (lane / kPatchW);
  const bool active = (w < g.fW) && (h < g.fH);
  const int wc = min(w, g.fW - 1), hc = min(h, g.fH - 1);
  const int S = g.D - 1;
  const float u = __ldg(t.us + wc), vv = __ldg(t.vs + hc);
  float rayA[2] = {0.0f, 0.0f};
  if (affine) frustum_ray_affine(s_m, u, vv, rayA);
  auto point = [&](int d, float (&p)[3]) {
    if (affine) frustum_point_affine(s_m, has_bda, rayA, __ldg(t.ds + d), p);
    else frustum_point<false>(s_m, has_bda, u, vv, __ldg(t.ds + d), p);
    if (!(fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]) <= 3.402823466e+38f)) {      // BV2:612
#pragma unroll
      for (int a = 0; a < 3; ++a) p[a] = nan_to_num(p[a], -1e3f);
    }
  };
  const size_t ray = (size_t)(n * npatch + patch);
  uint4* so = steps + (size_t)b * rays_per_sample * S + ray * S * 32 + lane;
  float* dl = delta + (size_t)b * rays_per_sample * S + ray * S * 32 + lane;
  uint2* bx = box + ((size_t)b * (rays_per_sample / 32) + ray) * S;
  float p0[3], p1[3];
  point(0, p0);
  int last_valid = -1;
  for (int i = 0; i < S; ++i) {
    point(i + 1, p1);
    const RenderCoord rc = render_coord<FASTDIV>(g, p0, &dv);
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    const bool ok = rc.valid && active;
    const int x0 = min(rc.x0, g.vX - 2), y0 = min(rc.y0, g.vY - 2), z0 = min(rc.z0, g.vZ - 2);
    // the warp's box: base corners of its in-volume rays, plus one for the far corners
    const int big = 1 << 20;
    const int xl = __reduce_min_sync(0xffffffffu, ok ? x0 : big), xh = __reduce_max_sync(0xffffffffu, ok ? x0 : -big);
    const int yl = __reduce_min_sync(0xffffffffu, ok ? y0 : big), yh = __reduce_max_sync(0xffffffffu, ok ? y0 : -big);
    const int zl = __reduce_min_sync(0xffffffffu, ok ? z0 : big), zh = __reduce_max_sync(0xffffffffu, ok ? z0 : -big);
    const int nx = xh - xl + 2, ny = yh - yl + 2, nz = zh - zl + 2;
    const bool any = xh >= xl;
    const bool staged = any && nx * ny * nz <= kBoxCap && ny <= 8 && nz <= 4;   // lane = dz * 8 + dy copies one x-row
    if (ok) {
      const uint32_t rel = staged ? (uint32_t)(((z0 - zl) * ny + (y0 - yl)) * nx + (x0 - xl)) : 0u;
      r.x = kPlanValid | (rel << kPlanRelShift) | (uint32_t)((z0 * g.vY + y0) * g.vX + x0);
      r.y = __float_as_uint(rc.ix - (float)x0);
      r.z = __float_as_uint(rc.iy - (float)y0);
      r.w = __float_as_uint(rc.iz - (float)z0);
      last_valid = i;
    }
    so[(size_t)i * 32] = r;
    if (lane == 0)
      bx[i] = any ? make_uint2((staged ? kBoxStaged : 0u) | (uint32_t)((zl * g.vY + yl) * g.vX + xl),
                               (uint32_t)nx | ((uint32_t)ny << 8) | ((uint32_t)nz << 16))
                  : make_uint2(0u, 0u);
    const float dx = p1[0] - p0[0], dy = p1[1] - p0[1], dz = p1[2] - p0[2];
    dl[(size_t)i * 32] = sqrtf(dx * dx + dy * dy + dz * dz);                    // BV2:426
#pragma unroll
    for (int a = 0; a < 3; ++a) p0[a] = p1[a];
  }
  last[(size_t)b * rays_per_sample + ray * 32 + lane] = (int16_t)last_valid;
}

// ---- the march ---------------------------------------------------------------------------------------------------
#ifndef VB_MARCH_PLANNED_MINB
#define VB_MARCH_PLANNED_MINB 5
#endif
// SPLIT: launched as clusters of n CTAs along x, CTA r of a cluster marches the r-th n-th of the samples and
// march_cluster_fold() composes the segments (vb_render_common.cuh).
template <typename T, int K, bool SPLIT>
__global__ void __launch_bounds__(kMarchThreads, VB_MARCH_PLANNED_MINB) march_fwd_planned_kernel(
    VbGrid g, VbTables t, const VbRenderPlan* __restrict__ plans, const T* __restrict__ packed,
    const int* __restrict__ nonfinite_flag, const float* __restrict__ beta_ptr, float* __restrict__ o_rgb,
    float* __restrict__ o_seg, float* __restrict__ o_depth, int b0) {
  constexpr int CP = packed_channels(K);
  if (*nonfinite_flag != 0) return;   // the NaN-safe variant of march_fwd_kernel takes over (recomputes the geometry)
  const int b = b0 + blockIdx.z, n = blockIdx.y;
  const int patches_x = (g.fW + kPatchW - 1) / kPatchW;
  const int npatch = march_patches(g);
  const int S = g.D - 1, HW = g.fH * g.fW;
  const int nseg = SPLIT ? (int)cooperative_groups::this_cluster().num_blocks() : 1;
  const int seg = SPLIT ? (int)cooperative_groups::this_cluster().block_rank() : 0;
  const int xblock = SPLIT ? blockIdx.x / nseg : blockIdx.x;
  const int i_lo = SPLIT ? seg * S / nseg : 0, i_hi = SPLIT ? (seg + 1) * S / nseg : S;
  const int patch_raw = xblock * (kMarchThreads / 32) + (threadIdx.x >> 5);
  if (!SPLIT && patch_raw >= npatch) return;  // whole warp leaves together (a split warp stays for the cluster barriers)
  const int patch = SPLIT ? min(patch_raw, npatch - 1) : patch_raw;
  const int lane = threadIdx.x & 31;
  const int w = (patch % patches_x) * kPatchW + (lane % kPatchW);
  const int h = (patch / patches_x) * kPatchH + (lane / kPatchW);
  const bool active = (w < g.fW) && (h < g.fH) && (!SPLIT || patch_raw < npatch);
  const int nvox = g.vZ * g.vY * g.vX;
  const T* vol = packed + (size_t)blockIdx.z * nvox * CP;  // packed holds only this launch's samples
  const size_t ray = (size_t)(n * npatch + patch);
  const uint4* __restrict__ rec = reinterpret_cast<const uint4*>(plans[b].steps) + ray * S * 32 + lane;
  const float* __restrict__ dl = plans[b].delta + ray * S * 32 + lane;
  const int last = (int)__ldg(plans[b].last + ray * 32 + lane);

  const float beta = fabsf(__ldg(beta_ptr)) + g.beta_min;
  const float inv_beta = 1.0f / beta;
  const float sigma_masked = vb_density_rcp(g, 0.0f, inv_beta);   // feature 0 outside the volume
  // corner offsets are launch constants (inward-shifted base, see march_fwd_kernel): x-pairs = one pointer + immediate
  const int c_sy = g.vX * CP, c_sz = g.vY * g.vX * CP;

  float acc = 0.0f, dep = 0.0f, trans = 1.0f;
  float ch[K + 3];
#pragma unroll
  for (int c = 0; c < K + 3; ++c) ch[c] = 0.0f;

  // The plan is streamed from HBM (never re-used), so its latency must be hidden explicitly: an L2 prefetch runs
  // kPrefetchAhead samples ahead of the march (a warp's records are one contiguous 640-byte run per sample), the
  // record of sample i+2 is requested into registers, and the record of sample i+1 has arrived so that its density
  // gathers are issued before sample i is composited.
#ifndef VB_MARCH_PREFETCH
#define VB_MARCH_PREFETCH 4    // measured: 0.444 ms at 4 samples ahead, 0.455 at 12, 0.463 at 32
#endif
  constexpr int kPrefetchAhead = VB_MARCH_PREFETCH;
  auto prefetch = [&](int i) {
    if (i < i_hi) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (size_t)i * 32));
      if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(dl + (size_t)i * 32 + lane * 8 - lane));
    }
  };
#pragma unroll 1
  for (int i = 0; i < kPrefetchAhead; ++i) prefetch(i_lo + i);
  uint4 r_n = __ldg(rec + (size_t)i_lo * 32);      // every segment holds at least one sample (launcher: >= 4)
  float d_n = __ldg(dl + (size_t)i_lo * 32);
  uint4 r_n2 = make_uint4(0u, 0u, 0u, 0u);
  float d_n2 = 0.0f;
  if (i_lo + 1 < i_hi) {
    r_n2 = __ldg(rec + (size_t)(i_lo + 1) * 32);
    d_n2 = __ldg(dl + (size_t)(i_lo + 1) * 32);
  }
  T raw_n[8];
  auto gather_density = [&](const uint4& r, T (&raw)[8]) {
    if (r.x & kPlanValid) {
      const T* p = vol + (size_t)(r.x & kPlanVoxMask) * CP;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        raw[q] = __ldg(p + ((q & 2) ? c_sy : 0) + ((q & 4) ? c_sz : 0) + ((q & 1) ? CP : 0));
    }
  };
  gather_density(r_n, raw_n);

  for (int i = i_lo; i < i_hi; ++i) {
    if (g.term_eps > 0.0f) {
      const bool done = !active || trans < g.term_eps;
      if (__all_sync(0xffffffffu, done)) break;
      if (__all_sync(0xffffffffu, done || i > last)) {
        // every remaining sample of every live ray is outside the volume: feature 0, sigma(0), only delta_i and
        // mid_i enter the compositing -- a geometry-free tail that still reads the exact per-sample step lengths
        if (!done) {
          for (int ii = i; ii < i_hi; ++ii) {
            const float sd = sigma_masked * __ldg(dl + (size_t)ii * 32);
            const float e = expf(-sd);
            const float wgt = (1.0f - e) * trans;
            acc += wgt;
            dep = fmaf(wgt, __ldg(t.mids + ii), dep);
            trans *= e;
          }
        }
        break;
      }
    }
    const uint4 r = r_n;
    const float delta = d_n;
    T raw[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) raw[q] = raw_n[q];
    prefetch(i + kPrefetchAhead);
    r_n = r_n2;
    d_n = d_n2;
    if (i + 2 < i_hi) {
      r_n2 = __ldg(rec + (size_t)(i + 2) * 32);
      d_n2 = __ldg(dl + (size_t)(i + 2) * 32);
    }
    if (i + 1 < i_hi) gather_density(r_n, raw_n);
    const bool live = (r.x & kPlanValid) != 0u;     // build wrote valid = 0 for rays outside the image
    float sigma = sigma_masked;
    float cw[8];
    if (live) {
      const float fx = __uint_as_float(r.y), fy = __uint_as_float(r.z), fz = __uint_as_float(r.w);
      const float wx[2] = {1.0f - fx, fx}, wy[2] = {1.0f - fy, fy}, wz[2] = {1.0f - fz, fz};
      float s0 = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        cw[q] = wx[q & 1] * wy[(q >> 1) & 1] * wz[q >> 2];
        s0 = fmaf(cw[q], widen_elem(raw[q]), s0);
      }
      sigma = vb_density_rcp(g, s0, inv_beta);                   // BV2:423
    }
    const float sd = sigma * delta;                                           // BV2:429
    const float e = expf(-sd);
    const float wgt = (1.0f - e) * trans;                                     // BV2:430-434
    acc += wgt;
    dep = fmaf(wgt, __ldg(t.mids + i), dep);
    // the 21 value channels, only where they can contribute (alpha is exactly 0.0f in free space)
    if (live && wgt != 0.0f) {
      const T* p = vol + (size_t)(r.x & kPlanVoxMask) * CP;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        PackedLoad<T, CP>::template fma_values<K + 3>(p + ((q & 2) ? c_sy : 0) + ((q & 4) ? c_sz : 0) + ((q & 1) ? CP : 0),
                                                      cw[q] * wgt, ch);
    }
    // exp(-cumsum) of BV2:431-433 as a running product: one exp per sample instead of two
    trans *= e;
  }
  if (SPLIT) {
    if (!march_cluster_fold<K + 3>(acc, dep, trans, ch)) return;
  }
  if (!active) return;
  const size_t pix = (size_t)h * g.fW + w;
  const size_t bn = (size_t)b * g.N + n;
  o_depth[bn * HW + pix] = dep + (1.0f - acc) * g.bg_depth;                   // BV2:436, 440
#pragma unroll
  for (int k = 0; k < K; ++k) o_seg[(bn * K + k) * HW + pix] = ch[k];
#pragma unroll
  for (int j = 0; j < 3; ++j) o_rgb[(bn * 3 + j) * HW + pix] = ch[K + j];
}

}  // namespace
