// vb_render_bwd.cu -- Bk for R1-R6/T4: backward of the volume rendering.
//
// Reference: autograd of volume_rendering_from_multiple_views (BV2:391-467).  Closed forms
// (verified against autograd in fp64, B.7), per ray / BEV column, with
//   G'_i = sum_c g_c v_{i,c} + g_depth (mid_i - bg)          (bg = 0 and g_height for BEV)
//   w_i  = (1 - e^{-sd_i}) T_i,  T_{i+1} = T_i e^{-sd_i},  sd_i = sigma_i delta_i
//   dL/dv_{i,c} = g_c w_i
//   dL/dsd_i    = G'_i T_{i+1} - sum_{j>i} w_j G'_j
// The suffix sum is formed as (total - prefix) where total = sum_c g_c out_c + g_depth (depth - bg)
// comes from the forward's saved outputs, so the backward is a single front-to-back re-march.
//   dsigma/ds   = -e^{-|x|/beta} / (2 beta^2)  (x = s - bias, 0 at x = 0)
//   dsigma/dbeta = -sigma/beta + x e^{-|x|/beta} / (2 beta^3),   dbeta/dparam = sign(param)
//
//   march_bwd          camera branch: re-march + trilinear scatter of the 22 channel gradients into a
//                      channels-last fp32 volume with 128-bit vector atomics (red.global.add.v4.f32;
//                      order-dependent in the last bits, tolerance-checked -- SURVEY §7.1 d)
//   bev_bwd_partials/  BEV branch, per column: channel-group partial sums of g_c * sample, then the compositing
//   bev_bwd_composite  backward -> dL/d(sampled density feature), compositing weights, d beta partials
//   unpack_gather      per INPUT voxel: reads the camera-branch gradient (channels-last), GATHERS the BEV
//                      branch gradient from the <= 2x2x2 output samples whose stencil covers the voxel
//                      (deterministic, inverse index tables), writes the four NCDHW gradients
//   beta_reduce        fixed-order sum of the per-block d beta partials
#include "vb_render_common.cuh"
#include "vb_trace.cuh"

namespace {

// 64-thread blocks: at B=1 the 65,536 BEV columns are only 256 blocks of 256 threads (0.3 waves)
constexpr int kBevBwdThreads = 64;

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float tot = 0.0f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
  return tot;   // valid in thread 0
}

// slow path of march_bwd (a non-finite value somewhere in the 8 corners): the reference's per-channel
// nan_to_num(interpolated value) (BV2:421), kept out of line so that its 24-float temporary costs the hot path no registers
template <typename T, int K>
__device__ __noinline__ void nan_safe_values(const T* __restrict__ packed, const int (&cidx)[8], const float (&cw)[8],
                                             const float (&gc)[K + 3], float& s0, float& Gv) {
  constexpr int CP = packed_channels(K);
  float v[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) v[c] = 0.0f;
  for (int q = 0; q < 8; ++q) PackedLoad<T, CP>::fma_corner(packed + cidx[q], cw[q], v);
  s0 = nan_to_num(v[0], 0.0f);
  Gv = 0.0f;
#pragma unroll
  for (int c = 0; c < K + 3; ++c) Gv = fmaf(gc[c], nan_to_num(v[1 + c], 0.0f), Gv);
}

// ---- camera branch ---------------------------------------------------------------------------------------
#ifndef VB_MARCH_BWD_THREADS
#define VB_MARCH_BWD_THREADS 128
#endif
constexpr int kMarchBwdThreads = VB_MARCH_BWD_THREADS;
#ifndef VB_MARCH_BWD_MINB
#define VB_MARCH_BWD_MINB 3   // measured B=1 fp32: 0.593 ms at 3 blocks/SM, 0.68 at 4, 0.76 at 5 (B=8: 4.35 ms either way: L2-bound)
                              // block shape (B=1 is 1.19 waves of 128-thread blocks): 96 x 5 (one wave, 128 regs) 0.652,
                              // 64 x 7 0.693, 64 x 6 (168 regs) 0.594 vs 0.605 -- nothing to gain, 128 x 3 stays.
                              // L1 prefetch of the next sample's 8 corner records (prefetch.global.L1, coordinates
                              // computed one iteration early): B=1 0.617 -> 0.649, B=8 4.49 -> 4.64 -- slower, removed
#endif
template <typename T, int K, bool FROM_MATS>
__global__ void __launch_bounds__(kMarchBwdThreads, VB_MARCH_BWD_MINB) march_bwd_kernel(
    VbGrid g, VbTables t, VbRenderDiv dv, const float* __restrict__ d_mats, const float* __restrict__ d_geom,
    const T* __restrict__ packed, const float* __restrict__ beta_ptr, const float* __restrict__ o_rgb,
    const float* __restrict__ o_seg, const float* __restrict__ o_depth, const float* __restrict__ g_rgb,
    const float* __restrict__ g_seg, const float* __restrict__ g_depth, float* __restrict__ gpacked,
    float* __restrict__ beta_partials, size_t packed_stride, size_t gpacked_stride) {
  constexpr int CP = packed_channels(K);
  __shared__ float s_m[VB200_MAT_SLOTS * 16];
  __shared__ float s_red[kMarchBwdThreads / 32];
  // grid = (patch blocks, cameras, samples): the whole batch in ONE launch (a sample alone is 1.2 waves at 3 blocks
  // per SM, i.e. 40 % of its time is a tail) -- each sample has its own packed copy and gradient accumulator
  const int n = blockIdx.y, b = blockIdx.z;
  packed += (size_t)b * packed_stride;
  gpacked += (size_t)b * gpacked_stride;
  for (int i = threadIdx.x; i < VB200_MAT_SLOTS * 16; i += blockDim.x)
    s_m[i] = __ldg(d_mats + (size_t)(b * g.N + n) * VB200_MAT_SLOTS * 16 + i);
  __syncthreads();
  const bool has_bda = (g.has_bda != 0) && !block_is_identity(s_m + 5 * 16);   // slot 5 = bda
  // the forward march's exact shortcuts (vb_render.cu): affine ida^-1, launch-constant divisors
  const bool affine = FROM_MATS && block_ida_inv_affine(s_m + 3 * 16);
  const bool fastdiv = FROM_MATS && vb_render_div_ok(dv);

  const int patches_x = (g.fW + kPatchW - 1) / kPatchW;
  const int patches_y = (g.fH + kPatchH - 1) / kPatchH;
  const int patch = blockIdx.x * (kMarchBwdThreads / 32) + (threadIdx.x >> 5);
  const bool warp_live = patch < patches_x * patches_y;
  const int lane = threadIdx.x & 31;
  const int w = (patch % patches_x) * kPatchW + (lane % kPatchW);
  const int h = (patch / patches_x) * kPatchH + (lane / kPatchW);
  const bool active = warp_live && (w < g.fW) && (h < g.fH);
  const int wc = min(w, g.fW - 1), hc = min(max(h, 0), g.fH - 1);

  const int S = g.D - 1, HW = g.fH * g.fW;
  const float beta = fabsf(__ldg(beta_ptr)) + g.beta_min;
  const float u = __ldg(t.us + wc), vv = __ldg(t.vs + hc);
  const float* gsrc = FROM_MATS ? nullptr : d_geom + ((size_t)(b * g.N + n) * g.D * HW + (size_t)hc * g.fW + wc) * 3;
  float rayA[2] = {0.0f, 0.0f};
  if (affine) frustum_ray_affine(s_m, u, vv, rayA);
  auto point = [&](int d, float (&p)[3]) {
    if (FROM_MATS) {
      if (affine) frustum_point_affine(s_m, has_bda, rayA, __ldg(t.ds + d), p);
      else frustum_point<false>(s_m, has_bda, u, vv, __ldg(t.ds + d), p);
      if (!(fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]) <= 3.402823466e+38f)) {   // BV2:612, see the forward march
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = nan_to_num(p[a], -1e3f);
      }
    } else {
      const float* q = gsrc + (size_t)d * HW * 3;
      p[0] = __ldg(q); p[1] = __ldg(q + 1); p[2] = __ldg(q + 2);
    }
  };

  // cotangents and saved forward outputs of this ray
  const size_t pix = (size_t)hc * g.fW + wc;
  const size_t bn = (size_t)b * g.N + n;
  float gc[K + 3];
  float gd = 0.0f, omega = 0.0f;
  if (active) {
    gd = g_depth ? __ldg(g_depth + bn * HW + pix) : 0.0f;
    omega = gd * (__ldg(o_depth + bn * HW + pix) - g.bg_depth);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    gc[k] = (active && g_seg) ? __ldg(g_seg + (bn * K + k) * HW + pix) : 0.0f;
    if (active) omega = fmaf(gc[k], __ldg(o_seg + (bn * K + k) * HW + pix), omega);
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    gc[K + j] = (active && g_rgb) ? __ldg(g_rgb + (bn * 3 + j) * HW + pix) : 0.0f;
    if (active) omega = fmaf(gc[K + j], __ldg(o_rgb + (bn * 3 + j) * HW + pix), omega);
  }

  // which 4-channel vectors of the record carry a non-zero cotangent (bit f: record channels 4f .. 4f+3)
  unsigned gc_nz = 0u;
#pragma unroll
  for (int c = 0; c < K + 3; ++c)
    if (gc[c] != 0.0f) gc_nz |= 1u << ((c + 1) >> 2);

  float tau = 0.0f, prefix = 0.0f, dbeta = 0.0f;
  float p0[3], p1[3];
  if (warp_live) {
    point(0, p0);
    for (int i = 0; i < S; ++i) {
      const float trans = expf(-tau);
      if (g.term_eps > 0.0f && __all_sync(0xffffffffu, !active || trans < g.term_eps)) break;
      point(i + 1, p1);
      const float dx = p1[0] - p0[0], dy = p1[1] - p0[1], dz = p1[2] - p0[2];
      const float delta = sqrtf(dx * dx + dy * dy + dz * dz);
      const RenderCoord rc = fastdiv ? render_coord<true>(g, p0, &dv) : render_coord<false>(g, p0);
      const bool live = rc.valid && active;
      // The interpolated channel values are needed only through two scalars: s0 (the density feature) and
      // Gv = sum_c g_c v_c.  Both are accumulated corner by corner -- Gv as sum_q cw_q (sum_c g_c t_qc) -- so no
      // 24-float v[] (nor a 24-float dv[] below) stays live across the gather and the scatter: 150 -> under 128
      // registers, one more resident block per SM for a kernel that waits on memory 80 % of the time.
      float s0 = 0.0f, Gv = 0.0f;
      int cidx[8];
      float cw[8];
      if (live) {
        const float wx[2] = {(float)(rc.x0 + 1) - rc.ix, rc.x0 + 1 < g.vX ? rc.ix - (float)rc.x0 : 0.0f};
        const float wy[2] = {(float)(rc.y0 + 1) - rc.iy, rc.y0 + 1 < g.vY ? rc.iy - (float)rc.y0 : 0.0f};
        const float wz[2] = {(float)(rc.z0 + 1) - rc.iz, rc.z0 + 1 < g.vZ ? rc.iz - (float)rc.z0 : 0.0f};
        const int xs_[2] = {rc.x0, min(rc.x0 + 1, g.vX - 1)};
        const int ys_[2] = {rc.y0, min(rc.y0 + 1, g.vY - 1)};
        const int zs_[2] = {rc.z0, min(rc.z0 + 1, g.vZ - 1)};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int cx = q & 1, cy = (q >> 1) & 1, cz = q >> 2;
          cw[q] = wx[cx] * wy[cy] * wz[cz];
          cidx[q] = ((zs_[cz] * g.vY + ys_[cy]) * g.vX + xs_[cx]) * CP;
          float den_q, dot_q;
          PackedLoad<T, CP>::template dot_values<K + 3>(packed + cidx[q], gc, den_q, dot_q);
          s0 = fmaf(cw[q], den_q, s0);
          Gv = fmaf(cw[q], dot_q, Gv);
        }
        // torch.nan_to_num of the interpolated features (BV2:421) is the identity unless something is non-finite, and
        // then s0 or Gv is non-finite too: the exact per-channel form runs only in that case
        if (!(fabsf(s0) + fabsf(Gv) <= 3.402823466e+38f)) nan_safe_values<T, K>(packed, cidx, cw, gc, s0, Gv);
      }
      const DensityD dd = vb_density_with_grads(g, s0, beta);
      const float sd = dd.sigma * delta;
      const float e_sd = expf(-sd);
      const float wgt = (1.0f - e_sd) * trans;
      const float t_next = trans * e_sd;
      const float Gp = fmaf(gd, __ldg(t.mids + i) - g.bg_depth, Gv);
      prefix = fmaf(wgt, Gp, prefix);
      const float dsd = Gp * t_next - (omega - prefix);
      const float dsig = dsd * delta;
      if (active) dbeta = fmaf(dsig, dd.dbeta, dbeta);
      if (live) {
        // d v_c = g_c w (c >= 1), d v_0 = dsig * dsigma/ds; scattered as cw_q * (.) with 128-bit vector atomics.
        // Exact zeros need no atomic at all: in free space w = 0 and only the density vector is live; a zero rgb
        // loss weight (the target experiment, base_exp.py loss_weights) zeroes a whole vector for the entire ray
        // (gc_nz, decided once per ray).
        const float d0 = dsig * dd.ds;
        const bool w_nz = wgt != 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (cw[q] != 0.0f) {
            float4* dst = reinterpret_cast<float4*>(gpacked + cidx[q]);
            const float sw = cw[q] * wgt;
            if (d0 != 0.0f || (w_nz && (gc_nz & 1u)))
              atomicAdd(dst, make_float4(cw[q] * d0, sw * gc[0], sw * gc[1], sw * gc[2]));
#pragma unroll
            for (int f = 1; f < CP / 4; ++f) {
              if (w_nz && ((gc_nz >> f) & 1u)) {
                // channels 4f .. 4f+3 of the record = gc[4f-1 .. 4f+2]; beyond K+3 the record is padding
                const float g0 = gc[4 * f - 1];
                const float g1 = (4 * f < K + 3) ? gc[4 * f < K + 3 ? 4 * f : 0] : 0.0f;
                const float g2 = (4 * f + 1 < K + 3) ? gc[4 * f + 1 < K + 3 ? 4 * f + 1 : 0] : 0.0f;
                const float g3 = (4 * f + 2 < K + 3) ? gc[4 * f + 2 < K + 3 ? 4 * f + 2 : 0] : 0.0f;
                atomicAdd(dst + f, make_float4(sw * g0, sw * g1, sw * g2, sw * g3));
              }
            }
          }
        }
      }
      tau += sd;
      p0[0] = p1[0]; p0[1] = p1[1]; p0[2] = p1[2];
    }
  }
  const float tot = block_sum(dbeta, s_red);
  if (threadIdx.x == 0) beta_partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// ---- BEV branch: per-column compositing backward -------------------------------------------------------------
// Two stages so that the 22 channel planes are walked by 6 thread groups in parallel (a single thread
// per column walking all of them left the GPU at 0.3 waves / 14 % issue):
//   bev_bwd_partials   grid.z = channel group: G_part[grp][l] = sum_{c in grp} g_c v_{l,c}; group 0 also
//                      samples the density plane (-> S, parked in ds_ws) and adds g_height * mid_l
//   bev_bwd_composite  G = sum of the partials in fixed group order, then the compositing recurrences
constexpr int kBevBwdGroup = 4;   // composited channels per group

template <typename T, int K>
__global__ void __launch_bounds__(kBevBwdThreads) bev_bwd_partials_kernel(
    VbGrid g, VbTables t, const T* __restrict__ den, const T* __restrict__ sem, const T* __restrict__ rgb,
    const float* __restrict__ g_bev_rgb, const float* __restrict__ g_bev_seg, const float* __restrict__ g_bev_height,
    float* __restrict__ gpart_ws, float* __restrict__ s_ws) {
  __shared__ BevLevel s_lv[kMaxLevels];
  extern __shared__ float s_col[];            // G[kMaxLevels][threads]
  constexpr int TB = kBevBwdThreads;
  float* sG = s_col;
  bev_level_table(g, t, s_lv);
  const int b = blockIdx.y, grp = blockIdx.z;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = g.oY * g.oX;
  if (col >= ncol) return;
  const BevColumn bc = bev_column(g, t, col % g.oX, col / g.oX);
  const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
  const int tid = threadIdx.x;
  auto walk = [&](const T* plane, auto&& sink) {
    float prev_lo = 0.0f;
    int prev_z0 = -1000000;
    for (int l = 0; l < g.oZ; ++l) {
      const BevLevel L = s_lv[l];
      const float hi = (L.z0 + 1 == prev_z0) ? prev_lo : bev_row<T>(g, bc, plane, L.z0 + 1);
      const float lo = bev_row<T>(g, bc, plane, L.z0);
      prev_z0 = L.z0;
      prev_lo = lo;
      sink(l, L.wz0 * lo + L.wz1 * hi);
    }
  };
  if (grp == 0) {
    const float gh = g_bev_height ? __ldg(g_bev_height + (size_t)b * ncol + col) : 0.0f;
    walk(den + (size_t)b * nvox, [&](int l, float v) {
      s_ws[((size_t)b * g.oZ + l) * ncol + col] = v;
      sG[l * TB + tid] = gh * __ldg(t.bev_mids + l);
    });
  } else {
    for (int l = 0; l < g.oZ; ++l) sG[l * TB + tid] = 0.0f;
  }
  const int j0 = grp * kBevBwdGroup, j1 = min(K + 3, j0 + kBevBwdGroup);
  for (int j = j0; j < j1; ++j) {
    const float* gsrc = j < K ? g_bev_seg : g_bev_rgb;
    if (!gsrc) continue;
    const float gcv = j < K ? __ldg(gsrc + ((size_t)b * K + j) * ncol + col)
                            : __ldg(gsrc + ((size_t)b * 3 + (j - K)) * ncol + col);
    const T* plane = j < K ? sem + ((size_t)b * K + j) * nvox : rgb + ((size_t)b * 3 + (j - K)) * nvox;
    walk(plane, [&](int l, float v) { sG[l * TB + tid] = fmaf(gcv, v, sG[l * TB + tid]); });
  }
  float* out = gpart_ws + (((size_t)grp * g.B + b) * g.oZ) * ncol + col;
  for (int l = 0; l < g.oZ; ++l) out[(size_t)l * ncol] = sG[l * TB + tid];
}

// Two sweeps over the column's levels, nothing kept in per-level register arrays (the first version held G[16] and S[16]:
// 113 registers, 24 % occupancy for a kernel that only streams): sweep 1 sums the channel-group partials of each level
// in fixed group order, parks the sum in group 0's slot and accumulates omega; sweep 2 re-reads S_l and G_l.
template <int K>
__global__ void __launch_bounds__(256) bev_bwd_composite_kernel(
    VbGrid g, const float* __restrict__ beta_ptr, const float* __restrict__ g_vd, float* __restrict__ gpart_ws,
    int ngroups, float* __restrict__ wl_ws, float* __restrict__ ds_ws, float* __restrict__ beta_partials) {
  __shared__ float s_red[8];
  const int b = blockIdx.y;
  const int col_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = g.oY * g.oX;
  const bool live = col_raw < ncol;
  const int col = live ? col_raw : ncol - 1;
  const float beta = fabsf(__ldg(beta_ptr)) + g.beta_min;
  const size_t gstride = (size_t)g.B * g.oZ * ncol;
  // total = sum_l w_l G_l, top-down
  float omega = 0.0f, tau = 0.0f;
  for (int l = 0; l < g.oZ; ++l) {
    const size_t o = ((size_t)b * g.oZ + l) * ncol + col;
    float G = 0.0f;
    for (int grp = 0; grp < ngroups; ++grp) G += gpart_ws[(size_t)grp * gstride + o];
    const float sigma = vb_density(g, ds_ws[o], beta);                                   // S_l parked by stage 1
    const float sd = sigma * g.bev_delta;
    const float w = (1.0f - expf(-sd)) * expf(-tau);
    tau += sd;
    omega = fmaf(w, G, omega);
    if (live) {
      wl_ws[o] = w;
      gpart_ws[o] = G;           // group 0's slot now holds the level's total (this thread is its only reader)
    }
  }
  float prefix = 0.0f, dbeta = 0.0f;
  tau = 0.0f;
  for (int l = 0; l < g.oZ; ++l) {
    const size_t o = ((size_t)b * g.oZ + l) * ncol + col;
    const float G = gpart_ws[o];     // (a dead thread reads another column's slot: its result is discarded)
    const DensityD dd = vb_density_with_grads(g, ds_ws[o], beta);
    const float sd = dd.sigma * g.bev_delta;
    const float trans = expf(-tau), e_sd = expf(-sd);
    const float w = (1.0f - e_sd) * trans;
    prefix = fmaf(w, G, prefix);
    const float dsd = G * (trans * e_sd) - (omega - prefix);
    const float dsig = dsd * g.bev_delta + (g_vd ? __ldg(g_vd + o) : 0.0f);
    if (live) {
      ds_ws[o] = dsig * dd.ds;
      dbeta = fmaf(dsig, dd.dbeta, dbeta);
    }
    tau += sd;
  }
  const float tot = block_sum(dbeta, s_red);
  if (threadIdx.x == 0) beta_partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
}

// ---- inverse index tables for the BEV gather ------------------------------------------------------------------
struct BevTables {      // all device pointers into the workspace
  int *xi0, *yi0, *li0;          // [oX], [oY], [oZ] base input index of each output (levels: top first)
  float *xw0, *xw1, *yw0, *yw1, *lw0, *lw1;
  int *xlo, *xhi, *ylo, *yhi, *llo, *lhi;   // [vX], [vY], [vZ] inclusive output ranges touching an input index
};

__global__ void __launch_bounds__(256) bev_tables_kernel(VbGrid g, VbTables t, BevTables bt) {
  for (int o = threadIdx.x; o < g.oX; o += blockDim.x)
    axis_coord(__ldg(t.oxs + o), g.seg_lo[0], g.seg_ext[0], g.vX, bt.xi0[o], bt.xw0[o], bt.xw1[o]);
  for (int o = threadIdx.x; o < g.oY; o += blockDim.x)
    axis_coord(__ldg(t.oys + o), g.seg_lo[1], g.seg_ext[1], g.vY, bt.yi0[o], bt.yw0[o], bt.yw1[o]);
  for (int l = threadIdx.x; l < g.oZ; l += blockDim.x)
    axis_coord(__ldg(t.ozs + (g.oZ - 1 - l)), g.seg_lo[2], g.seg_ext[2], g.vZ, bt.li0[l], bt.lw0[l], bt.lw1[l]);
  __syncthreads();
  auto invert = [&](const int* i0, int nout, int nin, int* lo, int* hi) {
    for (int i = threadIdx.x; i < nin; i += blockDim.x) {
      int a = nout, bmax = -1;
      for (int o = 0; o < nout; ++o) {
        const int v = i0[o];
        if (v == i || v + 1 == i) { a = min(a, o); bmax = max(bmax, o); }
      }
      lo[i] = a;
      hi[i] = bmax;
    }
  };
  invert(bt.xi0, g.oX, g.vX, bt.xlo, bt.xhi);
  invert(bt.yi0, g.oY, g.vY, bt.ylo, bt.yhi);
  invert(bt.li0, g.oZ, g.vZ, bt.llo, bt.lhi);
}

// ---- per input voxel: camera-branch gradient (channels-last) + gathered BEV gradient -> NCDHW outputs ---------
template <typename T, int K, int C>
#ifndef VB_UNPACK_MINB
#define VB_UNPACK_MINB 3     // 80 registers; measured B=8 / B=1 fp32 (scope incl. the BEV partial kernels): 1.634 / 0.262 ms;
#endif                       // 4 blocks (64 regs, 92 B spilled): 1.671 / 0.269; 2 blocks (128 regs): 2.271 / 0.340
__global__ void __launch_bounds__(256, VB_UNPACK_MINB) unpack_gather_kernel(
    VbGrid g, BevTables bt, const float* __restrict__ gpacked, const float* __restrict__ wl_ws,
    const float* __restrict__ ds_ws, const float* __restrict__ g_bev_rgb, const float* __restrict__ g_bev_seg,
    const T* __restrict__ g_vo, T* __restrict__ o_den, T* __restrict__ o_sem, T* __restrict__ o_rgb,
    T* __restrict__ o_feat, size_t gpacked_stride, int do_bev) {
  constexpr int CP = packed_channels(K);
  const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (vox >= (int)nvox) return;
  if (gpacked) gpacked += (size_t)b * gpacked_stride;
  const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
  float cam[CP];
  if (gpacked) {
    const float4* src = reinterpret_cast<const float4*>(gpacked + (size_t)vox * CP);
#pragma unroll
    for (int f = 0; f < CP / 4; ++f) {
      const float4 v = src[f];
      cam[4 * f] = v.x; cam[4 * f + 1] = v.y; cam[4 * f + 2] = v.z; cam[4 * f + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CP; ++c) cam[c] = 0.0f;
  }
  float accF[C];
#pragma unroll
  for (int c = 0; c < C; ++c) accF[c] = 0.0f;
  if (do_bev) {
    const int l0 = __ldg(bt.llo + z), l1 = __ldg(bt.lhi + z);
    const int y0 = __ldg(bt.ylo + y), y1 = __ldg(bt.yhi + y);
    const int x0 = __ldg(bt.xlo + x), x1 = __ldg(bt.xhi + x);
    const int ncol = g.oY * g.oX;
    if (l0 <= l1) {
      // the z weights depend on the level only: looked up once per voxel, not once per (oy, ox, level)
      constexpr int kMaxL = 4;
      float wzl[kMaxL];
#pragma unroll
      for (int j = 0; j < kMaxL; ++j) {
        const int l = min(l0 + j, l1);
        wzl[j] = (__ldg(bt.li0 + l) == z) ? __ldg(bt.lw0 + l) : __ldg(bt.lw1 + l);
      }
      for (int oy = y0; oy <= y1; ++oy) {
        const float wy = (__ldg(bt.yi0 + oy) == y) ? __ldg(bt.yw0 + oy) : __ldg(bt.yw1 + oy);
        for (int ox = x0; ox <= x1; ++ox) {
          const float wxy = wy * ((__ldg(bt.xi0 + ox) == x) ? __ldg(bt.xw0 + ox) : __ldg(bt.xw1 + ox));
          const int col = oy * g.oX + ox;
          float A = 0.0f;
          for (int l = l0; l <= l1; ++l) {
            const int j = l - l0;
            const float wz = j < kMaxL ? wzl[j < kMaxL ? j : 0]
                                       : ((__ldg(bt.li0 + l) == z) ? __ldg(bt.lw0 + l) : __ldg(bt.lw1 + l));
            const size_t o = ((size_t)b * g.oZ + l) * ncol + col;
            cam[0] = fmaf(wxy * wz, __ldg(ds_ws + o), cam[0]);
            A = fmaf(wz, __ldg(wl_ws + o), A);
            if (g_vo) {
#pragma unroll
              for (int c = 0; c < C; ++c)
                accF[c] = fmaf(wxy * wz, VbType<T>::ld(g_vo + (((size_t)b * C + c) * g.oZ + l) * ncol + col), accF[c]);
            }
          }
          A *= wxy;
          if (g_bev_seg) {
#pragma unroll
            for (int k = 0; k < K; ++k) cam[1 + k] = fmaf(A, __ldg(g_bev_seg + ((size_t)b * K + k) * ncol + col), cam[1 + k]);
          }
          if (g_bev_rgb) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
              cam[1 + K + j] = fmaf(A, __ldg(g_bev_rgb + ((size_t)b * 3 + j) * ncol + col), cam[1 + K + j]);
          }
        }
      }
    }
  }
  o_den[(size_t)b * nvox + vox] = VbType<T>::cvt(cam[0]);
#pragma unroll
  for (int k = 0; k < K; ++k) o_sem[((size_t)b * K + k) * nvox + vox] = VbType<T>::cvt(cam[1 + k]);
#pragma unroll
  for (int j = 0; j < 3; ++j) o_rgb[((size_t)b * 3 + j) * nvox + vox] = VbType<T>::cvt(cam[1 + K + j]);
#pragma unroll
  for (int c = 0; c < C; ++c) o_feat[((size_t)b * C + c) * nvox + vox] = VbType<T>::cvt(accF[c]);
}

__global__ void beta_reduce_kernel(const float* __restrict__ partials, int n, const float* __restrict__ beta_ptr,
                                   float* __restrict__ g_beta) {
  // one warp, fixed order: lane-strided partial sums, then a shuffle tree
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += 32) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    const float p = *beta_ptr;
    *g_beta = s * ((p > 0.0f) ? 1.0f : ((p < 0.0f) ? -1.0f : 0.0f));   // d|p|/dp
  }
}

struct BwdLayout {
  size_t packed, gpacked, wl, ds, gpart, tables, partials, total;
  int n_bev_blocks, n_march_blocks;
};
BwdLayout bwd_layout(const VbGrid* g, int dtype) {
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX, ncol = (size_t)g->oY * g->oX;
  const int cp = packed_channels(g->K);
  BwdLayout l;
  size_t o = 0;
  // one packed copy + one fp32 gradient accumulator PER SAMPLE: the whole batch is packed, marched and unpacked by
  // single launches (B = 8 fp32: 2 GB of the 180 GB)
  l.packed = o;   o += (size_t)g->B * vb_align256(nvox * cp * vb_elem_size(dtype));
  l.gpacked = o;  o += (size_t)g->B * vb_align256(nvox * cp * 4);
  l.wl = o;       o += vb_align256((size_t)g->B * g->oZ * ncol * 4);
  l.ds = o;       o += vb_align256((size_t)g->B * g->oZ * ncol * 4);
  l.gpart = o;    o += vb_align256((size_t)vb_ceil_div(g->K + 3, kBevBwdGroup) * g->B * g->oZ * ncol * 4);
  l.tables = o;   o += vb_align256((size_t)(3 * (g->oX + g->oY + g->oZ) + 2 * (g->vX + g->vY + g->vZ)) * 4 + 64);
  l.n_bev_blocks = vb_ceil_div(ncol, 256) * g->B;
  const int patches = vb_ceil_div(g->fW, kPatchW) * vb_ceil_div(g->fH, kPatchH);
  l.n_march_blocks = vb_ceil_div(patches, kMarchBwdThreads / 32) * g->N;
  l.partials = o; o += vb_align256((size_t)(l.n_bev_blocks + (size_t)l.n_march_blocks * g->B) * 4);
  l.total = o;
  return l;
}

template <typename T>
int launch_render_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbRenderIn* in,
                      const VbRenderOut* out, const VbRenderGrad* gr, int branches, char* ws, cudaStream_t st) {
  constexpr int K = 18, C = 16;
  if (g->K != K || g->C != C || g->oZ > kMaxLevels) return VB200_ERR_ARG;
  const BwdLayout l = bwd_layout(g, VbType<T>::code);
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX;
  const int ncol = g->oY * g->oX;
  const int cp = packed_channels(K);
  const T* den = reinterpret_cast<const T*>(in->density);
  const T* sem = reinterpret_cast<const T*>(in->sem);
  const T* rgb = reinterpret_cast<const T*>(in->rgb);
  T* packed = reinterpret_cast<T*>(ws + l.packed);
  float* gpacked = reinterpret_cast<float*>(ws + l.gpacked);
  float* wl_ws = reinterpret_cast<float*>(ws + l.wl);
  float* ds_ws = reinterpret_cast<float*>(ws + l.ds);
  float* partials = reinterpret_cast<float*>(ws + l.partials);
  const bool cam = branches & VB200_BRANCH_CAM, bev = branches & VB200_BRANCH_BEV;

  BevTables bt;
  {
    int* p = reinterpret_cast<int*>(ws + l.tables);
    bt.xi0 = p; p += g->oX;  bt.yi0 = p; p += g->oY;  bt.li0 = p; p += g->oZ;
    bt.xlo = p; p += g->vX;  bt.xhi = p; p += g->vX;
    bt.ylo = p; p += g->vY;  bt.yhi = p; p += g->vY;
    bt.llo = p; p += g->vZ;  bt.lhi = p; p += g->vZ;
    float* f = reinterpret_cast<float*>(p);
    bt.xw0 = f; f += g->oX;  bt.xw1 = f; f += g->oX;
    bt.yw0 = f; f += g->oY;  bt.yw1 = f; f += g->oY;
    bt.lw0 = f; f += g->oZ;  bt.lw1 = f; f += g->oZ;
  }
  int n_partials = 0;
  if (bev) {
    VbTraceScope tr(VB_K_UNPACK_BEV_BWD, st, 3);
    bev_tables_kernel<<<1, 256, 0, st>>>(*g, *t, bt);
    VB_LAUNCH_CHECK();
    const size_t smem = (size_t)kMaxLevels * kBevBwdThreads * sizeof(float);
    const int ngroups = vb_ceil_div(K + 3, kBevBwdGroup);
    float* gpart_ws = reinterpret_cast<float*>(ws + l.gpart);
    bev_bwd_partials_kernel<T, K><<<dim3(vb_ceil_div(ncol, kBevBwdThreads), g->B, ngroups), kBevBwdThreads, smem, st>>>(
        *g, *t, den, sem, rgb, gr->g_bev_rgb, gr->g_bev_seg, gr->g_bev_height, gpart_ws, ds_ws);
    VB_LAUNCH_CHECK();
    bev_bwd_composite_kernel<K><<<dim3(vb_ceil_div(ncol, 256), g->B), 256, 0, st>>>(
        *g, in->beta, gr->g_voxel_density, gpart_ws, ngroups, wl_ws, ds_ws, partials);
    VB_LAUNCH_CHECK();
    n_partials += l.n_bev_blocks;
  }
  const int patches = vb_ceil_div(g->fW, kPatchW) * vb_ceil_div(g->fH, kPatchH);
  const size_t packed_stride = vb_align256(nvox * cp * sizeof(T)) / sizeof(T);
  const size_t gpacked_stride = vb_align256(nvox * cp * 4) / 4;
  if (cam) {
    if (cudaMemsetAsync(gpacked, 0, (size_t)g->B * gpacked_stride * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
    if (in->packed) {
      packed = const_cast<T*>(reinterpret_cast<const T*>(in->packed));     // the forward's copy (read only here)
    } else {
      VbTraceScope tr(VB_K_PACK, st);
      pack_cam_volume_kernel<T, K><<<dim3(vb_ceil_div(nvox, kPackThreads * PackVox<T>::n), g->B), kPackThreads, 0, st>>>(
          den, sem, rgb, packed, (int)nvox, packed_stride, nullptr);
      VB_LAUNCH_CHECK();
    }
    VbTraceScope tr(VB_K_MARCH_BWD, st);
    dim3 grid(vb_ceil_div(patches, kMarchBwdThreads / 32), g->N, g->B);
    if (in->geom)
      march_bwd_kernel<T, K, false><<<grid, kMarchBwdThreads, 0, st>>>(
          *g, *t, vb_render_div(g), d_mats, in->geom, packed, in->beta, out->rgb, out->seg, out->depth, gr->g_rgb, gr->g_seg,
          gr->g_depth, gpacked, partials + n_partials, packed_stride, gpacked_stride);
    else
      march_bwd_kernel<T, K, true><<<grid, kMarchBwdThreads, 0, st>>>(
          *g, *t, vb_render_div(g), d_mats, nullptr, packed, in->beta, out->rgb, out->seg, out->depth, gr->g_rgb, gr->g_seg,
          gr->g_depth, gpacked, partials + n_partials, packed_stride, gpacked_stride);
    VB_LAUNCH_CHECK();
    n_partials += l.n_march_blocks * g->B;
  }
  {
    VbTraceScope tr(VB_K_UNPACK_BEV_BWD, st);
    unpack_gather_kernel<T, K, C><<<dim3(vb_ceil_div(nvox, 256), g->B), 256, 0, st>>>(
        *g, bt, cam ? gpacked : nullptr, wl_ws, ds_ws, gr->g_bev_rgb, gr->g_bev_seg,
        reinterpret_cast<const T*>(gr->g_voxel_output), reinterpret_cast<T*>(gr->g_density),
        reinterpret_cast<T*>(gr->g_sem), reinterpret_cast<T*>(gr->g_rgb_in), reinterpret_cast<T*>(gr->g_feat),
        gpacked_stride, bev ? 1 : 0);
    VB_LAUNCH_CHECK();
  }
  {
    VbTraceScope tr(VB_K_MISC, st);
    beta_reduce_kernel<<<1, 32, 0, st>>>(partials, n_partials, in->beta, gr->g_beta);
    VB_LAUNCH_CHECK();
  }
  return VB200_OK;
}

}  // namespace

extern "C" size_t vb200_render_bwd_workspace(const VbGrid* g, int dtype) {
  if (!g) return 0;
  return bwd_layout(g, dtype).total;
}

extern "C" int vb200_render_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbRenderIn* in,
                                int dtype, const VbRenderOut* out, const VbRenderGrad* grad, int branches,
                                void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && in && out && grad && d_workspace);
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS && g->D >= 2);
  VB_CHECK_ARG(g->density_mode == VB200_DENSITY_SDF || g->density_mode == VB200_DENSITY_NAIVE);
  VB_CHECK_ARG(in->density && in->sem && in->rgb && in->feat && in->beta);
  VB_CHECK_ARG(grad->g_density && grad->g_sem && grad->g_rgb_in && grad->g_feat && grad->g_beta);
  VB_CHECK_ARG((branches & (VB200_BRANCH_CAM | VB200_BRANCH_BEV)) != 0);
  if (branches & VB200_BRANCH_CAM) VB_CHECK_ARG(out->rgb && out->seg && out->depth);
  if (workspace_bytes < vb200_render_bwd_workspace(g, dtype)) return VB200_ERR_WORKSPACE;
  if ((uintptr_t)d_workspace & 15) return VB200_ERR_ALIGN;
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(d_workspace);
  switch (dtype) {
    case VB200_F32: return launch_render_bwd<float>(g, t, d_mats, in, out, grad, branches, ws, st);
    case VB200_BF16: return launch_render_bwd<__nv_bfloat16>(g, t, d_mats, in, out, grad, branches, ws, st);
    case VB200_F16: return launch_render_bwd<__half>(g, t, d_mats, in, out, grad, branches, ws, st);
    default: return VB200_ERR_DTYPE;
  }
}
