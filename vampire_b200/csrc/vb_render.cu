// vb_render.cu -- R1-R6 (+ T4): volume rendering of the 3-D feature volume into the six cameras
// (depth / semantics / rgb) and top-down into BEV.
//
// Reference semantics: volume_rendering_from_multiple_views (BV2:391-467):
//   R1  vol = cat[density_feature, semantic_logits, rgb, base_features]            (199 MB copy)
//   R2  g = (geom[:, :, :D-1] - lo) / ext * 2 - 1 ; mask = all(-1 <= g <= 1)
//   R3  grid_sample(vol, g, align_corners=True) * mask -> nan_to_num                (873 MB)
//   R4  sigma = Laplace(ch0); w_i = (1 - e^{-sigma_i delta_i}) e^{-tau_i}; sum w*(rgb, sem, mid) + bg
//   R5  grid_sample(vol, det-grid centres), flip z          R6  same compositing over oZ levels
//
// Here:
//   pack_cam_volume  = R1 restricted to the 1+K+3 channels the camera branch consumes, written
//                      channels-last so one voxel corner is 3 (bf16) / 6 (fp32) 128-bit loads;
//                      run sample by sample right before the march so the packed volume
//                      (63 / 126 MB) is still L2-resident (126 MB L2) when the rays gather it.
//   march_fwd        = R2+R3+R4 fused: one thread per ray, a warp = an 8x4 pixel patch marching in
//                      lock-step (neighbouring rays share voxel corners => L1 hits), geometry
//                      recomputed bit-exactly from 6 matrices instead of reading 70 MB/sample,
//                      warp-level early termination when every ray's transmittance < term_eps.
//   bev_fwd          = R5+R6 fused: reads the four NCDHW tensors directly (regular stencil, fully
//                      coalesced), never materialises the 38-channel cat.
#include "vb_render_common.cuh"
#include "vb_march_planned.cuh"
#include "vb_march_staged.cuh"
#include "vb_trace.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace {

// The BEV branch and the camera branch of one render call are independent: fork the BEV kernels
// onto a per-device side stream and join before returning, so the two instruction-bound kernel
// families fill each other's idle issue slots.  Everything stays stream-ordered w.r.t. the caller's
// stream (fork event before, join event after), so caller-owned buffers remain valid.
std::atomic<int> g_render_fork_disabled{0};
std::atomic<int> g_march_split{0};      // 0 = default (VB200_MARCH_SPLIT, else never), 1 = never, -1 = auto, 2 / 4 / 8 = forced

constexpr int kMaxSplit = 8;
struct SideStream {
  cudaStream_t stream = nullptr;      // low priority: BEV branch
  cudaStream_t stream_hi = nullptr;   // high priority: packs of the later sample groups (see launch_render_fwd)
  cudaEvent_t fork = nullptr, join = nullptr, first_pack = nullptr;
  cudaEvent_t packed[kMaxSplit] = {};
  // The events are shared by every caller on this device.  A cudaStreamWaitEvent captures the event's state when it
  // is ENQUEUED, so reuse across calls is fine -- what must not happen is two host threads interleaving their
  // record / wait pairs.  `use` is held by a render call from its first record to its join (host-side enqueue only,
  // microseconds), which makes concurrent calls from several threads / streams of one device safe.
  std::mutex use;
};
// holds SideStream::use and guarantees the caller's stream is re-joined on EVERY exit path once the side stream has
// been given work (the Python side frees the workspace and outputs as soon as the call returns)
struct SideSession {
  SideStream* side = nullptr;
  cudaStream_t caller = nullptr;
  bool forked = false;
  std::unique_lock<std::mutex> lock;
  void acquire(SideStream* s, cudaStream_t st) {
    if (!side && s) {
      side = s;
      caller = st;
      lock = std::unique_lock<std::mutex>(s->use);
    }
  }
  ~SideSession() {
    if (side && forked) cudaStreamWaitEvent(caller, side->join, 0);
  }
};
SideStream* side_stream_for_current_device() {
  static std::mutex mu;
  static SideStream table[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  SideStream& s = table[dev];
  if (!s.stream) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // lowest priority: never starve the caller's stream
    if (cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&s.stream_hi, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.first_pack, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (auto& e : s.packed)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}

// ---- R2+R3+R4: camera ray march, one sample (blockIdx.z = sample within this launch) ---------
// NANSAFE: the packed volume holds a non-finite value somewhere (flag raised by the pack): keep
// torch.nan_to_num of the interpolated features (BV2:421) per sample.  Otherwise every interpolated value is a
// convex combination of finite values, nan_to_num is the identity and the corners are accumulated directly
// into the ray's channel sums (24 fewer live registers, no per-sample finiteness test).  Both variants are
// launched back to back; the one whose turn it is not returns at once.
// SPLIT: clusters of n CTAs along x, CTA r marches the r-th n-th of the samples (march_cluster_fold composes them).
template <typename T, int K, bool FROM_MATS, bool FASTDIV, bool NANSAFE, bool SPLIT = false>
__global__ void __launch_bounds__(kMarchThreads, VB_MARCH_MINB) march_fwd_kernel(VbGrid g, VbTables t, VbRenderDiv dv,
                                                                  const float* __restrict__ d_mats,
                                                                  const float* __restrict__ d_geom,
                                                                  const T* __restrict__ packed,
                                                                  const int* __restrict__ nonfinite_flag,
                                                                  const float* __restrict__ beta_ptr,
                                                                  float* __restrict__ o_rgb, float* __restrict__ o_seg,
                                                                  float* __restrict__ o_depth, int b0) {
  constexpr int CP = packed_channels(K);
  if ((*nonfinite_flag != 0) != NANSAFE) return;
  __shared__ float s_m[VB200_MAT_SLOTS * 16];
  const int b = b0 + blockIdx.z, n = blockIdx.y;
  for (int i = threadIdx.x; i < VB200_MAT_SLOTS * 16; i += blockDim.x)
    s_m[i] = __ldg(d_mats + (size_t)(b * g.N + n) * VB200_MAT_SLOTS * 16 + i);
  __syncthreads();
  const bool has_bda = (g.has_bda != 0) && !block_is_identity(s_m + 5 * 16);   // slot 5 = bda
  const bool affine = FROM_MATS && block_ida_inv_affine(s_m + 3 * 16);          // slot 3 = ida^-1

  const int patches_x = (g.fW + kPatchW - 1) / kPatchW;
  const int patches_y = (g.fH + kPatchH - 1) / kPatchH;
  const int S = g.D - 1, HW = g.fH * g.fW;
  const int nseg = SPLIT ? (int)cooperative_groups::this_cluster().num_blocks() : 1;
  const int seg = SPLIT ? (int)cooperative_groups::this_cluster().block_rank() : 0;
  const int xblock = SPLIT ? blockIdx.x / nseg : blockIdx.x;
  const int i_lo = SPLIT ? seg * S / nseg : 0, i_hi = SPLIT ? (seg + 1) * S / nseg : S;
  const int patch_raw = xblock * (kMarchThreads / 32) + (threadIdx.x >> 5);
  if (!SPLIT && patch_raw >= patches_x * patches_y) return;  // whole warp leaves together (split: stays for the barriers)
  const int patch = SPLIT ? min(patch_raw, patches_x * patches_y - 1) : patch_raw;
  const int lane = threadIdx.x & 31;
  const int w = (patch % patches_x) * kPatchW + (lane % kPatchW);
  const int h = (patch / patches_x) * kPatchH + (lane / kPatchW);
  const bool active = (w < g.fW) && (h < g.fH) && (!SPLIT || patch_raw < patches_x * patches_y);
  const int wc = min(w, g.fW - 1), hc = min(h, g.fH - 1);

  const int nvox = g.vZ * g.vY * g.vX;
  const float beta = fabsf(__ldg(beta_ptr)) + g.beta_min;
  const T* vol = packed + (size_t)blockIdx.z * nvox * CP;  // packed holds only this launch's samples
  const float u = __ldg(t.us + wc), vv = __ldg(t.vs + hc);
  const float* gsrc = FROM_MATS ? nullptr : d_geom + ((size_t)(b * g.N + n) * g.D * HW + (size_t)hc * g.fW + wc) * 3;

  float rayA[2] = {0.0f, 0.0f};
  if (affine) frustum_ray_affine(s_m, u, vv, rayA);
  auto point = [&](int d, float (&p)[3]) {
    if (FROM_MATS) {
      if (affine) frustum_point_affine(s_m, has_bda, rayA, __ldg(t.ds + d), p);
      else frustum_point<false>(s_m, has_bda, u, vv, __ldg(t.ds + d), p);
      // BV2:612 nan_to_num: one test guards the per-component fix-up (a non-finite component makes the sum
      // non-finite; a finite sum that overflows only takes the slow, still exact, path)
      if (!(fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]) <= 3.402823466e+38f)) {
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = nan_to_num(p[a], -1e3f);
      }
    } else {
      const float* q = gsrc + (size_t)d * HW * 3;
      p[0] = __ldg(q); p[1] = __ldg(q + 1); p[2] = __ldg(q + 2);
    }
  };

  float acc = 0.0f, dep = 0.0f, tau = 0.0f;
  float ch[K + 3];
#pragma unroll
  for (int c = 0; c < K + 3; ++c) ch[c] = 0.0f;

  // a masked / out-of-volume sample has feature 0: sigma(0) is a per-launch constant (~2.27e-4)
  const float inv_beta = 1.0f / beta;
  const float sigma_masked = vb_density_rcp(g, 0.0f, inv_beta);
  // One sample position, prepared ONE STEP AHEAD of its use: ncu attributed a quarter of all stall samples to
  // the first use of the eight density gathers (issued and consumed back to back, 5 warps per scheduler cannot
  // hide an L1-miss).  Now the geometry of sample i+1 is computed and its density loads are issued before
  // sample i is composited, so they have the whole value phase of sample i to land.
  struct Smp {
    bool valid, live;      // inside the volume; and the ray exists
    int v0;                // voxel index of the base corner
    int dxo, sy, sz;       // NANSAFE: per-sample corner strides (clamped); otherwise launch constants
    float fx, fy, fz;      // far-corner weights (near = 1 - far)
    T raw[8];              // density at the eight corners, not yet widened (widening would wait for the load)
  };
  const int c_sy = g.vX, c_sz = g.vY * g.vX;
  auto prepare = [&](const float (&pp)[3], Smp& sm) {
    const RenderCoord rc = render_coord<FROM_MATS && FASTDIV>(g, pp, &dv);
    sm.valid = rc.valid;
    sm.live = rc.valid && active;
    if (sm.live) {
      // valid => 0 <= i0 <= size-1, so only the far corner can leave the grid, and only when i == size-1
      // exactly, where its weight i - i0 is exactly 0.
      int x0 = rc.x0, y0 = rc.y0, z0 = rc.z0;
      int dxo = 1, sy = c_sy, sz = c_sz;
      if (!NANSAFE) {
        // All values are finite here, so a zero-weight corner may be ANY in-grid voxel: shift the base one
        // voxel inwards instead of clamping the far corner (weights become (0, 1) exactly), which makes the
        // eight corner offsets launch constants: no clamps, no selects, immediate x offset.  Needs every
        // grid dimension >= 2 (the launcher routes thinner grids to the NANSAFE variant).
        x0 = min(x0, g.vX - 2); y0 = min(y0, g.vY - 2); z0 = min(z0, g.vZ - 2);
      } else {
        // clamp the far corner's address (its weight is 0) instead of branching
        dxo = x0 + 1 < g.vX ? 1 : 0; sy = y0 + 1 < g.vY ? c_sy : 0; sz = z0 + 1 < g.vZ ? c_sz : 0;
        sm.dxo = dxo; sm.sy = sy; sm.sz = sz;
      }
      sm.fx = rc.ix - (float)x0; sm.fy = rc.iy - (float)y0; sm.fz = rc.iz - (float)z0;
      sm.v0 = (z0 * g.vY + y0) * g.vX + x0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int cx = q & 1, cy = (q >> 1) & 1, cz = q >> 2;
        sm.raw[q] = __ldg(vol + (size_t)(sm.v0 + (cy ? sy : 0) + (cz ? sz : 0)) * CP + (cx ? dxo * CP : 0));
      }
    }
  };

  float p0[3], p1[3], p2[3];
  point(i_lo, p0);
  point(i_lo + 1, p1);
  Smp nxt;
  prepare(p0, nxt);
  float delta_n;
  {
    const float dx = p1[0] - p0[0], dy = p1[1] - p0[1], dz = p1[2] - p0[2];
    delta_n = sqrtf(dx * dx + dy * dy + dz * dz);                           // BV2:426
  }
  // A ray is a straight line and the volume a convex box, so once a ray has LEFT the box it never
  // re-enters: every later sample is masked (feature 0, sigma_masked) and only delta_i and mid_i enter
  // the compositing.  When every ray of the warp has left decisively (outside by > 1 mm on an axis along
  // which it keeps moving outward -- far beyond fp32 rounding of the per-sample geometry) or is already
  // opaque, the warp finishes with a geometry-free tail loop (delta_i = the ray's constant step length,
  // equal to the reference's per-sample norm up to ~1e-7 relative).
  bool was_valid = false, exited = false;
  for (int i = i_lo; i < i_hi; ++i) {
    const float trans = expf(-tau);
    const float delta = delta_n;
    if (g.term_eps > 0.0f) {
      const bool done = !active || trans < g.term_eps;
      if (__all_sync(0xffffffffu, done)) break;
      if (FROM_MATS && __all_sync(0xffffffffu, done || exited)) {   // a caller-supplied geom tensor need not be straight rays
        if (exited && !done) {
          float tr = trans;
          for (int ii = i; ii < i_hi; ++ii) {
            const float sd = sigma_masked * delta;
            const float wgt = (1.0f - expf(-sd)) * tr;
            acc += wgt;
            dep = fmaf(wgt, __ldg(t.mids + ii), dep);
            tau += sd;
            tr = expf(-tau);
          }
        }
        break;
      }
    }
    const Smp cur = nxt;
    if (cur.valid) {
      was_valid = true;
    } else if (was_valid && !exited) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float hi = g.seg_lo[a] + g.seg_ext[a];
        exited = exited || (p0[a] > hi + 1e-3f && p1[a] > p0[a]) || (p0[a] < g.seg_lo[a] - 1e-3f && p1[a] < p0[a]);
      }
    }
    if (i + 1 < i_hi) {                     // look ahead: geometry + density gathers of sample i+1
      point(i + 2, p2);
      prepare(p1, nxt);
      const float dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
      delta_n = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    float sigma = sigma_masked;
    float cw[8];
    const bool live = cur.live;
    const int sy = NANSAFE ? cur.sy : c_sy, sz = NANSAFE ? cur.sz : c_sz, dxo = NANSAFE ? cur.dxo : 1;
    if (live) {
      const float wx[2] = {1.0f - cur.fx, cur.fx}, wy[2] = {1.0f - cur.fy, cur.fy}, wz[2] = {1.0f - cur.fz, cur.fz};
      // phase 1: density channel only -> sigma, alpha
      float s0 = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int cx = q & 1, cy = (q >> 1) & 1, cz = q >> 2;
        cw[q] = wx[cx] * wy[cy] * wz[cz];
        s0 = fmaf(cw[q], widen_elem(cur.raw[q]), s0);
      }
      if (NANSAFE) s0 = nan_to_num(s0, 0.0f);                                  // BV2:421
      sigma = vb_density_rcp(g, s0, inv_beta);                   // BV2:423
    }
    const float sd = sigma * delta;                                           // BV2:429
    const float wgt = (1.0f - expf(-sd)) * trans;                           // BV2:430-434
    acc += wgt;
    dep = fmaf(wgt, __ldg(t.mids + i), dep);
    // phase 2: the 21 value channels, only where they can contribute.  In free space alpha = 1 - exp(-sd)
    // is exactly 0.0f in fp32 (the reference's too), so w * v == 0 exactly: skipping the fetch is bit-neutral.
    if (!NANSAFE) {
      if (live && wgt != 0.0f) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          PackedLoad<T, CP>::template fma_values<K + 3>(
              vol + (size_t)(cur.v0 + ((q & 2) ? sy : 0) + ((q & 4) ? sz : 0)) * CP + ((q & 1) ? dxo * CP : 0),
              cw[q] * wgt, ch);
      }
    } else if (live && wgt != 0.0f) {
      float v[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) v[c] = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        PackedLoad<T, CP>::fma_corner(
            vol + (size_t)(cur.v0 + ((q & 2) ? sy : 0) + ((q & 4) ? sz : 0)) * CP + ((q & 1) ? dxo * CP : 0), cw[q], v);
      // torch.nan_to_num (BV2:421): any NaN/inf channel makes the channel sum non-finite, so one
      // test guards the per-channel fix-up
      float chk = 0.0f;
#pragma unroll
      for (int c = 1; c < K + 4; ++c) chk += v[c];
      if (!(fabsf(chk) <= 3.402823466e+38f)) {
#pragma unroll
        for (int c = 1; c < K + 4; ++c) v[c] = nan_to_num(v[c], 0.0f);
      }
#pragma unroll
      for (int c = 0; c < K + 3; ++c) ch[c] = fmaf(wgt, v[1 + c], ch[c]);
    }
    tau += sd;
#pragma unroll
    for (int a = 0; a < 3; ++a) { p0[a] = p1[a]; p1[a] = p2[a]; }
  }
  if (SPLIT) {
    float trans = expf(-tau);
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

// ---- R5+R6: BEV branch ---------------------------------------------------------------------------
// The sample position of output voxel (oz, oy, ox) is its centre normalised by the SEG bounds
// (BV2:408-417): the x/y terms are level-independent and the z terms column-independent, so a
// column needs (oZ + 1) xy-bilinear rows per channel instead of 8 corners x oZ levels, and the z
// terms live in a tiny per-block table.  Two kernels:
//   bev_weights   one thread per column: density plane -> sigma (= the voxel_density output),
//                 compositing weights w_l (workspace, 2.6 MB/sample, L2-resident) and bev_height
//   bev_channels  grid = (column tiles, 37 channels, samples): every other channel plane is read
//                 exactly once, coalesced along x, straight from the NCDHW tensors -- the 38-channel
//                 cat of the reference is never materialised.
template <typename T>
__global__ void __launch_bounds__(256) bev_weights_kernel(VbGrid g, VbTables t, const T* __restrict__ den,
                                                          const float* __restrict__ beta_ptr,
                                                          float* __restrict__ o_height, float* __restrict__ o_density,
                                                          float* __restrict__ wl_ws, float* __restrict__ th_ws) {
  __shared__ BevLevel s_lv[kMaxLevels];
  bev_level_table(g, t, s_lv);
  const int b = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = g.oY * g.oX;
  if (col >= ncol) return;
  const BevColumn bc = bev_column(g, t, col % g.oX, col / g.oX);
  const T* plane = den + (size_t)b * g.vZ * g.vY * g.vX;
  const float beta = fabsf(__ldg(beta_ptr)) + g.beta_min;
  float tau = 0.0f, height = 0.0f, prev_lo = 0.0f;
  int prev_z0 = -1000000;
  for (int l = 0; l < g.oZ; ++l) {
    const BevLevel L = s_lv[l];
    const float hi = (L.z0 + 1 == prev_z0) ? prev_lo : bev_row<T>(g, bc, plane, L.z0 + 1);
    const float lo = bev_row<T>(g, bc, plane, L.z0);
    prev_z0 = L.z0;
    prev_lo = lo;
    const float sigma = vb_density(g, L.wz0 * lo + L.wz1 * hi, beta);      // BV2:445
    const float sd = sigma * g.bev_delta;
    const float w = (1.0f - expf(-sd)) * expf(-tau);                                       // BV2:454-458
    tau += sd;
    height = fmaf(w, __ldg(t.bev_mids + l), height);
    const size_t o = ((size_t)b * g.oZ + l) * ncol + col;
    o_density[o] = sigma;
    wl_ws[o] = w;
    // BEV epilogue (BV2:627-630), once per voxel: voxel_output * bev_density.tanh() ('sdf') / * bev_density ('naive')
    if (th_ws) th_ws[o] = g.density_mode == VB200_DENSITY_NAIVE ? sigma : tanhf(sigma);
  }
  o_height[(size_t)b * ncol + col] = height;                                               // BV2:461
}

template <typename T, int K, int C>
__global__ void __launch_bounds__(256) bev_channels_kernel(VbGrid g, VbTables t, const T* __restrict__ sem,
                                                           const T* __restrict__ rgb, const T* __restrict__ feat,
                                                           const float* __restrict__ wl_ws, float* __restrict__ o_rgb,
                                                           float* __restrict__ o_seg, T* __restrict__ o_feat,
                                                           const float* __restrict__ th_ws) {
  __shared__ BevLevel s_lv[kMaxLevels];
  bev_level_table(g, t, s_lv);
  const int b = blockIdx.z, j = blockIdx.y;        // j: 0..K-1 sem | K..K+2 rgb | K+3.. feat
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = g.oY * g.oX;
  if (col >= ncol) return;
  const BevColumn bc = bev_column(g, t, col % g.oX, col / g.oX);
  const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
  float prev_lo = 0.0f;
  int prev_z0 = -1000000;
  if (j < K + 3) {   // composited channels                                                 BV2:459-460
    const T* plane = j < K ? sem + ((size_t)b * K + j) * nvox : rgb + ((size_t)b * 3 + (j - K)) * nvox;
    const float* wl = wl_ws + (size_t)b * g.oZ * ncol + col;
    float a = 0.0f;
    for (int l = 0; l < g.oZ; ++l) {
      const BevLevel L = s_lv[l];
      const float hi = (L.z0 + 1 == prev_z0) ? prev_lo : bev_row<T>(g, bc, plane, L.z0 + 1);
      const float lo = bev_row<T>(g, bc, plane, L.z0);
      prev_z0 = L.z0;
      prev_lo = lo;
      a = fmaf(__ldg(wl + (size_t)l * ncol), L.wz0 * lo + L.wz1 * hi, a);
    }
    if (j < K) o_seg[((size_t)b * K + j) * ncol + col] = a;
    else o_rgb[((size_t)b * 3 + (j - K)) * ncol + col] = a;
  } else {           // resampled base features, returned unweighted                        BV2:448
    const int c = j - (K + 3);
    const T* plane = feat + ((size_t)b * C + c) * nvox;
    T* o = o_feat + ((size_t)b * C + c) * g.oZ * ncol + col;
    for (int l = 0; l < g.oZ; ++l) {
      const BevLevel L = s_lv[l];
      const float hi = (L.z0 + 1 == prev_z0) ? prev_lo : bev_row<T>(g, bc, plane, L.z0 + 1);
      const float lo = bev_row<T>(g, bc, plane, L.z0);
      prev_z0 = L.z0;
      prev_lo = lo;
      float v = L.wz0 * lo + L.wz1 * hi;
      if (th_ws) v *= __ldg(th_ws + ((size_t)b * g.oZ + l) * ncol + col);                    // BV2:627-630
      o[(size_t)l * ncol] = VbType<T>::cvt(v);
    }
  }
}

// ---- bev_channels, vectorised: one thread = 4 consecutive output columns of one channel plane -------------
// Output column ox samples input columns x0(ox), x0(ox)+1 with x0(ox) in {ox-1, ox} whenever the det grid
// shares the seg grid's xy lattice (the reference config).  When that offset is uniform over a warp
// (MODE 1: x0 = ox, MODE 0: x0 = ox - 1) a thread needs in[ox0 .. ox0+4] resp. in[ox0-1 .. ox0+3]: one
// 64/128-bit load plus one value from the neighbouring lane.  Anything else takes MODE 2 (scalar
// gathers, exact for any grid).  ~5x fewer instructions per output than the scalar kernel.
template <typename T> struct Vec4Load;
template <> struct Vec4Load<float> {
  __device__ __forceinline__ static void ld(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  __device__ __forceinline__ static void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4Load<__nv_bfloat16> {
  __device__ __forceinline__ static void ld(const __nv_bfloat16* p, float (&o)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  __device__ __forceinline__ static void st(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <> struct Vec4Load<__half> {
  __device__ __forceinline__ static void ld(const __half* p, float (&o)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  __device__ __forceinline__ static void st(__half* p, const float (&v)[4]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

struct BevQuadX {
  int ox0;                 // first output column, multiple of 4
  int x0[4];               // base input column per output column (MODE 2 only)
  float wx0[4], wx1[4];    // zeroed where the corner leaves the grid
};

// x-interpolated values of the input row at `row` for the 4 columns, accumulated as out += wy * value
template <typename T, int MODE>
__device__ __forceinline__ void quad_row(const VbGrid& g, const BevQuadX& q, const T* __restrict__ row, int lane,
                                         float wy, float (&out)[4]) {
  float a[4], b[4];
  if (MODE == 2) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xa = min(max(q.x0[c], 0), g.vX - 1), xb = min(max(q.x0[c] + 1, 0), g.vX - 1);
      a[c] = VbType<T>::ld(row + xa);
      b[c] = VbType<T>::ld(row + xb);
    }
  } else {
    float v[4];
    Vec4Load<T>::ld(row + q.ox0, v);
    if (MODE == 1) {
      float right = __shfl_down_sync(0xffffffffu, v[0], 1);
      if (lane == 31) right = VbType<T>::ld(row + min(q.ox0 + 4, g.vX - 1));   // weight is 0 if outside
      a[0] = v[0]; a[1] = v[1]; a[2] = v[2]; a[3] = v[3];
      b[0] = v[1]; b[1] = v[2]; b[2] = v[3]; b[3] = right;
    } else {
      float left = __shfl_up_sync(0xffffffffu, v[3], 1);
      if (lane == 0) left = VbType<T>::ld(row + max(q.ox0 - 1, 0));
      a[0] = left; a[1] = v[0]; a[2] = v[1]; a[3] = v[2];
      b[0] = v[0]; b[1] = v[1]; b[2] = v[2]; b[3] = v[3];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = fmaf(wy, fmaf(q.wx1[c], b[c], q.wx0[c] * a[c]), out[c]);
}

template <typename T, int MODE>
__device__ __forceinline__ void quad_zrow(const VbGrid& g, const BevQuadX& q, const T* __restrict__ plane, int z,
                                          int y0, float wy0, float wy1, int lane, float (&out)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = 0.0f;
  if (z < 0 || z >= g.vZ) return;                                   // zeros padding (uniform)
  const T* zr = plane + z * g.vY * g.vX;
  if (y0 >= 0 && y0 < g.vY) quad_row<T, MODE>(g, q, zr + y0 * g.vX, lane, wy0, out);
  if (y0 + 1 >= 0 && y0 + 1 < g.vY) quad_row<T, MODE>(g, q, zr + (y0 + 1) * g.vX, lane, wy1, out);
}

// ---- fast path (MODE 0 / 1): per input z-row ONE xy-interpolated 4-vector per channel -----------------------
// A thread owns 4 output columns and loops over a chunk of channel planes, so the coordinate prologue
// (5 strict axis_coord's with IEEE divisions) is paid once per ~8 channels instead of once per channel.
// Per channel and level it fetches at most one new z-row: 2 vector loads (rows y0, y0+1) + 2 scalar edge
// loads (the 5th column), combined with 16 pre-multiplied xy weights (4 FMAs per column).  The row of the
// NEXT level is requested before the current level is composited, so two rows are always in flight.
template <typename T> struct Raw4;                       // raw (un-widened) 4 consecutive elements
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
template <> struct Raw4<__half> { using type = uint2; };

template <typename T> __device__ __forceinline__ void widen4(const typename Raw4<T>::type& r, float (&o)[4]);
template <> __device__ __forceinline__ void widen4<float>(const float4& r, float (&o)[4]) {
  o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
template <> __device__ __forceinline__ void widen4<__nv_bfloat16>(const uint2& r, float (&o)[4]) {
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void widen4<__half>(const uint2& r, float (&o)[4]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

struct BevQuadFast {
  int off_a, off_b;        // element offsets of (row y0, column ox0) / (row y0+1, column ox0) inside a z-row (rows clamped)
  int off_ea, off_eb;      // ... of the 5th column (MODE 1: ox0+4, MODE 0: ox0-1; clamped)
  float w[4][4];           // [column][a0, a1, b0, b1] = wy * wx, zero where the tap leaves the grid
};

template <typename T> struct RawRow {
  typename Raw4<T>::type a, b;
  T ea, eb;
};

template <typename T>
__device__ __forceinline__ RawRow<T> bev_load_row(const T* __restrict__ plane, const BevQuadFast& q, int zoff) {
  RawRow<T> r;
  using V = typename Raw4<T>::type;
  r.a = __ldg(reinterpret_cast<const V*>(plane + (zoff + q.off_a)));
  r.b = __ldg(reinterpret_cast<const V*>(plane + (zoff + q.off_b)));
  r.ea = __ldg(plane + (zoff + q.off_ea));
  r.eb = __ldg(plane + (zoff + q.off_eb));
  return r;
}

template <typename T> __device__ __forceinline__ float widen1(T v);
template <> __device__ __forceinline__ float widen1<float>(float v) { return v; }
template <> __device__ __forceinline__ float widen1<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float widen1<__half>(__half v) { return __half2float(v); }

template <typename T, int MODE>
__device__ __forceinline__ void bev_finish_row(const RawRow<T>& r, const BevQuadFast& q, float (&out)[4]) {
  float va[4], vb[4];
  widen4<T>(r.a, va);
  widen4<T>(r.b, vb);
  const float ea = widen1<T>(r.ea), eb = widen1<T>(r.eb);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float a0 = MODE == 1 ? va[c] : (c > 0 ? va[c > 0 ? c - 1 : 0] : ea);
    const float a1 = MODE == 1 ? (c < 3 ? va[c < 3 ? c + 1 : 3] : ea) : va[c];
    const float b0 = MODE == 1 ? vb[c] : (c > 0 ? vb[c > 0 ? c - 1 : 0] : eb);
    const float b1 = MODE == 1 ? (c < 3 ? vb[c < 3 ? c + 1 : 3] : eb) : vb[c];
    out[c] = fmaf(q.w[c][0], a0, fmaf(q.w[c][1], a1, fmaf(q.w[c][2], b0, q.w[c][3] * b1)));
  }
}

struct BevLevelX {       // per level: row offsets of z0 / z0+1 inside a plane, their validity, reuse flag
  int zoff, zoff_hi;     // z * vY * vX (z clamped so the address is always valid)
  float wz0, wz1;        // zeroed when the row is outside the grid (zeros padding)
  int flags;             // bit0: row z0+1 of this level == row z0 of the previous level (reuse it)
                         // bit1: row z0+1 is inside the grid      bit2: row z0 is inside the grid
};

// the same per level as one 16-byte shared-memory word (one LDS.128 per level instead of five loads and a select; ncu
// attributed 11 % of the kernel's instructions to fetching BevLevelX in the level loop).  The padding of row z0 rides
// in wz0 = 0 -- its address is clamped, the value finite -- so the fast path needs no flag for it.
struct __align__(16) BevLevelF {
  int zoff;              // z0 * vY * vX (z0 clamped)
  float wz0, wz1;        // zero when the row is outside the grid
  int flags;             // bit0: row z0+1 == row z0 of the previous level; bit1: row z0+1 is inside the grid
};

// One channel plane, all levels, MODE 0/1 (mode1, warp-uniform).  MAP (block-uniform): composite with the
// level weights into a (oY,oX) map; otherwise store the resampled rows as T.
template <typename T, bool MAP, bool EPI = false>
__device__ __forceinline__ void bev_fast_channel(const BevLevelF* __restrict__ lv, const int* __restrict__ zhi, int oZ,
                                                 const bool mode1, const BevQuadFast& q, const T* __restrict__ plane,
                                                 bool live, const float* __restrict__ wl, float* __restrict__ o_map,
                                                 T* __restrict__ o_feat, int ncol) {
  float prev[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  BevLevelF L = lv[0];
  RawRow<T> nxt = bev_load_row<T>(plane, q, L.zoff);
  const float4* wp = reinterpret_cast<const float4*>(wl);
  const int wstep = ncol >> 2;                             // ncol % 4 == 0 (oX % 4 == 0)
#pragma unroll 2
  for (int l = 0; l < oZ; ++l, wp += wstep, o_feat += ncol) {
    float hi[4], lo[4];
    if (L.flags & 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) hi[c] = prev[c];
    } else if (L.flags & 2) {
      const RawRow<T> h = bev_load_row<T>(plane, q, zhi[l]);
      if (mode1) bev_finish_row<T, 1>(h, q, hi);
      else bev_finish_row<T, 0>(h, q, hi);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) hi[c] = 0.0f;
    }
    const RawRow<T> cur = nxt;
    const BevLevelF Ln = lv[l + 1];                        // the table has one entry past the last level
    nxt = bev_load_row<T>(plane, q, Ln.zoff);              // in flight while this level is composited
    if (mode1) bev_finish_row<T, 1>(cur, q, lo);
    else bev_finish_row<T, 0>(cur, q, lo);
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = fmaf(L.wz1, hi[c], L.wz0 * lo[c]);
      prev[c] = lo[c];
    }
    if (MAP) {                                                                              // BV2:459-460
      const float4 w = __ldg(wp);
      acc[0] = fmaf(w.x, v[0], acc[0]); acc[1] = fmaf(w.y, v[1], acc[1]);
      acc[2] = fmaf(w.z, v[2], acc[2]); acc[3] = fmaf(w.w, v[3], acc[3]);
    } else if (live) {                                                                      // BV2:448
      if (EPI) {                                 // fused BEV epilogue: x tanh(sigma) of the same voxels (BV2:627-630)
        const float4 th = __ldg(wp);
        v[0] *= th.x; v[1] *= th.y; v[2] *= th.z; v[3] *= th.w;
      }
      Vec4Load<T>::st(o_feat, v);
    }
    L = Ln;
  }
  if (MAP && live) *reinterpret_cast<float4*>(o_map) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// generic fallback (MODE 2): scalar gathers, exact for any det grid
template <typename T, int K, int C, bool EPI = false>
__device__ __forceinline__ void bev_quad_channel_generic(const VbGrid& g, const BevLevel* __restrict__ lv,
                                                         const BevQuadX& q, const T* __restrict__ plane, int y0,
                                                         float wy0, float wy1, int lane, bool live,
                                                         const float* __restrict__ wl, float* __restrict__ o_map,
                                                         T* __restrict__ o_feat, int ncol) {
  float prev_lo[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int prev_z0 = -1000000;
  for (int l = 0; l < g.oZ; ++l) {
    const BevLevel L = lv[l];
    float hi[4], lo[4];
    if (L.z0 + 1 == prev_z0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) hi[c] = prev_lo[c];
    } else {
      quad_zrow<T, 2>(g, q, plane, L.z0 + 1, y0, wy0, wy1, lane, hi);
    }
    quad_zrow<T, 2>(g, q, plane, L.z0, y0, wy0, wy1, lane, lo);
    prev_z0 = L.z0;
    float4 w = make_float4(0, 0, 0, 0);
    if (o_map && live) w = __ldg(reinterpret_cast<const float4*>(wl + (size_t)l * ncol));
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = fmaf(L.wz1, hi[c], L.wz0 * lo[c]);
      prev_lo[c] = lo[c];
    }
    if (o_map) {
      acc[0] = fmaf(w.x, v[0], acc[0]); acc[1] = fmaf(w.y, v[1], acc[1]);
      acc[2] = fmaf(w.z, v[2], acc[2]); acc[3] = fmaf(w.w, v[3], acc[3]);
    } else if (live) {
      if (EPI) {                                 // fused BEV epilogue (wl = tanh(sigma) workspace in this mode)
        const float4 th = __ldg(reinterpret_cast<const float4*>(wl + (size_t)l * ncol));
        v[0] *= th.x; v[1] *= th.y; v[2] *= th.z; v[3] *= th.w;
      }
      Vec4Load<T>::st(o_feat + (size_t)l * ncol, v);
    }
  }
  if (o_map && live) *reinterpret_cast<float4*>(o_map) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// Measured alternative (round 2, removed again): a y-strip walk -- a thread owns 4 columns of one plane over 16
// consecutive output rows and keeps the x-interpolated input row y0 + 1 of every z-row in registers as row y0 of the
// next output row, so an output row costs one vector + one edge load per z-row instead of two.  ncu, R50 B=8 bf16:
// 510 us vs 266 us for this kernel: the loads halved but the instruction count did not fall (179 M vs 164 M warp
// instructions: the loads were never the bulk, the widen / interpolate / composite arithmetic is, and this kernel already
// issues 57 % of its slots), while 44 persistent registers halved the occupancy (128 regs, 25 % vs 50 %) of a kernel
// that lives on hiding latency (long scoreboard 7-8 warps per issue in both).
//
// channel chunks: the K+3 composited planes and the C feature planes are split into chunks of <= 8 planes,
// one chunk per blockIdx.y
__host__ __device__ constexpr int bev_chunks(int n) { return (n + 7) / 8; }
__host__ __device__ constexpr int bev_chunk_len(int n) { return (n + bev_chunks(n) - 1) / bev_chunks(n); }
__host__ __device__ constexpr int bev_groups(int K, int C) { return bev_chunks(K + 3) + bev_chunks(C); }

template <typename T, int K, int C, bool EPI>
__global__ void __launch_bounds__(64, 16) bev_channels_vec4_kernel(VbGrid g, VbTables t, const T* __restrict__ sem,
                                                               const T* __restrict__ rgb, const T* __restrict__ feat,
                                                               const float* __restrict__ wl_ws,
                                                               float* __restrict__ o_rgb, float* __restrict__ o_seg,
                                                               T* __restrict__ o_feat, const float* __restrict__ th_ws) {
  __shared__ BevLevel s_lv[kMaxLevels];
  __shared__ BevLevelF s_lx[kMaxLevels + 1];       // + one entry the look-ahead of the last level may read
  __shared__ int s_zhi[kMaxLevels];
  bev_level_table(g, t, s_lv);
  if (threadIdx.x <= g.oZ) {
    const int l = min((int)threadIdx.x, g.oZ - 1);   // entry oZ repeats the last level: its row is requested, never used
    const BevLevel L = s_lv[l];
    BevLevelF X;
    const bool in0 = L.z0 >= 0 && L.z0 < g.vZ, in1 = L.z0 + 1 >= 0 && L.z0 + 1 < g.vZ;
    X.zoff = min(max(L.z0, 0), g.vZ - 1) * g.vY * g.vX;
    X.wz0 = in0 ? L.wz0 : 0.0f;
    X.wz1 = in1 ? L.wz1 : 0.0f;
    X.flags = ((l > 0 && L.z0 + 1 == s_lv[l - 1].z0) ? 1 : 0) | (in1 ? 2 : 0);
    s_lx[threadIdx.x] = X;
    if (threadIdx.x < g.oZ) s_zhi[l] = min(max(L.z0 + 1, 0), g.vZ - 1) * g.vY * g.vX;
  }
  __syncthreads();
  const int b = blockIdx.z, grp = blockIdx.y;
  const int tiles_x = (g.oX + 255) / 256;
  const int oy = blockIdx.x / tiles_x;
  const int lane = threadIdx.x & 31;
  const int ox_raw = (blockIdx.x % tiles_x) * 256 + threadIdx.x * 4;
  const bool live = ox_raw < g.oX;                     // oX % 4 == 0 guaranteed by the launcher
  BevQuadX q;
  q.ox0 = live ? ox_raw : 0;
  bool all1 = q.ox0 + 3 < g.vX, all0 = all1;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    axis_coord(__ldg(t.oxs + q.ox0 + c), g.seg_lo[0], g.seg_ext[0], g.vX, q.x0[c], q.wx0[c], q.wx1[c]);
    all1 = all1 && (q.x0[c] == q.ox0 + c);
    all0 = all0 && (q.x0[c] == q.ox0 + c - 1);
    if (!(q.x0[c] >= 0 && q.x0[c] < g.vX)) q.wx0[c] = 0.0f;
    if (!(q.x0[c] + 1 >= 0 && q.x0[c] + 1 < g.vX)) q.wx1[c] = 0.0f;
  }
  int y0;
  float wy0, wy1;
  axis_coord(__ldg(t.oys + oy), g.seg_lo[1], g.seg_ext[1], g.vY, y0, wy0, wy1);
  const bool w1 = __all_sync(0xffffffffu, all1), w0 = __all_sync(0xffffffffu, all0);   // one path per warp

  const int ncol = g.oY * g.oX;
  const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
  const int col0 = oy * g.oX + q.ox0;
  constexpr int NM = K + 3, GM = bev_chunks(NM), LM = bev_chunk_len(NM), LF = bev_chunk_len(C);
  const bool is_map = grp < GM;
  const int c_begin = is_map ? grp * LM : (grp - GM) * LF;
  const int c_end = is_map ? min(c_begin + LM, NM) : min(c_begin + LF, C);
  // composited planes: the level weights; feature planes: tanh(sigma) when the BEV epilogue is fused, else nothing
  // composited planes: the level weights; feature planes: tanh(sigma) when the BEV epilogue is fused (else unused)
  const float* wl = ((EPI && !is_map) ? th_ws : wl_ws) + (size_t)b * g.oZ * ncol + col0;

  auto planes = [&](int j, const T*& plane, float*& o_map, T*& o_f) {
    o_map = nullptr;
    o_f = nullptr;
    if (is_map) {
      plane = j < K ? sem + ((size_t)b * K + j) * nvox : rgb + ((size_t)b * 3 + (j - K)) * nvox;
      o_map = (j < K ? o_seg + ((size_t)b * K + j) * ncol : o_rgb + ((size_t)b * 3 + (j - K)) * ncol) + col0;
    } else {
      plane = feat + ((size_t)b * C + j) * nvox;
      o_f = o_feat + ((size_t)b * C + j) * g.oZ * ncol + col0;
    }
  };
  if (w1 || w0) {
    BevQuadFast f;
    const bool y0in = y0 >= 0 && y0 < g.vY, y1in = y0 + 1 >= 0 && y0 + 1 < g.vY;
    const int ya = min(max(y0, 0), g.vY - 1), yb = min(max(y0 + 1, 0), g.vY - 1);
    const int xe = w1 ? min(q.ox0 + 4, g.vX - 1) : max(q.ox0 - 1, 0);
    f.off_a = ya * g.vX + q.ox0; f.off_b = yb * g.vX + q.ox0;
    f.off_ea = ya * g.vX + xe;   f.off_eb = yb * g.vX + xe;
    const float ya_w = y0in ? wy0 : 0.0f, yb_w = y1in ? wy1 : 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      f.w[c][0] = ya_w * q.wx0[c]; f.w[c][1] = ya_w * q.wx1[c];
      f.w[c][2] = yb_w * q.wx0[c]; f.w[c][3] = yb_w * q.wx1[c];
    }
    for (int j = c_begin; j < c_end; ++j) {
      const T* plane;
      float* o_map;
      T* o_f;
      planes(j, plane, o_map, o_f);
      if (is_map) bev_fast_channel<T, true>(s_lx, s_zhi, g.oZ, w1, f, plane, live, wl, o_map, o_f, ncol);
      else bev_fast_channel<T, false, EPI>(s_lx, s_zhi, g.oZ, w1, f, plane, live, wl, o_map, o_f, ncol);
    }
  } else {
    for (int j = c_begin; j < c_end; ++j) {
      const T* plane;
      float* o_map;
      T* o_f;
      planes(j, plane, o_map, o_f);
      if (is_map) bev_quad_channel_generic<T, K, C, false>(g, s_lv, q, plane, y0, wy0, wy1, lane, live, wl, o_map, o_f, ncol);
      else bev_quad_channel_generic<T, K, C, EPI>(g, s_lv, q, plane, y0, wy0, wy1, lane, live, wl, o_map, o_f, ncol);
    }
  }
}

// ---- bev_channels, TMA-staged ---------------------------------------------------------------------------------
// Same arithmetic as the fast path above, but the input rows travel HBM -> shared memory as 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx) through a ring of kStages row-pair slots.  The direct-load version keeps
// one level (512 B per warp) in flight and ncu shows it latency-bound (long scoreboard on the first use of every
// row, ~2.4 TB/s = bytes in flight / latency); the ring keeps kStages x 1 KB per block in flight without holding
// registers, the 5th column is simply the next shared-memory element, and the 64-bit global address arithmetic
// disappears from the consumers' level loop.
//   block = 2 consumer warps (128 columns each) + 1 producer warp whose lane 0 issues every copy
//   item  = one z-row of one channel plane: its two y-rows [ya, yb] x [cs, ce) (16-byte aligned column range
//           covering the block's 256 columns +- 8), copied to slot (item % kStages) at element offset 8
//   order = per channel, per level: the hi row when it is neither reused nor outside the grid, then the lo row
//           (the same for every channel, tabulated once per block in s_item_z)
//   sync  = full[slot]: producer arrive.expect_tx + 2 copies complete_tx, consumers try_wait;
//           empty[slot]: one arrive per consumer warp once the slot's values sit in registers, producer try_wait.
// Measured on B200 (R50, B=8, bf16; direct-load kernel: 0.291 ms): first version (one thread of a 2-warp block
// issuing, a block barrier per item, generic shared addressing) 0.435 ms at twice the instruction count; this
// lean ring 0.335 ms, the same with one merged 1 KB copy per item instead of two -- so neither the copy rate nor
// the issue overhead of the ring is what holds it back; with 10 blocks x 2 consumer warps per SM it simply has
// fewer warps to cover the arithmetic than the direct kernel's 32.  Opt-in (VB200_BEV_TMA=1), bit-identical.
__device__ __forceinline__ uint32_t vb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vb_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void vb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vb_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void vb_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void vb_mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a lost transaction must fail loudly, never hang
  }
}
template <typename T> struct SmemRow;      // 4 consecutive elements + 1 scalar from shared memory, by 32-bit address
template <> struct SmemRow<float> {
  __device__ __forceinline__ static float4 ld4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
  }
  __device__ __forceinline__ static float ld1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
  }
};
template <> struct SmemRow<__nv_bfloat16> {
  __device__ __forceinline__ static uint2 ld4(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
  }
  __device__ __forceinline__ static __nv_bfloat16 ld1(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return __ushort_as_bfloat16(v);
  }
};
template <> struct SmemRow<__half> {
  __device__ __forceinline__ static uint2 ld4(uint32_t a) { return SmemRow<__nv_bfloat16>::ld4(a); }
  __device__ __forceinline__ static __half ld1(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return __ushort_as_half(v);
  }
};

constexpr int kBevRowPad = 288;   // 8 zero | up to 272 copied | >= 8 zero  (elements)
constexpr int kBevTmaThreads = 96;
template <typename T> struct BevStages { static constexpr int n = sizeof(T) == 4 ? 6 : 8; };

template <typename T, int K, int C>
__global__ void __launch_bounds__(kBevTmaThreads, sizeof(T) == 4 ? 8 : 10) bev_channels_tma_kernel(
    VbGrid g, VbTables t, const T* __restrict__ sem, const T* __restrict__ rgb, const T* __restrict__ feat,
    const float* __restrict__ wl_ws, float* __restrict__ o_rgb, float* __restrict__ o_seg, T* __restrict__ o_feat) {
  constexpr int S = BevStages<T>::n;
  constexpr uint32_t kRowBytes = kBevRowPad * sizeof(T), kSlotBytes = 2 * kRowBytes;
  __shared__ __align__(128) T s_ring[S][2][kBevRowPad];
  __shared__ __align__(8) uint64_t s_full[S], s_empty[S];
  __shared__ BevLevel s_lv[kMaxLevels];
  __shared__ BevLevelX s_lx[kMaxLevels];
  __shared__ int s_item_z[2 * kMaxLevels];
  __shared__ int s_nitems;
  const uint32_t ring_u32 = vb_smem_u32(&s_ring[0][0][0]);
  const uint32_t full_u32 = vb_smem_u32(&s_full[0]), empty_u32 = vb_smem_u32(&s_empty[0]);
  bev_level_table(g, t, s_lv);
  if (threadIdx.x < g.oZ) {
    const int l = threadIdx.x;
    const BevLevel L = s_lv[l];
    BevLevelX X;
    const bool in0 = L.z0 >= 0 && L.z0 < g.vZ, in1 = L.z0 + 1 >= 0 && L.z0 + 1 < g.vZ;
    X.zoff = min(max(L.z0, 0), g.vZ - 1) * g.vY * g.vX;
    X.zoff_hi = min(max(L.z0 + 1, 0), g.vZ - 1) * g.vY * g.vX;
    X.wz0 = in0 ? L.wz0 : 0.0f;
    X.wz1 = in1 ? L.wz1 : 0.0f;
    X.flags = ((l > 0 && L.z0 + 1 == s_lv[l - 1].z0) ? 1 : 0) | (in1 ? 2 : 0) | (in0 ? 4 : 0);
    s_lx[l] = X;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      vb_mbar_init(full_u32 + 8 * i, 1);
      vb_mbar_init(empty_u32 + 8 * i, 2);     // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // the per-channel item order
    int n = 0;
    for (int l = 0; l < g.oZ; ++l) {
      const BevLevelX L = s_lx[l];
      if (!(L.flags & 1) && (L.flags & 2)) s_item_z[n++] = L.zoff_hi;
      s_item_z[n++] = L.zoff;
    }
    s_nitems = n;
  }
  const int b = blockIdx.z, grp = blockIdx.y;
  const int tiles_x = (g.oX + 255) / 256;
  const int oy = blockIdx.x / tiles_x;
  const int lane = threadIdx.x & 31;
  const bool producer = threadIdx.x >= 64;
  const int tile_x0 = (blockIdx.x % tiles_x) * 256;
  const int ox_raw = tile_x0 + (threadIdx.x & 63) * 4;
  const bool live = !producer && ox_raw < g.oX;        // oX % 4 == 0 guaranteed by the launcher
  BevQuadX q;
  q.ox0 = live ? ox_raw : 0;
  bool all1 = q.ox0 + 3 < g.vX, all0 = all1;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    axis_coord(__ldg(t.oxs + q.ox0 + c), g.seg_lo[0], g.seg_ext[0], g.vX, q.x0[c], q.wx0[c], q.wx1[c]);
    all1 = all1 && (q.x0[c] == q.ox0 + c);
    all0 = all0 && (q.x0[c] == q.ox0 + c - 1);
    if (!(q.x0[c] >= 0 && q.x0[c] < g.vX)) q.wx0[c] = 0.0f;
    if (!(q.x0[c] + 1 >= 0 && q.x0[c] + 1 < g.vX)) q.wx1[c] = 0.0f;
  }
  int y0;
  float wy0, wy1;
  axis_coord(__ldg(t.oys + oy), g.seg_lo[1], g.seg_ext[1], g.vY, y0, wy0, wy1);
  const bool w1 = __all_sync(0xffffffffu, all1 || !live), w0 = __all_sync(0xffffffffu, all0 || !live);
  // the copied column range [cs, ce) must hold every live thread's 4 columns and its 5th (else: zero pad)
  const int cs = max(tile_x0 - 8, 0), ce = min(tile_x0 + 264, g.vX);
  const bool fast_block = __syncthreads_and((w1 || w0) && (!live || (q.ox0 >= cs && q.ox0 + 4 <= ce))) != 0;

  const int ncol = g.oY * g.oX;
  const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
  const int col0 = oy * g.oX + q.ox0;
  constexpr int NM = K + 3, GM = bev_chunks(NM), LM = bev_chunk_len(NM), LF = bev_chunk_len(C);
  const bool is_map = grp < GM;
  const int c_begin = is_map ? grp * LM : (grp - GM) * LF;
  const int c_end = is_map ? min(c_begin + LM, NM) : min(c_begin + LF, C);
  const float* wl = wl_ws + (size_t)b * g.oZ * ncol + col0;
  auto plane_of = [&](int j) -> const T* {
    return is_map ? (j < K ? sem + ((size_t)b * K + j) * nvox : rgb + ((size_t)b * 3 + (j - K)) * nvox)
                  : feat + ((size_t)b * C + j) * nvox;
  };
  auto outputs_of = [&](int j, float*& o_map, T*& o_f) {
    o_map = nullptr;
    o_f = nullptr;
    if (is_map) o_map = (j < K ? o_seg + ((size_t)b * K + j) * ncol : o_rgb + ((size_t)b * 3 + (j - K)) * ncol) + col0;
    else o_f = o_feat + ((size_t)b * C + j) * g.oZ * ncol + col0;
  };
  if (!fast_block) {   // irregular det grid: exact scalar gathers, no staging
    if (producer) return;
    for (int j = c_begin; j < c_end; ++j) {
      float* o_map;
      T* o_f;
      outputs_of(j, o_map, o_f);
      bev_quad_channel_generic<T, K, C>(g, s_lv, q, plane_of(j), y0, wy0, wy1, lane, live, wl, o_map, o_f, ncol);
    }
    return;
  }

  // ---- staged path ----
  const int ya = min(max(y0, 0), g.vY - 1), yb = min(max(y0 + 1, 0), g.vY - 1);
  const int ncopy = ce - cs;                                   // elements per row copy, multiple of 16 bytes
  // Adjacent y-rows that are copied whole are ONE contiguous 2*vX-element range: one bulk copy per item instead
  // of two (the copy engine is rate-limited on ~0.5 KB transfers).  Row 1 then starts right behind row 0; its
  // "column -1" / row 0's "column vX" are the neighbouring row's finite edge values, weight 0 like the zero pads.
  const bool merged = (yb == ya + 1) && cs == 0 && ce == g.vX && (2 * g.vX + 16 <= 2 * kBevRowPad);
  const uint32_t row1_off = merged ? (uint32_t)g.vX * sizeof(T) : kRowBytes;
  // zero the whole ring once (the copies never touch the pads); order these generic-proxy writes before the
  // async-proxy copies
  for (int i = threadIdx.x; i < S * 2 * kBevRowPad; i += blockDim.x) (&s_ring[0][0][0])[i] = VbType<T>::cvt(0.0f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int nitems = s_nitems;
  const int total = (c_end - c_begin) * nitems;

  if (producer) {
    if (lane != 0) return;
    const uint32_t row_bytes = (uint32_t)ncopy * sizeof(T);
    int j = c_begin, r = 0;
    const T* plane = plane_of(j) + cs;
    const int oa = ya * g.vX, ob = yb * g.vX;
    for (int it = 0; it < total; ++it) {
      const int slot = it % S;
      if (it >= S) vb_mbar_wait(empty_u32 + 8 * slot, (uint32_t)(it / S - 1) & 1u);
      const T* src = plane + s_item_z[r];
      const uint32_t dst = ring_u32 + slot * kSlotBytes + 8 * sizeof(T), bar = full_u32 + 8 * slot;
      vb_mbar_expect_tx(bar, 2u * row_bytes);
      if (merged) {
        vb_bulk_load(dst, src + oa, 2u * row_bytes, bar);
      } else {
        vb_bulk_load(dst, src + oa, row_bytes, bar);
        vb_bulk_load(dst + kRowBytes, src + ob, row_bytes, bar);
      }
      if (++r == nitems) {
        r = 0;
        ++j;
        if (j < c_end) plane = plane_of(j) + cs;
      }
    }
    return;
  }

  // ---- consumers ----
  BevQuadFast f;
  {
    const bool y0in = y0 >= 0 && y0 < g.vY, y1in = y0 + 1 >= 0 && y0 + 1 < g.vY;
    const float ya_w = y0in ? wy0 : 0.0f, yb_w = y1in ? wy1 : 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      f.w[c][0] = ya_w * q.wx0[c]; f.w[c][1] = ya_w * q.wx1[c];
      f.w[c][2] = yb_w * q.wx0[c]; f.w[c][3] = yb_w * q.wx1[c];
    }
    f.off_a = f.off_b = f.off_ea = f.off_eb = 0;
  }
  const int xi = live ? (q.ox0 - cs + 8) : 8;                  // this thread's first column inside a staged row
  const uint32_t a4 = ring_u32 + (uint32_t)xi * sizeof(T);     // ... its address in slot 0, row 0
  const uint32_t a1 = ring_u32 + (uint32_t)(w1 ? xi + 4 : xi - 1) * sizeof(T);   // its 5th column
  int consumed = 0;
  auto consume = [&](RawRow<T>& r) {
    const int slot = consumed % S;
    vb_mbar_wait(full_u32 + 8 * slot, (uint32_t)(consumed / S) & 1u);
    const uint32_t so = slot * kSlotBytes;
    r.a = SmemRow<T>::ld4(a4 + so);
    r.b = SmemRow<T>::ld4(a4 + so + row1_off);
    r.ea = SmemRow<T>::ld1(a1 + so);
    r.eb = SmemRow<T>::ld1(a1 + so + row1_off);
    __syncwarp();                                   // every lane's values are in registers
    if (lane == 0) vb_mbar_arrive(empty_u32 + 8 * slot);
    ++consumed;
  };
  const int wstep = ncol >> 2;
  for (int j = c_begin; j < c_end; ++j) {
    float* o_map;
    T* o_f;
    outputs_of(j, o_map, o_f);
    const float4* wp = reinterpret_cast<const float4*>(wl);
    float prev[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float4 wnext = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (is_map) wnext = __ldg(wp);
    for (int l = 0; l < g.oZ; ++l) {
      const BevLevelX L = s_lx[l];
      float hi[4], lo[4];
      if (L.flags & 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) hi[c] = prev[c];
      } else if (L.flags & 2) {
        RawRow<T> h;
        consume(h);
        if (w1) bev_finish_row<T, 1>(h, f, hi);
        else bev_finish_row<T, 0>(h, f, hi);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) hi[c] = 0.0f;
      }
      RawRow<T> cur;
      consume(cur);
      const float4 w = wnext;
      if (is_map && l + 1 < g.oZ) wnext = __ldg(wp + wstep);   // next level's weights, one level ahead
      if (w1) bev_finish_row<T, 1>(cur, f, lo);
      else bev_finish_row<T, 0>(cur, f, lo);
      if (!(L.flags & 4)) {                                    // row z0 outside the grid: zeros padding (uniform)
#pragma unroll
        for (int c = 0; c < 4; ++c) lo[c] = 0.0f;
      }
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[c] = fmaf(L.wz1, hi[c], L.wz0 * lo[c]);
        prev[c] = lo[c];
      }
      if (is_map) {                                                                         // BV2:459-460
        acc[0] = fmaf(w.x, v[0], acc[0]); acc[1] = fmaf(w.y, v[1], acc[1]);
        acc[2] = fmaf(w.z, v[2], acc[2]); acc[3] = fmaf(w.w, v[3], acc[3]);
      } else if (live) {                                                                    // BV2:448
        Vec4Load<T>::st(o_f, v);
      }
      wp += wstep;
      if (!is_map) o_f += ncol;
    }
    if (is_map && live) *reinterpret_cast<float4*>(o_map) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

size_t bev_level_bytes(const VbGrid* g) {
  const size_t n = (size_t)g->B * g->oZ * g->oY * g->oX * sizeof(float);
  return (n + 255) & ~(size_t)255;
}
// compositing weights | tanh(sigma) of the fused BEV epilogue | one 256-byte slot: the pack's non-finite flags
size_t bev_weight_bytes(const VbGrid* g) { return 2 * bev_level_bytes(g) + 256; }

size_t packed_bytes_per_sample(const VbGrid* g, int dtype) {
  const size_t n = (size_t)g->vZ * g->vY * g->vX * packed_channels(g->K) * vb_elem_size(dtype);
  return (n + 255) & ~(size_t)255;
}

template <typename T>
int launch_render_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbRenderIn* in,
                      const VbRenderOut* out, int branches, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int K = 18, C = 16;
  if (g->K != K || g->C != C || g->oZ > kMaxLevels) return VB200_ERR_ARG;
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX;
  const size_t per = packed_bytes_per_sample(g, VbType<T>::code);
  const size_t bev_bytes = bev_weight_bytes(g);
  if (ws_bytes < bev_bytes + per) return VB200_ERR_WORKSPACE;
  // workspace = [packed camera volumes of `group` samples][BEV compositing weights of all samples]
  const int group = (int)(((ws_bytes - bev_bytes) / per) < (size_t)g->B ? ((ws_bytes - bev_bytes) / per) : (size_t)g->B);
  const size_t cam_bytes = (size_t)group * per;
  const T* den = reinterpret_cast<const T*>(in->density);
  const T* sem = reinterpret_cast<const T*>(in->sem);
  const T* rgb = reinterpret_cast<const T*>(in->rgb);
  const T* feat = reinterpret_cast<const T*>(in->feat);
  const int patches = vb_ceil_div(g->fW, kPatchW) * vb_ceil_div(g->fH, kPatchH);
  const int ncol = g->oY * g->oX;
  SideSession session;          // declared first: its destructor (the join) runs on every return below
  bool& forked = session.forked;
  SideStream* side = nullptr;
  // BEV branch on stream `bst`
  auto launch_bev = [&](cudaStream_t bst) -> int {
    float* wl_ws = reinterpret_cast<float*>((char*)ws + cam_bytes);
    // VB200_RENDER_TANH_EPILOGUE: voxel_output leaves the kernel already multiplied by tanh(sigma) (BV2:627-630)
    float* th_ws = (in->flags & VB200_RENDER_TANH_EPILOGUE)
                       ? reinterpret_cast<float*>((char*)ws + cam_bytes + bev_level_bytes(g)) : nullptr;
    VbTraceScope tr(VB_K_BEV_FWD, bst, 2);
    bev_weights_kernel<T><<<dim3(vb_ceil_div(ncol, 256), g->B), 256, 0, bst>>>(
        *g, *t, den, in->beta, out->bev_height, out->voxel_density, wl_ws, th_ws);
    VB_LAUNCH_CHECK();
    const bool vec_ok = (g->vX % 4 == 0) && (g->oX % 4 == 0) &&
                        ((((uintptr_t)sem | (uintptr_t)rgb | (uintptr_t)feat | (uintptr_t)out->voxel_output |
                           (uintptr_t)out->bev_rgb | (uintptr_t)out->bev_seg | (uintptr_t)wl_ws) & 15) == 0);
    // bulk copies need 16-byte aligned rows: vX a multiple of 16 bytes of elements (plane bases are checked above)
    // Opt-in (VB200_BEV_TMA=1): measured slower than the direct-load kernel on B200 (R50, B=8, bf16: 0.335 vs
    // 0.291 ms, see the kernel's header).  Kept as a measured alternative, parity-tested bit-identical.
    static const bool tma_env = getenv("VB200_BEV_TMA") != nullptr;
    const bool tma_ok = vec_ok && tma_env && !th_ws && ((size_t)g->vX * sizeof(T)) % 16 == 0 && g->vX >= 8;
    if (tma_ok) {
      static VbPerDeviceFlag carve;
      if (vb_func_attr_per_device(bev_channels_tma_kernel<T, K, C>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared, carve) != VB200_OK)
        return VB200_ERR_CUDA;
      bev_channels_tma_kernel<T, K, C><<<dim3(g->oY * vb_ceil_div(g->oX, 256), bev_groups(K, C), g->B), kBevTmaThreads, 0, bst>>>(
          *g, *t, sem, rgb, feat, wl_ws, out->bev_rgb, out->bev_seg, reinterpret_cast<T*>(out->voxel_output));
    } else if (vec_ok)
    {
      const dim3 vgrid(g->oY * vb_ceil_div(g->oX, 256), bev_groups(K, C), g->B);
      if (th_ws)
        bev_channels_vec4_kernel<T, K, C, true><<<vgrid, 64, 0, bst>>>(
            *g, *t, sem, rgb, feat, wl_ws, out->bev_rgb, out->bev_seg, reinterpret_cast<T*>(out->voxel_output), th_ws);
      else
        bev_channels_vec4_kernel<T, K, C, false><<<vgrid, 64, 0, bst>>>(
            *g, *t, sem, rgb, feat, wl_ws, out->bev_rgb, out->bev_seg, reinterpret_cast<T*>(out->voxel_output), th_ws);
    }
    else
      bev_channels_kernel<T, K, C><<<dim3(vb_ceil_div(ncol, 256), K + 3 + C, g->B), 256, 0, bst>>>(
          *g, *t, sem, rgb, feat, wl_ws, out->bev_rgb, out->bev_seg, reinterpret_cast<T*>(out->voxel_output), th_ws);
    VB_LAUNCH_CHECK();
    return VB200_OK;
  };
  if (!(branches & VB200_BRANCH_CAM)) return launch_bev(st);
  if (per != nvox * packed_channels(K) * sizeof(T)) return VB200_ERR_ARG;  // march indexes densely
  // thin grids (a dimension of one voxel) cannot shift the corner base inwards: raise the flag up front,
  // the pack never lowers it, and the exact clamp-and-zero variant of the march runs
  const int flag_init = (g->vX < 2 || g->vY < 2 || g->vZ < 2) ? 1 : 0;
  int* nf_flags = reinterpret_cast<int*>((char*)ws + cam_bytes + bev_bytes - 256);   // one int per sub-round
  const VbRenderDiv dv = vb_render_div(g);
  // The fork is an event edge caller stream -> side stream -> caller stream, so it is also legal while the caller's
  // stream is being captured into a CUDA graph (the side stream joins the capture and is joined back before this
  // call returns): a replayed graph keeps the BEV branch beside the march (dp.GraphedTrainStep).
  static const bool no_fork_env = getenv("VB200_NO_FORK") != nullptr;   // measurement aid
  const bool no_fork = no_fork_env || g_render_fork_disabled.load(std::memory_order_relaxed) != 0;
  const bool may_fork = !no_fork;

  auto pack_round = [&](int b0, int nb, T* region, int* flag, cudaStream_t ps) -> int {
    VbTraceScope tr(VB_K_PACK, ps);
    if (cudaMemsetAsync(flag, flag_init, sizeof(int), ps) != cudaSuccess) return VB200_ERR_CUDA;
    pack_cam_volume_kernel<T, K><<<dim3(vb_ceil_div(nvox, kPackThreads * PackVox<T>::n), nb), kPackThreads, 0, ps>>>(
        den + (size_t)b0 * nvox, sem + (size_t)b0 * K * nvox, rgb + (size_t)b0 * 3 * nvox, region, (int)nvox,
        per / sizeof(T), flag);
    VB_LAUNCH_CHECK();
    return VB200_OK;
  };
  auto march_round = [&](int b0, int nb, const T* region, const int* flag) -> int {
    const int xblocks = vb_ceil_div(patches, kMarchThreads / 32);
    dim3 grid(xblocks, g->N, nb);
    VbTraceScope tr(VB_K_MARCH_FWD, st, 2);
    // depth split (march_cluster_fold): clusters of nseg CTAs share each group of patches; opt-in through
    // vb200_render_set_march_split() / VB200_MARCH_SPLIT.
    int nseg = 1;
    {
      static const int split_env = getenv("VB200_MARCH_SPLIT") ? atoi(getenv("VB200_MARCH_SPLIT")) : 0;
      const int set = g_march_split.load(std::memory_order_relaxed);
      const int mode = set != 0 ? set : split_env;       // 0 / 1: never, -1: auto, 2 / 4 / 8: forced
      const long long blocks = (long long)xblocks * g->N * nb, slots = (long long)VB_SM_COUNT_B200 * VB_MARCH_MINB;
      // measured on B200 (R50, bf16 / fp32, render call incl. pack): B = 1: 0.209 / 0.301 ms unsplit, 0.186 / 0.254 with
      // 2 segments, 0.209 / 0.280 with 4, 0.257 / 0.328 with 8; B = 2: 0.274 unsplit, 0.305 with 2 -- a later segment
      // cannot see that the ray is already opaque and gathers what the unsplit march skips, so "auto" splits only
      // while the launch is under one wave, and only in two.  Off by default: the fold changes the summation order, and
      // a sample's result must not depend on the batch it rides in (the sharding invariant of the data-parallel path).
      if (mode > 1) nseg = mode;
      else if (mode < 0 && blocks < slots) nseg = 2;
      while (nseg > 1 && (nseg > 8 || (g->D - 1) / nseg < 4)) nseg /= 2;    // portable cluster size; >= 4 samples each
      nseg = nseg >= 8 ? 8 : nseg >= 4 ? 4 : nseg >= 2 ? 2 : 1;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(xblocks * nseg, g->N, nb);
    cfg.blockDim = dim3(kMarchThreads);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nseg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const float* geom_null = nullptr;
#define VB_MARCH(FM, FD, NS)                                                                                        \
  march_fwd_kernel<T, K, FM, FD, NS><<<grid, kMarchThreads, 0, st>>>(*g, *t, dv, d_mats, in->geom, region, flag,    \
                                                                     in->beta, out->rgb, out->seg, out->depth, b0)
#define VB_MARCH_SPLIT(FD)                                                                                          \
  do {                                                                                                              \
    if (cudaLaunchKernelEx(&cfg, march_fwd_kernel<T, K, true, FD, false, true>, *g, *t, dv, d_mats, geom_null,      \
                           region, flag, (const float*)in->beta, out->rgb, out->seg, out->depth, b0) != cudaSuccess) \
      return VB200_ERR_CUDA;                                                                                        \
  } while (0)
    if (in->geom) {
      VB_MARCH(false, false, false);
      VB_MARCH(false, false, true);
    } else if (in->plans) {
      // cached geometry; a non-finite packed volume (flag raised by the pack) takes the recomputing NaN-safe variant
      // VB200_MARCH_STAGED=1: the shared-memory staged variant (bulk asynchronous copies of the warp's voxel box, see
      // vb_march_staged.cuh); bit-identical to the direct-gather kernel below
      static const bool staged_env = getenv("VB200_MARCH_STAGED") != nullptr;
      bool launched = false;
      if constexpr (sizeof(T) == 2) {
        if (staged_env) {
          march_fwd_staged_kernel<T, K><<<grid, kMarchThreads, 0, st>>>(*g, *t, in->plans, region, flag, in->beta,
                                                                        out->rgb, out->seg, out->depth, b0);
          launched = true;
        }
      }
      if (!launched && nseg > 1) {
        if (cudaLaunchKernelEx(&cfg, march_fwd_planned_kernel<T, K, true>, *g, *t, in->plans, region, flag,
                               (const float*)in->beta, out->rgb, out->seg, out->depth, b0) != cudaSuccess)
          return VB200_ERR_CUDA;
        launched = true;
      }
      if (!launched)
        march_fwd_planned_kernel<T, K, false><<<grid, kMarchThreads, 0, st>>>(*g, *t, in->plans, region, flag, in->beta,
                                                                              out->rgb, out->seg, out->depth, b0);
      if (vb_render_div_ok(dv)) VB_MARCH(true, true, true);
      else VB_MARCH(true, false, true);
    } else if (vb_render_div_ok(dv)) {
      if (nseg > 1) VB_MARCH_SPLIT(true);
      else VB_MARCH(true, true, false);
      VB_MARCH(true, true, true);
    } else {
      if (nseg > 1) VB_MARCH_SPLIT(false);
      else VB_MARCH(true, false, false);
      VB_MARCH(true, false, true);
    }
#undef VB_MARCH
#undef VB_MARCH_SPLIT
    VB_LAUNCH_CHECK();
    return VB200_OK;
  };
  auto fork_bev = [&]() -> int {
    // fork AFTER the first pack: the pack is HBM-bound, the march issue-bound -- the BEV kernels
    // (low-priority side stream) fill the march's idle issue slots instead of fighting the pack for DRAM
    cudaStream_t bst = st;
    if (may_fork && (side || (side = side_stream_for_current_device()) != nullptr)) {
      session.acquire(side, st);
      if (cudaEventRecord(side->fork, st) == cudaSuccess && cudaStreamWaitEvent(side->stream, side->fork, 0) == cudaSuccess) {
        bst = side->stream;
        // record the join point right away so that an early return below still waits for whatever was enqueued
        forked = cudaEventRecord(side->join, bst) == cudaSuccess;
        if (!forked) bst = st;
      }
    }
    const int rc = launch_bev(bst);
    if (forked && cudaEventRecord(side->join, bst) != cudaSuccess) return VB200_ERR_CUDA;
    return rc;
  };

  // Sample groups.  Optional experiment (VB200_RENDER_SPLIT=n, default off): when the workspace holds every
  // sample's packed volume, split the batch into n sub-rounds whose packs (HBM-bound) run on a high-priority side
  // stream under the march of the previous sub-round (L1 / issue-bound, 8 % DRAM).  Measured on B200 (R50, B=8,
  // bf16): 1.26 ms/step unsplit, 1.34 (n=2), 1.57 (n=4), 2.15 (n=8) -- a march over fewer samples loses more to
  // partial waves and to the co-running pack than hiding the pack gains, so one round over the whole batch stays
  // the default.  Everything stays stream-ordered for the caller either way.
  int nsplit = 1;
  if (group >= g->B && may_fork && g->B >= 4 && (side || (side = side_stream_for_current_device()) != nullptr)) {
    static const int split_env = getenv("VB200_RENDER_SPLIT") ? atoi(getenv("VB200_RENDER_SPLIT")) : 0;
    nsplit = split_env > 0 ? split_env : 1;
    if (nsplit > kMaxSplit) nsplit = kMaxSplit;
    if (nsplit > g->B) nsplit = g->B;
  }
  if (nsplit > 1) {
    session.acquire(side, st);
    const int sub = vb_ceil_div(g->B, nsplit);
    int rc = pack_round(0, sub < g->B ? sub : g->B, reinterpret_cast<T*>(ws), nf_flags, st);
    if (rc) return rc;
    if (cudaEventRecord(side->first_pack, st) != cudaSuccess ||
        cudaStreamWaitEvent(side->stream_hi, side->first_pack, 0) != cudaSuccess)
      return VB200_ERR_CUDA;
    for (int k = 1; k * sub < g->B; ++k) {
      const int b0 = k * sub, nb = (g->B - b0) < sub ? (g->B - b0) : sub;
      rc = pack_round(b0, nb, reinterpret_cast<T*>((char*)ws + (size_t)b0 * per), nf_flags + k, side->stream_hi);
      if (rc) return rc;
      if (cudaEventRecord(side->packed[k], side->stream_hi) != cudaSuccess) return VB200_ERR_CUDA;
    }
    if (branches & VB200_BRANCH_BEV) {
      rc = fork_bev();
      if (rc) return rc;
    }
    for (int k = 0; k * sub < g->B; ++k) {
      const int b0 = k * sub, nb = (g->B - b0) < sub ? (g->B - b0) : sub;
      if (k > 0 && cudaStreamWaitEvent(st, side->packed[k], 0) != cudaSuccess) return VB200_ERR_CUDA;
      rc = march_round(b0, nb, reinterpret_cast<const T*>((char*)ws + (size_t)b0 * per), nf_flags + k);
      if (rc) return rc;
    }
  } else {
    // Where and how the BEV branch forks was measured (R50, B=8, bf16, whole step, two runs each): after the first
    // pack on the low-priority side stream (this code) 1.077 / 1.081 ms; before the pack 1.148 / 1.157; after the pack
    // at high priority 1.173; both 1.165; no fork at all 1.190.
    for (int b0 = 0; b0 < g->B; b0 += group) {
      const int nb = (g->B - b0) < group ? (g->B - b0) : group;
      int rc = pack_round(b0, nb, reinterpret_cast<T*>(ws), nf_flags, st);
      if (rc) return rc;
      if (b0 == 0 && (branches & VB200_BRANCH_BEV)) {
        rc = fork_bev();
        if (rc) return rc;
      }
      rc = march_round(b0, nb, reinterpret_cast<const T*>(ws), nf_flags);
      if (rc) return rc;
    }
  }
  return VB200_OK;     // ~SideSession joins the side stream into the caller's
}

}  // namespace

extern "C" size_t vb200_render_fwd_workspace(const VbGrid* g, int dtype) {
  if (!g) return 0;
  // minimum: BEV weights + one packed sample per pack/march round; every further
  // vb200_render_packed_bytes() lets one more sample share a round
  return bev_weight_bytes(g) + packed_bytes_per_sample(g, dtype);
}

extern "C" size_t vb200_render_plan_rays(const VbGrid* g) {
  return g ? (size_t)g->N * march_patches(*g) * 32 : 0;
}

extern "C" int vb200_render_plan_build(const VbGrid* g, const VbTables* t, const float* d_mats, void* d_steps,
                                       float* d_delta, int16_t* d_last, void* d_box, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && d_steps && d_delta && d_last && d_box);
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS && g->D >= 2 && g->D - 1 < 32767);
  VB_CHECK_ARG((size_t)g->vZ * g->vY * g->vX <= (size_t)kPlanVoxMask + 1);
  VB_CHECK_ARG(g->vX >= 2 && g->vY >= 2 && g->vZ >= 2);     // the plan stores the inward-shifted base corner
  if ((uintptr_t)d_steps & 15) return VB200_ERR_ALIGN;
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const VbRenderDiv dv = vb_render_div(g);
  const int npatch = march_patches(*g);
  dim3 grid(vb_ceil_div(npatch, kMarchThreads / 32), g->N, g->B);
  VbTraceScope tr(VB_K_GET_GEOMETRY, st);
  if (vb_render_div_ok(dv))
    render_plan_build_kernel<true><<<grid, kMarchThreads, 0, st>>>(*g, *t, dv, d_mats, reinterpret_cast<uint4*>(d_steps),
                                                                    d_delta, d_last, reinterpret_cast<uint2*>(d_box),
                                                                    vb200_render_plan_rays(g));
  else
    render_plan_build_kernel<false><<<grid, kMarchThreads, 0, st>>>(*g, *t, dv, d_mats, reinterpret_cast<uint4*>(d_steps),
                                                                     d_delta, d_last, reinterpret_cast<uint2*>(d_box),
                                                                     vb200_render_plan_rays(g));
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" int vb200_render_set_march_split(int segments) {
  VB_CHECK_ARG(segments == -1 || segments == 0 || segments == 1 || segments == 2 || segments == 4 || segments == 8);
  g_march_split.store(segments);
  return VB200_OK;
}

extern "C" int vb200_render_set_fork(int enable) {
  g_render_fork_disabled.store(enable ? 0 : 1);
  return VB200_OK;
}

extern "C" size_t vb200_render_packed_bytes(const VbGrid* g, int dtype) {
  return g ? packed_bytes_per_sample(g, dtype) : 0;
}

extern "C" int vb200_render_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const VbRenderIn* in,
                                int dtype, const VbRenderOut* out, int branches, void* d_workspace,
                                size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && in && out);
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS && g->D >= 2);
  VB_CHECK_ARG(g->density_mode == VB200_DENSITY_SDF || g->density_mode == VB200_DENSITY_NAIVE);
  VB_CHECK_ARG(in->density && in->sem && in->rgb && in->feat && in->beta);
  VB_CHECK_ARG((branches & (VB200_BRANCH_CAM | VB200_BRANCH_BEV)) != 0);
  if (branches & VB200_BRANCH_CAM) VB_CHECK_ARG(out->rgb && out->seg && out->depth && d_workspace);
  if (branches & VB200_BRANCH_BEV)
    VB_CHECK_ARG(d_workspace && out->bev_rgb && out->bev_seg && out->bev_height && out->voxel_density &&
                 out->voxel_output);
  if ((uintptr_t)d_workspace & 15) return VB200_ERR_ALIGN;
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case VB200_F32: return launch_render_fwd<float>(g, t, d_mats, in, out, branches, d_workspace, workspace_bytes, st);
    case VB200_BF16:
      return launch_render_fwd<__nv_bfloat16>(g, t, d_mats, in, out, branches, d_workspace, workspace_bytes, st);
    case VB200_F16:
      return launch_render_fwd<__half>(g, t, d_mats, in, out, branches, d_workspace, workspace_bytes, st);
    default: return VB200_ERR_DTYPE;
  }
}
