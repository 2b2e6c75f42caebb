// vb_lift.cu -- L1-L4 (+ backward): depth (x) context lift fused with the camera->voxel gather/pool.
//
// Reference semantics (BV2:553 + get_voxel_feats BV2:483-516): build the (B,N,C,D,fH,fW) frustum
// tensor depth[d,h,w]*ctx[c,h,w] (372 MB/sample), project every voxel centre into every camera
// (get_pixel), trilinear grid_sample(align_corners=False, zeros padding) the frustum there, mask,
// and average over the cameras with a per-channel non-zero count.
//
// Here the outer product is never formed: trilinear interpolation is linear, and the frustum
// value factorises, so per (voxel, camera)
//     f[c] = sum_{4 pixel corners (j,k)} w_j w_k ctx[c,j,k] * (sum_{2 depth bins i} w_i depth[i,j,k])
// with 16 register accumulators per thread.  Context rows are re-laid out
// channels-last by a tiny pre-pass so one pixel's C channels are one or two 128-bit loads.
//
// HBM roofline: read depth + ctx once (L2-resident afterwards: 27.6 MB fp32 per sample), write the
// (B,C,vZ,vY,vX) volume once => 111.5 MB/sample fp32, 55.7 MB bf16.
#include <cstdlib>

#include "vb_lift_common.cuh"
#include "vb_trace.cuh"

namespace {

#ifndef VB_LIFT_THREADS
#define VB_LIFT_THREADS 128
#endif
#ifndef VB_LIFT_MINB
#define VB_LIFT_MINB 7
#endif
constexpr int kLiftThreads = VB_LIFT_THREADS;

// Trilinear weights exactly as ATen forms them (GridSampler: corner weight = product of
// (x_far - ix) terms), in the tolerance zone (FMA allowed).
struct TriW {
  float wx0, wx1, wy0, wy1, wz0, wz1;
};
__device__ __forceinline__ TriW tri_weights(float ix, float iy, float iz, int x0, int y0, int z0) {
  TriW w;
  w.wx1 = ix - (float)x0;
  w.wx0 = (float)(x0 + 1) - ix;
  w.wy1 = iy - (float)y0;
  w.wy0 = (float)(y0 + 1) - iy;
  w.wz1 = iz - (float)z0;
  w.wz0 = (float)(z0 + 1) - iz;
  return w;
}

// ---- forward: one thread per voxel, loop over cameras ---------------------------------------
template <typename TD, typename TC, int C, int OUT_LAYOUT, bool FASTDIV>
__global__ void __launch_bounds__(kLiftThreads, VB_LIFT_MINB) lift_pool_fwd_kernel(VbGrid g, VbTables t, VbLiftDiv dv,
                                                                     const float* __restrict__ d_mats,
                                                                     const TD* __restrict__ depth,
                                                                     const TC* __restrict__ ctx_nhwc,
                                                                     TD* __restrict__ out, uint64_t* __restrict__ cnt_out, int zrun) {
  static_assert(C <= 16, "per-channel counts are packed 4 bits each into 64 bits");
  __shared__ float s_m[VB_MAX_CAMS * VB200_MAT_SLOTS * 16];
  __shared__ float s_q[VB_MAX_CAMS * 16];   // fast cull: (K.E^-1)(bda^-1), FMA-composed
  const int b = blockIdx.y;
  stage_mats(s_m, d_mats, b, g.N);
  __syncthreads();
  const bool has_bda = (g.has_bda != 0) && !block_is_identity(s_m);   // slot 0 of camera 0 = bda^-1 (same for all cameras)
  const bool affine = block_pixel_affine(s_m, g.N, has_bda);
  for (int i = threadIdx.x; i < g.N * 16; i += blockDim.x) {
    const int n = i / 16, r = (i % 16) / 4, c = i % 4;
    const float* A = s_m + n * VB200_MAT_SLOTS * 16 + 16;   // K.E^-1
    const float* Bm = s_m + n * VB200_MAT_SLOTS * 16;       // bda^-1
    float v = A[r * 4 + c];
    if (has_bda) {
      v = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) v = fmaf(A[r * 4 + k], Bm[k * 4 + c], v);
    }
    s_q[i] = v;
  }
  __syncthreads();
  const int nvox = g.vZ * g.vY * g.vX;
  const int nplane = g.vY * g.vX;
  // a block owns kLiftThreads consecutive (y, x) positions and walks a run of `zrun` z-levels, so the
  // prologue above (matrix staging, cull-matrix composition) is paid once per run, not once per voxel
  const int pos_raw = blockIdx.x * kLiftThreads + threadIdx.x;
  const bool pos_live = pos_raw < nplane;       // no early return: the warp-level cull below shuffles
  const int pos = pos_live ? pos_raw : nplane - 1;
  const int x = pos % g.vX, y = pos / g.vX;
  const float px = __ldg(t.xs + x), py = __ldg(t.ys + y);
  const int HW = g.fH * g.fW;
  const int z_begin = blockIdx.z * zrun, z_end = min(g.vZ, z_begin + zrun);

  // per-channel non-zero camera count (BV2:509-512).  A seeing camera almost always contributes to all
  // C channels, so the common case is one shared counter; exact zeros (dead ctx channel, both depth
  // bins outside) take the packed per-channel path: 4 bits per channel, count of cameras that were ZERO.

  // ---- conservative camera culling (tolerance zone; the strict projection below alone decides `valid`) ----
  // 84 % of (voxel, camera) pairs are invisible by a wide margin.  Two levels, both with FMA/approximate
  // arithmetic and a 1 px / 5 cm guard band (fp32 rounding differences are < 0.02 px / 1e-4 m here):
  //  (1) per warp and z-run: the warp's voxels form a planar rectangle (32 x-positions times the z-run);
  //      a projective map sends it to a convex quadrilateral when all corners are in front of the
  //      camera, so if all 4 corners are rejected by the SAME half-space every voxel is.
  //      Lanes 0..4N-1 test (camera, corner); one ballot gives the camera mask (~1.5 of 6 survive).
  //  (2) per thread, for surviving cameras: the same test on the voxel itself.
  auto cull_codes = [&](int n, float qx, float qy, float qz) -> unsigned {
    // bit0 z<lo, bit1 z>hi, bit2 x<min, bit3 x>max, bit4 y<min, bit5 y>max (x/y bits only when z >= lo)
    const float* q = s_q + n * 16;
    const float cz = fmaf(q[8], qx, fmaf(q[9], qy, fmaf(q[10], qz, q[11])));
    if (!(cz >= g.d_lo - 0.05f)) return 1u;            // also catches NaN
    // The half-space argument needs the point strictly in front of the camera: the guard band reaches behind it when
    // d_lo = 0 (the 2-D lift's depth test is z > 0), where 1 / cz flips the sign of the projection.  Too close to the
    // camera plane to decide here: keep the pair, the strict projection decides.
    if (cz < 1e-3f) return 0u;
    unsigned code = (cz > g.d_hi + 0.05f) ? 2u : 0u;
    const float cx = fmaf(q[0], qx, fmaf(q[1], qy, fmaf(q[2], qz, q[3])));
    const float cy = fmaf(q[4], qx, fmaf(q[5], qy, fmaf(q[6], qz, q[7])));
    const float cw = fmaf(q[12], qx, fmaf(q[13], qy, fmaf(q[14], qz, q[15])));
    const float rz = __frcp_rn(cz);
    const float ux = cx * rz, uy = cy * rz;
    const float* I = s_m + n * VB200_MAT_SLOTS * 16 + 2 * 16;   // ida
    const float ax = fmaf(I[0], ux, fmaf(I[1], uy, fmaf(I[2], cz, I[3] * cw)));
    const float ay = fmaf(I[4], ux, fmaf(I[5], uy, fmaf(I[6], cz, I[7] * cw)));
    code |= (ax < -1.5f) ? 4u : 0u;
    code |= (ax > g.x_hi + 1.0f) ? 8u : 0u;
    code |= (ay < -1.5f) ? 16u : 0u;
    code |= (ay > g.y_hi + 1.0f) ? 32u : 0u;
    if (!(ax == ax) || !(ay == ay)) code = 0u;          // NaN: cannot reject here
    return code;
  };
  // warp-level mask for the whole run: the warp's voxels form a planar rectangle (32 x-positions times
  // the z-run); it is rejected for a camera when its 4 corners are rejected by the same half-space.
  unsigned cam_mask;
  {
    const int lane = threadIdx.x & 31;
    const int pos_a = __shfl_sync(0xffffffffu, pos, 0), pos_b = __shfl_sync(0xffffffffu, pos, 31);
    const bool straight = (pos_b - pos_a == 31) && (pos_a / g.vX == pos_b / g.vX);
    unsigned code = 0u;
    if (lane < 4 * g.N) {
      const float ex = __ldg(t.xs + ((lane & 1) ? (pos_b % g.vX) : (pos_a % g.vX)));
      const float ez = __ldg(t.zs + ((lane & 2) ? (z_end - 1) : z_begin));
      code = cull_codes(lane >> 2, ex, py, ez);
    }
    code &= __shfl_xor_sync(0xffffffffu, code, 1);
    code &= __shfl_xor_sync(0xffffffffu, code, 2);
    const bool rejected = straight && (lane < 4 * g.N) && (code != 0u);
    const unsigned rej = __ballot_sync(0xffffffffu, rejected);   // all 4 lanes of a camera's quad agree
    cam_mask = 0u;
    for (int n = 0; n < g.N; ++n)
      if (!((rej >> (4 * n)) & 1u)) cam_mask |= 1u << n;
  }

  for (int z = z_begin; z < z_end; ++z) {
  const int vox = z * nplane + pos;
  const float pz = __ldg(t.zs + z);
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0f;
  int cams_seen = 0;
  uint64_t zero_cnt = 0;
  for (unsigned todo = cam_mask; todo != 0u; todo &= todo - 1u) {   // only the cameras that survived the warp cull
    const int n = __ffs(todo) - 1;
    if (cull_codes(n, px, py, pz) != 0u) continue;
    float pix[3];
    if (affine) project_voxel_affine(s_m + n * VB200_MAT_SLOTS * 16, has_bda, px, py, pz, pix);
    else project_voxel<false>(s_m + n * VB200_MAT_SLOTS * 16, has_bda, px, py, pz, pix);
    const LiftCoord lc = lift_coord<FASTDIV>(g, pix, &dv);
    if (!lc.valid) continue;  // f = grid_sample * 0: adds nothing to numer nor to the count
    const TriW w = tri_weights(lc.ix, lc.iy, lc.iz, lc.x0, lc.y0, lc.z0);
    const TD* dcam = depth + (size_t)(b * g.N + n) * g.D * HW;
    const TC* ccam = ctx_nhwc + (size_t)(b * g.N + n) * HW * C;
    float f[C];
    lift_pair_gather<TD, TC, C>(g, dcam, ccam, HW, lc.x0, lc.y0, lc.z0, w.wx0, w.wx1, w.wy0, w.wy1, w.wz0, w.wz1, f);
    lift_accumulate<C>(f, acc, cams_seen, zero_cnt);
  }

  if (!pos_live) continue;
  lift_store<TD, C, OUT_LAYOUT>(out, cnt_out, b, nvox, vox, acc, cams_seen, zero_cnt);
  }   // z-run
}

template <typename TD, typename TC>
int launch_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_depth, const void* d_ctx,
               void* d_out, int out_layout, uint64_t* d_cnt, void* ws, cudaStream_t st) {
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
  // z-run per thread: longer runs amortise the prologue, shorter ones keep the warp-level camera mask selective.
  // Measured on B200 (R50, B=8, bf16) with 128-thread blocks: 0.342 ms at 7 blocks/SM (72 registers) and 10
  // levels, 0.349 at 8 blocks/SM, 0.355 at 5 levels (256-thread blocks: 0.356 / 0.362); VB200_LIFT_ZRUN overrides
  const int plane_blocks = vb_ceil_div(g->vY * g->vX, kLiftThreads);
  int zrun = g->vZ < 10 ? g->vZ : 10;
  {
    const char* env = getenv("VB200_LIFT_ZRUN");
    if (env && atoi(env) > 0) zrun = atoi(env);
  }
  dim3 grid(plane_blocks, g->B, vb_ceil_div(g->vZ, zrun));
  VbTraceScope tr(VB_K_LIFT_FWD, st);
  const VbLiftDiv dv = vb_lift_div(g);
#define VB_LIFT(LAYOUT, FD)                                                                                   \
  lift_pool_fwd_kernel<TD, TC, C, LAYOUT, FD><<<grid, kLiftThreads, 0, st>>>(                                  \
      *g, *t, dv, d_mats, reinterpret_cast<const TD*>(d_depth), ctx_nhwc, reinterpret_cast<TD*>(d_out), d_cnt, zrun)
  if (vb_lift_div_ok(dv)) {
    if (out_layout == VB200_NCDHW) VB_LIFT(VB200_NCDHW, true);
    else VB_LIFT(VB200_NDHWC, true);
  } else {
    if (out_layout == VB200_NCDHW) VB_LIFT(VB200_NCDHW, false);
    else VB_LIFT(VB200_NDHWC, false);
  }
#undef VB_LIFT
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

}  // namespace

extern "C" size_t vb200_lift_pool_fwd_workspace(const VbGrid* g, int ctx_dtype) {
  if (!g) return 0;
  // channels-last copy of ctx (its own dtype), rounded up to 256 B
  const size_t n = (size_t)g->B * g->N * g->fH * g->fW * g->C * vb_lift_elem_size(ctx_dtype);
  return (n + 255) & ~(size_t)255;
}

extern "C" int vb200_lift_pool_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_depth,
                                   const void* d_ctx, int dtype, int ctx_dtype, void* d_out, int out_layout,
                                   uint64_t* d_cnt, void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && d_depth && d_ctx && d_out && d_workspace);
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS);
  VB_CHECK_ARG(g->D >= 1);
  VB_CHECK_ARG(out_layout == VB200_NCDHW || out_layout == VB200_NDHWC);
  if (workspace_bytes < vb200_lift_pool_fwd_workspace(g, ctx_dtype)) return VB200_ERR_WORKSPACE;
  if (((uintptr_t)d_out | (uintptr_t)d_ctx | (uintptr_t)d_depth) & 15) return VB200_ERR_ALIGN;
  if ((uintptr_t)d_workspace & 31) return VB200_ERR_ALIGN;   // the channels-last ctx copy is read with 256-bit loads
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
#define VB_CALL(TD, TC) launch_fwd<TD, TC>(g, t, d_mats, d_depth, d_ctx, d_out, out_layout, d_cnt, d_workspace, st)
  VB_LIFT_DISPATCH(dtype, ctx_dtype, VB_CALL);
#undef VB_CALL
}

// ---- backward: implemented in vb_lift_bwd.cu; plan-driven variants in vb_lift_plan.cu ------------
