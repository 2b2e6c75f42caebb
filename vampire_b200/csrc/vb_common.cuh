// vb_common.cuh -- shared device helpers for libvb200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/vb200.h"

#define VB_MAX_CAMS 8
#define VB_SM_COUNT_B200 148

#define VB_CHECK_ARG(cond) \
  do {                     \
    if (!(cond)) return VB200_ERR_ARG; \
  } while (0)

#define VB_LAUNCH_CHECK()                                   \
  do {                                                      \
    if (cudaGetLastError() != cudaSuccess) return VB200_ERR_CUDA; \
  } while (0)

static inline int vb_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// dtype helpers: features are stored as T, always widened to fp32 for arithmetic
// ---------------------------------------------------------------------------------------------
template <typename T> struct VbType;
template <> struct VbType<float> {
  static constexpr int code = VB200_F32;
  __device__ __forceinline__ static float ld(const float* p) { return __ldg(p); }
  __device__ __forceinline__ static float cvt(float v) { return v; }
};
template <> struct VbType<__nv_bfloat16> {
  static constexpr int code = VB200_BF16;
  __device__ __forceinline__ static float ld(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
  }
  __device__ __forceinline__ static __nv_bfloat16 cvt(float v) { return __float2bfloat16_rn(v); }
};
template <> struct VbType<__half> {
  static constexpr int code = VB200_F16;
  __device__ __forceinline__ static float ld(const __half* p) { return __half2float(__ldg(p)); }
  __device__ __forceinline__ static __half cvt(float v) { return __float2half_rn(v); }
};

// load NV consecutive elements of T starting at p (16-byte aligned group), widened to fp32
template <typename T, int NV> struct VbVec;
template <> struct VbVec<float, 4> {
  __device__ __forceinline__ static void ld(const float* p, float* o) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct VbVec<__nv_bfloat16, 8> {
  __device__ __forceinline__ static void ld(const __nv_bfloat16* p, float* o) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct VbVec<__half, 8> {
  __device__ __forceinline__ static void ld(const __half* p, float* o) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 f = __half22float2(h);
      o[2 * i] = f.x;
      o[2 * i + 1] = f.y;
    }
  }
};
template <typename T> struct VbLanes { static constexpr int n = 16 / sizeof(T); };  // elements per 128-bit

// ---------------------------------------------------------------------------------------------
// strict fp32 zone: everything that feeds floor() or a validity compare is written with the
// round-to-nearest intrinsics so that nvcc can never contract a*b+c into an FMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float smul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ssub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float sdiv(float a, float b) { return __fdiv_rn(a, b); }

// row-major 4x4 (16 floats at M) times 4-vector, ATen native bmm order:
// acc = 0; acc += M[i][k] * p[k] for k = 0..3, separate multiply / add roundings.
// SIGNED_ZERO=false drops the leading "0 +" (it only turns a -0 product into +0): value-identical
// for every compare / floor downstream, used by the fused kernels that never output coordinates.
template <bool SIGNED_ZERO = true>
__device__ __forceinline__ void mv_strict(const float* __restrict__ M, const float (&p)[4], float (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float acc = SIGNED_ZERO ? sadd(0.0f, smul(M[i * 4 + 0], p[0])) : smul(M[i * 4 + 0], p[0]);
    acc = sadd(acc, smul(M[i * 4 + 1], p[1]));
    acc = sadd(acc, smul(M[i * 4 + 2], p[2]));
    acc = sadd(acc, smul(M[i * 4 + 3], p[3]));
    r[i] = acc;
  }
}

// G1 get_pixel for one voxel centre (BV2:367-388). M = this camera's 6 prepared matrices.
template <bool SIGNED_ZERO = true>
__device__ __forceinline__ void project_voxel(const float* __restrict__ M, bool has_bda, float x, float y,
                                              float z, float (&pix)[3]) {
  float p[4] = {x, y, z, 1.0f}, q[4];
  if (has_bda) {
    mv_strict<SIGNED_ZERO>(M + 0 * 16, p, q);  // bda^-1                        BV2:374
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = q[i];
  }
  mv_strict<SIGNED_ZERO>(M + 1 * 16, p, q);    // K . E^-1                      BV2:380
  const float zc = q[2] < 1e-6f ? 1e-6f : q[2];  // torch.clamp(min=eps), NaN-propagating   BV2:385
  p[0] = sdiv(q[0], zc);
  p[1] = sdiv(q[1], zc);
  p[2] = q[2];
  p[3] = q[3];
  mv_strict<SIGNED_ZERO>(M + 2 * 16, p, q);    // ida                           BV2:387
  pix[0] = q[0]; pix[1] = q[1]; pix[2] = q[2];
}

// G2 get_geometry for one frustum lattice point (BV2:332-349).
template <bool SIGNED_ZERO = true>
__device__ __forceinline__ void frustum_point(const float* __restrict__ M, bool has_bda, float u, float v,
                                              float d, float (&xyz)[3]) {
  float p[4] = {u, v, d, 1.0f}, q[4];
  mv_strict<SIGNED_ZERO>(M + 3 * 16, p, q);    // ida^-1                        BV2:334
  p[0] = smul(q[0], q[2]);        //                                            BV2:336-338
  p[1] = smul(q[1], q[2]);
  p[2] = q[2];
  p[3] = q[3];
  mv_strict<SIGNED_ZERO>(M + 4 * 16, p, q);    // E . K^-1                      BV2:341-342
  if (has_bda) {
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = q[i];
    mv_strict<SIGNED_ZERO>(M + 5 * 16, p, q);  // bda                           BV2:346
  }
  xyz[0] = q[0]; xyz[1] = q[1]; xyz[2] = q[2];
}

// torch.nan_to_num(x, nan) : nan -> `nan`, +-inf -> +-FLT_MAX
__device__ __forceinline__ float nan_to_num(float x, float nanv) {
  if (isnan(x)) return nanv;
  if (isinf(x)) return x > 0.0f ? 3.402823466e+38f : -3.402823466e+38f;
  return x;
}

// L2: validity + normalised/clamped coords + ATen unnormalise (align_corners=False)  BV2:493-507
struct LiftCoord {
  bool valid;
  int x0, y0, z0;
  float ix, iy, iz;  // unnormalised sample position
};
__device__ __forceinline__ LiftCoord lift_coord(const VbGrid& g, const float (&pix)[3]) {
  LiftCoord c;
  const float x = pix[0], y = pix[1], z = pix[2];
  c.valid = (x > -0.5f) && (x < g.x_hi) && (y > -0.5f) && (y < g.y_hi) && (z > g.d_lo) && (z < g.d_hi);
  float nx = ssub(smul(2.0f, sdiv(x, g.img_w_m1)), 1.0f);
  float ny = ssub(smul(2.0f, sdiv(y, g.img_h_m1)), 1.0f);
  float nz = ssub(smul(2.0f, sdiv(ssub(z, g.d_lo), g.d_ext)), 1.0f);
  // torch.clamp(min=-2, max=2) propagates NaN; fminf/fmaxf would not, so select explicitly
  nx = nx < -2.0f ? -2.0f : (nx > 2.0f ? 2.0f : nx);
  ny = ny < -2.0f ? -2.0f : (ny > 2.0f ? 2.0f : ny);
  nz = nz < -2.0f ? -2.0f : (nz > 2.0f ? 2.0f : nz);
  // ATen divides by 2; x * 0.5f is bit-identical (power of two) and avoids an IEEE-division sequence
  c.ix = smul(ssub(smul(sadd(nx, 1.0f), (float)g.fW), 1.0f), 0.5f);
  c.iy = smul(ssub(smul(sadd(ny, 1.0f), (float)g.fH), 1.0f), 0.5f);
  c.iz = smul(ssub(smul(sadd(nz, 1.0f), (float)g.D), 1.0f), 0.5f);
  c.x0 = (int)floorf(c.ix);
  c.y0 = (int)floorf(c.iy);
  c.z0 = (int)floorf(c.iz);
  return c;
}

// R2: normalise + inclusive mask + ATen unnormalise (align_corners=True)  BV2:397-407, 419
struct RenderCoord {
  bool valid;
  int x0, y0, z0;
  float ix, iy, iz;
};
__device__ __forceinline__ RenderCoord render_coord(const VbGrid& g, const float (&p)[3]) {
  RenderCoord c;
  const float gx = ssub(smul(sdiv(ssub(p[0], g.seg_lo[0]), g.seg_ext[0]), 2.0f), 1.0f);
  const float gy = ssub(smul(sdiv(ssub(p[1], g.seg_lo[1]), g.seg_ext[1]), 2.0f), 1.0f);
  const float gz = ssub(smul(sdiv(ssub(p[2], g.seg_lo[2]), g.seg_ext[2]), 2.0f), 1.0f);
  c.valid = (gx >= -1.0f) && (gx <= 1.0f) && (gy >= -1.0f) && (gy <= 1.0f) && (gz >= -1.0f) && (gz <= 1.0f);
  c.ix = smul(smul(sadd(gx, 1.0f), 0.5f), (float)(g.vX - 1));   // (g + 1) / 2 == (g + 1) * 0.5f exactly
  c.iy = smul(smul(sadd(gy, 1.0f), 0.5f), (float)(g.vY - 1));
  c.iz = smul(smul(sadd(gz, 1.0f), 0.5f), (float)(g.vZ - 1));
  // floorf of a huge / non-finite value is only ever used when !valid; clamp so the int cast is defined
  const float fx = floorf(c.ix), fy = floorf(c.iy), fz = floorf(c.iz);
  c.x0 = c.valid ? (int)fx : 0;
  c.y0 = c.valid ? (int)fy : 0;
  c.z0 = c.valid ? (int)fz : 0;
  return c;
}

// T4 ModifyLaplaceDensity (render_utils.py:37-42); beta = |beta_param| + beta_min
// expm1(t) is formed as expf(t) - 1: the cancellation near t = 0 costs < 6e-8 absolute on a term
// that is added to 0.5, i.e. < 1.2e-7 relative on sigma -- below fp32 resolution of the reference.
__device__ __forceinline__ float laplace_density(float s, float bias, float beta) {
  const float x = s - bias;
  const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
  return (1.0f / beta) * (0.5f + 0.5f * sgn * (expf(-fabsf(x) / beta) - 1.0f));
}

// true iff the 4x4 at M is exactly the identity.  mv(I, p) == p for every finite p, so the fused
// kernels skip the two bda products then (the reference's default bda_aug_conf IS the identity,
// base_exp.py:113-120) without changing a single compare or floor.  Block-uniform.
__device__ __forceinline__ bool block_is_identity(const float* M) {
  const int i = threadIdx.x;
  return __syncthreads_and(i >= 16 || M[i] == ((i % 5 == 0) ? 1.0f : 0.0f)) != 0;
}

__device__ __forceinline__ void stage_mats(float* s_m, const float* __restrict__ d_mats, int b, int N) {
  const float* src = d_mats + (size_t)b * N * VB200_MAT_SLOTS * 16;
  for (int i = threadIdx.x; i < N * VB200_MAT_SLOTS * 16; i += blockDim.x) s_m[i] = __ldg(src + i);
}
