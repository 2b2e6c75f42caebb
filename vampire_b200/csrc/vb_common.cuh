// vb_common.cuh -- shared device helpers for libvb200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

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

// cudaFuncSetAttribute applies to the CURRENT device only, so "done once" must be remembered per device (a
// process-wide flag would leave the attribute unset on every device but the first one to launch).
// Usage: `static VbPerDeviceFlag flag;` next to the launch (one per kernel instantiation).
struct VbPerDeviceFlag {
  volatile unsigned char done[64];
};
template <typename F>
static inline int vb_func_attr_per_device(F fn, cudaFuncAttribute attr, int value, VbPerDeviceFlag& flag) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VB200_ERR_CUDA;
  const bool tracked = dev >= 0 && dev < 64;
  if (!tracked || !flag.done[dev]) {           // idempotent call, so a race between host threads is benign
    if (cudaFuncSetAttribute(fn, attr, value) != cudaSuccess) return VB200_ERR_CUDA;
    if (tracked) flag.done[dev] = 1;
  }
  return VB200_OK;
}

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

// 256-bit global load (sm_100: LDG.E.256): p must be 32-byte aligned.  The gather kernels are bound by L1 data-pipe
// wavefronts, which are paid per instruction and per distinct 128-byte line -- one 32-byte load costs what a
// 16-byte load costs, so a 32-byte pixel / record chunk should be ONE instruction.
struct __align__(32) VbU8 { uint32_t v[8]; };
__device__ __forceinline__ VbU8 vb_ldg256(const void* p) {
  VbU8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
// widen 32 bytes of T (8 fp32 / 16 bf16 / 16 fp16) to fp32
template <typename T> struct VbWiden32;
template <> struct VbWiden32<float> {
  static constexpr int n = 8;
  __device__ __forceinline__ static void cvt(const VbU8& r, float* o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(r.v[i]);
  }
};
template <> struct VbWiden32<__nv_bfloat16> {
  static constexpr int n = 16;
  __device__ __forceinline__ static void cvt(const VbU8& r, float* o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[2 * i] = __uint_as_float(r.v[i] << 16);
      o[2 * i + 1] = __uint_as_float(r.v[i] & 0xffff0000u);
    }
  }
};
template <> struct VbWiden32<__half> {
  static constexpr int n = 16;
  __device__ __forceinline__ static void cvt(const VbU8& r, float* o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r.v[i]));
      o[2 * i] = f.x;
      o[2 * i + 1] = f.y;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// strict fp32 zone: everything that feeds floor() or a validity compare is written with the
// round-to-nearest intrinsics so that nvcc can never contract a*b+c into an FMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float smul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ssub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float sdiv(float a, float b) { return __fdiv_rn(a, b); }

// x / y for a launch-constant divisor y with r = RN(1/y) precomputed on the host: Markstein's
// two-step FMA correction (the same sequence __fdiv_rn runs after refining its own reciprocal estimate;
// with an exactly rounded r and a faithful q1 the result is the correctly rounded quotient, Markstein 1990 /
// Handbook of Floating-Point Arithmetic thm. 4.x).  Preconditions, checked on the host (vb_div_const):
// y normal, finite, mantissa not all ones.  x must be 0, NaN, or |x| in [2^-100, FLT_MAX]; the callers'
// operands are differences against / multiples of O(1..1e3) lattice constants.
struct VbDivConst {
  float y, r;
  int ok;      // 0: fall back to __fdiv_rn
};
static inline VbDivConst vb_div_const(float y) {
  VbDivConst d;
  d.y = y;
  d.r = 1.0f / y;
  uint32_t bits;
  memcpy(&bits, &y, sizeof(bits));
  const uint32_t ex = (bits >> 23) & 0xffu, man = bits & 0x7fffffu;
  d.ok = (ex >= 32u && ex <= 222u && man != 0x7fffffu) ? 1 : 0;   // |y| in [2^-95, 2^95], not 2^k (2 - 2^-23)
  return d;
}
template <bool FAST>
__device__ __forceinline__ float sdiv_const(float x, const VbDivConst& d) {
  if (!FAST) return __fdiv_rn(x, d.y);
  const float q0 = __fmul_rn(x, d.r);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, d.y, x), d.r, q0);
  return __fmaf_rn(__fmaf_rn(-q1, d.y, x), d.r, q1);
}

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

// The same collapse for get_pixel: with ida = 2-D affine (rows 2, 3 = e_z, e_w, zero z-column in rows 0, 1)
// and homogeneous last rows e_w in K.E^-1 (and bda^-1), the w component is exactly 1 and the third mat-vec is
// pix = ((I00 u + I01 v) + I03, (I10 u + I11 v) + I13, z): every dropped term is a +-0 product or a x*1.
// (A non-finite z would differ -- NaN instead of a finite pixel -- but then z itself fails the depth test.)
// Block-uniform test over all N cameras; call from every thread of the block.
__device__ __forceinline__ bool block_pixel_affine(const float* s_m, int N, bool has_bda) {
  bool ok = true;
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    const float* M = s_m + (i >> 4) * VB200_MAT_SLOTS * 16;
    const int e = i & 15;
    const float ida = M[2 * 16 + e], ke = M[1 * 16 + e], bi = M[e];
    if (e == 2 || e == 6) ok = ok && (ida == 0.0f);
    if (e >= 8) ok = ok && (ida == ((e == 10 || e == 15) ? 1.0f : 0.0f));
    if (e >= 12) ok = ok && (ke == (e == 15 ? 1.0f : 0.0f)) && (!has_bda || bi == (e == 15 ? 1.0f : 0.0f));
  }
  return __syncthreads_and(ok) != 0;
}
__device__ __forceinline__ void project_voxel_affine(const float* __restrict__ M, bool has_bda, float x, float y,
                                                     float z, float (&pix)[3]) {
  float p[3] = {x, y, z};
  if (has_bda) {
    const float* B = M;
    float q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      q[i] = sadd(sadd(sadd(smul(B[i * 4], p[0]), smul(B[i * 4 + 1], p[1])), smul(B[i * 4 + 2], p[2])), B[i * 4 + 3]);
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = q[i];
  }
  const float* E = M + 16;
  float q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    q[i] = sadd(sadd(sadd(smul(E[i * 4], p[0]), smul(E[i * 4 + 1], p[1])), smul(E[i * 4 + 2], p[2])), E[i * 4 + 3]);
  const float zc = q[2] < 1e-6f ? 1e-6f : q[2];
  const float u = sdiv(q[0], zc), v = sdiv(q[1], zc);
  const float* I = M + 32;
  pix[0] = sadd(sadd(smul(I[0], u), smul(I[1], v)), I[3]);
  pix[1] = sadd(sadd(smul(I[4], u), smul(I[5], v)), I[7]);
  pix[2] = q[2];
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

// ida^-1 (slot 3) as the dataset builds it (nusc_det_seg_dataset.py:118-146: a 2-D rotation/scale/flip plus a
// translation in column 3) has rows 2, 3 = e_z, e_w and a zero z-column in rows 0, 1.  Then, with finite
// lattice values, the first mat-vec of get_geometry collapses EXACTLY (every dropped term is a +-0 product
// or a x*1): q = (A0, A1, d, 1) with A_i = (m_i0 u + m_i1 v) + m_i3 constant along the ray.  Block-uniform.
__device__ __forceinline__ bool block_ida_inv_affine(const float* M) {
  const int i = threadIdx.x;
  bool ok = true;
  if (i < 16) {
    const float m = M[i];
    if (i == 2 || i == 6) ok = (m == 0.0f);
    else if (i >= 8) ok = (m == ((i == 10 || i == 15) ? 1.0f : 0.0f));
  }
  return __syncthreads_and(ok) != 0;
}
// ray constants (A0, A1) of the affine case
__device__ __forceinline__ void frustum_ray_affine(const float* __restrict__ M, float u, float v, float (&A)[2]) {
  const float* I = M + 3 * 16;
  A[0] = sadd(sadd(smul(I[0], u), smul(I[1], v)), I[3]);
  A[1] = sadd(sadd(smul(I[4], u), smul(I[5], v)), I[7]);
}
// G2 for the affine case: identical values to frustum_point<false>
__device__ __forceinline__ void frustum_point_affine(const float* __restrict__ M, bool has_bda, const float (&A)[2],
                                                     float d, float (&xyz)[3]) {
  const float p0 = smul(A[0], d), p1 = smul(A[1], d);
  const float* E = M + 4 * 16;
  float q[4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    q[i] = sadd(sadd(sadd(smul(E[i * 4 + 0], p0), smul(E[i * 4 + 1], p1)), smul(E[i * 4 + 2], d)), E[i * 4 + 3]);
  if (has_bda) {
    q[3] = sadd(sadd(sadd(smul(E[12], p0), smul(E[13], p1)), smul(E[14], d)), E[15]);
    float r[4];
    mv_strict<false>(M + 5 * 16, q, r);
    xyz[0] = r[0]; xyz[1] = r[1]; xyz[2] = r[2];
  } else {
    xyz[0] = q[0]; xyz[1] = q[1]; xyz[2] = q[2];
  }
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
struct VbLiftDiv { VbDivConst a[3]; };   // divisions by img_w_m1, img_h_m1, d_ext
static inline VbLiftDiv vb_lift_div(const VbGrid* g) {
  VbLiftDiv d;
  d.a[0] = vb_div_const(g->img_w_m1);
  d.a[1] = vb_div_const(g->img_h_m1);
  d.a[2] = vb_div_const(g->d_ext);
  return d;
}
static inline bool vb_lift_div_ok(const VbLiftDiv& d) { return d.a[0].ok && d.a[1].ok && d.a[2].ok; }

// FAST: only for operands that passed the conservative frustum cull of the fused kernels (finite, O(1e3)
// pixel coordinates or NaN) -- see sdiv_const.  A |quotient| < 2^-100 may be mis-rounded, but it is then
// absorbed by the "- 1" that follows.
template <bool FAST = false>
__device__ __forceinline__ LiftCoord lift_coord(const VbGrid& g, const float (&pix)[3], const VbLiftDiv* dv = nullptr) {
  LiftCoord c;
  const float x = pix[0], y = pix[1], z = pix[2];
  c.valid = (x > -0.5f) && (x < g.x_hi) && (y > -0.5f) && (y < g.y_hi) && (z > g.d_lo) && (z < g.d_hi);
  const float qx = FAST ? sdiv_const<true>(x, dv->a[0]) : sdiv(x, g.img_w_m1);
  const float qy = FAST ? sdiv_const<true>(y, dv->a[1]) : sdiv(y, g.img_h_m1);
  const float qz = FAST ? sdiv_const<true>(ssub(z, g.d_lo), dv->a[2]) : sdiv(ssub(z, g.d_lo), g.d_ext);
  float nx = ssub(smul(2.0f, qx), 1.0f);
  float ny = ssub(smul(2.0f, qy), 1.0f);
  float nz = ssub(smul(2.0f, qz), 1.0f);
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
  if (g.D == 1) {
    // BaseBiLinear (base_bilinear.py:484, 505-507): z_valid = z > 0, and the sampled volume has ONE depth plane
    // addressed at normalised z = 0, i.e. iz = ((0 + 1) * 1 - 1) / 2 = 0: plane 0 with weight 1 (launch-uniform)
    c.valid = (x > -0.5f) && (x < g.x_hi) && (y > -0.5f) && (y < g.y_hi) && (z > 0.0f);
    c.iz = 0.0f;
    c.z0 = 0;
  }
  return c;
}

// strict projection + sampling coordinates of one (voxel, camera) pair, with the exact shortcuts where the
// block-uniform flags allow them (affine ida / homogeneous last rows; launch-constant divisors).  Only for
// pairs that passed a conservative frustum cull (see lift_coord<FAST>).  Value-identical on every path.
__device__ __forceinline__ LiftCoord pair_strict(const VbGrid& g, const float* __restrict__ Mcam, bool has_bda,
                                                 bool affine, const VbLiftDiv& dv, float px, float py, float pz) {
  float pix[3];
  if (affine) project_voxel_affine(Mcam, has_bda, px, py, pz, pix);
  else project_voxel<false>(Mcam, has_bda, px, py, pz, pix);
  const bool fast = dv.a[0].ok && dv.a[1].ok && dv.a[2].ok;
  return fast ? lift_coord<true>(g, pix, &dv) : lift_coord<false>(g, pix);
}

// R2: normalise + inclusive mask + ATen unnormalise (align_corners=True)  BV2:397-407, 419
struct RenderCoord {
  bool valid;
  int x0, y0, z0;
  float ix, iy, iz;
};
struct VbRenderDiv { VbDivConst a[3]; };   // divisions by seg_ext[0..2]
static inline VbRenderDiv vb_render_div(const VbGrid* g) {
  VbRenderDiv d;
  for (int a = 0; a < 3; ++a) d.a[a] = vb_div_const(g->seg_ext[a]);
  return d;
}
__host__ __device__ static inline bool vb_render_div_ok(const VbRenderDiv& d) { return d.a[0].ok && d.a[1].ok && d.a[2].ok; }

template <bool FAST = false>
__device__ __forceinline__ RenderCoord render_coord(const VbGrid& g, const float (&p)[3],
                                                    const VbRenderDiv* dv = nullptr) {
  RenderCoord c;
  const float qx = FAST ? sdiv_const<true>(ssub(p[0], g.seg_lo[0]), dv->a[0]) : sdiv(ssub(p[0], g.seg_lo[0]), g.seg_ext[0]);
  const float qy = FAST ? sdiv_const<true>(ssub(p[1], g.seg_lo[1]), dv->a[1]) : sdiv(ssub(p[1], g.seg_lo[1]), g.seg_ext[1]);
  const float qz = FAST ? sdiv_const<true>(ssub(p[2], g.seg_lo[2]), dv->a[2]) : sdiv(ssub(p[2], g.seg_lo[2]), g.seg_ext[2]);
  const float gx = ssub(smul(qx, 2.0f), 1.0f);
  const float gy = ssub(smul(qy, 2.0f), 1.0f);
  const float gz = ssub(smul(qz, 2.0f), 1.0f);
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

// same with a per-launch reciprocal of beta: |x| * (1/beta) differs from |x| / beta by one rounding (2^-24
// relative on the exponent t), i.e. by <= t e^-t 2^-24 <= 2.2e-8 absolute on a term that is added to 0.5.
__device__ __forceinline__ float laplace_density_rcp(float s, float bias, float inv_beta) {
  const float x = s - bias;
  const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
  return inv_beta * (0.5f + 0.5f * sgn * (expf(-fabsf(x) * inv_beta) - 1.0f));
}

// density_mode = 'naive': self.density = nn.Sigmoid() (BV2:191-192); ATen's fp32 form 1 / (1 + exp(-x))
__device__ __forceinline__ float sigmoid_density(float s) { return 1.0f / (1.0f + expf(-s)); }

// sigma of a density feature under the grid's density mode (the branch is launch-uniform)
__device__ __forceinline__ float vb_density(const VbGrid& g, float s, float beta) {
  if (g.density_mode == VB200_DENSITY_NAIVE) return sigmoid_density(s);
  return laplace_density(s, g.sdf_bias, beta);
}
__device__ __forceinline__ float vb_density_rcp(const VbGrid& g, float s, float inv_beta) {
  if (g.density_mode == VB200_DENSITY_NAIVE) return sigmoid_density(s);
  return laplace_density_rcp(s, g.sdf_bias, inv_beta);
}

// sigma with its derivatives (backward kernels).  'sdf' (render_utils.py:30-46, x = s - bias):
//   dsigma/ds = -e^{-|x|/beta} / (2 beta^2) (0 at x = 0),  dsigma/dbeta = -sigma/beta + x e^{-|x|/beta} / (2 beta^3);
// 'naive': dsigma/ds = sigma (1 - sigma), no beta.
struct DensityD {
  float sigma, ds, dbeta;
};
__device__ __forceinline__ DensityD vb_density_with_grads(const VbGrid& g, float s, float beta) {
  DensityD d;
  if (g.density_mode == VB200_DENSITY_NAIVE) {
    d.sigma = sigmoid_density(s);
    d.ds = d.sigma * (1.0f - d.sigma);
    d.dbeta = 0.0f;
    return d;
  }
  const float x = s - g.sdf_bias, ax = fabsf(x);
  const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
  const float e = expf(-ax / beta);
  d.sigma = (1.0f / beta) * (0.5f + 0.5f * sgn * (e - 1.0f));
  d.ds = (x != 0.0f) ? -e / (2.0f * beta * beta) : 0.0f;
  d.dbeta = -d.sigma / beta + x * e / (2.0f * beta * beta * beta);
  return d;
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
