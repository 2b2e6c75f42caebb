// vb_lift_common.cuh -- device code shared by the fused lift+pool forward (vb_lift.cu: projection recomputed
// per call) and its plan-driven twin (vb_lift_plan.cu: projection read from a cached plan).
//
// Feature dtypes: TD = dtype of the depth distribution AND of the pooled volume, TC = dtype of the context
// features.  TD == TC is the plain case; TD = float with a 16-bit TC is the reference under AMP, where
// softmax is autocast to fp32 and `depth.unsqueeze(2) * ctx.unsqueeze(3)` (BV2:553) promotes to fp32, so the
// frustum, the grid_sample and the pooled volume are fp32 there.
#pragma once
#include "vb_common.cuh"

namespace {

// ---- ctx (B,N,C,fH,fW) -> (B,N,fH,fW,C): one block per (b*n, h) row -------------------------
// (the copy keeps the feature dtype: the gather is L1-wavefront bound -- ncu: l1tex 87 % with an fp32 copy --
//  so a bf16 pixel = one 32-byte sector beats saving the 16 unpack instructions of an fp32 copy)
template <typename T, int C>
__global__ void __launch_bounds__(256) ctx_to_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int fH,
                                                          int fW) {
  extern __shared__ unsigned char s_raw[];
  T* s = reinterpret_cast<T*>(s_raw);  // [C][fW + 1]
  const int h = blockIdx.x, bn = blockIdx.y;
  const int ld = fW + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // read: a warp per channel row, lanes along w (coalesced, no integer division)
  for (int c = wid; c < C; c += nw) {
    const T* row = src + (((size_t)bn * C + c) * fH + h) * fW;
    for (int w = lane; w < fW; w += 32) s[c * ld + w] = row[w];
  }
  __syncthreads();
  // write: one 128-bit store per thread = 16 / sizeof(T) consecutive channels of one pixel
  constexpr int L = 16 / sizeof(T), PARTS = C / L;
  static_assert(C % L == 0, "a pixel's channels must be whole 128-bit groups");
  uint4* out = reinterpret_cast<uint4*>(dst + ((size_t)bn * fH + h) * fW * C);
  for (int p = threadIdx.x; p < fW * PARTS; p += blockDim.x) {
    const int w = p / PARTS, c0 = (p % PARTS) * L;   // PARTS is a compile-time power of two
    __align__(16) T v[L];
#pragma unroll
    for (int e = 0; e < L; ++e) v[e] = s[(c0 + e) * ld + w];
    out[p] = *reinterpret_cast<const uint4*>(v);
  }
}

// ---- one (voxel, camera) pair: f[c] = sum_{4 pixels} w_jk ctx[c,j,k] * (sum_{2 depth bins} w_i depth[i,j,k]) ----
// (x0, y0, z0) = base corner (valid => each in [-1, size-1]); w?0 / w?1 = near / far weights as ATen forms them.
// Zeros padding without branches: clamp the address, zero the weight.
template <typename TD, typename TC, int C>
__device__ __forceinline__ void lift_pair_gather(const VbGrid& g, const TD* __restrict__ dcam,
                                                 const TC* __restrict__ ccam, int HW, int x0, int y0, int z0,
                                                 float wx0, float wx1, float wy0, float wy1, float wz0, float wz1,
                                                 float (&f)[C]) {
  const int xa = max(x0, 0), xb = min(x0 + 1, g.fW - 1);
  const int ya = max(y0, 0), yb = min(y0 + 1, g.fH - 1);
  const int za = max(z0, 0), zb = min(z0 + 1, g.D - 1);
  const float wxa = x0 >= 0 ? wx0 : 0.0f, wxb = x0 + 1 < g.fW ? wx1 : 0.0f;
  const float wya = y0 >= 0 ? wy0 : 0.0f, wyb = y0 + 1 < g.fH ? wy1 : 0.0f;
  const float wza = z0 >= 0 ? wz0 : 0.0f, wzb = z0 + 1 < g.D ? wz1 : 0.0f;
  const int pxl[4] = {ya * g.fW + xa, ya * g.fW + xb, yb * g.fW + xa, yb * g.fW + xb};
  const float wxy[4] = {wxa * wya, wxb * wya, wxa * wyb, wxb * wyb};
  const TD* d0 = dcam + za * HW;
  const TD* d1 = dcam + zb * HW;
  float wgt[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    wgt[k] = wxy[k] * fmaf(wzb, VbType<TD>::ld(d1 + pxl[k]), wza * VbType<TD>::ld(d0 + pxl[k]));
#pragma unroll
  for (int c = 0; c < C; ++c) f[c] = 0.0f;
  // a pixel's C channels are C * sizeof(TC) contiguous, 32-byte aligned bytes: 256-bit loads (one per bf16 pixel)
  constexpr int L = VbWiden32<TC>::n;
  static_assert(C % L == 0, "a pixel's channels must be whole 256-bit groups");
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const TC* cp = ccam + pxl[k] * C;
#pragma unroll
    for (int q8 = 0; q8 < C / L; ++q8) {
      float cv[L];
      VbWiden32<TC>::cvt(vb_ldg256(cp + q8 * L), cv);
#pragma unroll
      for (int e = 0; e < L; ++e) f[q8 * L + e] = fmaf(cv[e], wgt[k], f[q8 * L + e]);
    }
  }
}

// accumulate one camera's f into the voxel's sums + the per-channel non-zero bookkeeping (BV2:509-512).
// A seeing camera almost always contributes to all C channels, so the common case is one shared counter;
// exact zeros (dead ctx channel, both depth bins outside) take the packed per-channel path: 4 bits per
// channel, count of cameras that were ZERO.
template <int C>
__device__ __forceinline__ void lift_accumulate(const float (&f)[C], float (&acc)[C], int& cams_seen,
                                                uint64_t& zero_cnt) {
  float fmin_abs = fabsf(f[0]);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    acc[c] += f[c];
    fmin_abs = fminf(fmin_abs, fabsf(f[c]));
  }
  cams_seen += 1;
  if (!(fmin_abs > 0.0f)) {   // some channel is exactly 0 (or NaN): voxel_mask = |f| > 0  BV2:509
#pragma unroll
    for (int c = 0; c < C; ++c) zero_cnt += (uint64_t)(fabsf(f[c]) > 0.0f ? 0 : 1) << (4 * c);
  }
}

// mean = numer / (count + 1e-6)  (BV2:512-514) and the store; reciprocal-multiply is within 2 ulp of the division
template <typename TO, int C, int OUT_LAYOUT>
__device__ __forceinline__ void lift_store(TO* __restrict__ out, uint64_t* __restrict__ cnt_out, int b, int nvox,
                                           int vox, const float (&acc)[C], int cams_seen, uint64_t zero_cnt) {
  static_assert(C <= 16, "per-channel counts are packed 4 bits each into 64 bits");
  const uint64_t seen_all = 0x1111111111111111ull * (uint64_t)cams_seen;   // cams_seen in every 4-bit field
  if (cnt_out) cnt_out[(size_t)b * nvox + vox] = seen_all - zero_cnt;      // saved for the backward
  float inv[C];
  if (zero_cnt == 0) {
    const float r = __fdividef(1.0f, (float)cams_seen + 1e-6f);
#pragma unroll
    for (int c = 0; c < C; ++c) inv[c] = r;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c)
      inv[c] = __fdividef(1.0f, (float)(cams_seen - (int)((zero_cnt >> (4 * c)) & 0xf)) + 1e-6f);
  }
  if (OUT_LAYOUT == VB200_NCDHW) {
    TO* o = out + (size_t)b * C * nvox + vox;
#pragma unroll
    for (int c = 0; c < C; ++c) o[(size_t)c * nvox] = VbType<TO>::cvt(acc[c] * inv[c]);
  } else {
    TO* o = out + ((size_t)b * nvox + vox) * C;
    __align__(16) TO v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = VbType<TO>::cvt(acc[c] * inv[c]);
    constexpr int L = VbLanes<TO>::n;
#pragma unroll
    for (int q = 0; q < C / L; ++q) reinterpret_cast<uint4*>(o)[q] = reinterpret_cast<const uint4*>(v)[q];
  }
}

inline size_t vb_lift_elem_size(int dtype) { return dtype == VB200_F32 ? 4 : 2; }

// dispatch over the supported (depth/out dtype, ctx dtype) pairs: equal, or fp32 depth with 16-bit ctx (AMP)
#define VB_LIFT_DISPATCH(depth_dtype, ctx_dtype, CALL)                                         \
  do {                                                                                         \
    if ((depth_dtype) == VB200_F32 && (ctx_dtype) == VB200_F32) return CALL(float, float);     \
    if ((depth_dtype) == VB200_BF16 && (ctx_dtype) == VB200_BF16) return CALL(__nv_bfloat16, __nv_bfloat16); \
    if ((depth_dtype) == VB200_F16 && (ctx_dtype) == VB200_F16) return CALL(__half, __half);   \
    if ((depth_dtype) == VB200_F32 && (ctx_dtype) == VB200_BF16) return CALL(float, __nv_bfloat16); \
    if ((depth_dtype) == VB200_F32 && (ctx_dtype) == VB200_F16) return CALL(float, __half);    \
    return VB200_ERR_DTYPE;                                                                    \
  } while (0)

}  // namespace
