// vb_march_staged.cuh -- north-star kernel (c) as written: the camera march with its voxel tiles STAGED in shared
// memory by bulk asynchronous copies (cp.async.bulk + mbarrier complete_tx, the TMA engine's 1-D form).
//
// The cached render plan gives, for every (warp, sample), the axis-aligned box of voxels that covers all 8 trilinear
// corners of the warp's in-volume rays (tools/tma_box_analysis.py: median 32 voxels = 1.5 KB, 1.38x the voxels really
// touched; 89 % of the boxes are <= 96 voxels).  Two samples ahead of the march each lane of the warp issues ONE bulk
// copy -- one x-row of the box, nx * 48 bytes -- into the warp's private two-stage ring; the copies complete on the
// stage's mbarrier while the warp composites the samples in between, and the gathers of the sample (8 density scalars,
// then 24 x 128-bit value loads where the weight is non-zero) read shared memory at a base + the record's in-box
// index.  Boxes over the cap (far range, rays spread over many voxels) take the global-load path of
// march_fwd_planned_kernel for that sample.  Same arithmetic in the same order as march_fwd_planned_kernel:
// bit-identical outputs (tests/test_gpu_plan.py).
//
// 16-bit features only (a 96-voxel stage is 4.6 KB; 4 warps x 2 stages = 37 KB per block, 5 blocks per SM).
#pragma once
#include "vb_march_planned.cuh"

namespace {

__device__ __forceinline__ uint32_t vs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vs_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void vs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vs_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool vs_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void vs_mbar_wait(uint32_t bar, uint32_t parity) {
  if (vs_mbar_try(bar, parity)) return;          // the copies were issued two samples ago: the common case
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
__device__ __forceinline__ uint4 vs_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t vs_lds16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
template <typename T> __device__ __forceinline__ float vs_widen16(uint32_t bits);
template <> __device__ __forceinline__ float vs_widen16<__nv_bfloat16>(uint32_t bits) { return __uint_as_float(bits << 16); }
template <> __device__ __forceinline__ float vs_widen16<__half>(uint32_t bits) {
  return __half2float(__ushort_as_half((unsigned short)bits));
}
template <typename T> __device__ __forceinline__ void vs_widen8(const uint4& v, float* o);
template <> __device__ __forceinline__ void vs_widen8<__nv_bfloat16>(const uint4& v, float* o) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <> __device__ __forceinline__ void vs_widen8<__half>(const uint4& v, float* o) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}

#ifndef VB_MARCH_STAGED_MINB
#define VB_MARCH_STAGED_MINB 5
#endif
template <typename T, int K>
__global__ void __launch_bounds__(kMarchThreads, VB_MARCH_STAGED_MINB) march_fwd_staged_kernel(
    VbGrid g, VbTables t, const VbRenderPlan* __restrict__ plans, const T* __restrict__ packed,
    const int* __restrict__ nonfinite_flag, const float* __restrict__ beta_ptr, float* __restrict__ o_rgb,
    float* __restrict__ o_seg, float* __restrict__ o_depth, int b0) {
  static_assert(sizeof(T) == 2, "the staged march is built for 16-bit features");
  constexpr int CP = packed_channels(K);
  constexpr uint32_t kRec = CP * sizeof(T);                 // 48 bytes
  constexpr uint32_t kStageBytes = kBoxCap * kRec;          // 4608 bytes
  constexpr int kWarps = kMarchThreads / 32;
  __shared__ __align__(128) unsigned char s_stage[kWarps][2][kStageBytes];
  __shared__ __align__(8) uint64_t s_bar[kWarps][2];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = vs_smem_u32(&s_bar[wid][0]), bar1 = vs_smem_u32(&s_bar[wid][1]);
  const uint32_t stage0 = vs_smem_u32(&s_stage[wid][0][0]);
  if (lane == 0) {
    vs_mbar_init(bar0, 1);
    vs_mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (*nonfinite_flag != 0) return;   // the NaN-safe variant of march_fwd_kernel takes over (recomputes the geometry)
  const int b = b0 + blockIdx.z, n = blockIdx.y;
  const int patches_x = (g.fW + kPatchW - 1) / kPatchW;
  const int npatch = march_patches(g);
  const int patch = blockIdx.x * kWarps + wid;
  if (patch >= npatch) return;  // whole warp leaves together
  const int w = (patch % patches_x) * kPatchW + (lane % kPatchW);
  const int h = (patch / patches_x) * kPatchH + (lane / kPatchW);
  const bool active = (w < g.fW) && (h < g.fH);
  const int S = g.D - 1, HW = g.fH * g.fW;
  const int nvox = g.vZ * g.vY * g.vX;
  const T* vol = packed + (size_t)blockIdx.z * nvox * CP;
  const size_t ray = (size_t)(n * npatch + patch);
  const uint4* __restrict__ rec = reinterpret_cast<const uint4*>(plans[b].steps) + ray * S * 32 + lane;
  const float* __restrict__ dl = plans[b].delta + ray * S * 32 + lane;
  const uint2* __restrict__ box = reinterpret_cast<const uint2*>(plans[b].box) + ray * S;
  const int last = (int)__ldg(plans[b].last + ray * 32 + lane);

  const float beta = fabsf(__ldg(beta_ptr)) + g.beta_min;
  const float inv_beta = 1.0f / beta;
  const float sigma_masked = vb_density_rcp(g, 0.0f, inv_beta);
  const int c_sy = g.vX * CP, c_sz = g.vY * g.vX * CP;     // global corner strides (elements)

  float acc = 0.0f, dep = 0.0f, trans = 1.0f;
  float ch[K + 3];
#pragma unroll
  for (int c = 0; c < K + 3; ++c) ch[c] = 0.0f;

  constexpr int kPrefetchAhead = VB_MARCH_PREFETCH;
  auto prefetch = [&](int i) {
    if (i < S) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (size_t)i * 32));
      if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(dl + (size_t)i * 32 + lane * 8 - lane));
    }
  };
  // issue the bulk copies of sample i's box into stage (i & 1): lane r copies row r = (dz, dy) of the box
  auto issue = [&](const uint2& bx, int i) -> bool {
    if (!(bx.x & kBoxStaged)) return false;
    const int nx = bx.y & 0xff, ny = (bx.y >> 8) & 0xff, nz = (bx.y >> 16) & 0xff;
    const uint32_t row_bytes = (uint32_t)nx * kRec;
    const uint32_t bar = (i & 1) ? bar1 : bar0;
    if (lane == 0) vs_mbar_expect_tx(bar, row_bytes * (uint32_t)(ny * nz));
    __syncwarp();
    const int dz = lane >> 3, dy = lane & 7;               // ny <= 8, nz <= 4 (plan build)
    if (dy < ny && dz < nz) {
      const T* src = vol + (size_t)((bx.x & kPlanVoxMask) * CP + dz * c_sz + dy * c_sy);
      vs_bulk_load(stage0 + (uint32_t)(i & 1) * kStageBytes + (uint32_t)(dz * ny + dy) * row_bytes, src, row_bytes, bar);
    }
    return true;
  };

#pragma unroll 1
  for (int i = 0; i < kPrefetchAhead; ++i) prefetch(i);
  uint4 r_n = __ldg(rec);
  float d_n = __ldg(dl);
  uint2 b_cur = __ldg(box);
  uint2 b_n1 = S > 1 ? __ldg(box + 1) : make_uint2(0u, 0u);
  uint2 b_n2 = S > 2 ? __ldg(box + 2) : make_uint2(0u, 0u);
  // staged[s]: a copy is in flight into / sits in stage s; ph[s]: parity of that stage's next completion
  bool pend0 = issue(b_cur, 0), pend1 = S > 1 ? issue(b_n1, 1) : false;
  uint32_t ph0 = 0u, ph1 = 0u;

  int i = 0;
  for (; i < S; ++i) {
    if (g.term_eps > 0.0f) {
      const bool done = !active || trans < g.term_eps;
      if (__all_sync(0xffffffffu, done)) break;
      if (__all_sync(0xffffffffu, done || i > last)) {
        if (!done) {
          for (int ii = i; ii < S; ++ii) {
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
    const uint2 bx = b_cur;
    prefetch(i + kPrefetchAhead);
    if (i + 1 < S) {
      r_n = __ldg(rec + (size_t)(i + 1) * 32);
      d_n = __ldg(dl + (size_t)(i + 1) * 32);
    }
    const bool odd = (i & 1) != 0;
    const bool staged = odd ? pend1 : pend0;           // warp-uniform
    if (staged) {
      vs_mbar_wait(odd ? bar1 : bar0, odd ? ph1 : ph0);
      if (odd) { ph1 ^= 1u; pend1 = false; } else { ph0 ^= 1u; pend0 = false; }
    }
    const bool live = (r.x & kPlanValid) != 0u;
    float sigma = sigma_masked;
    float cw[8];
    if (live) {
      const float fx = __uint_as_float(r.y), fy = __uint_as_float(r.z), fz = __uint_as_float(r.w);
      const float wx[2] = {1.0f - fx, fx}, wy[2] = {1.0f - fy, fy}, wz[2] = {1.0f - fz, fz};
#pragma unroll
      for (int q = 0; q < 8; ++q) cw[q] = wx[q & 1] * wy[(q >> 1) & 1] * wz[q >> 2];
    }
    float e, wgt;
    if (staged) {
      // ---- shared-memory path: the record's index inside the box, uniform corner strides ----
      const int nx = bx.y & 0xff, ny = (bx.y >> 8) & 0xff;
      const uint32_t s_sy = (uint32_t)nx * kRec, s_sz = (uint32_t)(nx * ny) * kRec;
      const uint32_t a0 = stage0 + (odd ? kStageBytes : 0u) + ((r.x >> kPlanRelShift) & kPlanRelMask) * kRec;
      const uint32_t a[8] = {a0, a0 + kRec, a0 + s_sy, a0 + s_sy + kRec,
                             a0 + s_sz, a0 + s_sz + kRec, a0 + s_sz + s_sy, a0 + s_sz + s_sy + kRec};
      if (live) {
        float s0 = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s0 = fmaf(cw[q], vs_widen16<T>(vs_lds16(a[q])), s0);
        sigma = vb_density_rcp(g, s0, inv_beta);                 // BV2:423
      }
      const float sd = sigma * delta;                                         // BV2:429
      e = expf(-sd);
      wgt = (1.0f - e) * trans;                                               // BV2:430-434
      if (live && wgt != 0.0f) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float cwq = cw[q] * wgt;
#pragma unroll
          for (int part = 0; part < CP / 8; ++part) {
            float tmp[8];
            vs_widen8<T>(vs_lds128(a[q] + part * 16), tmp);
#pragma unroll
            for (int e8 = 0; e8 < 8; ++e8) {
              const int c = part * 8 + e8 - 1;
              if (c >= 0 && c < K + 3) ch[c] = fmaf(tmp[e8], cwq, ch[c]);
            }
          }
        }
      }
    } else {
      // ---- global path (box over the cap): as march_fwd_planned_kernel ----
      const T* gp = vol + (size_t)(r.x & kPlanVoxMask) * CP;
      if (live) {
        float s0 = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          s0 = fmaf(cw[q], widen_elem(__ldg(gp + ((q & 2) ? c_sy : 0) + ((q & 4) ? c_sz : 0) + ((q & 1) ? CP : 0))), s0);
        sigma = vb_density_rcp(g, s0, inv_beta);
      }
      const float sd = sigma * delta;
      e = expf(-sd);
      wgt = (1.0f - e) * trans;
      if (live && wgt != 0.0f) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          PackedLoad<T, CP>::template fma_values<K + 3>(gp + ((q & 2) ? c_sy : 0) + ((q & 4) ? c_sz : 0) + ((q & 1) ? CP : 0),
                                                        cw[q] * wgt, ch);
      }
    }
    acc += wgt;
    dep = fmaf(wgt, __ldg(t.mids + i), dep);
    trans *= e;
    // stage (i & 1) is free again: every lane's reads of it have been consumed above
    __syncwarp();
    b_cur = b_n1;
    b_n1 = b_n2;
    if (i + 3 < S) b_n2 = __ldg(box + i + 3);
    if (i + 2 < S) {
      const bool p = issue(b_n1, i + 2);
      if (odd) pend1 = p; else pend0 = p;
    }
  }
  // never leave with a bulk copy still in flight into this block's shared memory
  if (pend0) vs_mbar_wait(bar0, ph0);
  if (pend1) vs_mbar_wait(bar1, ph1);
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
