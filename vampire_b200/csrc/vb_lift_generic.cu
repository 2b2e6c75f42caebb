// vb_lift_generic.cu -- compatibility path for the reference's get_voxel_feats SIGNATURE
// (BV2:483-516), which receives the already materialised (B,N,C,D,fH,fW) frustum tensor.
//
// The product path is vb200_lift_pool_* (the frustum tensor is never formed); these two kernels
// exist only so that code calling `get_voxel_feats(frustum_feats, sweep_index, mats_dict)` keeps
// working unchanged.  Same projection / validity / trilinear / non-zero-mean semantics, but a
// generic 8-corner x C-channel gather from the 6-D tensor (and an atomicAdd scatter for its
// backward, like ATen's grid_sampler_3d_backward) -- not tuned, not on the benchmarked path.
#include "vb_common.cuh"
#include "vb_trace.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxC = 16;

template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(kThreads) gather_pool_kernel(VbGrid g, VbTables t, const float* __restrict__ d_mats,
                                                               const T* __restrict__ frustum, T* __restrict__ out,
                                                               uint64_t* __restrict__ cnt_io,
                                                               const T* __restrict__ gout, float* __restrict__ gfrustum) {
  __shared__ float s_m[VB_MAX_CAMS * VB200_MAT_SLOTS * 16];
  const int b = blockIdx.y;
  stage_mats(s_m, d_mats, b, g.N);
  __syncthreads();
  const int nvox = g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * kThreads + threadIdx.x;
  if (vox >= nvox) return;
  const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
  const float px = __ldg(t.xs + x), py = __ldg(t.ys + y), pz = __ldg(t.zs + z);
  const int C = g.C;
  const size_t HW = (size_t)g.fH * g.fW, DHW = HW * g.D;
  float acc[kMaxC], cntf[kMaxC], gp[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) { acc[c] = 0.0f; cntf[c] = 0.0f; gp[c] = 0.0f; }
  if (BACKWARD) {
    const uint64_t cw = cnt_io[(size_t)b * nvox + vox];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) gp[c] = VbType<T>::ld(gout + ((size_t)b * C + c) * nvox + vox) / ((float)((cw >> (4 * c)) & 0xf) + 1e-6f);
  }
  for (int n = 0; n < g.N; ++n) {
    float pix[3];
    project_voxel<false>(s_m + n * VB200_MAT_SLOTS * 16, g.has_bda != 0, px, py, pz, pix);
    const LiftCoord lc = lift_coord(g, pix);
    if (!lc.valid) continue;
    float f[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) f[c] = 0.0f;
    const size_t cam = ((size_t)b * g.N + n) * C * DHW;
    for (int q = 0; q < 8; ++q) {
      const int xx = lc.x0 + (q & 1), yy = lc.y0 + ((q >> 1) & 1), zz = lc.z0 + (q >> 2);
      if (xx < 0 || xx >= g.fW || yy < 0 || yy >= g.fH || zz < 0 || zz >= g.D) continue;   // zeros padding
      const float wx = (q & 1) ? lc.ix - (float)lc.x0 : (float)(lc.x0 + 1) - lc.ix;
      const float wy = ((q >> 1) & 1) ? lc.iy - (float)lc.y0 : (float)(lc.y0 + 1) - lc.iy;
      const float wz = (q >> 2) ? lc.iz - (float)lc.z0 : (float)(lc.z0 + 1) - lc.iz;
      const float wgt = wx * wy * wz;
      const size_t o = cam + (size_t)zz * HW + (size_t)yy * g.fW + xx;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        if (c < C) {
          if (BACKWARD) atomicAdd(gfrustum + o + (size_t)c * DHW, wgt * gp[c]);
          else f[c] = fmaf(wgt, VbType<T>::ld(frustum + o + (size_t)c * DHW), f[c]);
        }
      }
    }
    if (!BACKWARD) {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        acc[c] += f[c];
        cntf[c] += (fabsf(f[c]) > 0.0f) ? 1.0f : 0.0f;
      }
    }
  }
  if (!BACKWARD) {
    uint64_t cnt = 0;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {
      if (c < C) {
        out[((size_t)b * C + c) * nvox + vox] = VbType<T>::cvt(acc[c] / (cntf[c] + 1e-6f));
        cnt |= (uint64_t)(uint32_t)cntf[c] << (4 * c);
      }
    }
    if (cnt_io) cnt_io[(size_t)b * nvox + vox] = cnt;
  }
}

template <typename T>
__global__ void cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = VbType<T>::cvt(src[i]);
}

template <typename T>
int launch(const VbGrid* g, const VbTables* t, const float* d_mats, const void* frustum, void* out, uint64_t* cnt,
           const void* gout, void* gfrustum, float* ws, cudaStream_t st) {
  const int nvox = g->vZ * g->vY * g->vX;
  dim3 grid(vb_ceil_div(nvox, kThreads), g->B);
  VbTraceScope tr(VB_K_MISC, st);
  if (!gout) {
    gather_pool_kernel<T, false><<<grid, kThreads, 0, st>>>(*g, *t, d_mats, reinterpret_cast<const T*>(frustum),
                                                            reinterpret_cast<T*>(out), cnt, nullptr, nullptr);
  } else {
    const size_t n = (size_t)g->B * g->N * g->C * g->D * g->fH * g->fW;
    float* accum = sizeof(T) == 4 ? reinterpret_cast<float*>(gfrustum) : ws;
    if (cudaMemsetAsync(accum, 0, n * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
    gather_pool_kernel<T, true><<<grid, kThreads, 0, st>>>(*g, *t, d_mats, nullptr, nullptr, cnt,
                                                           reinterpret_cast<const T*>(gout), accum);
    if (sizeof(T) != 4) cast_from_f32_kernel<T><<<VB_SM_COUNT_B200 * 8, 256, 0, st>>>(accum, reinterpret_cast<T*>(gfrustum), n);
  }
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

}  // namespace

extern "C" size_t vb200_gather_pool_bwd_workspace(const VbGrid* g, int dtype) {
  if (!g || dtype == VB200_F32) return 0;
  return (size_t)g->B * g->N * g->C * g->D * g->fH * g->fW * 4;
}

extern "C" int vb200_gather_pool_fwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_frustum,
                                     int dtype, void* d_out, uint64_t* d_cnt, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && d_frustum && d_out);
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS && g->C > 0 && g->C <= kMaxC);
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case VB200_F32: return launch<float>(g, t, d_mats, d_frustum, d_out, d_cnt, nullptr, nullptr, nullptr, st);
    case VB200_BF16: return launch<__nv_bfloat16>(g, t, d_mats, d_frustum, d_out, d_cnt, nullptr, nullptr, nullptr, st);
    case VB200_F16: return launch<__half>(g, t, d_mats, d_frustum, d_out, d_cnt, nullptr, nullptr, nullptr, st);
    default: return VB200_ERR_DTYPE;
  }
}

extern "C" int vb200_gather_pool_bwd(const VbGrid* g, const VbTables* t, const float* d_mats, const void* d_gout,
                                     const uint64_t* d_cnt, int dtype, void* d_gfrustum, void* d_workspace,
                                     size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && t && d_mats && d_gout && d_cnt && d_gfrustum);
  VB_CHECK_ARG(g->B > 0 && g->N > 0 && g->N <= VB_MAX_CAMS && g->C > 0 && g->C <= kMaxC);
  if (workspace_bytes < vb200_gather_pool_bwd_workspace(g, dtype)) return VB200_ERR_WORKSPACE;
  if (dtype != VB200_F32) VB_CHECK_ARG(d_workspace);
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t* cnt = const_cast<uint64_t*>(d_cnt);
  float* ws = reinterpret_cast<float*>(d_workspace);
  switch (dtype) {
    case VB200_F32: return launch<float>(g, t, d_mats, nullptr, nullptr, cnt, d_gout, d_gfrustum, ws, st);
    case VB200_BF16: return launch<__nv_bfloat16>(g, t, d_mats, nullptr, nullptr, cnt, d_gout, d_gfrustum, ws, st);
    case VB200_F16: return launch<__half>(g, t, d_mats, nullptr, nullptr, cnt, d_gout, d_gfrustum, ws, st);
    default: return VB200_ERR_DTYPE;
  }
}
