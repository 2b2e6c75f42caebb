// vb_geometry.cu -- G1 / G2 / L2 / R2: the bit-exact projective geometry of the path.
//
// These kernels materialise what the fused lift / render kernels recompute on the fly with the
// very same device functions (vb_common.cuh), so the integers they emit are the integers the
// fused kernels use.  They exist for the drop-in get_pixel / get_geometry methods and for the
// index-parity tests.
#include "vb_common.cuh"
#include "vb_trace.cuh"

namespace {

// one thread per (b, n, voxel); x fastest => coalesced 12-byte-per-thread stores
__global__ void __launch_bounds__(256) get_pixel_kernel(VbGrid g, VbTables t, const float* __restrict__ d_mats,
                                                        float* __restrict__ d_pix, uint8_t* __restrict__ d_valid,
                                                        int16_t* __restrict__ d_i0, float* __restrict__ d_frac) {
  __shared__ float s_m[VB200_MAT_SLOTS * 16];
  const int bn = blockIdx.y;
  for (int i = threadIdx.x; i < VB200_MAT_SLOTS * 16; i += blockDim.x)
    s_m[i] = __ldg(d_mats + (size_t)bn * VB200_MAT_SLOTS * 16 + i);
  __syncthreads();
  const int nvox = g.vZ * g.vY * g.vX;
  const int vox = blockIdx.x * blockDim.x + threadIdx.x;
  if (vox >= nvox) return;
  const int x = vox % g.vX, y = (vox / g.vX) % g.vY, z = vox / (g.vX * g.vY);
  float pix[3];
  project_voxel(s_m, g.has_bda != 0, __ldg(t.xs + x), __ldg(t.ys + y), __ldg(t.zs + z), pix);
  const size_t o = (size_t)bn * nvox + vox;
  if (d_pix) {
    d_pix[o * 3 + 0] = pix[0];
    d_pix[o * 3 + 1] = pix[1];
    d_pix[o * 3 + 2] = pix[2];
  }
  if (d_valid || d_i0 || d_frac) {
    const LiftCoord c = lift_coord(g, pix);
    if (d_valid) d_valid[o] = c.valid ? 1 : 0;
    if (d_i0) {
      d_i0[o * 3 + 0] = (int16_t)c.x0;
      d_i0[o * 3 + 1] = (int16_t)c.y0;
      d_i0[o * 3 + 2] = (int16_t)c.z0;
    }
    if (d_frac) {
      d_frac[o * 3 + 0] = c.ix - floorf(c.ix);
      d_frac[o * 3 + 1] = c.iy - floorf(c.iy);
      d_frac[o * 3 + 2] = c.iz - floorf(c.iz);
    }
  }
}

// one thread per (b, n, d, h, w); w fastest
__global__ void __launch_bounds__(256) get_geometry_kernel(VbGrid g, VbTables t, const float* __restrict__ d_mats,
                                                           const float* __restrict__ d_geom_in,
                                                           float* __restrict__ d_geom, int do_nan_to_num,
                                                           uint8_t* __restrict__ d_mask, int16_t* __restrict__ d_i0,
                                                           float* __restrict__ d_frac) {
  __shared__ float s_m[VB200_MAT_SLOTS * 16];
  const int bn = blockIdx.y;
  for (int i = threadIdx.x; i < VB200_MAT_SLOTS * 16; i += blockDim.x)
    s_m[i] = __ldg(d_mats + (size_t)bn * VB200_MAT_SLOTS * 16 + i);
  __syncthreads();
  const int npts = g.D * g.fH * g.fW;
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= npts) return;
  const int w = pt % g.fW, h = (pt / g.fW) % g.fH, d = pt / (g.fW * g.fH);
  float p[3];
  const size_t o = (size_t)bn * npts + pt;
  if (d_geom_in) {
    p[0] = __ldg(d_geom_in + o * 3 + 0);
    p[1] = __ldg(d_geom_in + o * 3 + 1);
    p[2] = __ldg(d_geom_in + o * 3 + 2);
  } else {
    frustum_point(s_m, g.has_bda != 0, __ldg(t.us + w), __ldg(t.vs + h), __ldg(t.ds + d), p);
    if (do_nan_to_num) {
#pragma unroll
      for (int a = 0; a < 3; ++a) p[a] = nan_to_num(p[a], -1e3f);
    }
  }
  if (d_geom) {
    d_geom[o * 3 + 0] = p[0];
    d_geom[o * 3 + 1] = p[1];
    d_geom[o * 3 + 2] = p[2];
  }
  if ((d_mask || d_i0 || d_frac) && d < g.D - 1) {
    const RenderCoord c = render_coord(g, p);
    const size_t s = ((size_t)bn * (g.D - 1) + d) * g.fH * g.fW + (size_t)h * g.fW + w;
    if (d_mask) d_mask[s] = c.valid ? 1 : 0;
    if (d_i0) {
      d_i0[s * 3 + 0] = (int16_t)c.x0;
      d_i0[s * 3 + 1] = (int16_t)c.y0;
      d_i0[s * 3 + 2] = (int16_t)c.z0;
    }
    if (d_frac) {
      d_frac[s * 3 + 0] = c.valid ? c.ix - floorf(c.ix) : 0.0f;
      d_frac[s * 3 + 1] = c.valid ? c.iy - floorf(c.iy) : 0.0f;
      d_frac[s * 3 + 2] = c.valid ? c.iz - floorf(c.iz) : 0.0f;
    }
  }
}

// min_D: 1 for the voxel-side entry points (D == 1 = the 2-D lift), 2 for the frustum-side ones (rays need 2 planes)
int check_grid(const VbGrid* g, int min_D = 2) {
  if (!g) return VB200_ERR_ARG;
  if (g->B <= 0 || g->N <= 0 || g->N > VB_MAX_CAMS) return VB200_ERR_ARG;
  if (g->D < min_D || g->fH <= 0 || g->fW <= 0 || g->vZ <= 0 || g->vY <= 0 || g->vX <= 0) return VB200_ERR_ARG;
  return VB200_OK;
}

}  // namespace

extern "C" int vb200_get_pixel(const VbGrid* g, const VbTables* t, const float* d_mats, float* d_pix,
                               void* stream) {
  int rc = check_grid(g, 1);
  if (rc) return rc;
  VB_CHECK_ARG(t && d_mats && d_pix);
  if ((rc = vb200_device_check())) return rc;
  const int nvox = g->vZ * g->vY * g->vX;
  dim3 grid(vb_ceil_div(nvox, 256), g->B * g->N);
  VbTraceScope tr(VB_K_GET_PIXEL, (cudaStream_t)stream);
  get_pixel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *t, d_mats, d_pix, nullptr, nullptr, nullptr);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" int vb200_lift_indices(const VbGrid* g, const VbTables* t, const float* d_mats, uint8_t* d_valid,
                                  int16_t* d_i0, float* d_frac, void* stream) {
  int rc = check_grid(g, 1);
  if (rc) return rc;
  VB_CHECK_ARG(t && d_mats);
  if ((rc = vb200_device_check())) return rc;
  const int nvox = g->vZ * g->vY * g->vX;
  dim3 grid(vb_ceil_div(nvox, 256), g->B * g->N);
  VbTraceScope tr(VB_K_GET_PIXEL, (cudaStream_t)stream);
  get_pixel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *t, d_mats, nullptr, d_valid, d_i0, d_frac);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" int vb200_get_geometry(const VbGrid* g, const VbTables* t, const float* d_mats, float* d_geom,
                                  int nan_to_num_flag, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  VB_CHECK_ARG(t && d_mats && d_geom);
  if ((rc = vb200_device_check())) return rc;
  const int npts = g->D * g->fH * g->fW;
  dim3 grid(vb_ceil_div(npts, 256), g->B * g->N);
  VbTraceScope tr(VB_K_GET_GEOMETRY, (cudaStream_t)stream);
  get_geometry_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *t, d_mats, nullptr, d_geom, nan_to_num_flag,
                                                             nullptr, nullptr, nullptr);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" int vb200_render_indices(const VbGrid* g, const VbTables* t, const float* d_mats, const float* d_geom,
                                    uint8_t* d_mask, int16_t* d_i0, float* d_frac, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  VB_CHECK_ARG(t && d_mats);
  if ((rc = vb200_device_check())) return rc;
  const int npts = g->D * g->fH * g->fW;
  dim3 grid(vb_ceil_div(npts, 256), g->B * g->N);
  VbTraceScope tr(VB_K_GET_GEOMETRY, (cudaStream_t)stream);
  get_geometry_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *t, d_mats, d_geom, nullptr, 1, d_mask, d_i0,
                                                             d_frac);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}
