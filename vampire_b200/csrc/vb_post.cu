// vb_post.cu -- the callers immediately AFTER the path (SURVEY §8f "next" rows 2 and 3):
//
//   upsample_bilinear   nn.UpsamplingBilinear2d(scale_factor=upsample_factor) applied to the rendered
//                       rgb / semantic / depth maps (BV2:210, 616-626): align_corners=True bilinear,
//                       forward + deterministic gather backward
//   query_points        F.grid_sample of semantic_logits (padding_mode='border') / density_feature /
//                       density at LiDAR points and at the Occ3D grid rotated by bda (BV2:576-609):
//                       align_corners=True trilinear gather straight from the NCDHW volumes, forward +
//                       atomicAdd backward (what ATen does)
//
// Same conventions as the rest of libvb200 (caller-owned buffers, async on the given stream).
#include "vb_common.cuh"
#include "vb_trace.cuh"

namespace {

// ATen upsample_bilinear2d, align_corners=True: src = dst * (in-1)/(out-1); i0 = (int)src;
// i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1   (UpSample.h area_pixel_compute_source_index)
__device__ __forceinline__ void up_axis(int o, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  const float src = scale * (float)o;
  i0 = (int)src;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int H, int W, int OH, int OW, float sy, float sx) {
  const int plane = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= OH * OW) return;
  const int ox = o % OW, oy = o / OW;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  up_axis(oy, sy, H, y0, y1, ly0, ly1);
  up_axis(ox, sx, W, x0, x1, lx0, lx1);
  const float* p = in + (size_t)plane * H * W;
  out[(size_t)plane * OH * OW + o] = ly0 * (lx0 * __ldg(p + y0 * W + x0) + lx1 * __ldg(p + y0 * W + x1)) +
                                     ly1 * (lx0 * __ldg(p + y1 * W + x0) + lx1 * __ldg(p + y1 * W + x1));
}

// The same, four consecutive output columns per thread and one 128-bit store: the op is a pure write stream (16
// output pixels per input pixel at factor 4), and scalar 4-byte stores left it at 16 % of the copy bandwidth.
// Same expression per output pixel as the scalar kernel (bit-identical).
__global__ void __launch_bounds__(256) upsample_fwd_vec4_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                int H, int W, int OH, int OW, float sy, float sx) {
  const int plane = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;      // quad index inside the plane
  const int qw = OW >> 2;
  if (q >= OH * qw) return;
  const int oy = q / qw, ox0 = (q % qw) << 2;
  int y0, y1;
  float ly0, ly1;
  up_axis(oy, sy, H, y0, y1, ly0, ly1);
  const float* r0 = in + (size_t)plane * H * W + (size_t)y0 * W;
  const float* r1 = in + (size_t)plane * H * W + (size_t)y1 * W;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int x0, x1;
    float lx0, lx1;
    up_axis(ox0 + k, sx, W, x0, x1, lx0, lx1);
    v[k] = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
  }
  *reinterpret_cast<float4*>(out + (size_t)plane * OH * OW + (size_t)oy * OW + ox0) = make_float4(v[0], v[1], v[2], v[3]);
}

// gather backward: one thread per INPUT pixel sums the <= (2f+1)^2 output pixels whose stencil touches it
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin,
                                                           int H, int W, int OH, int OW, float sy, float sx) {
  const int plane = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int x = i % W, y = i / W;
  // outputs with src in (y-1, y+1): conservative index window, exact membership tested per output
  const int oy_lo = max(0, (int)floorf((float)(y - 1) / sy) - 1), oy_hi = min(OH - 1, (int)ceilf((float)(y + 1) / sy) + 1);
  const int ox_lo = max(0, (int)floorf((float)(x - 1) / sx) - 1), ox_hi = min(OW - 1, (int)ceilf((float)(x + 1) / sx) + 1);
  const float* g = gout + (size_t)plane * OH * OW;
  float acc = 0.0f;
  // the x weights do not depend on the output row: tabulate them once per thread (x4: the window is <= 12 columns)
  constexpr int kWin = 14;
  if (ox_hi - ox_lo < kWin) {
    float wxv[kWin];
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const int ox = ox_lo + k;
      wxv[k] = 0.0f;
      if (ox <= ox_hi) {
        int x0, x1;
        float lx0, lx1;
        up_axis(ox, sx, W, x0, x1, lx0, lx1);
        wxv[k] = (x0 == x ? lx0 : 0.0f) + (x1 == x ? lx1 : 0.0f);
      }
    }
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly0, ly1;
      up_axis(oy, sy, H, y0, y1, ly0, ly1);
      const float wy = (y0 == y ? ly0 : 0.0f) + (y1 == y ? ly1 : 0.0f);
      if (wy == 0.0f) continue;
      const float* row = g + (size_t)oy * OW + ox_lo;
#pragma unroll
      for (int k = 0; k < kWin; ++k)      // same order and the same products as the generic loop below
        if (wxv[k] != 0.0f) acc = fmaf(wy * wxv[k], __ldg(row + k), acc);
    }
    gin[(size_t)plane * H * W + i] = acc;
    return;
  }
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    int y0, y1;
    float ly0, ly1;
    up_axis(oy, sy, H, y0, y1, ly0, ly1);
    const float wy = (y0 == y ? ly0 : 0.0f) + (y1 == y ? ly1 : 0.0f);
    if (wy == 0.0f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      int x0, x1;
      float lx0, lx1;
      up_axis(ox, sx, W, x0, x1, lx0, lx1);
      const float wx = (x0 == x ? lx0 : 0.0f) + (x1 == x ? lx1 : 0.0f);
      if (wx != 0.0f) acc = fmaf(wy * wx, __ldg(g + oy * OW + ox), acc);
    }
  }
  gin[(size_t)plane * H * W + i] = acc;
}

// ---- point / occupancy queries -------------------------------------------------------------------------------
struct QueryCoord {
  bool valid;            // all three normalised coordinates inside [-1, 1] (BV2:587-589)
  int x0, y0, z0;
  float ix, iy, iz;
};

// norm = (p - lo) / ext * 2 - 1 (strict, BV2:581-586 / 602-606); unnormalise align_corners=True;
// border padding clips the unnormalised coordinate to [0, size-1] before the floor (ATen clip_coordinates)
__device__ __forceinline__ QueryCoord query_coord(const VbGrid& g, const float (&p)[3], bool border) {
  QueryCoord c;
  const float gx = ssub(smul(sdiv(ssub(p[0], g.seg_lo[0]), g.seg_ext[0]), 2.0f), 1.0f);
  const float gy = ssub(smul(sdiv(ssub(p[1], g.seg_lo[1]), g.seg_ext[1]), 2.0f), 1.0f);
  const float gz = ssub(smul(sdiv(ssub(p[2], g.seg_lo[2]), g.seg_ext[2]), 2.0f), 1.0f);
  c.valid = (gx >= -1.0f) && (gx <= 1.0f) && (gy >= -1.0f) && (gy <= 1.0f) && (gz >= -1.0f) && (gz <= 1.0f);
  c.ix = smul(smul(sadd(gx, 1.0f), 0.5f), (float)(g.vX - 1));
  c.iy = smul(smul(sadd(gy, 1.0f), 0.5f), (float)(g.vY - 1));
  c.iz = smul(smul(sadd(gz, 1.0f), 0.5f), (float)(g.vZ - 1));
  if (border) {
    c.ix = fminf((float)(g.vX - 1), fmaxf(c.ix, 0.0f));
    c.iy = fminf((float)(g.vY - 1), fmaxf(c.iy, 0.0f));
    c.iz = fminf((float)(g.vZ - 1), fmaxf(c.iz, 0.0f));
  }
  // keep the int casts defined for absurd coordinates (their corners are all out of range anyway)
  const float lim = 1.0e8f;
  c.x0 = (int)floorf(fminf(fmaxf(c.ix, -lim), lim));
  c.y0 = (int)floorf(fminf(fmaxf(c.iy, -lim), lim));
  c.z0 = (int)floorf(fminf(fmaxf(c.iz, -lim), lim));
  return c;
}

struct Corners {
  int off[8];
  float w[8];
};
__device__ __forceinline__ Corners query_corners(const VbGrid& g, const QueryCoord& c) {
  Corners k;
  const float fx = c.ix - (float)c.x0, fy = c.iy - (float)c.y0, fz = c.iz - (float)c.z0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int xx = c.x0 + (q & 1), yy = c.y0 + ((q >> 1) & 1), zz = c.z0 + (q >> 2);
    const bool in = xx >= 0 && xx < g.vX && yy >= 0 && yy < g.vY && zz >= 0 && zz < g.vZ;   // zeros padding
    const float wx = (q & 1) ? fx : (float)(c.x0 + 1) - c.ix;
    const float wy = ((q >> 1) & 1) ? fy : (float)(c.y0 + 1) - c.iy;
    const float wz = (q >> 2) ? fz : (float)(c.z0 + 1) - c.iz;
    k.w[q] = in ? wx * wy * wz : 0.0f;
    k.off[q] = in ? (zz * g.vY + yy) * g.vX + xx : 0;
  }
  return k;
}

__device__ __forceinline__ void load_point(const float* __restrict__ pts, const float* __restrict__ rot, int b,
                                           size_t idx, float (&p)[3]) {
  const float* q = pts + idx * 3;
  const float a[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
  if (rot) {   // bda[:3,:3] @ p, ATen native bmm order (BV2:598-601)
    const float* R = rot + (size_t)b * 9;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      p[i] = sadd(sadd(sadd(0.0f, smul(__ldg(R + 3 * i), a[0])), smul(__ldg(R + 3 * i + 1), a[1])),
                  smul(__ldg(R + 3 * i + 2), a[2]));
  } else {
    p[0] = a[0]; p[1] = a[1]; p[2] = a[2];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) query_fwd_kernel(VbGrid g, const T* __restrict__ vol, int CH,
                                                        const float* __restrict__ pts, const float* __restrict__ rot,
                                                        int P, int pts_batched, int border, int apply_density,
                                                        int mask_invalid, const float* __restrict__ beta_ptr,
                                                        float* __restrict__ out, uint8_t* __restrict__ valid_out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float p[3];
  load_point(pts, rot, b, (pts_batched ? (size_t)b * P : 0) + i, p);
  const QueryCoord c = query_coord(g, p, border != 0);
  const Corners k = query_corners(g, c);
  const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
  const float beta = apply_density ? fabsf(__ldg(beta_ptr)) + g.beta_min : 1.0f;
  const float m = (mask_invalid && !c.valid) ? 0.0f : 1.0f;
  if (valid_out) valid_out[(size_t)b * P + i] = c.valid ? 1 : 0;
  for (int ch = 0; ch < CH; ++ch) {
    const T* plane = vol + ((size_t)b * CH + ch) * nvox;
    float v = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float s = VbType<T>::ld(plane + k.off[q]);
      if (apply_density) s = vb_density(g, s, beta);   // sigma volume is sampled (BV2:609)
      v = fmaf(k.w[q], s, v);
    }
    out[((size_t)b * CH + ch) * P + i] = v * m;
  }
}

// d_vol += w * g (atomicAdd, like grid_sampler_3d_backward); with apply_density the chain rule through sigma
template <typename T>
__global__ void __launch_bounds__(256) query_bwd_kernel(VbGrid g, const T* __restrict__ vol, int CH,
                                                        const float* __restrict__ pts, const float* __restrict__ rot,
                                                        int P, int pts_batched, int border, int apply_density,
                                                        int mask_invalid, const float* __restrict__ beta_ptr,
                                                        const float* __restrict__ gout, float* __restrict__ gvol,
                                                        float* __restrict__ gbeta_partial) {
  __shared__ float s_red[8];
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float dbeta = 0.0f;
  if (i < P) {
    float p[3];
    load_point(pts, rot, b, (pts_batched ? (size_t)b * P : 0) + i, p);
    const QueryCoord c = query_coord(g, p, border != 0);
    const Corners k = query_corners(g, c);
    const size_t nvox = (size_t)g.vZ * g.vY * g.vX;
    const float beta = apply_density ? fabsf(__ldg(beta_ptr)) + g.beta_min : 1.0f;
    const float m = (mask_invalid && !c.valid) ? 0.0f : 1.0f;
    for (int ch = 0; ch < CH; ++ch) {
      const float go = m * __ldg(gout + ((size_t)b * CH + ch) * P + i);
      if (go == 0.0f) continue;
      float* gp = gvol + ((size_t)b * CH + ch) * nvox;
      const T* plane = vol + ((size_t)b * CH + ch) * nvox;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (k.w[q] == 0.0f) continue;
        float d = k.w[q] * go;
        if (apply_density) {
          const DensityD dd = vb_density_with_grads(g, VbType<T>::ld(plane + k.off[q]), beta);
          dbeta = fmaf(d, dd.dbeta, dbeta);
          d *= dd.ds;
        }
        atomicAdd(gp + k.off[q], d);
      }
    }
  }
  if (gbeta_partial) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dbeta += __shfl_down_sync(0xffffffffu, dbeta, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dbeta;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.0f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
      gbeta_partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
  }
}

template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = VbType<T>::cvt(src[i]);
}

__global__ void beta_sum_kernel(const float* __restrict__ partials, int n, const float* __restrict__ beta_ptr,
                                float* __restrict__ g_beta) {
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += 32) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    const float p = *beta_ptr;
    *g_beta = s * ((p > 0.0f) ? 1.0f : ((p < 0.0f) ? -1.0f : 0.0f));
  }
}

}  // namespace

extern "C" int vb200_upsample_bilinear_fwd(const float* d_in, float* d_out, int planes, int H, int W, int factor,
                                           void* stream) {
  VB_CHECK_ARG(d_in && d_out && planes > 0 && H > 0 && W > 0 && factor >= 1 && planes <= 65535);
  int rc = vb200_device_check();
  if (rc) return rc;
  const int OH = H * factor, OW = W * factor;
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.0f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.0f;
  cudaStream_t st = (cudaStream_t)stream;
  VbTraceScope tr(VB_K_MISC, st);
  if ((OW & 3) == 0 && ((uintptr_t)d_out & 15) == 0)
    upsample_fwd_vec4_kernel<<<dim3(vb_ceil_div((long long)OH * (OW >> 2), 256), planes), 256, 0, st>>>(d_in, d_out, H, W,
                                                                                                         OH, OW, sy, sx);
  else
    upsample_fwd_kernel<<<dim3(vb_ceil_div((long long)OH * OW, 256), planes), 256, 0, st>>>(d_in, d_out, H, W, OH, OW, sy, sx);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" int vb200_upsample_bilinear_bwd(const float* d_gout, float* d_gin, int planes, int H, int W, int factor,
                                           void* stream) {
  VB_CHECK_ARG(d_gout && d_gin && planes > 0 && H > 0 && W > 0 && factor >= 1 && planes <= 65535);
  int rc = vb200_device_check();
  if (rc) return rc;
  const int OH = H * factor, OW = W * factor;
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.0f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.0f;
  cudaStream_t st = (cudaStream_t)stream;
  VbTraceScope tr(VB_K_MISC, st);
  upsample_bwd_kernel<<<dim3(vb_ceil_div((long long)H * W, 256), planes), 256, 0, st>>>(d_gout, d_gin, H, W, OH, OW, sy, sx);
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

extern "C" size_t vb200_query_points_bwd_workspace(const VbGrid* g, int channels, int P, int dtype) {
  if (!g) return 0;
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX;
  size_t n = (size_t)vb_ceil_div(P, 256) * g->B * 4 + 256;                  // d beta partials
  if (dtype != VB200_F32) n += (size_t)g->B * channels * nvox * 4;          // fp32 accumulator
  return (n + 255) & ~(size_t)255;
}

extern "C" int vb200_query_points_fwd(const VbGrid* g, const void* d_vol, int dtype, int channels, const float* d_pts,
                                      int P, int pts_batched, const float* d_rot3x3, int border, int apply_density,
                                      int mask_invalid, const float* d_beta, float* d_out, uint8_t* d_valid,
                                      void* stream) {
  VB_CHECK_ARG(g && d_vol && d_pts && d_out && P > 0 && channels > 0 && g->B > 0);
  if (apply_density) VB_CHECK_ARG(d_beta);
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(vb_ceil_div(P, 256), g->B);
  VbTraceScope tr(VB_K_MISC, st);
  switch (dtype) {
    case VB200_F32:
      query_fwd_kernel<float><<<grid, 256, 0, st>>>(*g, (const float*)d_vol, channels, d_pts, d_rot3x3, P, pts_batched,
                                                    border, apply_density, mask_invalid, d_beta, d_out, d_valid);
      break;
    case VB200_BF16:
      query_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*g, (const __nv_bfloat16*)d_vol, channels, d_pts, d_rot3x3,
                                                            P, pts_batched, border, apply_density, mask_invalid, d_beta,
                                                            d_out, d_valid);
      break;
    case VB200_F16:
      query_fwd_kernel<__half><<<grid, 256, 0, st>>>(*g, (const __half*)d_vol, channels, d_pts, d_rot3x3, P, pts_batched,
                                                     border, apply_density, mask_invalid, d_beta, d_out, d_valid);
      break;
    default: return VB200_ERR_DTYPE;
  }
  VB_LAUNCH_CHECK();
  return VB200_OK;
}

namespace {
template <typename T>
int launch_query_bwd(const VbGrid* g, const void* d_vol, int channels, const float* d_pts, int P, int pts_batched,
                     const float* d_rot, int border, int apply_density, int mask_invalid, const float* d_beta,
                     const float* d_gout, void* d_gvol, float* d_gbeta, char* ws, cudaStream_t st) {
  const size_t nvox = (size_t)g->vZ * g->vY * g->vX, n = (size_t)g->B * channels * nvox;
  const int nblk = vb_ceil_div(P, 256) * g->B;
  float* partials = reinterpret_cast<float*>(ws);
  float* accum = sizeof(T) == 4 ? reinterpret_cast<float*>(d_gvol)
                                : reinterpret_cast<float*>(ws + (((size_t)nblk * 4 + 255) & ~(size_t)255));
  if (cudaMemsetAsync(accum, 0, n * 4, st) != cudaSuccess) return VB200_ERR_CUDA;
  VbTraceScope tr(VB_K_MISC, st);
  query_bwd_kernel<T><<<dim3(vb_ceil_div(P, 256), g->B), 256, 0, st>>>(
      *g, reinterpret_cast<const T*>(d_vol), channels, d_pts, d_rot, P, pts_batched, border, apply_density,
      mask_invalid, d_beta, d_gout, accum, (apply_density && d_gbeta) ? partials : nullptr);
  VB_LAUNCH_CHECK();
  if (sizeof(T) != 4) {
    cast_f32_kernel<T><<<VB_SM_COUNT_B200 * 8, 256, 0, st>>>(accum, reinterpret_cast<T*>(d_gvol), n);
    VB_LAUNCH_CHECK();
  }
  if (apply_density && d_gbeta) {
    beta_sum_kernel<<<1, 32, 0, st>>>(partials, nblk, d_beta, d_gbeta);
    VB_LAUNCH_CHECK();
  }
  return VB200_OK;
}
}  // namespace

extern "C" int vb200_query_points_bwd(const VbGrid* g, const void* d_vol, int dtype, int channels, const float* d_pts,
                                      int P, int pts_batched, const float* d_rot3x3, int border, int apply_density,
                                      int mask_invalid, const float* d_beta, const float* d_gout, void* d_gvol,
                                      float* d_gbeta, void* d_workspace, size_t workspace_bytes, void* stream) {
  VB_CHECK_ARG(g && d_vol && d_pts && d_gout && d_gvol && d_workspace && P > 0 && channels > 0 && g->B > 0);
  if (apply_density) VB_CHECK_ARG(d_beta);
  if (workspace_bytes < vb200_query_points_bwd_workspace(g, channels, P, dtype)) return VB200_ERR_WORKSPACE;
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(d_workspace);
  switch (dtype) {
    case VB200_F32:
      return launch_query_bwd<float>(g, d_vol, channels, d_pts, P, pts_batched, d_rot3x3, border, apply_density,
                                     mask_invalid, d_beta, d_gout, d_gvol, d_gbeta, ws, st);
    case VB200_BF16:
      return launch_query_bwd<__nv_bfloat16>(g, d_vol, channels, d_pts, P, pts_batched, d_rot3x3, border, apply_density,
                                             mask_invalid, d_beta, d_gout, d_gvol, d_gbeta, ws, st);
    case VB200_F16:
      return launch_query_bwd<__half>(g, d_vol, channels, d_pts, P, pts_batched, d_rot3x3, border, apply_density,
                                      mask_invalid, d_beta, d_gout, d_gvol, d_gbeta, ws, st);
    default: return VB200_ERR_DTYPE;
  }
}
