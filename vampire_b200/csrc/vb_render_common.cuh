// vb_render_common.cuh -- device code shared by the render forward (vb_render.cu) and backward
// (vb_render_bwd.cu): the channels-last packing of the camera-branch volume, the trilinear corner
// fetch, and the BEV sampling helpers.
#pragma once
#include "vb_common.cuh"

namespace {

constexpr int kPackThreads = 256;
constexpr int kMarchThreads = 128;
constexpr int kPatchW = 8, kPatchH = 4;  // a warp marches an 8x4 patch of feature-map pixels
constexpr int kMaxLevels = 16;

__host__ __device__ constexpr int packed_channels(int K) { return ((K + 4) + 7) / 8 * 8; }

// ---- R1: pack density | sem | rgb of ONE sample into channels-last ----------------------------
// grid = (voxel tiles, samples of this round).  den/sem/rgb point at the FIRST sample of the round;
// `packed_stride` = elements between consecutive samples' packed copies.
template <typename T, int K>
__global__ void __launch_bounds__(kPackThreads) pack_cam_volume_kernel(const T* __restrict__ den,
                                                                       const T* __restrict__ sem,
                                                                       const T* __restrict__ rgb, T* __restrict__ packed,
                                                                       int nvox, size_t packed_stride) {
  constexpr int NCH = K + 4, CP = packed_channels(K), LD = CP + 1;
  __shared__ T s[kPackThreads * LD];
  const int i_s = blockIdx.y;
  den += (size_t)i_s * nvox;
  sem += (size_t)i_s * K * nvox;
  rgb += (size_t)i_s * 3 * nvox;
  packed += (size_t)i_s * packed_stride;
  const int v0 = blockIdx.x * kPackThreads;
  const int v = v0 + threadIdx.x;
  if (v < nvox) {
    s[threadIdx.x * LD] = den[v];
#pragma unroll
    for (int k = 0; k < K; ++k) s[threadIdx.x * LD + 1 + k] = sem[(size_t)k * nvox + v];
#pragma unroll
    for (int j = 0; j < 3; ++j) s[threadIdx.x * LD + 1 + K + j] = rgb[(size_t)j * nvox + v];
  }
  __syncthreads();
  const int nv = min(kPackThreads, nvox - v0);
  T* out = packed + (size_t)v0 * CP;
  for (int i = threadIdx.x; i < nv * CP; i += kPackThreads) {
    const int vv = i / CP, c = i % CP;
    out[i] = c < NCH ? s[vv * LD + c] : VbType<T>::cvt(0.0f);
  }
}

template <typename T, int CP> struct PackedLoad {
  __device__ __forceinline__ static void fma_corner(const T* p, float wgt, float (&v)[CP]) {
    constexpr int L = VbLanes<T>::n;
#pragma unroll
    for (int q = 0; q < CP / L; ++q) {
      float tmp[L];
      VbVec<T, L>::ld(p + q * L, tmp);
#pragma unroll
      for (int e = 0; e < L; ++e) v[q * L + e] = fmaf(tmp[e], wgt, v[q * L + e]);
    }
  }
};

__device__ __forceinline__ void axis_coord(float centre, float lo, float ext, int size, int& i0, float& w0,
                                           float& w1) {
  const float gn = ssub(smul(sdiv(ssub(centre, lo), ext), 2.0f), 1.0f);
  const float i = smul(smul(sadd(gn, 1.0f), 0.5f), (float)(size - 1));   // align_corners=True; /2 == *0.5 exactly
  const float fl = floorf(i);
  i0 = (int)fl;
  w1 = i - fl;
  w0 = (fl + 1.0f) - i;
}

struct BevLevel {     // per output level (top first): base z-row and weights -- identical for all columns
  int z0;
  float wz0, wz1;
};

struct BevColumn {
  int o00, o01, o10, o11;      // offsets of the 4 xy corners inside one z-row (clamped)
  float w00, w01, w10, w11;    // their weights (0 where the corner is outside the grid)
};

__device__ __forceinline__ void bev_level_table(const VbGrid& g, const VbTables& t, BevLevel* s_lv) {
  if (threadIdx.x < g.oZ) {   // level l samples output voxel oz = oZ-1-l (torch.flip BV2:443)
    BevLevel L;
    axis_coord(__ldg(t.ozs + (g.oZ - 1 - threadIdx.x)), g.seg_lo[2], g.seg_ext[2], g.vZ, L.z0, L.wz0, L.wz1);
    s_lv[threadIdx.x] = L;
  }
  __syncthreads();
}

__device__ __forceinline__ BevColumn bev_column(const VbGrid& g, const VbTables& t, int ox, int oy) {
  BevColumn bc;
  int x0, y0;
  float wx0, wx1, wy0, wy1;
  axis_coord(__ldg(t.oxs + ox), g.seg_lo[0], g.seg_ext[0], g.vX, x0, wx0, wx1);
  axis_coord(__ldg(t.oys + oy), g.seg_lo[1], g.seg_ext[1], g.vY, y0, wy0, wy1);
  const bool x0in = x0 >= 0 && x0 < g.vX, x1in = x0 + 1 >= 0 && x0 + 1 < g.vX;
  const bool y0in = y0 >= 0 && y0 < g.vY, y1in = y0 + 1 >= 0 && y0 + 1 < g.vY;
  bc.w00 = (x0in && y0in) ? wx0 * wy0 : 0.0f;
  bc.w01 = (x1in && y0in) ? wx1 * wy0 : 0.0f;
  bc.w10 = (x0in && y1in) ? wx0 * wy1 : 0.0f;
  bc.w11 = (x1in && y1in) ? wx1 * wy1 : 0.0f;
  const int xa = min(max(x0, 0), g.vX - 1), xb = min(max(x0 + 1, 0), g.vX - 1);
  const int ya = min(max(y0, 0), g.vY - 1), yb = min(max(y0 + 1, 0), g.vY - 1);
  bc.o00 = ya * g.vX + xa; bc.o01 = ya * g.vX + xb; bc.o10 = yb * g.vX + xa; bc.o11 = yb * g.vX + xb;
  return bc;
}

template <typename T>
__device__ __forceinline__ float bev_row(const VbGrid& g, const BevColumn& bc, const T* __restrict__ plane, int z) {
  if (z < 0 || z >= g.vZ) return 0.0f;   // zeros padding (uniform branch)
  const int base = z * g.vY * g.vX;      // 32-bit element index: one plane is < 2^31 elements
  return bc.w00 * VbType<T>::ld(plane + (base + bc.o00)) + bc.w01 * VbType<T>::ld(plane + (base + bc.o01)) +
         bc.w10 * VbType<T>::ld(plane + (base + bc.o10)) + bc.w11 * VbType<T>::ld(plane + (base + bc.o11));
}


inline size_t vb_elem_size(int dtype) { return dtype == VB200_F32 ? 4 : 2; }
inline size_t vb_align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace
