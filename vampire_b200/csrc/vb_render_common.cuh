// vb_render_common.cuh -- device code shared by the render forward (vb_render.cu) and backward
// (vb_render_bwd.cu): the channels-last packing of the camera-branch volume, the trilinear corner
// fetch, and the BEV sampling helpers.
#pragma once
#include "vb_common.cuh"

#include <cooperative_groups.h>

namespace {

constexpr int kPackThreads = 256;
#ifndef VB_MARCH_THREADS
#define VB_MARCH_THREADS 128
#endif
constexpr int kMarchThreads = VB_MARCH_THREADS;
#ifndef VB_MARCH_MINB
#define VB_MARCH_MINB 5
#endif
#ifndef VB_PATCH_W
#define VB_PATCH_W 8
#endif
constexpr int kPatchW = VB_PATCH_W, kPatchH = 32 / VB_PATCH_W;  // a warp marches an 8x4 patch of feature-map pixels
constexpr int kMaxLevels = 16;

__host__ __device__ constexpr int packed_channels(int K) { return ((K + 4) + 7) / 8 * 8; }

// ---- R1: pack density | sem | rgb of ONE sample into channels-last ----------------------------
// grid = (voxel tiles, samples of this round).  den/sem/rgb point at the FIRST sample of the round;
// `packed_stride` = elements between consecutive samples' packed copies.
// Transposition in registers: a thread owns 16/sizeof(T)/... consecutive voxels (2 for 16-bit features, 1 for
// fp32), loads one 32-bit word per channel plane (coalesced 128 B per warp) and stores its voxels'
// records with 128-bit stores -- no shared memory, no barrier, 4x fewer LSU instructions than a
// scalar shared-memory transpose (ncu showed the LSU pipe, not HBM, limiting the first version).
template <typename T> struct PackVox { static constexpr int n = 4 / sizeof(T); };   // voxels per thread
__device__ __forceinline__ float widen_elem(float v) { return v; }
__device__ __forceinline__ float widen_elem(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float widen_elem(__half v) { return __half2float(v); }

#ifndef VB_PACK_MINB
#define VB_PACK_MINB 5   // 51 registers: measured 0.150 ms = 98 % of the copy peak (unbounded: 128 regs, 0.180 ms)
#endif
template <typename T, int K>
__global__ void __launch_bounds__(kPackThreads, VB_PACK_MINB) pack_cam_volume_kernel(const T* __restrict__ den,
                                                                       const T* __restrict__ sem,
                                                                       const T* __restrict__ rgb, T* __restrict__ packed,
                                                                       int nvox, size_t packed_stride,
                                                                       int* __restrict__ nonfinite_flag) {
  constexpr int NCH = K + 4, CP = packed_channels(K), V = PackVox<T>::n;
  __shared__ uint4 s_rec[kPackThreads * 6];
  const int i_s = blockIdx.y;
  den += (size_t)i_s * nvox;
  sem += (size_t)i_s * K * nvox;
  rgb += (size_t)i_s * 3 * nvox;
  packed += (size_t)i_s * packed_stride;
  const int v = (blockIdx.x * kPackThreads + threadIdx.x) * V;
  // the vector path needs the whole warp in range (cooperative stores) and, for 16-bit features, an even
  // plane size (32-bit words of two voxels must be aligned); anything else takes the scalar path
  const bool vec_ok = (v + V <= nvox) && (V == 1 || !(nvox & 1));
  if (!__all_sync(0xffffffffu, vec_ok)) {
    bool bad = false;
    for (int vv = v; vv < min(v + V, nvox); ++vv) {
      T* o = packed + (size_t)vv * CP;
      o[0] = den[vv];
      for (int k = 0; k < K; ++k) o[1 + k] = sem[(size_t)k * nvox + vv];
      for (int j = 0; j < 3; ++j) o[1 + K + j] = rgb[(size_t)j * nvox + vv];
      for (int c = NCH; c < CP; ++c) o[c] = VbType<T>::cvt(0.0f);
      for (int c = 0; c < NCH; ++c) bad = bad || !(fabsf(widen_elem(o[c])) <= 3.402823466e+38f);
    }
    if (bad && nonfinite_flag) *nonfinite_flag = 1;
    return;
  }
  uint32_t w[CP];                          // one 32-bit word per channel: V voxels side by side
  w[0] = __ldg(reinterpret_cast<const uint32_t*>(den + v));
#pragma unroll
  for (int k = 0; k < K; ++k) w[1 + k] = __ldg(reinterpret_cast<const uint32_t*>(sem + (size_t)k * nvox + v));
#pragma unroll
  for (int j = 0; j < 3; ++j) w[1 + K + j] = __ldg(reinterpret_cast<const uint32_t*>(rgb + (size_t)j * nvox + v));
#pragma unroll
  for (int c = NCH; c < CP; ++c) w[c] = 0u;
  // non-finite detector for the march (torch.nan_to_num of the sampled features, BV2:421, only matters then):
  // adding one to an all-ones exponent field carries into the sign position of that element
  {
    constexpr uint32_t EXP = sizeof(T) == 4 ? 0x7f800000u : (VbType<T>::code == VB200_BF16 ? 0x7f807f80u : 0x7c007c00u);
    constexpr uint32_t ONE = sizeof(T) == 4 ? 0x00800000u : (VbType<T>::code == VB200_BF16 ? 0x00800080u : 0x04000400u);
    constexpr uint32_t SGN = sizeof(T) == 4 ? 0x80000000u : 0x80008000u;
    uint32_t bad = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) bad |= ((w[c] & EXP) + ONE) & SGN;
    if (bad != 0u && nonfinite_flag) *nonfinite_flag = 1;   // benign race: every writer stores 1
  }
  // assemble this thread's records (96 contiguous bytes) in registers ...
  uint4 rec[6];
  static_assert(CP * sizeof(T) * V == 96, "record staging below assumes 96 bytes per thread");
  if (V == 1) {
#pragma unroll
    for (int q = 0; q < 6; ++q) rec[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  } else {
    // record of voxel v: low halves of consecutive channel words; voxel v+1: high halves
    uint32_t r0[CP / 2], r1[CP / 2];
#pragma unroll
    for (int q = 0; q < CP / 2; ++q) {
      r0[q] = __byte_perm(w[2 * q], w[2 * q + 1], 0x5410);
      r1[q] = __byte_perm(w[2 * q], w[2 * q + 1], 0x7632);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      rec[q] = make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]);
      rec[3 + q] = make_uint4(r1[4 * q], r1[4 * q + 1], r1[4 * q + 2], r1[4 * q + 3]);
    }
  }
  // ... and bounce them through shared memory so that every 128-bit store instruction of a warp covers
  // 512 contiguous bytes (16 full sectors) instead of 32 half-written sectors at a 96-byte stride
  s_rec[threadIdx.x * 6 + 0] = rec[0]; s_rec[threadIdx.x * 6 + 1] = rec[1]; s_rec[threadIdx.x * 6 + 2] = rec[2];
  s_rec[threadIdx.x * 6 + 3] = rec[3]; s_rec[threadIdx.x * 6 + 4] = rec[4]; s_rec[threadIdx.x * 6 + 5] = rec[5];
  __syncwarp();
  const int lane = threadIdx.x & 31, warp0 = threadIdx.x - lane;
  uint4* out = reinterpret_cast<uint4*>(packed + (size_t)(v - lane * V) * CP);   // the warp's first record
#pragma unroll
  for (int q = 0; q < 6; ++q) out[q * 32 + lane] = s_rec[warp0 * 6 + q * 32 + lane];
}

template <typename T, int CP> struct PackedLoad;
template <typename T, int CP> struct PackedLoad {
  // acc[c] += wgt * p[1 + c] for c < NV (channel 0 = density is skipped)
  template <int NV>
  __device__ __forceinline__ static void fma_values(const T* p, float wgt, float (&acc)[NV]) {
    constexpr int L = VbLanes<T>::n;
#pragma unroll
    for (int q = 0; q < CP / L; ++q) {
      float tmp[L];
      VbVec<T, L>::ld(p + q * L, tmp);
#pragma unroll
      for (int e = 0; e < L; ++e) {
        const int c = q * L + e - 1;
        if (c >= 0 && c < NV) acc[c] = fmaf(tmp[e], wgt, acc[c]);
      }
    }
  }
  // den = p[0]; dot = sum_{c < NV} g[c] * p[1 + c]   (one corner of the backward march: no per-channel temporary)
  template <int NV>
  __device__ __forceinline__ static void dot_values(const T* p, const float (&g)[NV], float& den, float& dot) {
    constexpr int L = VbLanes<T>::n;
    den = 0.0f;
    dot = 0.0f;
#pragma unroll
    for (int q = 0; q < CP / L; ++q) {
      float tmp[L];
      VbVec<T, L>::ld(p + q * L, tmp);
#pragma unroll
      for (int e = 0; e < L; ++e) {
        const int c = q * L + e - 1;
        if (c == -1) den = tmp[e];
        else if (c < NV) dot = fmaf(tmp[e], g[c < NV ? c : 0], dot);
      }
    }
  }
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

// ---- depth-split march: a thread-block CLUSTER marches one group of ray patches, CTA r of the cluster the samples
// [r S / n, (r + 1) S / n) with a transmittance that starts at 1.  Volume rendering composes front to back:
//   out = out_0 + T_0 out_1 + T_0 T_1 out_2 + ...,   T = T_0 T_1 ...
// so every CTA leaves (acc, dep, T, ch[NV]) of its segment in its own shared memory and CTA 0 folds the others in
// through distributed shared memory (cluster.map_shared_rank), in segment order.  Meant for launches that do not fill
// the GPU (B = 1 at the R50 config: 528 blocks on 740 slots, and the kernel then lasts as long as its longest ray):
// n times more, n times shorter blocks.  A segment cannot know that an earlier one already made the ray opaque, so it
// composites samples the unsplit march skips (their weight is below term_eps after the fold) -- which is why two
// segments win at B = 1 (-11 % bf16, -16 % fp32) and more segments, or any split of a full launch, lose.
// Returns true in the CTA that holds the folded result (cluster rank 0).
template <int NV>
__device__ __forceinline__ bool march_cluster_fold(float& acc, float& dep, float& trans, float (&ch)[NV]) {
  namespace cg = cooperative_groups;
  __shared__ float s_part[(NV + 3) * kMarchThreads];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned seg = cl.block_rank(), nseg = cl.num_blocks();
  float* mine = s_part + threadIdx.x;
  if (seg != 0) {
    mine[0] = acc;
    mine[kMarchThreads] = dep;
    mine[2 * kMarchThreads] = trans;
#pragma unroll
    for (int c = 0; c < NV; ++c) mine[(3 + c) * kMarchThreads] = ch[c];
  }
  cl.sync();
  if (seg == 0) {
    for (unsigned r = 1; r < nseg; ++r) {
      const float* rp = cl.map_shared_rank(s_part, r) + threadIdx.x;
      acc = fmaf(trans, rp[0], acc);
      dep = fmaf(trans, rp[kMarchThreads], dep);
#pragma unroll
      for (int c = 0; c < NV; ++c) ch[c] = fmaf(trans, rp[(3 + c) * kMarchThreads], ch[c]);
      trans *= rp[2 * kMarchThreads];
    }
  }
  cl.sync();      // the other CTAs' shared memory must outlive the reads above
  return seg == 0;
}

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
