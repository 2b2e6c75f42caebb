// vb_lift_pairs.cuh -- the valid (voxel, camera) pairs of the lift and their destination pixel cells:
// shared by the per-call backward plan (vb_lift_bwd.cu) and the cached plan (vb_lift_plan.cu).
#pragma once
#include "vb_common.cuh"

namespace {

constexpr int kVoxBits = 21;   // cell record key = (z0 + 1) << 21 | voxel  (nvox <= 2^21, D + 1 < 2^11)

struct CellDims {
  int ncy, ncx, nc;   // cells per camera row / col, cells per sample = N * ncy * ncx
};
__host__ __device__ inline CellDims cell_dims(const VbGrid& g) {
  CellDims c;
  c.ncy = g.fH + 1;   // y0 in [-1, fH-1]
  c.ncx = g.fW + 1;
  c.nc = g.N * c.ncy * c.ncx;
  return c;
}

// shared by plan + backward: cull + strict projection of (voxel, camera n); returns validity
__device__ __forceinline__ bool pair_coord(const VbGrid& g, const float* s_m, const float* s_q, bool has_bda,
                                           bool affine, const VbLiftDiv& dv, int n, float px, float py, float pz,
                                           LiftCoord& lc) {
  const float* q = s_q + n * 16;
  const float cz = fmaf(q[8], px, fmaf(q[9], py, fmaf(q[10], pz, q[11])));
  if (!(cz > g.d_lo - 0.05f && cz < g.d_hi + 0.05f)) return false;
  const float cx = fmaf(q[0], px, fmaf(q[1], py, fmaf(q[2], pz, q[3])));
  const float cy = fmaf(q[4], px, fmaf(q[5], py, fmaf(q[6], pz, q[7])));
  const float rz = __frcp_rn(cz);
  const float ux = cx * rz, uy = cy * rz;
  const float* I = s_m + n * VB200_MAT_SLOTS * 16 + 2 * 16;
  const float cw = fmaf(q[12], px, fmaf(q[13], py, fmaf(q[14], pz, q[15])));
  const float ax = fmaf(I[0], ux, fmaf(I[1], uy, fmaf(I[2], cz, I[3] * cw)));
  const float ay = fmaf(I[4], ux, fmaf(I[5], uy, fmaf(I[6], cz, I[7] * cw)));
  // the image-bounds cull is only meaningful strictly in front of the camera: with d_lo = 0 (2-D lift) the guard band
  // reaches behind it, where 1 / cz flips the projection's sign -- there the strict projection alone decides
  if (cz >= 1e-3f && !(ax > -1.5f && ax < g.x_hi + 1.0f && ay > -1.5f && ay < g.y_hi + 1.0f)) return false;
  lc = pair_strict(g, s_m + n * VB200_MAT_SLOTS * 16, has_bda, affine, dv, px, py, pz);
  return lc.valid;
}

__device__ __forceinline__ void stage_cull(float* s_q, const float* s_m, int N, bool has_bda) {
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    const int n = i / 16, r = (i % 16) / 4, c = i % 4;
    const float* A = s_m + n * VB200_MAT_SLOTS * 16 + 16;
    const float* Bm = s_m + n * VB200_MAT_SLOTS * 16;
    float v = A[r * 4 + c];
    if (has_bda) {
      v = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) v = fmaf(A[r * 4 + k], Bm[k * 4 + c], v);
    }
    s_q[i] = v;
  }
}

}  // namespace
