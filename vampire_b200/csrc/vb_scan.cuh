// vb_scan.cuh -- chunked 3-pass exclusive scan of int32 counts (one independent scan per batch sample),
// shared by the lift backward plan (vb_lift_bwd.cu) and the cached lift plan (vb_lift_plan.cu).
#pragma once
#include "vb_common.cuh"

namespace {

// ---- exclusive scan of the per-cell counts: chunked 3-pass scan (a single block per sample took 63 us) ----
constexpr int kScanThreads = 1024;
constexpr int kScanPerThread = 4;
constexpr int kScanChunk = kScanThreads * kScanPerThread;

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += u;
  }
  if (lane == 31) s_warp[wid] = s;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  total = s_warp[31];
  return (wid ? s_warp[wid - 1] : 0) + s - v;
}

// pass 1: per-chunk exclusive scan + chunk totals.  grid = (chunks, B)
__global__ void __launch_bounds__(kScanThreads) scan_chunks_kernel(const int* __restrict__ counts, int* __restrict__ offsets,
                                                                   int* __restrict__ chunk_sums, int nc, int nchunks) {
  __shared__ int s_warp[32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int* in = counts + (size_t)b * nc;
  int* out = offsets + (size_t)b * (nc + 1);
  const int base = chunk * kScanChunk + threadIdx.x * kScanPerThread;
  int v[kScanPerThread], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    v[k] = (base + k < nc) ? in[base + k] : 0;
    sum += v[k];
  }
  int total;
  int excl = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    if (base + k < nc) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == 0) chunk_sums[(size_t)b * nchunks + chunk] = total;
}

// pass 2: exclusive scan of the chunk totals (<= 1024 chunks), one block per sample; also the grand total
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int* __restrict__ chunk_sums, int* __restrict__ offsets,
                                                                 int nc, int nchunks) {
  __shared__ int s_warp[32];
  const int b = blockIdx.x;
  int* sums = chunk_sums + (size_t)b * nchunks;
  const int v = threadIdx.x < nchunks ? sums[threadIdx.x] : 0;
  int total;
  const int excl = block_exclusive_scan(v, s_warp, total);
  if (threadIdx.x < nchunks) sums[threadIdx.x] = excl;
  if (threadIdx.x == 0) offsets[(size_t)b * (nc + 1) + nc] = total;
}

// pass 3: add the chunk base.  grid = (chunks, B)
__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ offsets, const int* __restrict__ chunk_sums,
                                                                int nc, int nchunks) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int add = chunk_sums[(size_t)b * nchunks + chunk];
  int* out = offsets + (size_t)b * (nc + 1);
  const int base = chunk * kScanChunk + threadIdx.x * kScanPerThread;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k)
    if (base + k < nc) out[base + k] += add;
}


// offsets[b][0..n] = exclusive scan of counts[b][0..n-1] (offsets[b][n] = total); chunk_sums: B * 1024 ints of scratch.
// Returns VB200_ERR_ARG when n needs more than 1024 chunks (> 4 M entries per sample).
inline int vb_exclusive_scan(const int* counts, int* offsets, int* chunk_sums, int n, int B, cudaStream_t st) {
  const int nchunks = vb_ceil_div(n, kScanChunk);
  if (nchunks > kScanThreads) return VB200_ERR_ARG;
  scan_chunks_kernel<<<dim3(nchunks, B), kScanThreads, 0, st>>>(counts, offsets, chunk_sums, n, nchunks);
  scan_sums_kernel<<<B, kScanThreads, 0, st>>>(chunk_sums, offsets, n, nchunks);
  scan_add_kernel<<<dim3(nchunks, B), kScanThreads, 0, st>>>(offsets, chunk_sums, n, nchunks);
  if (cudaGetLastError() != cudaSuccess) return VB200_ERR_CUDA;
  return VB200_OK;
}

}  // namespace
