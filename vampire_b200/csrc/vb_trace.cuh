// vb_trace.cuh -- launch counter + optional per-kernel CUDA-event timing inside libvb200.
//
// bench.py needs (a) how many of OUR kernels ran in the timed region and (b) the device time of
// each kernel family measured with CUDA events on the launching stream (the roofline numerator
// must be timed live, not under a profiler).  Counting is always on (one relaxed atomic per
// launch); event timing only while vb200_trace_enable(1).
#pragma once
#include <cuda_runtime.h>

enum VbKernelId {
  VB_K_GET_PIXEL = 0,
  VB_K_GET_GEOMETRY,
  VB_K_CTX_NHWC,
  VB_K_LIFT_FWD,
  VB_K_LIFT_PLAN,
  VB_K_LIFT_BWD,
  VB_K_PACK,
  VB_K_MARCH_FWD,
  VB_K_BEV_FWD,
  VB_K_MARCH_BWD,
  VB_K_UNPACK_BEV_BWD,
  VB_K_MISC,
  VB_K_COUNT
};

void vb_trace_begin(int kernel_id, cudaStream_t st, int launches = 1);
void vb_trace_end(int kernel_id, cudaStream_t st);

struct VbTraceScope {
  int id;
  cudaStream_t st;
  // `launches` = kernels launched inside the scope (the launch counter reports kernels, not scopes)
  VbTraceScope(int id_, cudaStream_t st_, int launches = 1) : id(id_), st(st_) { vb_trace_begin(id, st, launches); }
  ~VbTraceScope() { vb_trace_end(id, st); }
};
