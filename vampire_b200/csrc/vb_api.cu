// vb_api.cu -- version / error strings / device check of libvb200.
#include "vb_common.cuh"

extern "C" int vb200_version(void) { return VB200_VERSION; }

extern "C" const char* vb200_strerror(int status) {
  switch (status) {
    case VB200_OK: return "ok";
    case VB200_ERR_ARG: return "vb200: invalid argument (null pointer, bad size or unsupported channel count)";
    case VB200_ERR_DTYPE: return "vb200: unsupported dtype code";
    case VB200_ERR_ARCH: return "vb200: current CUDA device is not sm_100 (B200); there is no fallback path";
    case VB200_ERR_WORKSPACE: return "vb200: workspace too small";
    case VB200_ERR_CUDA: return "vb200: CUDA runtime / launch failure";
    case VB200_ERR_ALIGN: return "vb200: pointer not 16-byte aligned";
    default: return "vb200: unknown status";
  }
}

extern "C" int vb200_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VB200_ERR_CUDA;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return VB200_ERR_CUDA;
  return major == 10 ? VB200_OK : VB200_ERR_ARCH;
}
