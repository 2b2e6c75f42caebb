// vb_softmax.cu -- SURVEY §8f "next" row 1: the producer of the lift's depth input.
//
// Reference: depth_softmax_features = mapping_along_depth(source_features).softmax(dim=1)   (BV2:551)
// on (B*N, D, fH, fW) logits: a softmax over the D = 86 depth planes, i.e. over a dimension whose stride is
// fH*fW elements.  Under the reference's AMP the conv emits fp16 logits and softmax is an autocast-to-fp32 op,
// so: logits in fp32 / bf16 / fp16, arithmetic in fp32, probabilities out in the logits' dtype or in fp32.
//
// HBM roofline: read the logits once, write the probabilities once (2 x 23.25 MB per sample in fp32).  A thread
// owns 4 adjacent pixels (one 8- or 16-byte vector per plane); its D vectors are staged into shared memory with
// cp.async (all D copies of a thread in flight at once, no registers held), then three passes run over the
// thread-private shared-memory column: max, sum of exp(x - max), write exp(x - max) / sum -- the same
// formulation as ATen's softmax, one DRAM read.  The backward stages y and dy the same way and writes
// dx = y * (dy - sum_d y dy).  Shapes that cannot be vectorised (fH*fW % 4 != 0, unaligned pointers, a D too
// large for shared memory) take a scalar kernel that re-reads the planes (L2) instead of staging them.
#include "vb_common.cuh"
#include "vb_trace.cuh"

namespace {

constexpr int kSmxMaxSmem = 96 * 1024;

template <typename T> struct Vec4T;                       // 4 consecutive elements of T as one vector
template <> struct Vec4T<float> { using type = float4; };
template <> struct Vec4T<__nv_bfloat16> { using type = uint2; };
template <> struct Vec4T<__half> { using type = uint2; };

__device__ __forceinline__ void unpack4(const float4& r, float (&o)[4]) { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
template <typename T> __device__ __forceinline__ void unpack4(const uint2& r, float (&o)[4]);
template <> __device__ __forceinline__ void unpack4<__nv_bfloat16>(const uint2& r, float (&o)[4]) {
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack4<__half>(const uint2& r, float (&o)[4]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T> __device__ __forceinline__ void load4(const typename Vec4T<T>::type& r, float (&o)[4]);
template <> __device__ __forceinline__ void load4<float>(const float4& r, float (&o)[4]) { unpack4(r, o); }
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const uint2& r, float (&o)[4]) { unpack4<__nv_bfloat16>(r, o); }
template <> __device__ __forceinline__ void load4<__half>(const uint2& r, float (&o)[4]) { unpack4<__half>(r, o); }

template <typename T> __device__ __forceinline__ typename Vec4T<T>::type pack4(const float (&v)[4]);
template <> __device__ __forceinline__ float4 pack4<float>(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ uint2 pack4<__nv_bfloat16>(const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <> __device__ __forceinline__ uint2 pack4<__half>(const float (&v)[4]) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// exp(v - m) as 2^(v*log2e - m*log2e): one FMA + MUFU.EX2 (2 ulp; results below 2^-126 flush to 0).  The rounding of
// m*log2e is a common factor of every term of a pixel and cancels in the normalisation; the FMA's own rounding is
// <= 4e-8 t relative on a term 2^-t, i.e. <= 2e-8 of the largest term.
__device__ __forceinline__ float exp2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- forward, staged: grid = (pixel-quad tiles, outer), dynamic smem = blockDim.x * D vectors --------------------
template <typename TI, typename TO>
__global__ void depth_softmax_fwd_staged(const TI* __restrict__ x, TO* __restrict__ y, int D, int inner) {
  using VI = typename Vec4T<TI>::type;
  using VO = typename Vec4T<TO>::type;
  extern __shared__ __align__(16) unsigned char s_raw[];
  VI* s = reinterpret_cast<VI*>(s_raw);                        // [D][blockDim.x]
  const int quads = inner >> 2;
  const int qd = blockIdx.x * blockDim.x + threadIdx.x;
  if (qd >= quads) return;
  const size_t base = (size_t)blockIdx.y * D * inner + (size_t)qd * 4;
  const int nt = blockDim.x;
  for (int d = 0; d < D; ++d) cp_async<sizeof(VI)>(&s[d * nt + threadIdx.x], x + base + (size_t)d * inner);
  cp_async_wait_all();
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int d = 0; d < D; ++d) {
    float v[4];
    load4<TI>(s[d * nt + threadIdx.x], v);
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = fmaxf(m[k], v[k]);
  }
  float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, ml[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ml[k] = -m[k] * kLog2e;
  for (int d = 0; d < D; ++d) {
    float v[4];
    load4<TI>(s[d * nt + threadIdx.x], v);
#pragma unroll
    for (int k = 0; k < 4; ++k) sum[k] += exp2_approx(fmaf(v[k], kLog2e, ml[k]));
  }
  float inv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) inv[k] = 1.0f / sum[k];
  for (int d = 0; d < D; ++d) {
    float v[4], o[4];
    load4<TI>(s[d * nt + threadIdx.x], v);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = exp2_approx(fmaf(v[k], kLog2e, ml[k])) * inv[k];
    *reinterpret_cast<VO*>(y + base + (size_t)d * inner) = pack4<TO>(o);
  }
}

// ---- forward, scalar fallback: one thread per pixel, planes re-read -----------------------------------------------
template <typename TI, typename TO>
__global__ void depth_softmax_fwd_scalar(const TI* __restrict__ x, TO* __restrict__ y, int D, int inner) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= inner) return;
  const size_t base = (size_t)blockIdx.y * D * inner + p;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, VbType<TI>::ld(x + base + (size_t)d * inner));
  float sum = 0.0f;
  for (int d = 0; d < D; ++d) sum += expf(VbType<TI>::ld(x + base + (size_t)d * inner) - m);
  for (int d = 0; d < D; ++d)
    y[base + (size_t)d * inner] = VbType<TO>::cvt(expf(VbType<TI>::ld(x + base + (size_t)d * inner) - m) / sum);
}

// ---- backward: dx = y * (dy - sum_d y dy) -----------------------------------------------------------------------------
template <typename TY, typename TO>
__global__ void depth_softmax_bwd_staged(const TY* __restrict__ y, const TY* __restrict__ dy, TO* __restrict__ dx, int D,
                                         int inner) {
  using VY = typename Vec4T<TY>::type;
  using VO = typename Vec4T<TO>::type;
  extern __shared__ __align__(16) unsigned char s_raw[];
  VY* sy = reinterpret_cast<VY*>(s_raw);                       // [D][blockDim.x]
  VY* sg = sy + (size_t)D * blockDim.x;                        // [D][blockDim.x]
  const int quads = inner >> 2;
  const int qd = blockIdx.x * blockDim.x + threadIdx.x;
  if (qd >= quads) return;
  const size_t base = (size_t)blockIdx.y * D * inner + (size_t)qd * 4;
  const int nt = blockDim.x;
  for (int d = 0; d < D; ++d) {
    cp_async<sizeof(VY)>(&sy[d * nt + threadIdx.x], y + base + (size_t)d * inner);
    cp_async<sizeof(VY)>(&sg[d * nt + threadIdx.x], dy + base + (size_t)d * inner);
  }
  cp_async_wait_all();
  float dot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int d = 0; d < D; ++d) {
    float a[4], b[4];
    load4<TY>(sy[d * nt + threadIdx.x], a);
    load4<TY>(sg[d * nt + threadIdx.x], b);
#pragma unroll
    for (int k = 0; k < 4; ++k) dot[k] = fmaf(a[k], b[k], dot[k]);
  }
  for (int d = 0; d < D; ++d) {
    float a[4], b[4], o[4];
    load4<TY>(sy[d * nt + threadIdx.x], a);
    load4<TY>(sg[d * nt + threadIdx.x], b);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = a[k] * (b[k] - dot[k]);
    *reinterpret_cast<VO*>(dx + base + (size_t)d * inner) = pack4<TO>(o);
  }
}

template <typename TY, typename TO>
__global__ void depth_softmax_bwd_scalar(const TY* __restrict__ y, const TY* __restrict__ dy, TO* __restrict__ dx, int D,
                                         int inner) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= inner) return;
  const size_t base = (size_t)blockIdx.y * D * inner + p;
  float dot = 0.0f;
  for (int d = 0; d < D; ++d)
    dot = fmaf(VbType<TY>::ld(y + base + (size_t)d * inner), VbType<TY>::ld(dy + base + (size_t)d * inner), dot);
  for (int d = 0; d < D; ++d)
    dx[base + (size_t)d * inner] = VbType<TO>::cvt(VbType<TY>::ld(y + base + (size_t)d * inner) *
                                                   (VbType<TY>::ld(dy + base + (size_t)d * inner) - dot));
}

// threads per block such that `per_thread` bytes of staging per thread fit the shared-memory budget; 0 = cannot stage
int staged_threads(size_t per_thread) {
  const size_t t = kSmxMaxSmem / per_thread;
  if (t < 32) return 0;
  return t >= 128 ? 128 : (int)(t / 32) * 32;
}

template <typename TI, typename TO>
int launch_softmax_fwd(const void* x, void* y, long long outer, int D, int inner, cudaStream_t st) {
  using VI = typename Vec4T<TI>::type;
  const int nt = staged_threads((size_t)D * sizeof(VI));
  const bool vec = (inner % 4 == 0) && nt > 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  for (long long o0 = 0; o0 < outer; o0 += 65535) {
    const int no = (int)((outer - o0) < 65535 ? (outer - o0) : 65535);
    const TI* xi = reinterpret_cast<const TI*>(x) + (size_t)o0 * D * inner;
    TO* yo = reinterpret_cast<TO*>(y) + (size_t)o0 * D * inner;
    if (vec) {
      const size_t smem = (size_t)nt * D * sizeof(VI);
      static VbPerDeviceFlag attr_set;   // per instantiation AND per device
      if (vb_func_attr_per_device(depth_softmax_fwd_staged<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSmxMaxSmem, attr_set) != VB200_OK)
        return VB200_ERR_CUDA;
      depth_softmax_fwd_staged<TI, TO><<<dim3(vb_ceil_div(inner / 4, nt), no), nt, smem, st>>>(xi, yo, D, inner);
    } else {
      depth_softmax_fwd_scalar<TI, TO><<<dim3(vb_ceil_div(inner, 128), no), 128, 0, st>>>(xi, yo, D, inner);
    }
    VB_LAUNCH_CHECK();
  }
  return VB200_OK;
}

template <typename TY, typename TO>
int launch_softmax_bwd(const void* y, const void* dy, void* dx, long long outer, int D, int inner, cudaStream_t st) {
  using VY = typename Vec4T<TY>::type;
  const int nt = staged_threads((size_t)2 * D * sizeof(VY));
  const bool vec = (inner % 4 == 0) && nt > 0 && (((uintptr_t)y | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0;
  for (long long o0 = 0; o0 < outer; o0 += 65535) {
    const int no = (int)((outer - o0) < 65535 ? (outer - o0) : 65535);
    const size_t off = (size_t)o0 * D * inner;
    const TY* yi = reinterpret_cast<const TY*>(y) + off;
    const TY* gi = reinterpret_cast<const TY*>(dy) + off;
    TO* xo = reinterpret_cast<TO*>(dx) + off;
    if (vec) {
      const size_t smem = (size_t)2 * nt * D * sizeof(VY);
      static VbPerDeviceFlag attr_set;
      if (vb_func_attr_per_device(depth_softmax_bwd_staged<TY, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSmxMaxSmem, attr_set) != VB200_OK)
        return VB200_ERR_CUDA;
      depth_softmax_bwd_staged<TY, TO><<<dim3(vb_ceil_div(inner / 4, nt), no), nt, smem, st>>>(yi, gi, xo, D, inner);
    } else {
      depth_softmax_bwd_scalar<TY, TO><<<dim3(vb_ceil_div(inner, 128), no), 128, 0, st>>>(yi, gi, xo, D, inner);
    }
    VB_LAUNCH_CHECK();
  }
  return VB200_OK;
}

}  // namespace

extern "C" int vb200_depth_softmax_fwd(const void* d_logits, int in_dtype, void* d_probs, int out_dtype,
                                       long long outer, int D, int inner, void* stream) {
  VB_CHECK_ARG(d_logits && d_probs && outer > 0 && D > 0 && inner > 0);
  VB_CHECK_ARG(out_dtype == in_dtype || out_dtype == VB200_F32);
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  VbTraceScope tr(VB_K_MISC, st);
  if (in_dtype == VB200_F32) return launch_softmax_fwd<float, float>(d_logits, d_probs, outer, D, inner, st);
  if (in_dtype == VB200_BF16)
    return out_dtype == VB200_F32 ? launch_softmax_fwd<__nv_bfloat16, float>(d_logits, d_probs, outer, D, inner, st)
                                  : launch_softmax_fwd<__nv_bfloat16, __nv_bfloat16>(d_logits, d_probs, outer, D, inner, st);
  if (in_dtype == VB200_F16)
    return out_dtype == VB200_F32 ? launch_softmax_fwd<__half, float>(d_logits, d_probs, outer, D, inner, st)
                                  : launch_softmax_fwd<__half, __half>(d_logits, d_probs, outer, D, inner, st);
  return VB200_ERR_DTYPE;
}

extern "C" int vb200_depth_softmax_bwd(const void* d_probs, const void* d_gprobs, int dtype, void* d_glogits,
                                       int out_dtype, long long outer, int D, int inner, void* stream) {
  VB_CHECK_ARG(d_probs && d_gprobs && d_glogits && outer > 0 && D > 0 && inner > 0);
  VB_CHECK_ARG(out_dtype == dtype || dtype == VB200_F32);
  int rc = vb200_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  VbTraceScope tr(VB_K_MISC, st);
  if (dtype == VB200_F32) {
    if (out_dtype == VB200_F32) return launch_softmax_bwd<float, float>(d_probs, d_gprobs, d_glogits, outer, D, inner, st);
    if (out_dtype == VB200_BF16)
      return launch_softmax_bwd<float, __nv_bfloat16>(d_probs, d_gprobs, d_glogits, outer, D, inner, st);
    if (out_dtype == VB200_F16) return launch_softmax_bwd<float, __half>(d_probs, d_gprobs, d_glogits, outer, D, inner, st);
    return VB200_ERR_DTYPE;
  }
  if (dtype == VB200_BF16)
    return launch_softmax_bwd<__nv_bfloat16, __nv_bfloat16>(d_probs, d_gprobs, d_glogits, outer, D, inner, st);
  if (dtype == VB200_F16) return launch_softmax_bwd<__half, __half>(d_probs, d_gprobs, d_glogits, outer, D, inner, st);
  return VB200_ERR_DTYPE;
}
