// Shared helpers for libstfb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/stfb200.h"

namespace stfb {

void set_error(const char* fmt, ...);
int check_device();          // STFB_OK or STFB_ENODEV (cached)
int num_sms();
void count_launch(int n = 1);
int post_launch(const char* what);   // cudaGetLastError -> status

#define STFB_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      stfb::set_error(__VA_ARGS__);             \
      return STFB_EINVAL;                       \
    }                                           \
  } while (0)

#define STFB_DEVICE_OR_RETURN()                 \
  do {                                          \
    int _st = stfb::check_device();             \
    if (_st != STFB_OK) return _st;             \
  } while (0)

// ---- 4-element vector load/store with conversion to float ------------------------------------
struct f4 { float v[4]; };

__device__ __forceinline__ f4 ld4(const float* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  return f4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ f4 ld4(const __nv_bfloat16* p) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return f4{{fa.x, fa.y, fb.x, fb.y}};
}
__device__ __forceinline__ void st4(float* p, const f4& a) {
  *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const f4& a) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a.v[0], a.v[1]);
  __nv_bfloat162 hi = __floats2bfloat162_rn(a.v[2], a.v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&lo);
  t.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// 8-element (16 B) bf16 / 2x16 B fp32 loads
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  return f8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ f8 ld8(const __nv_bfloat16* p) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  f8 r;
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    float2 f = __bfloat1622float2(h);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}

__device__ __forceinline__ void st8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 t;
  t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  t.z = *reinterpret_cast<uint32_t*>(&c); t.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// A kernel launched with the programmatic-serialization attribute may START while the previous kernel of its stream is still
// draining (block scheduling, barrier init, TMEM allocation, tensor-map prefetch overlap the predecessor's last wave); it must
// call pdl_wait() before it touches anything the predecessor wrote.  pdl_trigger() lets the successor start early.  Both are
// no-ops for a kernel launched without the attribute.  STFB_PDL = bit mask of the kernel families whose launches carry the
// attribute (1 = convolutions, 2 = BatchNorm apply / backward, 3 = both); 0 = off, the default.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled(int family);   // family bit: 1 = tensor-core convolution kernels, 2 = BatchNorm elementwise kernels

template <int FAMILY, typename... KArgs, typename... Args>
static inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int n = 0;
  if (pdl_enabled(FAMILY)) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    n = 1;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace stfb
