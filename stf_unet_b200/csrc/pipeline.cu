// The two steps either side of the hot path, both pure HBM streaming:
//
//  * stfb_pack_series_u8: the reference's loader turns 8-bit grey images into normalised fp32 tensors on the CPU, one PIL
//    image per phase (/root/reference/my_dataset.py:143-232, transforms.py:120-134 ToTensor + Normalize, train.py:147-148) and the
//    model then re-lays them out.  Here the uint8 series [B,T,H,W] goes to the device as it is (a quarter of the fp32
//    bytes over PCIe) and ONE pass does x/255 -> (v - mean)/std -> dtype, written straight in the encoder's time-major
//    NHWC order [T*B, H, W, 1].  The arithmetic is the reference's, operation for operation (IEEE division, no
//    reciprocal folding), so the fp32 result is bit-identical to ToTensor + Normalize.
//  * stfb_adamw_flat: torch.optim.AdamW(fused=True) (train.py:227-237) walks ~190 parameter tensors through
//    multi-tensor-apply chunks; with parameters, gradients and both moments each in ONE flat fp32 buffer the whole
//    update is a single launch of 16-byte accesses (28 bytes per parameter), fed directly by the all-reduced gradient.
#include "common.cuh"

namespace stfb {

// one CTA row per (b, t) plane: blockIdx.y = b * T + t  ->  destination plane t * B + b.  16 pixels per thread.
template <typename T>
__global__ void __launch_bounds__(256) pack_series_u8_kernel(const unsigned char* __restrict__ x, T* __restrict__ y, int B, int Tn,
                                                             int HW, int nvec, float mean, float stdv) {
  const int plane = blockIdx.y;
  const int b = plane / Tn, t = plane - b * Tn;
  const unsigned char* __restrict__ src = x + (long long)plane * HW;
  T* __restrict__ dst = y + ((long long)t * B + b) * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + (long long)i * 16);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float u = (float)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
      v[j] = ((u / 255.0f) - mean) / stdv;               // ToTensor's div(255), then Normalize's sub(mean).div(std)
    }
    st8(dst + (long long)i * 16, v);
    st8(dst + (long long)i * 16 + 8, v + 8);
  }
  // tail / planes that are not a multiple of 16 pixels (nvec == 0: their starts are not 16-byte aligned)
  for (int i = (nvec << 4) + blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
    st1(dst + i, (((float)src[i] / 255.0f) - mean) / stdv);
}

// p, m, v updated in place from g; all four are flat fp32 arrays of n elements.
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, float lr, float omb1, float beta2, float omb2,
                                                         float eps, float weight_decay, float step_size, float bc2_sqrt,
                                                         float grad_scale) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * grad_scale;
      float pj = P[j];
      pj -= lr * weight_decay * pj;                       // decoupled weight decay
      const float mj = M[j] + omb1 * (gr - M[j]);         // lerp(m, g, 1 - beta1)
      const float vj = beta2 * V[j] + omb2 * gr * gr;
      const float denom = sqrtf(vj) / bc2_sqrt + eps;
      pj -= step_size * mj / denom;
      P[j] = pj; M[j] = mj; V[j] = vj;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * grad_scale;
    float pj = p[i];
    pj -= lr * weight_decay * pj;
    const float mj = m[i] + omb1 * (gr - m[i]);
    const float vj = beta2 * v[i] + omb2 * gr * gr;
    pj -= step_size * mj / (sqrtf(vj) / bc2_sqrt + eps);
    p[i] = pj; m[i] = mj; v[i] = vj;
  }
}

}  // namespace stfb

using namespace stfb;

extern "C" int stfb_pack_series_u8(const unsigned char* x, void* y, int B, int T, int H, int W, float mean, float stdv, int dtype,
                                   void* stream) {
  STFB_REQUIRE(x && y && B >= 0 && T > 0 && H > 0 && W > 0 && (dtype == STFB_F32 || dtype == STFB_BF16), "pack_series_u8: bad arguments");
  STFB_REQUIRE(stdv > 0.f, "pack_series_u8: std must be positive (got %g)", (double)stdv);
  STFB_REQUIRE((long long)H * W < (1LL << 31) && (long long)B * T <= 65535, "pack_series_u8: plane too large or more than 65535 planes");
  const long long HW = (long long)H * W;
  STFB_REQUIRE(HW % 16 == 0 ? ((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0) : true, "pack_series_u8: buffers must be 16-byte aligned");
  STFB_DEVICE_OR_RETURN();
  if (B == 0) return STFB_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int hw = (int)HW, nvec = HW % 16 == 0 ? (int)(HW / 16) : 0;   // unaligned planes take the scalar loop
  int gx = (int)(((nvec ? nvec : HW) + 255) / 256);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, (unsigned)(B * T));
  if (dtype == STFB_F32) pack_series_u8_kernel<float><<<grid, 256, 0, s>>>(x, (float*)y, B, T, hw, nvec, mean, stdv);
  else pack_series_u8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, (__nv_bfloat16*)y, B, T, hw, nvec, mean, stdv);
  return post_launch("pack_series_u8");
}

extern "C" int stfb_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                               double beta1, double beta2, double eps, double weight_decay, long long step, double grad_scale,
                               void* stream) {
  STFB_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "adamw_flat: bad arguments (step counts from 1)");
  STFB_REQUIRE(lr >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && weight_decay >= 0.0,
               "adamw_flat: bad hyper-parameters");
  STFB_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
               "adamw_flat: buffers must be 16-byte aligned");
  STFB_DEVICE_OR_RETURN();
  if (n == 0) return STFB_OK;
  // bias corrections on the host in double, as torch does for a host-side step count
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  // 1 - beta in double, then rounded once (torch passes the Python double 1 - beta as the lerp / addcmul weight)
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adamw_flat_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, (float)lr, omb1, (float)beta2, omb2, (float)eps, (float)weight_decay, step_size, bc2_sqrt,
      (float)grad_scale);
  return post_launch("adamw_flat");
}
