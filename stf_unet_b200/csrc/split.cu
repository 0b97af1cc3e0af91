// Split-precision operands (STFB_BF16X3): fp32-accurate convolutions on the bf16 tensor cores.
//
// An fp32 value is the exact sum of three bf16 numbers up to 2^-27: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)
// (both subtractions are exact in fp32).  A product x * w then needs the six partial products whose order is <= 2,
//     hi*hi + hi*mid + hi*lo + mid*hi + mid*mid + lo*hi            (dropped: mid*lo, lo*mid, lo*lo <= 2^-25 |x w|),
// each product exact in fp32.  The convolution kernels walk a K axis of six segments, activation planes
// [lo, hi, mid, mid, hi, hi] (a plane table in the TMA producer: three planes in memory, six visits) against weight blocks
// [hi, lo, mid, hi, mid, hi] (materialised: weights are small); the weight-gradient kernels walk their pixel range six times
// with the same plane pairs.  The ORDER matters: the fp32 accumulator of tcgen05.mma rounds toward zero once per K = 16
// instruction, a bias of ~half an ulp of the running sum per instruction, so the five correction terms (<= 2^-8 of the result)
// go first, while the accumulator is small, and the hi*hi chain last -- measured 4.3e-6 -> see tests/test_split_gpu.py for a
// 64-channel 3x3 layer (hi*hi first: 216 full-magnitude roundings; last: 36).  What remains is the rounding of the hi*hi chain
// itself (K/16 roundings), the accuracy the bf16 mode's accumulation has as well.
// This file holds the two producers of such operands.
#include "split.cuh"

namespace stfb {

// x [rows][C] fp32 -> y [rows][3C] bf16; one thread = 8 channels (32 B in, 3 x 16 B out) of every (256 / tpr)-th row: a fixed
// channel chunk per thread, rows by striding -- no 64-bit division per element (the first version's i / (C / 8) halved the rate)
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                           long long rows, int C, int tpr) {
  const int c8 = C / 8;
  const int lanes = 256 / tpr;                       // rows per CTA trip
  const int cl = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  if (rl >= lanes) return;
  for (int cv = cl; cv < c8; cv += tpr) {
    const int c = cv * 8;
    for (long long r = (long long)blockIdx.x * lanes + rl; r < rows; r += (long long)gridDim.x * lanes) {
      const f8 v = ld8(x + r * C + c);
      float h[8], m[8], l[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split3(v.v[e], h[e], m[e], l[e]);
      __nv_bfloat16* row = y + r * 3 * C + c;
      st8(row, h);
      st8(row + C, m);
      st8(row + 2 * C, l);
    }
  }
}

// destination [n][(tap, seg, k)] bf16; one work item = one (n, k) position with all its taps and segments
__global__ void __launch_bounds__(256) pack_weight_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int D0,
                                                                int D1, int khw, int k_is_dim1, int flip) {
  const int Kc = k_is_dim1 ? D1 : D0, Nc = k_is_dim1 ? D0 : D1;
  const long long total = (long long)Kc * Nc;
  const long long ld = (long long)khw * 6 * Kc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kc), n = (int)(i / Kc);
    const int d0 = k_is_dim1 ? n : k, d1 = k_is_dim1 ? k : n;
    const float* sp = w + ((long long)d0 * D1 + d1) * khw;
    for (int tap = 0; tap < khw; ++tap) {
      const float v = sp[flip ? (khw - 1 - tap) : tap];
      __nv_bfloat16* dst = wp + (long long)n * ld + (long long)tap * 6 * Kc + k;
#pragma unroll
      for (int seg = 0; seg < 6; ++seg) dst[seg * Kc] = __float2bfloat16_rn(weight_plane(v, seg));
    }
  }
}

}  // namespace stfb

using namespace stfb;

extern "C" int stfb_split_bf16x3(const float* x, void* y, long long rows, int C, void* stream) {
  STFB_REQUIRE(rows >= 0 && C > 0 && C % 8 == 0, "split_bf16x3: C (%d) must be a positive multiple of 8", C);
  if (rows == 0) return STFB_OK;
  STFB_REQUIRE(x && y && (reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(y) % 16) == 0,
               "split_bf16x3: null or misaligned pointer");
  STFB_DEVICE_OR_RETURN();
  const int c8 = C / 8;
  const int tpr = c8 < 256 ? c8 : 256;               // threads per row (a CTA trip covers 256 / tpr rows)
  const int lanes = 256 / tpr;
  long long blocks = (rows + lanes - 1) / lanes;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  split_bf16x3_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), rows, C, tpr);
  return post_launch("split_bf16x3");
}

extern "C" int stfb_pack_weight_split(const float* w, void* wp, int D0, int D1, int kh, int kw, int k_is_dim1, int flip,
                                      void* stream) {
  STFB_REQUIRE(w && wp && D0 > 0 && D1 > 0 && kh > 0 && kw > 0, "pack_weight_split: bad arguments");
  STFB_DEVICE_OR_RETURN();
  const long long total = (long long)D0 * D1;
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  pack_weight_split_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<__nv_bfloat16*>(wp), D0, D1, kh * kw, k_is_dim1, flip);
  return post_launch("pack_weight_split");
}
