// Split-precision helpers shared by csrc/split.cu and the batched weight pack (csrc/igemm_simt.cu); see csrc/split.cu.
#pragma once
#include "common.cuh"

namespace stfb {

__device__ __forceinline__ void split3(float x, float& h, float& m, float& l) {
  const __nv_bfloat16 bh = __float2bfloat16_rn(x);
  h = __bfloat162float(bh);
  const float r1 = x - h;                       // exact
  const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
  m = __bfloat162float(bm);
  const float r2 = r1 - m;                      // exact
  l = __bfloat162float(__float2bfloat16_rn(r2));
}

// weight planes that meet the activation planes [lo, hi, mid, mid, hi, hi]
__device__ __forceinline__ float weight_plane(float w, int seg) {
  float h, m, l;
  split3(w, h, m, l);
  // seg: 0 1 2 3 4 5 -> hi lo mid hi mid hi
  return (seg == 0 || seg == 3 || seg == 5) ? h : ((seg == 2 || seg == 4) ? m : l);
}

}  // namespace stfb
