// criterion = cross-entropy (mean) + multiclass softmax-Dice, forward and analytic backward.
// One pass over the logits: log-softmax, NLL sum and the per-(image, class) Dice sums reduced with warp
// shuffles, fp64 atomics across CTAs.  No host synchronisation (the reference syncs B*C times per step at
// train_utils/dice_coefficient_loss.py:34).
#include "common.cuh"

namespace stfb {

constexpr int LOSS_THREADS = 256;

template <int NC>
__global__ void __launch_bounds__(LOSS_THREADS) ce_dice_fwd_kernel(const float* __restrict__ logits,
                                                                    const long long* __restrict__ target, double* stats,
                                                                    int B, int HW) {
  const int b = blockIdx.y;
  const float* lg = logits + (long long)b * NC * HW;
  const long long* tg = target + (long long)b * HW;
  float I[NC], Sp[NC], St[NC], nll = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) I[c] = Sp[c] = St[c] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float l[NC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { l[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, l[c]); }
    float e[NC], sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { e[c] = expf(l[c] - mx); sum += e[c]; }
    const float inv = 1.f / sum;
    const int t = (int)tg[p];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float pc = e[c] * inv;
      Sp[c] += pc;
      if (c == t) { I[c] += pc; St[c] += 1.f; nll += logf(sum) - (l[c] - mx); }
    }
  }
  __shared__ float red[LOSS_THREADS / 32][3 * NC + 1];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float a = warp_sum(I[c]), s = warp_sum(Sp[c]), t = warp_sum(St[c]);
    if (lane == 0) { red[warp][3 * c] = a; red[warp][3 * c + 1] = s; red[warp][3 * c + 2] = t; }
  }
  nll = warp_sum(nll);
  if (lane == 0) red[warp][3 * NC] = nll;
  __syncthreads();
  if (threadIdx.x < 3 * NC + 1) {
    double v = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) v += (double)red[w][threadIdx.x];
    if (threadIdx.x < 3 * NC) atomicAdd(stats + (long long)b * NC * 3 + threadIdx.x, v);
    else atomicAdd(stats + (long long)B * NC * 3, v);
  }
}

__global__ void ce_dice_finalize_kernel(const double* __restrict__ stats, float* loss_out, int B, int C, int HW, float eps) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double ce = stats[(long long)B * C * 3] / ((double)B * HW);
  double dice = 0.0;
  for (int c = 0; c < C; ++c) {
    double d = 0.0;
    for (int b = 0; b < B; ++b) {
      const double* s = stats + ((long long)b * C + c) * 3;
      d += (2.0 * s[0] + (double)eps) / (s[1] + s[2] + (double)eps);
    }
    dice += d / B;
  }
  dice /= C;
  loss_out[0] = (float)(ce + 1.0 - dice);
  loss_out[1] = (float)ce;
  loss_out[2] = (float)(1.0 - dice);
}

template <int NC>
__global__ void __launch_bounds__(LOSS_THREADS) ce_dice_bwd_kernel(const float* __restrict__ logits,
                                                                    const long long* __restrict__ target,
                                                                    const double* __restrict__ stats,
                                                                    const float* __restrict__ dloss, float* __restrict__ dlogits,
                                                                    int B, int HW, float eps) {
  const int b = blockIdx.y;
  __shared__ float cA[NC], cB[NC];
  if (threadIdx.x < NC) {
    const double* s = stats + ((long long)b * NC + threadIdx.x) * 3;
    const double S = s[1] + s[2] + (double)eps;
    cA[threadIdx.x] = (float)(2.0 / S);
    cB[threadIdx.x] = (float)((2.0 * s[0] + (double)eps) / (S * S));
  }
  __syncthreads();
  const float up = dloss ? dloss[0] : 1.f;
  const float inv_P = 1.f / ((float)B * (float)HW);
  const float inv_CB = 1.f / ((float)NC * (float)B);
  const float* lg = logits + (long long)b * NC * HW;
  float* dl = dlogits + (long long)b * NC * HW;
  const long long* tg = target + (long long)b * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float l[NC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { l[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, l[c]); }
    float pr[NC], sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { pr[c] = expf(l[c] - mx); sum += pr[c]; }
    const float inv = 1.f / sum;
    const int t = (int)tg[p];
    float g[NC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      pr[c] *= inv;
      g[c] = -inv_CB * ((c == t ? cA[c] : 0.f) - cB[c]);
      dot += pr[c] * g[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float ce = (pr[c] - (c == t ? 1.f : 0.f)) * inv_P;
      dl[(long long)c * HW + p] = up * (ce + pr[c] * (g[c] - dot));
    }
  }
}

}  // namespace stfb

using namespace stfb;

#define LOSS_DISPATCH(C, ...)                 \
  switch (C) {                                \
    case 1: { constexpr int NC = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int NC = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int NC = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int NC = 4; __VA_ARGS__; break; } \
    case 5: { constexpr int NC = 5; __VA_ARGS__; break; } \
    case 6: { constexpr int NC = 6; __VA_ARGS__; break; } \
    case 7: { constexpr int NC = 7; __VA_ARGS__; break; } \
    case 8: { constexpr int NC = 8; __VA_ARGS__; break; } \
    default: set_error("ce_dice: num_classes %d not in [1, 8]", C); return STFB_ENOTSUP; \
  }

static dim3 loss_grid(int B, int HW) {
  int bx = (HW + LOSS_THREADS * 4 - 1) / (LOSS_THREADS * 4);
  const int cap = (4 * num_sms() + B - 1) / (B > 0 ? B : 1);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(bx, B);
}

extern "C" int stfb_ce_dice_fwd(const float* logits, const long long* target, double* stats, float* loss_out, int B, int C,
                                int HW, float eps, void* stream) {
  STFB_REQUIRE(logits && target && stats && loss_out && B > 0 && C > 0 && HW > 0, "ce_dice_fwd: bad arguments");
  STFB_REQUIRE(B <= 65535, "ce_dice_fwd: batch too large");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(stats, 0, sizeof(double) * ((size_t)B * C * 3 + 1), s);
  LOSS_DISPATCH(C, ce_dice_fwd_kernel<NC><<<loss_grid(B, HW), LOSS_THREADS, 0, s>>>(logits, target, stats, B, HW));
  int st = post_launch("ce_dice_fwd");
  if (st != STFB_OK) return st;
  ce_dice_finalize_kernel<<<1, 32, 0, s>>>(stats, loss_out, B, C, HW, eps);
  return post_launch("ce_dice_finalize");
}

extern "C" int stfb_ce_dice_bwd(const float* logits, const long long* target, const double* stats, const float* dloss,
                                float* dlogits, int B, int C, int HW, float eps, void* stream) {
  STFB_REQUIRE(logits && target && stats && dlogits && B > 0 && C > 0 && HW > 0, "ce_dice_bwd: bad arguments");
  STFB_REQUIRE(B <= 65535, "ce_dice_bwd: batch too large");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  LOSS_DISPATCH(C, ce_dice_bwd_kernel<NC><<<loss_grid(B, HW), LOSS_THREADS, 0, s>>>(logits, target, stats, dloss, dlogits, B, HW, eps));
  return post_launch("ce_dice_bwd");
}
