// criterion = cross-entropy (mean) + multiclass softmax-Dice, forward and analytic backward.
// One pass over the logits: log-softmax, NLL sum and the per-(image, class) Dice sums reduced with warp
// shuffles, fp64 atomics across CTAs.  No host synchronisation (the reference syncs B*C times per step at
// train_utils/dice_coefficient_loss.py:34).
#include "common.cuh"

namespace stfb {

constexpr int LOSS_THREADS = 256;

// stats layout (fp64): [B][NC][3] = {sum p*t, sum p, sum t} over the VALID pixels of image b, then three scalars:
// sum of w[t] * nll, sum of w[t] (the cross-entropy denominator), number of labels outside [0, NC) that are not
// ignore_index (the reference raises on those: here the loss turns NaN, loudly and without a host synchronisation).
template <int NC>
__global__ void __launch_bounds__(LOSS_THREADS) ce_dice_fwd_kernel(const float* __restrict__ logits,
                                                                    const long long* __restrict__ target,
                                                                    const float* __restrict__ weight, double* stats,
                                                                    int B, int HW, long long ignore_index) {
  const int b = blockIdx.y;
  const float* lg = logits + (long long)b * NC * HW;
  const long long* tg = target + (long long)b * HW;
  float I[NC], Sp[NC], St[NC], nll = 0.f, wsum = 0.f, bad = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) I[c] = Sp[c] = St[c] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const long long tl = tg[p];
    if (tl == ignore_index) continue;                      // ignored pixels enter neither loss (reference :303, dice :27-31)
    if (tl < 0 || tl >= NC) { bad += 1.f; continue; }
    float l[NC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { l[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, l[c]); }
    float e[NC], sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { e[c] = expf(l[c] - mx); sum += e[c]; }
    const float inv = 1.f / sum;
    const int t = (int)tl;
    const float w = weight ? __ldg(weight + t) : 1.f;
    wsum += w;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float pc = e[c] * inv;
      Sp[c] += pc;
      if (c == t) { I[c] += pc; St[c] += 1.f; nll += w * (logf(sum) - (l[c] - mx)); }
    }
  }
  __shared__ float red[LOSS_THREADS / 32][3 * NC + 3];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float a = warp_sum(I[c]), s = warp_sum(Sp[c]), t = warp_sum(St[c]);
    if (lane == 0) { red[warp][3 * c] = a; red[warp][3 * c + 1] = s; red[warp][3 * c + 2] = t; }
  }
  nll = warp_sum(nll); wsum = warp_sum(wsum); bad = warp_sum(bad);
  if (lane == 0) { red[warp][3 * NC] = nll; red[warp][3 * NC + 1] = wsum; red[warp][3 * NC + 2] = bad; }
  __syncthreads();
  if (threadIdx.x < 3 * NC + 3) {
    double v = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) v += (double)red[w][threadIdx.x];
    if (threadIdx.x < 3 * NC) atomicAdd(stats + (long long)b * NC * 3 + threadIdx.x, v);
    else atomicAdd(stats + (long long)B * NC * 3 + (threadIdx.x - 3 * NC), v);
  }
}

__global__ void ce_dice_finalize_kernel(const double* __restrict__ stats, float* loss_out, int B, int C, int HW, float eps,
                                        int with_dice) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double* tail = stats + (long long)B * C * 3;
  const double ce = tail[0] / tail[1];                     // weighted mean over the valid pixels (NaN when there are none,
                                                           // like F.cross_entropy)
  double dice = 0.0;
  for (int c = 0; c < C; ++c) {
    double d = 0.0;
    for (int b = 0; b < B; ++b) {
      const double* s = stats + ((long long)b * C + c) * 3;
      const double sets = s[1] + s[2];
      // dice_coefficient_loss.py:33-37: an empty region of interest (every pixel ignored) scores eps / eps = 1
      d += sets == 0.0 ? 1.0 : (2.0 * s[0] + (double)eps) / (sets + (double)eps);
    }
    dice += d / B;
  }
  dice /= C;
  const double dl = with_dice ? 1.0 - dice : 0.0;
  const bool bad = tail[2] > 0.0;
  const float nanv = __int_as_float(0x7fc00000);
  loss_out[0] = bad ? nanv : (float)(ce + dl);
  loss_out[1] = bad ? nanv : (float)ce;
  loss_out[2] = bad ? nanv : (float)(1.0 - dice);
  (void)HW;
}

template <int NC>
__global__ void __launch_bounds__(LOSS_THREADS) ce_dice_bwd_kernel(const float* __restrict__ logits,
                                                                    const long long* __restrict__ target,
                                                                    const float* __restrict__ weight,
                                                                    const double* __restrict__ stats,
                                                                    const float* __restrict__ dloss, float* __restrict__ dlogits,
                                                                    int B, int HW, float eps, long long ignore_index, int with_dice) {
  const int b = blockIdx.y;
  __shared__ float cA[NC], cB[NC];
  if (threadIdx.x < NC) {
    const double* s = stats + ((long long)b * NC + threadIdx.x) * 3;
    const double S = s[1] + s[2] + (double)eps;
    cA[threadIdx.x] = with_dice ? (float)(2.0 / S) : 0.f;
    cB[threadIdx.x] = with_dice ? (float)((2.0 * s[0] + (double)eps) / (S * S)) : 0.f;
  }
  __syncthreads();
  const float up = dloss ? dloss[0] : 1.f;
  const float inv_W = (float)(1.0 / stats[(long long)B * NC * 3 + 1]);     // 1 / sum of w[t] over the valid pixels
  const float inv_CB = 1.f / ((float)NC * (float)B);
  const float* lg = logits + (long long)b * NC * HW;
  float* dl = dlogits + (long long)b * NC * HW;
  const long long* tg = target + (long long)b * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const long long tl = tg[p];
    if (tl == ignore_index || tl < 0 || tl >= NC) {        // no loss term sees this pixel
#pragma unroll
      for (int c = 0; c < NC; ++c) dl[(long long)c * HW + p] = 0.f;
      continue;
    }
    float l[NC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { l[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, l[c]); }
    float pr[NC], sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { pr[c] = expf(l[c] - mx); sum += pr[c]; }
    const float inv = 1.f / sum;
    const int t = (int)tl;
    const float w = (weight ? __ldg(weight + t) : 1.f) * inv_W;
    float g[NC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      pr[c] *= inv;
      g[c] = -inv_CB * ((c == t ? cA[c] : 0.f) - cB[c]);
      dot += pr[c] * g[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float ce = (pr[c] - (c == t ? 1.f : 0.f)) * w;
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

extern "C" int stfb_ce_dice_fwd_ex(const float* logits, const long long* target, const float* class_weight, double* stats,
                                   float* loss_out, int B, int C, int HW, float eps, long long ignore_index, int with_dice,
                                   void* stream) {
  STFB_REQUIRE(logits && target && stats && loss_out && B > 0 && C > 0 && HW > 0, "ce_dice_fwd: bad arguments");
  STFB_REQUIRE(B <= 65535, "ce_dice_fwd: batch too large");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(stats, 0, sizeof(double) * ((size_t)B * C * 3 + 3), s);
  LOSS_DISPATCH(C, ce_dice_fwd_kernel<NC><<<loss_grid(B, HW), LOSS_THREADS, 0, s>>>(logits, target, class_weight, stats, B, HW,
                                                                                    ignore_index));
  int st = post_launch("ce_dice_fwd");
  if (st != STFB_OK) return st;
  ce_dice_finalize_kernel<<<1, 32, 0, s>>>(stats, loss_out, B, C, HW, eps, with_dice);
  return post_launch("ce_dice_finalize");
}

extern "C" int stfb_ce_dice_bwd_ex(const float* logits, const long long* target, const float* class_weight, const double* stats,
                                   const float* dloss, float* dlogits, int B, int C, int HW, float eps, long long ignore_index,
                                   int with_dice, void* stream) {
  STFB_REQUIRE(logits && target && stats && dlogits && B > 0 && C > 0 && HW > 0, "ce_dice_bwd: bad arguments");
  STFB_REQUIRE(B <= 65535, "ce_dice_bwd: batch too large");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  LOSS_DISPATCH(C, ce_dice_bwd_kernel<NC><<<loss_grid(B, HW), LOSS_THREADS, 0, s>>>(logits, target, class_weight, stats, dloss,
                                                                                    dlogits, B, HW, eps, ignore_index, with_dice));
  return post_launch("ce_dice_bwd");
}

/* the reference's training configuration: no class weights, ignore_index = -100, Dice on */
extern "C" int stfb_ce_dice_fwd(const float* logits, const long long* target, double* stats, float* loss_out, int B, int C,
                                int HW, float eps, void* stream) {
  return stfb_ce_dice_fwd_ex(logits, target, nullptr, stats, loss_out, B, C, HW, eps, -100, 1, stream);
}

extern "C" int stfb_ce_dice_bwd(const float* logits, const long long* target, const double* stats, const float* dloss,
                                float* dlogits, int B, int C, int HW, float eps, void* stream) {
  return stfb_ce_dice_bwd_ex(logits, target, nullptr, stats, dloss, dlogits, B, C, HW, eps, -100, 1, stream);
}
