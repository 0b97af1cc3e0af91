// Device-side evaluation metrics of the reference's evaluate() (train_utils/train_and_eval.py:322-336):
// one pass over the logits does argmax (first maximum, like torch.argmax), writes the uint8 tumour mask, adds the pixel
// into the confusion matrix (ConfusionMatrix.update, :30-39) and into the per-class Dice counts of this batch
// (DiceCoefficient.update, :80-118).  The reference spends ~10 launches, a bincount and two host syncs per batch here.
// Integer histogram work, HBM-bound: 4*C bytes of logits + 8 bytes of target per pixel in, 1 byte out.
#include "common.cuh"

namespace stfb {

constexpr int MET_THREADS = 256;
constexpr int MET_MAXC = 8;

// counts layout (unsigned long long): confmat[C*C] (row = target, col = prediction, accumulated across calls),
// then per call: inter[C], pred[C], targ[C] of the Dice update (ignore_index pixels count as class 0 on both sides, as the
// reference's `pred * mask`, `target * mask` does)
__global__ void __launch_bounds__(MET_THREADS) eval_metrics_kernel(const float* __restrict__ logits,
                                                                   const long long* __restrict__ target,
                                                                   unsigned char* __restrict__ mask, unsigned long long* confmat,
                                                                   unsigned long long* dice_counts, int B, int C, int HW,
                                                                   long long ignore_index, int use_ignore) {
  __shared__ unsigned int h[MET_MAXC * MET_MAXC + 3 * MET_MAXC];
  const int nh = C * C + 3 * C;
  for (int i = threadIdx.x; i < nh; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int p = (int)(i - (long long)b * HW);
    const float* lg = logits + (long long)b * C * HW + p;
    float best = lg[0];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = lg[(long long)c * HW];
      if (v > best) { best = v; arg = c; }
    }
    if (mask) mask[i] = (unsigned char)arg;
    const long long t = target[i];
    if (t >= 0 && t < C) atomicAdd(&h[(int)t * C + arg], 1u);
    long long td = t;
    int pd = arg;
    if (use_ignore && t == ignore_index) { td = 0; pd = 0; }
    atomicAdd(&h[C * C + C + pd], 1u);
    if (td >= 0 && td < C) {
      atomicAdd(&h[C * C + 2 * C + (int)td], 1u);
      if (td == pd) atomicAdd(&h[C * C + pd], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nh; i += blockDim.x) {
    const unsigned int v = h[i];
    if (v == 0u) continue;
    if (i < C * C) atomicAdd(confmat + i, (unsigned long long)v);
    else atomicAdd(dice_counts + (i - C * C), (unsigned long long)v);
  }
}

// cumulative_dice[c] += union > 0 ? 2 inter / union : 1 ; count += 1 ; the per-call counts are cleared for the next batch
__global__ void dice_accumulate_kernel(unsigned long long* dice_counts, float* cumulative, long long* count, int C) {
  const int c = threadIdx.x;
  if (c < C) {
    const double inter = (double)dice_counts[c], un = (double)dice_counts[C + c] + (double)dice_counts[2 * C + c];
    cumulative[c] += un > 0.0 ? (float)(2.0 * inter / un) : 1.f;
  }
  __syncthreads();
  if (c < 3 * C) dice_counts[c] = 0ull;
  if (c == 0) *count += 1;
}

}  // namespace stfb

using namespace stfb;

extern "C" int stfb_eval_metrics(const float* logits, const long long* target, unsigned char* mask, long long* confmat,
                                 long long* dice_counts, float* dice_cumulative, long long* dice_updates, int B, int C, int HW,
                                 long long ignore_index, int use_ignore, void* stream) {
  STFB_REQUIRE(((logits && target) || B == 0) && confmat && dice_counts && B >= 0 && C > 0 && C <= MET_MAXC && HW > 0,
               "eval_metrics: bad arguments (1 <= C <= %d)", MET_MAXC);
  STFB_REQUIRE((dice_cumulative == nullptr) == (dice_updates == nullptr), "eval_metrics: cumulative and update count come together");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * HW;
  if (total > 0) {
    long long blocks = (total + MET_THREADS * 8 - 1) / (MET_THREADS * 8);
    if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
    if (blocks < 1) blocks = 1;
    eval_metrics_kernel<<<(unsigned)blocks, MET_THREADS, 0, s>>>(logits, target, mask, reinterpret_cast<unsigned long long*>(confmat),
                                                               reinterpret_cast<unsigned long long*>(dice_counts), B, C, HW,
                                                               ignore_index, use_ignore);
    int rc = post_launch("eval_metrics");
    if (rc != STFB_OK) return rc;
  }
  if (dice_cumulative) {
    dice_accumulate_kernel<<<1, 32, 0, s>>>(reinterpret_cast<unsigned long long*>(dice_counts), dice_cumulative, dice_updates, C);
    return post_launch("eval_metrics(dice)");
  }
  return STFB_OK;
}
