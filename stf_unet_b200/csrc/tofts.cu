// Extended Tofts pharmacokinetic model, per pixel: forward model and the whole Adam fit as ONE kernel.
//
// Reference: ToftsModelFitter.extended_tofts_model_batch (/root/reference/pk_fitting.py:193-231) and the fitting loop of
// fit_volume_gpu (:233-420, the part between "initial parameter guess" :288 and "param_tensor" :372).
//
//   C(t_i) = vp * aif(t_i) + Ktrans * dt * sum_{j : tc_j < t_i} aif(tc_j) * exp(-Ktrans * (t_i - tc_j) / ve),   tc = arange(0, t_max, dt)
//
// The reference evaluates this with one [batch, <=700] exp tensor per time point and fits (Ktrans, ve, vp) of every tissue
// pixel with torch.optim.Adam over 100 epochs of 1024-pixel batches: ~50 k small launches per slice.  Pixels are
// independent, so here one thread owns one pixel for the entire fit: its three parameters and six Adam moments live in
// registers, the AIF tables in shared memory, and nothing but the final parameters is written.
//
// Faithful to a detail that is easy to miss: the reference drives ONE Adam instance over the full-length parameter
// vectors and steps it once per BATCH.  A pixel outside the current batch receives a zero gradient (the slice's autograd
// gradient is zero-padded), so its moments decay and the pixel still MOVES by the bias-corrected momentum at every batch
// step, and the bias corrections run on the global step count epoch * num_batches + batch + 1; the clamp
// (constrain_params, :302-306) follows every step.  The per-pixel loop below replays exactly that sequence; the loss of a
// batch is the mean over batch_pixels * T residuals (F.mse_loss, :335), which fixes the gradient scale of each pixel.
#include "common.cuh"

namespace stfb {

constexpr int TOFTS_MAX_T = 32;
constexpr int TOFTS_THREADS = 128;

struct ToftsTables {
  const float* t;         // [T] acquisition times
  const float* aif_t;     // [T] aif(t)
  const float* t_conv;    // [M] convolution grid
  const float* aif_conv;  // [M] aif(t_conv)
  const int* nvalid;      // [T] number of grid points with t_conv < t_i (the grid is ascending)
  int T, M;
  float dt;
};

// model value and (optionally) its partial derivatives at time index i; s_* are the shared-memory tables
__device__ __forceinline__ float tofts_point(const float* s_tc, const float* s_ac, int n, float ti, float aif_i, float dt, float K,
                                             float ve, float vp, float* dK, float* dve) {
  float s0 = 0.f, s1 = 0.f;
  for (int j = 0; j < n; ++j) {
    const float tau = ti - s_tc[j];
    const float e = expf((-K * tau) / ve);            // the reference's operation order: (-Ktrans * (ti - t_valid)) / ve
    const float ae = s_ac[j] * e;
    s0 += ae;
    s1 = fmaf(ae, tau, s1);
  }
  const float conv = s0 * dt;
  if (dK) {
    // d/dK [K * dt * sum a e] = dt * sum a e + K * dt * sum a e * (-tau / ve);   d/dve = K * dt * sum a e * (K tau / ve^2)
    const float c1 = s1 * dt;
    *dK = conv - K * c1 / ve;
    *dve = K * K * c1 / (ve * ve);
  }
  return fmaf(K, conv, vp * aif_i);
}

__global__ void __launch_bounds__(TOFTS_THREADS) tofts_forward_kernel(ToftsTables tb, const float* __restrict__ ktrans,
                                                                       const float* __restrict__ ve, const float* __restrict__ vp,
                                                                       float* __restrict__ out, long long N) {
  extern __shared__ float s_tab[];
  float* s_tc = s_tab;
  float* s_ac = s_tab + tb.M;
  for (int j = threadIdx.x; j < tb.M; j += blockDim.x) { s_tc[j] = tb.t_conv[j]; s_ac[j] = tb.aif_conv[j]; }
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const float K = ktrans[p], E = ve[p], V = vp[p];
    for (int i = 0; i < tb.T; ++i) {
      const int n = tb.nvalid[i];
      // a time point before the first grid point keeps the zero the reference initialises `result` with (:210, :216-217)
      out[p * tb.T + i] = n > 0 ? tofts_point(s_tc, s_ac, n, tb.t[i], tb.aif_t[i], tb.dt, K, E, V, nullptr, nullptr) : 0.f;
    }
  }
}

struct ToftsFit {
  int batch_size, epochs, num_batches;
  const float* step_size;   // [epochs * num_batches]  lr / (1 - beta1^s)
  const float* bc2_sqrt;    // [epochs * num_batches]  sqrt(1 - beta2^s)
  float omb1, beta2, omb2, eps;
  float lo[3], hi[3];       // clamp ranges of Ktrans, ve, vp
};

__global__ void __launch_bounds__(TOFTS_THREADS) tofts_fit_kernel(ToftsTables tb, ToftsFit f, const float* __restrict__ pixels,
                                                                   float* __restrict__ ktrans, float* __restrict__ ve,
                                                                   float* __restrict__ vp, float* __restrict__ epoch_loss, long long N) {
  extern __shared__ float s_tab[];
  float* s_tc = s_tab;
  float* s_ac = s_tab + tb.M;
  for (int j = threadIdx.x; j < tb.M; j += blockDim.x) { s_tc[j] = tb.t_conv[j]; s_ac[j] = tb.aif_conv[j]; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int my_batch = (int)(p / f.batch_size);
  const long long b0 = (long long)my_batch * f.batch_size;
  const long long bn = (N - b0 < f.batch_size) ? N - b0 : f.batch_size;       // the last batch is ragged
  const float inv_cnt = 1.f / ((float)bn * (float)tb.T);                      // mse_loss: mean over batch * T residuals
  float y[TOFTS_MAX_T];
  for (int i = 0; i < tb.T; ++i) y[i] = pixels[p * tb.T + i];
  float par[3] = {ktrans[p], ve[p], vp[p]};
  float m[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
  for (int ep = 0; ep < f.epochs; ++ep) {
    for (int b = 0; b < f.num_batches; ++b) {
      float g[3] = {0.f, 0.f, 0.f};
      if (b == my_batch) {
        float sq = 0.f;
        for (int i = 0; i < tb.T; ++i) {
          const int n = tb.nvalid[i];
          float dK = 0.f, dE = 0.f, pred = 0.f, dV = 0.f;
          if (n > 0) {
            pred = tofts_point(s_tc, s_ac, n, tb.t[i], tb.aif_t[i], tb.dt, par[0], par[1], par[2], &dK, &dE);
            dV = tb.aif_t[i];
          }
          const float r = pred - y[i];
          sq = fmaf(r, r, sq);
          const float w = 2.f * r * inv_cnt;
          g[0] = fmaf(w, dK, g[0]); g[1] = fmaf(w, dE, g[1]); g[2] = fmaf(w, dV, g[2]);
        }
        if (epoch_loss) atomicAdd(epoch_loss + ep, sq * inv_cnt / (float)f.num_batches);   // avg over batches of the batch means (:349)
      }
      const int s = ep * f.num_batches + b;
      const float ss = f.step_size[s], bc = f.bc2_sqrt[s];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        m[k] = m[k] + f.omb1 * (g[k] - m[k]);                                // exp_avg.lerp_(grad, 1 - beta1)
        v[k] = f.beta2 * v[k] + f.omb2 * g[k] * g[k];                        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v[k]) / bc + f.eps;
        par[k] -= ss * (m[k] / denom);                                       // param.addcdiv_(exp_avg, denom, value=-step_size)
        par[k] = fminf(fmaxf(par[k], f.lo[k]), f.hi[k]);                     // constrain_params after every optimizer step
      }
    }
  }
  ktrans[p] = par[0]; ve[p] = par[1]; vp[p] = par[2];
}

static int check_tables(const ToftsTables& tb, const char* what) {
  STFB_REQUIRE(tb.t && tb.aif_t && tb.t_conv && tb.aif_conv && tb.nvalid, "%s: null table", what);
  STFB_REQUIRE(tb.T > 0 && tb.T <= TOFTS_MAX_T, "%s: %d time points (1..%d supported)", what, tb.T, TOFTS_MAX_T);
  STFB_REQUIRE(tb.M >= 0 && (size_t)tb.M * 8 <= 200 * 1024, "%s: convolution grid of %d points does not fit shared memory", what, tb.M);
  return STFB_OK;
}

template <typename K>
static int reserve_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    set_error("%s: cannot reserve %zu bytes of shared memory: %s", what, bytes, cudaGetErrorString(cudaGetLastError()));
    return STFB_ECUDA;
  }
  return STFB_OK;
}

}  // namespace stfb

using namespace stfb;

extern "C" int stfb_tofts_forward(const float* t, const float* aif_t, const float* t_conv, const float* aif_conv, const int* nvalid,
                                  int T, int M, float dt, const float* ktrans, const float* ve, const float* vp, float* out,
                                  long long N, void* stream) {
  STFB_DEVICE_OR_RETURN();
  ToftsTables tb{t, aif_t, t_conv, aif_conv, nvalid, T, M, dt};
  int st = check_tables(tb, "tofts_forward");
  if (st != STFB_OK) return st;
  STFB_REQUIRE(N >= 0, "tofts_forward: negative pixel count");
  if (N == 0) return STFB_OK;                              // empty batch: nothing to launch (pointers may be null)
  STFB_REQUIRE(ktrans && ve && vp && out, "tofts_forward: null argument");
  const size_t smem = (size_t)M * 8;
  st = reserve_smem(tofts_forward_kernel, smem, "tofts_forward");
  if (st != STFB_OK) return st;
  long long blocks = (N + TOFTS_THREADS - 1) / TOFTS_THREADS;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  tofts_forward_kernel<<<(unsigned)blocks, TOFTS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tb, ktrans, ve, vp, out, N);
  return post_launch("tofts_forward");
}

extern "C" int stfb_tofts_fit(const float* pixels, const float* t, const float* aif_t, const float* t_conv, const float* aif_conv,
                              const int* nvalid, int T, int M, float dt, float* ktrans, float* ve, float* vp, long long N,
                              int batch_size, int epochs, const float* step_size, const float* bc2_sqrt, float beta1, float beta2,
                              float eps, const float* clamp_lo, const float* clamp_hi, float* epoch_loss, void* stream) {
  STFB_DEVICE_OR_RETURN();
  ToftsTables tb{t, aif_t, t_conv, aif_conv, nvalid, T, M, dt};
  int st = check_tables(tb, "tofts_fit");
  if (st != STFB_OK) return st;
  STFB_REQUIRE(N >= 0 && batch_size > 0 && epochs >= 0, "tofts_fit: bad sizes");
  if (N == 0 || epochs == 0) return STFB_OK;
  STFB_REQUIRE(pixels && ktrans && ve && vp && step_size && bc2_sqrt && clamp_lo && clamp_hi, "tofts_fit: null argument");
  ToftsFit f{};
  f.batch_size = batch_size; f.epochs = epochs; f.num_batches = (int)((N + batch_size - 1) / batch_size);
  f.step_size = step_size; f.bc2_sqrt = bc2_sqrt;
  f.omb1 = 1.f - beta1; f.beta2 = beta2; f.omb2 = 1.f - beta2; f.eps = eps;
  for (int k = 0; k < 3; ++k) { f.lo[k] = clamp_lo[k]; f.hi[k] = clamp_hi[k]; }   // host arrays: three floats each
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (epoch_loss) cudaMemsetAsync(epoch_loss, 0, sizeof(float) * epochs, s);
  const size_t smem = (size_t)M * 8;
  st = reserve_smem(tofts_fit_kernel, smem, "tofts_fit");
  if (st != STFB_OK) return st;
  const long long blocks = (N + TOFTS_THREADS - 1) / TOFTS_THREADS;
  STFB_REQUIRE(blocks <= 2147483647LL, "tofts_fit: too many pixels");
  tofts_fit_kernel<<<(unsigned)blocks, TOFTS_THREADS, smem, s>>>(tb, f, pixels, ktrans, ve, vp, epoch_loss, N);
  return post_launch("tofts_fit");
}
