// Memory-bound kernels of the STF-Unet hot path: BatchNorm statistics / apply / backward, column sums,
// max-pool, bilinear resize, LSTM cell update, layout adapters.  All NHWC, 4-wide vector accesses along the
// channel axis when C % 4 == 0, warp-shuffle-free smem column reductions with fp64 global accumulation.
#include "common.cuh"
#include <stdlib.h>

namespace stfb {

static inline bool aligned_to(const void* q, int b) { return (reinterpret_cast<uintptr_t>(q) % b) == 0; }
static inline int grid_for(long long work, int threads = 256, int max_waves = 16) {
  long long b = (work + threads - 1) / threads;
  long long cap = (long long)max_waves * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// =================================================================================================
// column reductions over rows grouped in G contiguous groups of R rows
// =================================================================================================
struct ColRedArgs {
  const void* a;       // MODE 0: x        MODE 1: dy      MODE 2: x
  const void* b;       //                  MODE 1: y (relu mask) or NULL
  const void* c;       //                  MODE 1: x (pre-BN)
  const float* mean;   // [G][C]  (MODE 1)
  const float* invstd; // [G][C]  (MODE 1)
  const float* scale;  // [G][C]  (MODE 1, relu without y: the mask is recomputed as fma(x, scale, shift) > 0)
  const float* shift;  // [G][C]
  float* partial;      // [nblk][2][G][C]  (MODE 0/1): per-CTA partial sums, combined in fp64 by the finalize kernels
  float* outf;         // [C]        (MODE 2)
  int G;
  long long R;
  int C;
  int relu;
  long long rows_per_block;
  int tpr;             // threads per row (power of two <= 256)
};

// VEC-wide (1 / 4 / 8 elements; 8 = 16-byte bf16 accesses) loads and stores with conversion to float
template <int VEC, typename T>
__device__ __forceinline__ void ldv(const T* p, float* v) {
  if constexpr (VEC == 8) {
    f8 t = ld8(p);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = t.v[j];
  } else if constexpr (VEC == 4) {
    f4 t = ld4(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = t.v[j];
  } else {
    v[0] = ld1(p);
  }
}
template <int VEC, typename T>
__device__ __forceinline__ void stv(T* p, const float* v) {
  if constexpr (VEC == 8) st8(p, v);
  else if constexpr (VEC == 4) st4(p, f4{{v[0], v[1], v[2], v[3]}});
  else st1(p, v[0]);
}

template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(256) colreduce_kernel(const ColRedArgs A) {
  __shared__ float red[256 * VEC * 2];
  const int tid = threadIdx.x;
  const int g = blockIdx.y;
  const int tpr = A.tpr, lanes = 256 / tpr;
  const int cl = tid % tpr, rl = tid / tpr;
  const int CVn = A.C / VEC;
  const long long r0 = (long long)blockIdx.x * A.rows_per_block;
  const long long r1 = min(A.R, r0 + A.rows_per_block);
  const T* __restrict__ pa = reinterpret_cast<const T*>(A.a);
  const T* __restrict__ pb = reinterpret_cast<const T*>(A.b);
  const T* __restrict__ pc = reinterpret_cast<const T*>(A.c);

  for (int cv0 = 0; cv0 < CVn; cv0 += tpr) {
    const int cv = cv0 + cl;
    const bool active = cv < CVn;
    float s[VEC], q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
    if (active) {
      const int c = cv * VEC;
      float mu[VEC], is[VEC], sc[VEC], sh[VEC];
      const bool mask_from_x = MODE == 1 && A.relu && pb == nullptr;
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          mu[j] = A.mean[g * A.C + c + j];
          is[j] = A.invstd[g * A.C + c + j];
          sc[j] = mask_from_x ? A.scale[g * A.C + c + j] : 0.f;
          sh[j] = mask_from_x ? A.shift[g * A.C + c + j] : 0.f;
        }
      }
      const long long gbase = (long long)g * A.R;
      // U rows per trip: U (x3 in MODE 1) independent 16-byte loads in flight per thread
      constexpr int U = (MODE == 1) ? 2 : 4;
      long long r = r0 + rl;
      for (; r + (long long)(U - 1) * lanes < r1; r += (long long)U * lanes) {
        float va[U][VEC], vb[U][VEC], vc[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long off = (gbase + r + (long long)u * lanes) * A.C + c;
          ldv<VEC>(pa + off, va[u]);
          if (MODE == 1) {
            ldv<VEC>(pc + off, vc[u]);
            if (A.relu && !mask_from_x) ldv<VEC>(pb + off, vb[u]);
          }
        }
        if (mask_from_x) {
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) vb[u][j] = fmaf(vc[u][j], sc[j], sh[j]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            if (MODE == 1) {
              const float dz = (A.relu && !(vb[u][j] > 0.f)) ? 0.f : va[u][j];
              s[j] += dz;
              q[j] = fmaf(dz, (vc[u][j] - mu[j]) * is[j], q[j]);
            } else {
              s[j] += va[u][j];
              if (MODE == 0) q[j] = fmaf(va[u][j], va[u][j], q[j]);
            }
          }
      }
      for (; r < r1; r += lanes) {
        const long long off = (gbase + r) * A.C + c;
        float va[VEC], vb[VEC], vc[VEC];
        ldv<VEC>(pa + off, va);
        if (MODE == 1) {
          ldv<VEC>(pc + off, vc);
          if (A.relu && !mask_from_x) ldv<VEC>(pb + off, vb);
          if (mask_from_x) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) vb[j] = fmaf(vc[j], sc[j], sh[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if (MODE == 1) {
            const float dz = (A.relu && !(vb[j] > 0.f)) ? 0.f : va[j];
            s[j] += dz;
            q[j] = fmaf(dz, (vc[j] - mu[j]) * is[j], q[j]);
          } else {
            s[j] += va[j];
            if (MODE == 0) q[j] = fmaf(va[j], va[j], q[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      red[(tid * VEC + j) * 2] = s[j];
      red[(tid * VEC + j) * 2 + 1] = q[j];
    }
    __syncthreads();
    // column sums over the row lanes: thread t < tpr*VEC owns column (t / VEC, t % VEC)
    for (int t = tid; t < tpr * VEC; t += 256) {
      const int ccl = t / VEC, j = t - ccl * VEC;
      if (cv0 + ccl < CVn) {
        double ds = 0.0, dq = 0.0;
        for (int l = 0; l < lanes; ++l) {
          ds += (double)red[((l * tpr + ccl) * VEC + j) * 2];
          dq += (double)red[((l * tpr + ccl) * VEC + j) * 2 + 1];
        }
        const int c = (cv0 + ccl) * VEC + j;
        if (MODE == 2) {
          atomicAdd(A.outf + c, (float)ds);
        } else {   // no atomics, no memset: one slot per (CTA, group, channel)
          float* pp = A.partial + ((long long)blockIdx.x * 2 * A.G + g) * A.C + c;
          pp[0] = (float)ds;
          pp[(long long)A.G * A.C] = (float)dq;
        }
      }
    }
    __syncthreads();
  }
}

template <typename T, int MODE>
static void launch_colreduce_t(const ColRedArgs& A, int VEC, dim3 grid, cudaStream_t st) {
  if (VEC == 8) colreduce_kernel<T, 8, MODE><<<grid, 256, 0, st>>>(A);
  else if (VEC == 4) colreduce_kernel<T, 4, MODE><<<grid, 256, 0, st>>>(A);
  else colreduce_kernel<T, 1, MODE><<<grid, 256, 0, st>>>(A);
}

// number of CTAs along the row axis (= number of partial-sum slots per (group, channel)); host-only, deterministic
static int colreduce_blocks(int G, long long R) {
  // few groups (UNet: one group per BatchNorm): 64 slots x G CTAs leave most SMs idle (1.2 TB/s on a 67 MB fp32 map,
  // profiles/r02_unet_fp32_families.txt) -- four CTAs per SM then, up to 512 slots, which the finalize kernels walk with
  // FIN_WIDE_SL lanes per (group, channel); >= 4 groups keep the 64-slot cap and the one-thread walk
  const bool few = G < 4;
  long long want = ((few ? 4LL : 2LL) * num_sms() + G - 1) / G;     // ~2 (4) waves in total
  long long cap = (R + 31) / 32;                      // at least 32 rows per CTA
  long long n = want < cap ? want : cap;
  const long long lim = few ? 512 : 64;
  if (n > lim) n = lim;
  if (n < 1) n = 1;
  return (int)n;
}

template <int MODE>
static int launch_colreduce(ColRedArgs A, int dtype, int VEC, cudaStream_t st, const char* what, int nblk = 0) {
  if (A.R == 0 || A.G == 0) return STFB_OK;
  const int CVn = A.C / VEC;
  int tpr = 1;
  while (tpr * 2 <= CVn && tpr * 2 <= 256) tpr *= 2;
  A.tpr = tpr;
  if (MODE == 2) {
    long long n = (A.R + 255) / 256;
    const long long cap = 4LL * num_sms();
    nblk = (int)(n < 1 ? 1 : (n > cap ? cap : n));
  } else if (nblk <= 0) {
    nblk = colreduce_blocks(A.G, A.R);
  }
  A.rows_per_block = (A.R + nblk - 1) / nblk;
  dim3 grid((unsigned)nblk, (unsigned)A.G);
  if (dtype == STFB_F32) launch_colreduce_t<float, MODE>(A, VEC, grid, st);
  else launch_colreduce_t<__nv_bfloat16, MODE>(A, VEC, grid, st);
  return post_launch(what);
}

// =================================================================================================
// BatchNorm finalize / fold / apply / backward apply
// =================================================================================================
// Finalize kernels: one thread per (group, channel) adds the nblk partial slots (coalesced along c, independent loads,
// fp64 accumulate), the block's G x 32 results meet in shared memory and the 32 threads of group 0 walk the groups in
// order (the reference updates the running stats once per time step, sequentially).  The first version gave each channel
// one warp with the slots strided over the lanes: 23 us per launch for a few KB of work (profiles/r01_launches_*).
constexpr int FIN_CH = 32;       // channels per block
constexpr int FIN_MAXG = 32;     // groups per block pass (blockDim = FIN_CH * min(G, FIN_MAXG))

__device__ __forceinline__ void sum_slots(const float* __restrict__ partial, int nblk, int G, int C, int g, int c, double& s,
                                          double& q) {
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  const long long slot = 2LL * G * C;
  const float* p = partial + (long long)g * C + c;
  if (nblk == 8) {
    // the conv epilogue's eight slots: all sixteen loads are issued before the first add (the rolled loop below pays one
    // L2 round trip per pair of slots: ~3 us of pure latency at the head of every BatchNorm-apply CTA); same add order
    float sv[8], qv[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) { sv[b] = p[(long long)b * slot]; qv[b] = p[(long long)b * slot + (long long)G * C]; }
#pragma unroll
    for (int b = 0; b < 8; b += 2) { a0 += (double)sv[b]; b0 += (double)qv[b]; a1 += (double)sv[b + 1]; b1 += (double)qv[b + 1]; }
    s = a0 + a1;
    q = b0 + b1;
    return;
  }
  int b = 0;
  for (; b + 1 < nblk; b += 2) {
    const float s0 = p[(long long)b * slot], q0 = p[(long long)b * slot + (long long)G * C];
    const float s1 = p[(long long)(b + 1) * slot], q1 = p[(long long)(b + 1) * slot + (long long)G * C];
    a0 += (double)s0; b0 += (double)q0; a1 += (double)s1; b1 += (double)q1;
  }
  if (b < nblk) { a0 += (double)p[(long long)b * slot]; b0 += (double)p[(long long)b * slot + (long long)G * C]; }
  s = a0 + a1;
  q = b0 + b1;
}

// scale / shift of (group g, channel c) from the partial slots: THE formula -- bn_finalize_train_kernel and the
// self-finalizing bn_apply_kernel both call it, so the two routes give bit-identical coefficients
__device__ __forceinline__ void bn_train_coefs(const float* __restrict__ partial, int nblk, int G, int C, int g, int c, double n,
                                               float gamma_c, float beta_c, float eps, float& sc, float& sh, double& m, double& var,
                                               float& is) {
  double s, q;
  sum_slots(partial, nblk, G, C, g, c, s, q);
  m = s / n;
  var = q / n - m * m;
  if (var < 0.0) var = 0.0;
  is = rsqrtf((float)var + eps);
  sc = gamma_c * is;
  sh = beta_c - (float)m * sc;
}

// more than 64 slots (few groups, see colreduce_blocks): FIN_WIDE_SL threads share the slot walk of one (group, channel)
constexpr int FIN_WIDE_SL = 32;     // one group per pass: blockDim = FIN_CH * FIN_WIDE_SL

// SL = 1: blockDim = FIN_CH * groups per pass.  SL = FIN_WIDE_SL: blockDim = FIN_CH * groups per pass * SL, lane sl of a
// (group, channel) sums slots sl, sl + SL, ... in fp64 and lane 0 adds the lane sums in order (deterministic).
template <int SL>
__global__ void __launch_bounds__(FIN_CH * FIN_MAXG) bn_finalize_train_kernel(
    const float* __restrict__ partial, int nblk, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* running_mean, float* running_var, long long* nbt, float* scale, float* shift, float* mean, float* invstd, int G,
    long long R, int C, float eps, float momentum) {
  __shared__ float sh_m[FIN_MAXG][FIN_CH], sh_v[FIN_MAXG][FIN_CH];
  __shared__ double lane_s[SL > 1 ? FIN_MAXG : 1][FIN_CH], lane_q[SL > 1 ? FIN_MAXG : 1][FIN_CH];
  const int gpb = blockDim.x / (FIN_CH * SL);
  const int cl = threadIdx.x % FIN_CH, gl = (threadIdx.x / FIN_CH) % gpb, sl = threadIdx.x / (FIN_CH * gpb);
  const int c = blockIdx.x * FIN_CH + cl;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += G;
  float rm = 0.f, rv = 1.f;
  if (gl == 0 && sl == 0 && c < C) { rm = running_mean ? running_mean[c] : 0.f; rv = running_var ? running_var[c] : 1.f; }
  const double n = (double)R;
  for (int g0 = 0; g0 < G; g0 += gpb) {
    const int g = g0 + gl;
    if constexpr (SL > 1) {
      double ls = 0.0, lq = 0.0;
      if (g < G && c < C) {
        const long long slot = 2LL * G * C;
        const float* p = partial + (long long)g * C + c;
        double s1 = 0.0, q1 = 0.0;
        int b = sl;
        for (; b + SL < nblk; b += 2 * SL) {
          const float a0 = p[(long long)b * slot], b0 = p[(long long)b * slot + (long long)G * C];
          const float a1 = p[(long long)(b + SL) * slot], b1 = p[(long long)(b + SL) * slot + (long long)G * C];
          ls += (double)a0; lq += (double)b0; s1 += (double)a1; q1 += (double)b1;
        }
        if (b < nblk) { ls += (double)p[(long long)b * slot]; lq += (double)p[(long long)b * slot + (long long)G * C]; }
        ls += s1; lq += q1;
      }
      lane_s[sl * gpb + gl][cl] = ls;
      lane_q[sl * gpb + gl][cl] = lq;
      __syncthreads();
    }
    if (sl == 0 && g < G && c < C) {
      float sc, sh, is;
      double m, var;
      if constexpr (SL > 1) {
        double s = 0.0, q = 0.0;
        for (int j = 0; j < SL; ++j) { s += lane_s[j * gpb + gl][cl]; q += lane_q[j * gpb + gl][cl]; }
        m = s / n;
        var = q / n - m * m;
        if (var < 0.0) var = 0.0;
        is = rsqrtf((float)var + eps);
        sc = gamma[c] * is;
        sh = beta[c] - (float)m * sc;
      } else {
        bn_train_coefs(partial, nblk, G, C, g, c, n, gamma[c], beta[c], eps, sc, sh, m, var, is);
      }
      scale[g * C + c] = sc;
      shift[g * C + c] = sh;
      mean[g * C + c] = (float)m;
      invstd[g * C + c] = is;
      sh_m[gl][cl] = (float)m;
      sh_v[gl][cl] = (float)(R > 1 ? var * n / (n - 1.0) : var);
    }
    __syncthreads();
    if (gl == 0 && sl == 0 && c < C) {
      for (int j = 0; j < gpb && g0 + j < G; ++j) {
        rm = (1.f - momentum) * rm + momentum * sh_m[j][cl];
        rv = (1.f - momentum) * rv + momentum * sh_v[j][cl];
      }
    }
    __syncthreads();
  }
  if (gl == 0 && sl == 0 && c < C) {
    if (running_mean) running_mean[c] = rm;
    if (running_var) running_var[c] = rv;
  }
}

__global__ void bn_fold_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv, float* scale,
                                    float* shift, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

// Row-tiled elementwise kernels: thread (rl, cl) of a CTA owns channels [cl*VEC, cl*VEC + VEC) of every (256/tpr)-th row
// of its row range inside ONE group (blockIdx.y), so the per-(group, channel) coefficients sit in registers and the
// loop body is U independent 16-byte loads per tensor followed by the stores.
// the coefficient table of one group in shared memory ([2][C]: scale, shift); a separate function so that its fp64 slot
// sums do not share a register allocation with the streaming loop of the caller
__device__ __noinline__ void bn_group_coefs_to_smem(const float* __restrict__ partial, int nblk, int G, int C, int g, long long R,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                    float* coef_sm) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sc_, sh_;
    if (partial) {
      float is_;
      double m_, v_;
      bn_train_coefs(partial, nblk, G, C, g, c, (double)R, gamma[c], beta[c], eps, sc_, sh_, m_, v_, is_);
    } else {
      sc_ = scale[g * C + c];
      sh_ = shift[g * C + c];
    }
    coef_sm[c] = sc_;
    coef_sm[C + c] = sh_;
  }
}

// partial != nullptr: the CTA derives scale/shift of its group itself from the (few) statistics slots the conv epilogue
// filled -- the finalize kernel (6 us of pure latency per layer, 44 layers) leaves the critical path and runs beside it,
// only to update the running statistics and to save mean / invstd / scale / shift for the backward pass.
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 4) bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const T* res, T* y, long long R, int C,
                                                       long long rows_per_block, int tpr, int relu,
                                                       const float* __restrict__ partial, int nblk,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  extern __shared__ float coef_sm[];                 // [2][C]: scale, shift of this CTA's group
  pdl_trigger();                                     // programmatic dependent launch (common.cuh): no-ops without the attribute
  pdl_wait();
  constexpr int U = 2;
  const int lanes = 256 / tpr;
  const int cl = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int g = blockIdx.y;
  const int CVn = C / VEC;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  const long long gbase = (long long)g * R;
  bn_group_coefs_to_smem(partial, nblk, gridDim.y, C, g, R, gamma, beta, eps, scale, shift, coef_sm);
  __syncthreads();
  for (int cv = cl; cv < CVn; cv += tpr) {
    const int c = cv * VEC;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sc[j] = coef_sm[c + j]; sh[j] = coef_sm[C + c + j]; }
    long long r = r0 + rl;
    for (; r + (long long)(U - 1) * lanes < r1; r += (long long)U * lanes) {
      float v[U][VEC], rr[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long off = (gbase + r + (long long)u * lanes) * C + c;
        ldv<VEC>(x + off, v[u]);
        if (res) ldv<VEC>(res + off, rr[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float t = fmaf(v[u][j], sc[j], sh[j]);
          if (res) t += rr[u][j];
          v[u][j] = relu ? fmaxf(t, 0.f) : t;
        }
        stv<VEC>(y + (gbase + r + (long long)u * lanes) * C + c, v[u]);
      }
    }
    for (; r < r1; r += lanes) {
      const long long off = (gbase + r) * C + c;
      float v[VEC], rr[VEC];
      ldv<VEC>(x + off, v);
      if (res) ldv<VEC>(res + off, rr);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float t = fmaf(v[j], sc[j], sh[j]);
        if (res) t += rr[j];
        v[j] = relu ? fmaxf(t, 0.f) : t;
      }
      stv<VEC>(y + off, v);
    }
  }
}

// backward finalize: up to 64 partial slots per (group, channel) -> FIN_SL threads share the slot walk
// (few groups with up to 512 slots: 32 lanes x 1 group per pass)
template <int FIN_SL, int FIN_BG>   // groups per block pass: blockDim = FIN_CH * FIN_BG * FIN_SL = 1024
__global__ void __launch_bounds__(FIN_CH * FIN_BG * FIN_SL) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk,
                                                                                  const float* __restrict__ gamma,
                                                                                  const float* __restrict__ invstd, float* dgamma,
                                                                                  float* dbeta, float* coef, int G, long long R,
                                                                                  int C) {
  __shared__ float sh_s[FIN_SL][FIN_BG][FIN_CH], sh_q[FIN_SL][FIN_BG][FIN_CH];
  const int cl = threadIdx.x % FIN_CH;
  const int gl = (threadIdx.x / FIN_CH) % FIN_BG;
  const int sl = threadIdx.x / (FIN_CH * FIN_BG);
  const int c = blockIdx.x * FIN_CH + cl;
  double dg = 0.0, db = 0.0;
  const double n = (double)R;
  const long long slot = 2LL * G * C;
  for (int g0 = 0; g0 < G; g0 += FIN_BG) {
    const int g = g0 + gl;
    float s = 0.f, q = 0.f;
    if (g < G && c < C) {
      const float* p = partial + (long long)g * C + c;
      float s1 = 0.f, q1 = 0.f;
      int b = sl;
      for (; b + FIN_SL < nblk; b += 2 * FIN_SL) {
        s += p[(long long)b * slot]; q += p[(long long)b * slot + (long long)G * C];
        s1 += p[(long long)(b + FIN_SL) * slot]; q1 += p[(long long)(b + FIN_SL) * slot + (long long)G * C];
      }
      if (b < nblk) { s += p[(long long)b * slot]; q += p[(long long)b * slot + (long long)G * C]; }
      s += s1; q += q1;
    }
    sh_s[sl][gl][cl] = s;
    sh_q[sl][gl][cl] = q;
    __syncthreads();
    if (sl == 0 && g < G && c < C) {
      double ds = 0.0, dq = 0.0;
#pragma unroll
      for (int j = 0; j < FIN_SL; ++j) { ds += (double)sh_s[j][gl][cl]; dq += (double)sh_q[j][gl][cl]; }
      float* k = coef + ((long long)g * C + c) * 3;
      k[0] = gamma[c] * invstd[g * C + c];
      k[1] = (float)(ds / n);
      k[2] = (float)(dq / n);
      sh_s[0][gl][cl] = (float)ds;
      sh_q[0][gl][cl] = (float)dq;
    }
    __syncthreads();
    if (sl == 0 && gl == 0 && c < C) {
      for (int j = 0; j < FIN_BG && g0 + j < G; ++j) { db += (double)sh_s[0][j][cl]; dg += (double)sh_q[0][j][cl]; }
    }
    __syncthreads();
  }
  if (sl == 0 && gl == 0 && c < C) {
    if (dgamma) dgamma[c] += (float)dg;
    if (dbeta) dbeta[c] += (float)db;
  }
}

// dx = k0 * (dz - k1 - xhat * k2) with xhat = (x - mean) * invstd, folded per channel into dx = k0*dz + cb*x + cc
// (three coefficient registers per channel instead of five: the kernel has to fit 4 CTAs of 256 threads per SM).
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 4) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                              const T* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ coef,
                                                              T* dx, T* dres, long long R, int C, long long rows_per_block, int tpr,
                                                              int relu, int accum_dres, const float* __restrict__ shift) {
  constexpr int U = 1;
  const int lanes = 256 / tpr;
  const int cl = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int g = blockIdx.y;
  const int CVn = C / VEC;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  const long long gbase = (long long)g * R;
  for (int cv = cl; cv < CVn; cv += tpr) {
    const int c = cv * VEC;
    // relu without y: the mask is fma(x, scale, shift) > 0 with scale = gamma * invstd = coef[0] (what bn_apply computed)
    const bool mask_from_x = relu && y == nullptr;
    float k0[VEC], cb[VEC], cc[VEC], sh[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int gc = g * C + c + j;
      const float mu = mean[gc], is = invstd[gc];
      const float a0 = coef[(long long)gc * 3], a1 = coef[(long long)gc * 3 + 1], a2 = coef[(long long)gc * 3 + 2];
      k0[j] = a0;
      cb[j] = -a0 * a2 * is;
      cc[j] = a0 * (mu * is * a2 - a1);
      sh[j] = mask_from_x ? shift[gc] : 0.f;
    }
    for (long long r = r0 + rl; r < r1; r += (long long)U * lanes) {
      float d[U][VEC], xv[U][VEC], yv[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * lanes;
        if (rr < r1) {
          const long long off = (gbase + rr) * C + c;
          ldv<VEC>(dy + off, d[u]);
          ldv<VEC>(x + off, xv[u]);
          if (relu && !mask_from_x) ldv<VEC>(y + off, yv[u]);
          if (mask_from_x) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) yv[u][j] = fmaf(xv[u][j], k0[j], sh[j]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * lanes;
        if (rr < r1) {
          const long long off = (gbase + rr) * C + c;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float dz = (relu && !(yv[u][j] > 0.f)) ? 0.f : d[u][j];
            d[u][j] = dz;
            xv[u][j] = fmaf(k0[j], dz, fmaf(cb[j], xv[u][j], cc[j]));
          }
          stv<VEC>(dx + off, xv[u]);
          if (dres) {
            if (accum_dres) {
              float old[VEC];
              ldv<VEC>(dres + off, old);
#pragma unroll
              for (int j = 0; j < VEC; ++j) d[u][j] += old[j];
            }
            stv<VEC>(dres + off, d[u]);
          }
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// BatchNorm backward in ONE launch: reduce -> per-group barrier -> finalize -> apply.
//
// The three-launch chain (colreduce<1> -> bn_bwd_finalize -> bn_bwd_apply) costs two extra kernel boundaries and a
// latency-bound finalize per layer: measured 29-36 us for the layer-3/4 maps whose HBM time is 7-10 us
// (tools/kernel_probe.py bn2), 44 times per step.  Here a single-wave grid (every CTA co-resident) does
//   phase 1: each CTA reduces (sum dz, sum dz*xhat) over ITS rows of its group and adds them into acc[g][2][C]
//            (fp32 red.add into a zeroed buffer), then arrives at the group's counter;
//   barrier: spins until all CTAs of the group arrived (the group's sums are complete);
//   phase 2: every thread forms the coefficients of its own channels from acc (what the finalize kernel did) and
//            applies dx = k0*dz + cb*x + cc over the SAME rows -- for the deep layers these are L2 hits.
// CTA 0 of each group adds the group's sums into dgamma / dbeta.  The spin is safe because the grid never exceeds the
// co-resident capacity the occupancy API reports; CTAs delayed by unrelated kernels on other streams only make the
// others wait.  Two such kernels spinning on the same device at the same time (two models trained concurrently on two
// streams of one process) could starve each other: STFB_NO_FUSED_BN_BWD=1 restores the three-launch chain.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// LEAN: no y and no dres are touched (mask recomputed from x) -> two tensors per pass, so four rows per trip stay in flight
template <typename T, int VEC, bool LEAN>
__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                              const T* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ shift, float* acc, unsigned* counters,
                                                              float* dgamma, float* dbeta, T* dx, T* dres, long long R, int C,
                                                              long long rows_per_block, int tpr, int relu, int accum_dres) {
  __shared__ float red[256 * VEC * 2];
  pdl_trigger();                                     // programmatic dependent launch (common.cuh): no-ops without the attribute
  pdl_wait();
  const int tid = threadIdx.x;
  const int lanes = 256 / tpr;
  const int cl = tid % tpr, rl = tid / tpr;
  const int g = blockIdx.y;
  const int CVn = C / VEC;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  const long long gbase = (long long)g * R;
  const bool mask_from_x = relu && y == nullptr;
  float* accs = acc + (long long)g * 2 * C;          // [2][C] of this group

  // ---------------- phase 1: partial sums over this CTA's rows ----------------
  for (int cv0 = 0; cv0 < CVn; cv0 += tpr) {
    const int cv = cv0 + cl;
    float s[VEC], q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
    if (cv < CVn) {
      const int c = cv * VEC;
      float mu[VEC], is[VEC], sc[VEC], sh[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int gc = g * C + c + j;
        mu[j] = mean[gc];
        is[j] = invstd[gc];
        sc[j] = mask_from_x ? gamma[c + j] * is[j] : 0.f;     // the forward's scale = gamma * invstd
        sh[j] = mask_from_x ? shift[gc] : 0.f;
      }
      constexpr int U = LEAN ? 4 : 2;
      long long r = r0 + rl;
      for (; r + (long long)(U - 1) * lanes < r1; r += (long long)U * lanes) {
        float va[U][VEC], vb[U][VEC], vc[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long off = (gbase + r + (long long)u * lanes) * C + c;
          ldv<VEC>(dy + off, va[u]);
          ldv<VEC>(x + off, vc[u]);
          if (!LEAN && relu && !mask_from_x) ldv<VEC>(y + off, vb[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float yy = mask_from_x ? fmaf(vc[u][j], sc[j], sh[j]) : ((!LEAN && relu) ? vb[u][j] : 1.f);
            const float dz = (relu && !(yy > 0.f)) ? 0.f : va[u][j];
            s[j] += dz;
            q[j] = fmaf(dz, (vc[u][j] - mu[j]) * is[j], q[j]);
          }
      }
      for (; r < r1; r += lanes) {
        const long long off = (gbase + r) * C + c;
        float va[VEC], vb[VEC], vc[VEC];
        ldv<VEC>(dy + off, va);
        ldv<VEC>(x + off, vc);
        if (!LEAN && relu && !mask_from_x) ldv<VEC>(y + off, vb);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float yy = mask_from_x ? fmaf(vc[j], sc[j], sh[j]) : ((!LEAN && relu) ? vb[j] : 1.f);
          const float dz = (relu && !(yy > 0.f)) ? 0.f : va[j];
          s[j] += dz;
          q[j] = fmaf(dz, (vc[j] - mu[j]) * is[j], q[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      red[(tid * VEC + j) * 2] = s[j];
      red[(tid * VEC + j) * 2 + 1] = q[j];
    }
    __syncthreads();
    for (int t = tid; t < tpr * VEC; t += 256) {
      const int ccl = t / VEC, j = t - ccl * VEC;
      if (cv0 + ccl < CVn) {
        double ds = 0.0, dq = 0.0;
        for (int l = 0; l < lanes; ++l) {
          ds += (double)red[((l * tpr + ccl) * VEC + j) * 2];
          dq += (double)red[((l * tpr + ccl) * VEC + j) * 2 + 1];
        }
        const int c = (cv0 + ccl) * VEC + j;
        atomicAdd(accs + c, (float)ds);
        atomicAdd(accs + C + c, (float)dq);
      }
    }
    __syncthreads();
  }

  // ---------------- barrier over the CTAs of this group ----------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(counters + g, 1u);
    while (ld_acquire_u32(counters + g) < gridDim.x) __nanosleep(40);
  }
  __syncthreads();

  // ---------------- phase 2: coefficients of my channels, then the apply pass over the same rows ----------------
  const float inv_n = 1.f / (float)R;
  if (blockIdx.x == 0) {
    for (int c = tid; c < C; c += 256) {
      const float sg = __ldcg(accs + c), qg = __ldcg(accs + C + c);
      if (dbeta) atomicAdd(dbeta + c, sg);
      if (dgamma) atomicAdd(dgamma + c, qg);
    }
  }
  for (int cv = cl; cv < CVn; cv += tpr) {
    const int c = cv * VEC;
    float k0[VEC], cb[VEC], cc[VEC], sh[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int gc = g * C + c + j;
      const float mu = mean[gc], is = invstd[gc];
      const float a0 = gamma[c + j] * is, a1 = __ldcg(accs + c + j) * inv_n, a2 = __ldcg(accs + C + c + j) * inv_n;
      k0[j] = a0;
      cb[j] = -a0 * a2 * is;
      cc[j] = a0 * (mu * is * a2 - a1);
      sh[j] = mask_from_x ? shift[gc] : 0.f;
    }
    constexpr int U = LEAN ? 4 : 2;
    for (long long r = r0 + rl; r < r1; r += (long long)U * lanes) {
      float d[U][VEC], xv[U][VEC], yv[U][VEC], old[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * lanes;
        if (rr < r1) {
          const long long off = (gbase + rr) * C + c;
          ldv<VEC>(dy + off, d[u]);
          ldv<VEC>(x + off, xv[u]);
          if (!LEAN && relu && !mask_from_x) ldv<VEC>(y + off, yv[u]);
          if (!LEAN && dres && accum_dres) ldv<VEC>(dres + off, old[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * lanes;
        if (rr < r1) {
          const long long off = (gbase + rr) * C + c;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float yy = mask_from_x ? fmaf(xv[u][j], k0[j], sh[j]) : ((!LEAN && relu) ? yv[u][j] : 1.f);
            const float dz = (relu && !(yy > 0.f)) ? 0.f : d[u][j];
            d[u][j] = dz;
            xv[u][j] = fmaf(k0[j], dz, fmaf(cb[j], xv[u][j], cc[j]));
          }
          stv<VEC>(dx + off, xv[u]);
          if (!LEAN && dres) {
            if (accum_dres) {
#pragma unroll
              for (int j = 0; j < VEC; ++j) d[u][j] += old[u][j];
            }
            stv<VEC>(dres + off, d[u]);
          }
        }
      }
    }
  }
}

// =================================================================================================
// max-pool
// =================================================================================================
template <typename T, int VEC>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo,
                                   int k, int stride, int pad, long long total_vec) {
  const int CVn = C / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CVn);
    long long r = i / CVn;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float best[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) best[j] = -INFINITY;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const long long off = (((long long)n * H + iy) * W + ix) * C + cv * VEC;
        if (VEC == 4) {
          f4 v = ld4(x + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], v.v[j]);
        } else {
          best[0] = fmaxf(best[0], ld1(x + off));
        }
      }
    }
    const long long oo = (((long long)n * Ho + oy) * Wo + ox) * C + cv * VEC;
    if (VEC == 4) st4(y + oo, f4{{best[0], best[1], best[2], best[3]}});
    else st1(y + oo, best[0]);
  }
}

// gather form: each input element sums dy of the windows whose FIRST maximum (scan order ky, kx) it is.
template <typename T, int VEC>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W,
                                   int C, int Ho, int Wo, int k, int stride, int pad, long long total_vec) {
  const int CVn = C / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CVn);
    long long r = i / CVn;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int n = (int)(r / H);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    // windows containing (iy, ix): oy*stride - pad <= iy <= oy*stride - pad + k - 1
    int oy_lo = iy + pad - k + 1; oy_lo = oy_lo <= 0 ? 0 : (oy_lo + stride - 1) / stride;
    int oy_hi = (iy + pad) / stride; if (oy_hi > Ho - 1) oy_hi = Ho - 1;
    int ox_lo = ix + pad - k + 1; ox_lo = ox_lo <= 0 ? 0 : (ox_lo + stride - 1) / stride;
    int ox_hi = (ix + pad) / stride; if (ox_hi > Wo - 1) ox_hi = Wo - 1;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        float best[VEC];
        int arg[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { best[j] = -INFINITY; arg[j] = -1; }
        for (int ky = 0; ky < k; ++ky) {
          const int yy = oy * stride - pad + ky;
          if (yy < 0 || yy >= H) continue;
          for (int kx = 0; kx < k; ++kx) {
            const int xx = ox * stride - pad + kx;
            if (xx < 0 || xx >= W) continue;
            const long long off = (((long long)n * H + yy) * W + xx) * C + cv * VEC;
            float v[VEC];
            if (VEC == 4) {
              f4 t = ld4(x + off);
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = t.v[j];
            } else {
              v[0] = ld1(x + off);
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              if (v[j] > best[j] || arg[j] < 0) { best[j] = v[j]; arg[j] = yy * W + xx; }
            }
          }
        }
        const long long oo = (((long long)n * Ho + oy) * Wo + ox) * C + cv * VEC;
        float g[VEC];
        if (VEC == 4) {
          f4 t = ld4(dy + oo);
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] = t.v[j];
        } else {
          g[0] = ld1(dy + oo);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (arg[j] == iy * W + ix) acc[j] += g[j];
      }
    }
    const long long io = (((long long)n * H + iy) * W + ix) * C + cv * VEC;
    if (VEC == 4) st4(dx + io, f4{{acc[0], acc[1], acc[2], acc[3]}});
    else st1(dx + io, acc[0]);
  }
}

// Work-item -> (image, y, x, channel vector) with a CTA pass covering a compact 2-D tile of pixels (all channel vectors
// of tw x th pixels), so that the overlapping pooling windows of neighbouring pixels are served by L1 instead of L2
// (the linear mapping re-read every window up to 9 times from L2: 338 us for the stem's max-pool backward).
struct PixTile { int tw, th, tiles_w, tiles_h; };
__host__ __device__ inline PixTile make_pix_tile(int CVn, int H, int W) {
  PixTile t;
  int ppb = 256 / CVn; if (ppb < 1) ppb = 1;
  t.tw = ppb < 8 ? ppb : 8;
  t.th = ppb / t.tw; if (t.th < 1) t.th = 1;
  t.tiles_w = (W + t.tw - 1) / t.tw;
  t.tiles_h = (H + t.th - 1) / t.th;
  return t;
}
// One CTA = one pixel tile: blockIdx = (tile column, tile row, image); a thread's slot r in the tile is (pixel, channel
// vector).  Only 32-bit divisions by small numbers are left (the flat-index form decoded a 64-bit linear index with five
// 64-bit div/mods per element: ~500 of the ~600 instructions a thread executed, 316 us for the stem's max-pool backward
// against ~57 us of HBM time).
// A CTA walks MP_TPB consecutive tile rows (unrolled: the loads of several tiles are in flight together; one tile per CTA
// was CTA-launch bound: 65 536 CTAs of one element per thread for the stem's map).
constexpr int MP_TPB = 4;
__device__ __forceinline__ bool tile_slot(unsigned r, unsigned CVn, const PixTile& t, int tile_row, int H, int W, int& y, int& x,
                                          int& cv) {
  const unsigned p = r / CVn;
  cv = (int)(r - p * CVn);
  const unsigned ly = p / (unsigned)t.tw, lx = p - ly * (unsigned)t.tw;
  x = (int)(blockIdx.x * (unsigned)t.tw + lx);
  y = (int)((unsigned)tile_row * (unsigned)t.th + ly);
  return x < W && y < H;
}

// Inference pool, bf16: tiled grid (no 64-bit index arithmetic), KC*KC 16-byte loads issued together, maxima taken two
// channels at a time (HMNMX2); padding taps contribute -inf.  The flat-index kernel above ran 173 us on the stem's 268 MB map
// (6 % of the 2.7 ms inference step) against 51 us of HBM time.
template <int KC>
__global__ void __launch_bounds__(256) maxpool_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                               int N, int H, int W, int C, int Ho, int Wo, int stride, int pad) {
  const int CVn = C >> 3;
  const PixTile pt = make_pix_tile(CVn, Ho, Wo);
  const unsigned per_tile = (unsigned)(pt.tw * pt.th * CVn);
  const uint32_t ninf2 = 0xFF80FF80u;                          // (-inf, -inf) in bf16x2
  for (int n = blockIdx.z; n < N; n += gridDim.z)
  for (unsigned r_ = threadIdx.x; r_ < per_tile; r_ += blockDim.x)
#pragma unroll
  for (int tr_ = 0; tr_ < MP_TPB; ++tr_) {
    int oy, ox, cv;
    if (!tile_slot(r_, (unsigned)CVn, pt, blockIdx.y * MP_TPB + tr_, Ho, Wo, oy, ox, cv)) continue;
    const __nv_bfloat16* __restrict__ xb = x + (long long)n * H * W * C + cv * 8;
    uint4 v[KC * KC];
#pragma unroll
    for (int t = 0; t < KC * KC; ++t) {
      const int iy = oy * stride - pad + t / KC, ix = ox * stride - pad + t % KC;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
      v[t] = ok ? *reinterpret_cast<const uint4*>(xb + ((long long)iy * W + ix) * C) : make_uint4(ninf2, ninf2, ninf2, ninf2);
    }
    uint32_t m[4] = {ninf2, ninf2, ninf2, ninf2};
#pragma unroll
    for (int t = 0; t < KC * KC; ++t) {
      const uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&m[q]), *reinterpret_cast<const __nv_bfloat162*>(&w[q]));
        m[q] = *reinterpret_cast<const uint32_t*>(&r);
      }
    }
    *reinterpret_cast<uint4*>(y + (((long long)n * Ho + oy) * Wo + ox) * C + cv * 8) = make_uint4(m[0], m[1], m[2], m[3]);
  }
}

// ---- indexed variants: forward stores the window position of the first maximum (uint8), backward gathers by index.
// KC > 0: compile-time window size -- the KC*KC loads of a window are issued back to back (predicated, not branched
// around), so one thread keeps KC*KC 16-byte loads in flight instead of a dependent chain.
template <typename T, int VEC, int KC>
__global__ void __launch_bounds__(256) maxpool_fwd_idx_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                              unsigned char* __restrict__ idx, int N, int H, int W, int C, int Ho,
                                                              int Wo, int k_, int stride, int pad) {
  const int k = KC > 0 ? KC : k_;
  const int CVn = C / VEC;
  const PixTile pt = make_pix_tile(CVn, Ho, Wo);
  const unsigned per_tile = (unsigned)(pt.tw * pt.th * CVn);
  for (int n = blockIdx.z; n < N; n += gridDim.z)
  for (unsigned r_ = threadIdx.x; r_ < per_tile; r_ += blockDim.x)
#pragma unroll
  for (int tr_ = 0; tr_ < MP_TPB; ++tr_) {
    int oy, ox, cv;
    if (!tile_slot(r_, (unsigned)CVn, pt, blockIdx.y * MP_TPB + tr_, Ho, Wo, oy, ox, cv)) continue;
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { best[j] = -INFINITY; arg[j] = -1; }
    const T* __restrict__ xb = x + (long long)n * H * W * C + cv * VEC;
    if constexpr (KC > 0) {
      float v[KC * KC][VEC];
      bool ok[KC * KC];
#pragma unroll
      for (int t = 0; t < KC * KC; ++t) {
        const int iy = oy * stride - pad + t / KC, ix = ox * stride - pad + t % KC;
        ok[t] = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const long long off = ok[t] ? ((long long)iy * W + ix) * C : 0;
        ldv<VEC>(xb + off, v[t]);
      }
#pragma unroll
      for (int t = 0; t < KC * KC; ++t)
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (ok[t] && (v[t][j] > best[j] || arg[j] < 0)) { best[j] = v[t][j]; arg[j] = t; }
    } else {
      for (int ky = 0; ky < k; ++ky) {
        const int iy = oy * stride - pad + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int ix = ox * stride - pad + kx;
          if (ix < 0 || ix >= W) continue;
          float v[VEC];
          ldv<VEC>(xb + ((long long)iy * W + ix) * C, v);
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (v[j] > best[j] || arg[j] < 0) { best[j] = v[j]; arg[j] = ky * k + kx; }
        }
      }
    }
    const long long oo = (((long long)n * Ho + oy) * Wo + ox) * C + cv * VEC;
    stv<VEC>(y + oo, best);
    if constexpr (VEC == 8) {
      uint2 pk;
      pk.x = (unsigned)arg[0] | ((unsigned)arg[1] << 8) | ((unsigned)arg[2] << 16) | ((unsigned)arg[3] << 24);
      pk.y = (unsigned)arg[4] | ((unsigned)arg[5] << 8) | ((unsigned)arg[6] << 16) | ((unsigned)arg[7] << 24);
      *reinterpret_cast<uint2*>(idx + oo) = pk;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) idx[oo + j] = (unsigned char)arg[j];
    }
  }
}

// BatchNorm-apply + ReLU + max-pool (+ argmax index) in ONE pass over the raw conv output -- the stem of the training step
// (src/stf_lstm_unet.py:177-180).  The post-BN map (268 MB at 128 x 128^2 x 64) is never written: the backward pass
// recomputes the ReLU mask from x (bn_bwd, mask_from_x) and routes the pool gradient by index.  Every tap is normalised,
// rectified and ROUNDED TO THE STORAGE TYPE before the comparison, so values, ties and indices are exactly those of
// bn_apply followed by maxpool_fwd_idx.  One CTA serves one image (blockIdx.z): its group's scale/shift come from the
// statistics slots by the finalize arithmetic (bn_group_coefs_to_smem).
template <int KC>
__global__ void __launch_bounds__(256) bn_relu_maxpool_idx_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                                  unsigned char* __restrict__ idx, const float* __restrict__ partial,
                                                                  int nblk, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, int G, long long R,
                                                                  int N, int H, int W, int C, int Ho, int Wo, int stride, int pad,
                                                                  int row_groups) {
  extern __shared__ float coef_sm[];                 // [2][C]
  constexpr int VEC = 8;
  const int n = blockIdx.z;
  const int g = n / (N / G);
  bn_group_coefs_to_smem(partial, nblk, G, C, g, R, gamma, beta, eps, nullptr, nullptr, coef_sm);
  __syncthreads();
  const int CVn = C / VEC;
  const PixTile pt = make_pix_tile(CVn, Ho, Wo);
  const unsigned per_tile = (unsigned)(pt.tw * pt.th * CVn);
  const __nv_bfloat16* __restrict__ xn = x + (long long)n * H * W * C;
  for (unsigned r_ = threadIdx.x; r_ < per_tile; r_ += blockDim.x) {
    float sc[VEC], sh[VEC];
    int cv_prev = -1;
    for (int tg = blockIdx.y; tg < row_groups; tg += gridDim.y)     // several groups of tile rows per CTA: the
#pragma unroll                                                    // coefficient preamble is paid once per CTA
    for (int tr_ = 0; tr_ < MP_TPB; ++tr_) {
      int oy, ox, cv;
      if (!tile_slot(r_, (unsigned)CVn, pt, tg * MP_TPB + tr_, Ho, Wo, oy, ox, cv)) continue;
      if (cv != cv_prev) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sc[j] = coef_sm[cv * VEC + j]; sh[j] = coef_sm[C + cv * VEC + j]; }
        cv_prev = cv;
      }
      float v[KC * KC][VEC];
      bool ok[KC * KC];
#pragma unroll
      for (int t = 0; t < KC * KC; ++t) {
        const int iy = oy * stride - pad + t / KC, ix = ox * stride - pad + t % KC;
        ok[t] = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const long long off = ok[t] ? ((long long)iy * W + ix) * C : 0;
        ldv<VEC>(xn + off + cv * VEC, v[t]);
      }
      float best[VEC];
      int arg[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) { best[j] = -INFINITY; arg[j] = -1; }
#pragma unroll
      for (int t = 0; t < KC * KC; ++t)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float a = __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(v[t][j], sc[j], sh[j]), 0.f)));
          if (ok[t] && (a > best[j] || arg[j] < 0)) { best[j] = a; arg[j] = t; }
        }
      const long long oo = (((long long)n * Ho + oy) * Wo + ox) * C + cv * VEC;
      stv<VEC>(y + oo, best);
      uint2 pk;
      pk.x = (unsigned)arg[0] | ((unsigned)arg[1] << 8) | ((unsigned)arg[2] << 16) | ((unsigned)arg[3] << 24);
      pk.y = (unsigned)arg[4] | ((unsigned)arg[5] << 8) | ((unsigned)arg[6] << 16) | ((unsigned)arg[7] << 24);
      *reinterpret_cast<uint2*>(idx + oo) = pk;
    }
  }
}

// gather form: input element (iy, ix) belongs to at most WMAX x WMAX windows (WMAX = ceil(k / stride)); their index words
// and dy vectors are all loaded up front (the branchy form waited on idx before it asked for dy: a dependent chain).
template <typename T, int VEC, int WMAX>
__global__ void __launch_bounds__(256) maxpool_bwd_idx_kernel(const unsigned char* __restrict__ idx, const T* __restrict__ dy,
                                                              T* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo, int k,
                                                              int stride, int pad) {
  const int CVn = C / VEC;
  const PixTile pt = make_pix_tile(CVn, H, W);
  const unsigned per_tile = (unsigned)(pt.tw * pt.th * CVn);
  for (int n = blockIdx.z; n < N; n += gridDim.z)
  for (unsigned r_ = threadIdx.x; r_ < per_tile; r_ += blockDim.x)
#pragma unroll
  for (int tr_ = 0; tr_ < MP_TPB; ++tr_) {
    int iy, ix, cv;
    if (!tile_slot(r_, (unsigned)CVn, pt, blockIdx.y * MP_TPB + tr_, H, W, iy, ix, cv)) continue;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    int oy_lo = iy + pad - k + 1; oy_lo = oy_lo <= 0 ? 0 : (oy_lo + stride - 1) / stride;
    int oy_hi = (iy + pad) / stride; if (oy_hi > Ho - 1) oy_hi = Ho - 1;
    int ox_lo = ix + pad - k + 1; ox_lo = ox_lo <= 0 ? 0 : (ox_lo + stride - 1) / stride;
    int ox_hi = (ix + pad) / stride; if (ox_hi > Wo - 1) ox_hi = Wo - 1;
    if constexpr (VEC == 8 && sizeof(T) == 2) {
      // bf16 fast path: the 8 argmax bytes of a window are compared against this pixel's window position with two SIMD
      // byte compares, the byte masks widened to halfword masks (PRMT) and ANDed onto the bf16x2 words of dy; the <= 4
      // masked contributions are added in bf16x2 (the unpack / compare / convert form was instruction bound: 380 us for
      // the stem's 268 MB map against ~60 us of HBM time)
      uint32_t accw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int t = 0; t < WMAX * WMAX; ++t) {
        const int oy = oy_lo + t / WMAX, ox = ox_lo + t % WMAX;
        const bool ok = oy <= oy_hi && ox <= ox_hi;
        const int p = ok ? (iy - (oy * stride - pad)) * k + (ix - (ox * stride - pad)) : 255;
        const long long oo = ok ? (((long long)n * Ho + oy) * Wo + ox) * C + cv * VEC : (long long)cv * VEC;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + oo);
        const uint4 dv = *reinterpret_cast<const uint4*>(dy + oo);
        const uint32_t p4 = (uint32_t)p * 0x01010101u;
        const uint32_t e0 = __vcmpeq4(pk.x, p4), e1 = __vcmpeq4(pk.y, p4);
        const uint32_t m0 = __byte_perm(e0, 0, 0x1100), m1 = __byte_perm(e0, 0, 0x3322);
        const uint32_t m2 = __byte_perm(e1, 0, 0x1100), m3 = __byte_perm(e1, 0, 0x3322);
        const uint32_t w[4] = {dv.x & m0, dv.y & m1, dv.z & m2, dv.w & m3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __nv_bfloat162 s2 = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&accw[q]), *reinterpret_cast<const __nv_bfloat162*>(&w[q]));
          accw[q] = *reinterpret_cast<const uint32_t*>(&s2);
        }
      }
      const long long io = (((long long)n * H + iy) * W + ix) * C + cv * VEC;
      *reinterpret_cast<uint4*>(dx + io) = make_uint4(accw[0], accw[1], accw[2], accw[3]);
      continue;
    }
    float g[WMAX * WMAX][VEC];
    unsigned char a[WMAX * WMAX][VEC];
    int pos[WMAX * WMAX];
#pragma unroll
    for (int t = 0; t < WMAX * WMAX; ++t) {
      const int oy = oy_lo + t / WMAX, ox = ox_lo + t % WMAX;
      const bool ok = oy <= oy_hi && ox <= ox_hi;
      pos[t] = ok ? (iy - (oy * stride - pad)) * k + (ix - (ox * stride - pad)) : -1;
      const long long oo = ok ? (((long long)n * Ho + oy) * Wo + ox) * C + cv * VEC : (long long)cv * VEC;
      if constexpr (VEC == 8) {
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + oo);
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[t][j] = (unsigned char)(pk.x >> (8 * j)); a[t][4 + j] = (unsigned char)(pk.y >> (8 * j)); }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[t][j] = idx[oo + j];
      }
      ldv<VEC>(dy + oo, g[t]);
    }
#pragma unroll
    for (int t = 0; t < WMAX * WMAX; ++t)
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if ((int)a[t][j] == pos[t]) acc[j] += g[t][j];
    const long long io = (((long long)n * H + iy) * W + ix) * C + cv * VEC;
    stv<VEC>(dx + io, acc);
  }
}

// 3x3 / stride 2 / pad 1 (the ResNet stem pool), bf16, even H and W: one thread owns a 2x2 input block x 8 channels.  The
// block lies in at most four windows -- (m, l), (m, l+1), (m+1, l), (m+1, l+1) -- whose index words and dy vectors are loaded
// ONCE and feed all four outputs through nine (window, position) byte compares (the generic kernel loads and tests four
// windows per input element: 16 window visits per block instead of 4).  Window (oy, ox) starts at input (2oy-1, 2ox-1), so
//   (2m, 2l)     <- (m,l) pos 4              (2m, 2l+1)   <- (m,l) pos 5, (m,l+1) pos 3
//   (2m+1, 2l)   <- (m,l) pos 7, (m+1,l) pos 1   (2m+1, 2l+1) <- (m,l) 8, (m,l+1) 6, (m+1,l) 2, (m+1,l+1) 0
__device__ __forceinline__ void mp_masked_add(uint32_t* acc, const uint2& pk, const uint4& dv, unsigned pos) {
  const uint32_t p4 = pos * 0x01010101u;
  const uint32_t e0 = __vcmpeq4(pk.x, p4), e1 = __vcmpeq4(pk.y, p4);
  const uint32_t w[4] = {dv.x & __byte_perm(e0, 0, 0x1100), dv.y & __byte_perm(e0, 0, 0x3322),
                         dv.z & __byte_perm(e1, 0, 0x1100), dv.w & __byte_perm(e1, 0, 0x3322)};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 s2 = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&acc[q]), *reinterpret_cast<const __nv_bfloat162*>(&w[q]));
    acc[q] = *reinterpret_cast<const uint32_t*>(&s2);
  }
}

constexpr int MP3_BW = 8, MP3_BH = 4, MP3_ROWS = 4;      // CTA: 8 x 4 blocks (x 8 channel vectors = 256 threads), 4 block rows deep
__global__ void __launch_bounds__(256) maxpool3s2_bwd_idx_kernel(const unsigned char* __restrict__ idx,
                                                                 const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                                 int N, int H, int W, int C, int Ho, int Wo) {
  const int CVn = C >> 3;                                 // launched with CVn == 8 channel vectors per CTA column slice
  const int cv = blockIdx.x % ((CVn + 7) / 8) * 8 + (threadIdx.x & 7);
  const int bxt = blockIdx.x / ((CVn + 7) / 8);
  const int l = bxt * MP3_BW + ((threadIdx.x >> 3) & (MP3_BW - 1));
  const int mrow = threadIdx.x >> 6;                      // 0..3
  if (cv >= CVn || 2 * l >= W) return;
  const bool ox1 = l + 1 < Wo;
  for (int n = blockIdx.z; n < N; n += gridDim.z) {
#pragma unroll
    for (int it = 0; it < MP3_ROWS; ++it) {
      const int m = (blockIdx.y * MP3_ROWS + it) * MP3_BH + mrow;
      if (2 * m >= H) continue;
      const bool oy1 = m + 1 < Ho;
      const long long o00 = (((long long)n * Ho + m) * Wo + l) * C + cv * 8;
      const long long o01 = ox1 ? o00 + C : o00, o10 = oy1 ? o00 + (long long)Wo * C : o00;
      const long long o11 = (ox1 && oy1) ? o00 + (long long)Wo * C + C : o00;
      uint2 i00 = *reinterpret_cast<const uint2*>(idx + o00), i01 = *reinterpret_cast<const uint2*>(idx + o01);
      uint2 i10 = *reinterpret_cast<const uint2*>(idx + o10), i11 = *reinterpret_cast<const uint2*>(idx + o11);
      const uint4 d00 = *reinterpret_cast<const uint4*>(dy + o00), d01 = *reinterpret_cast<const uint4*>(dy + o01);
      const uint4 d10 = *reinterpret_cast<const uint4*>(dy + o10), d11 = *reinterpret_cast<const uint4*>(dy + o11);
      // a missing neighbour window contributes nothing: 0xff never equals a window position
      if (!ox1) i01 = make_uint2(0xffffffffu, 0xffffffffu);
      if (!oy1) i10 = make_uint2(0xffffffffu, 0xffffffffu);
      if (!(ox1 && oy1)) i11 = make_uint2(0xffffffffu, 0xffffffffu);
      uint32_t a00[4] = {0u, 0u, 0u, 0u}, a01[4] = {0u, 0u, 0u, 0u}, a10[4] = {0u, 0u, 0u, 0u}, a11[4] = {0u, 0u, 0u, 0u};
      mp_masked_add(a00, i00, d00, 4);
      mp_masked_add(a01, i00, d00, 5); mp_masked_add(a01, i01, d01, 3);
      mp_masked_add(a10, i00, d00, 7); mp_masked_add(a10, i10, d10, 1);
      mp_masked_add(a11, i00, d00, 8); mp_masked_add(a11, i01, d01, 6); mp_masked_add(a11, i10, d10, 2); mp_masked_add(a11, i11, d11, 0);
      const long long x00 = (((long long)n * H + 2 * m) * W + 2 * l) * C + cv * 8;
      *reinterpret_cast<uint4*>(dx + x00) = make_uint4(a00[0], a00[1], a00[2], a00[3]);
      *reinterpret_cast<uint4*>(dx + x00 + C) = make_uint4(a01[0], a01[1], a01[2], a01[3]);
      *reinterpret_cast<uint4*>(dx + x00 + (long long)W * C) = make_uint4(a10[0], a10[1], a10[2], a10[3]);
      *reinterpret_cast<uint4*>(dx + x00 + (long long)W * C + C) = make_uint4(a11[0], a11[1], a11[2], a11[3]);
    }
  }
}

// =================================================================================================
// bilinear, align_corners=True
// =================================================================================================
__device__ __forceinline__ void bilin_coords(int o, int in, float scale, int& i0, int& i1, float& l1) {
  const float src = scale * (float)o;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

template <typename T>
__global__ void bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo,
                                    float sy, float sx, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    int y0, y1, x0, x1; float ly, lx;
    bilin_coords(oy, H, sy, y0, y1, ly);
    bilin_coords(ox, W, sx, x0, x1, lx);
    const T* b = x + (long long)n * H * W * C + c;
    const float v00 = ld1(b + ((long long)y0 * W + x0) * C), v01 = ld1(b + ((long long)y0 * W + x1) * C);
    const float v10 = ld1(b + ((long long)y1 * W + x0) * C), v11 = ld1(b + ((long long)y1 * W + x1) * C);
    const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    st1(y + i, v);
  }
}

template <typename T>
__global__ void bilinear_bwd_kernel(const T* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                    int Wo, float sy, float sx, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    int y0, y1, x0, x1; float ly, lx;
    bilin_coords(oy, H, sy, y0, y1, ly);
    bilin_coords(ox, W, sx, x0, x1, lx);
    const float g = ld1(dy + i);
    float* b = dx + (long long)n * H * W * C + c;
    atomicAdd(b + ((long long)y0 * W + x0) * C, g * (1.f - ly) * (1.f - lx));
    atomicAdd(b + ((long long)y0 * W + x1) * C, g * (1.f - ly) * lx);
    atomicAdd(b + ((long long)y1 * W + x0) * C, g * ly * (1.f - lx));
    atomicAdd(b + ((long long)y1 * W + x1) * C, g * ly * lx);
  }
}

// =================================================================================================
// LSTM cell (PyTorch gate order i, f, g, o)
// =================================================================================================
template <typename T>
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, T* acts, float* c_out,
                                     T* h_out, long long R, int C, long long total_vec) {
  const int CVn = C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / CVn;
    const int c = (int)(i - row * CVn) * 4;
    const float* gp = gates + row * 4 * C + c;
    const f4 gi = ld4(gp), gf = ld4(gp + C), gg = ld4(gp + 2 * C), go = ld4(gp + 3 * C);
    f4 cp{{0.f, 0.f, 0.f, 0.f}};
    if (c_prev) cp = ld4(c_prev + row * C + c);
    f4 ai, af, ag, ao, cn, hn;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ai.v[j] = sigmoidf_(gi.v[j]);
      af.v[j] = sigmoidf_(gf.v[j]);
      ag.v[j] = tanhf(gg.v[j]);
      ao.v[j] = sigmoidf_(go.v[j]);
      cn.v[j] = af.v[j] * cp.v[j] + ai.v[j] * ag.v[j];
      hn.v[j] = ao.v[j] * tanhf(cn.v[j]);
    }
    if (acts) {
      T* ap = acts + row * 4 * C + c;
      st4(ap, ai); st4(ap + C, af); st4(ap + 2 * C, ag); st4(ap + 3 * C, ao);
    }
    st4(c_out + row * C + c, cn);
    st4(h_out + row * C + c, hn);
  }
}

// acts_il: the saved activations are in the fused step's accumulator column order, (row, gate, u) at
// row*4C + (u/16)*64 + gate*16 + u%16 (stfb_lstm_step_fused), instead of gate-major [row][gate*C + u].
template <typename T>
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh, float* dc, const T* __restrict__ acts,
                                     const float* __restrict__ c_prev, const float* __restrict__ c_cur, T* dgates, long long R,
                                     int C, long long total_vec, int acts_il) {
  const int CVn = C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / CVn;
    const int c = (int)(i - row * CVn) * 4;
    const T* ap = acts + row * 4 * C + (acts_il ? (c >> 4) * 64 + (c & 15) : c);
    const int gs = acts_il ? 16 : C;
    const f4 ai = ld4(ap), af = ld4(ap + gs), ag = ld4(ap + 2 * gs), ao = ld4(ap + 3 * gs);
    const f4 vdh = ld4(dh + row * C + c), vdc = ld4(dc + row * C + c), cc = ld4(c_cur + row * C + c);
    f4 cp{{0.f, 0.f, 0.f, 0.f}};
    if (c_prev) cp = ld4(c_prev + row * C + c);
    f4 di, df, dg, dO, dcp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tc = tanhf(cc.v[j]);
      dO.v[j] = vdh.v[j] * tc * ao.v[j] * (1.f - ao.v[j]);
      const float dct = vdc.v[j] + vdh.v[j] * ao.v[j] * (1.f - tc * tc);
      di.v[j] = dct * ag.v[j] * ai.v[j] * (1.f - ai.v[j]);
      df.v[j] = dct * cp.v[j] * af.v[j] * (1.f - af.v[j]);
      dg.v[j] = dct * ai.v[j] * (1.f - ag.v[j] * ag.v[j]);
      dcp.v[j] = dct * af.v[j];
    }
    T* gp = dgates + row * 4 * C + c;
    st4(gp, di); st4(gp + C, df); st4(gp + 2 * C, dg); st4(gp + 3 * C, dO);
    st4(dc + row * C + c, dcp);
  }
}

// =================================================================================================
// layout adapters
// =================================================================================================
template <typename T>
__global__ void pack_series_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int Tn, int C, int H, int W,
                                   long long total) {
  // destination order (t, b, yy, xx, c)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int xx = (int)(r % W); r /= W;
    const int yy = (int)(r % H); r /= H;
    const int b = (int)(r % B);
    const int t = (int)(r / B);
    st1(y + i, x[((((long long)b * Tn + t) * C + c) * H + yy) * W + xx]);
  }
}

// time series + static per-sample maps -> one NHWC tensor: y[t*B+b][h][w][c] = c < Cx ? x[b][t][c][h][w] : m[b][c-Cx][h][w]
template <typename T>
__global__ void pack_series_maps_kernel(const float* __restrict__ x, const float* __restrict__ m, T* __restrict__ y, int B, int Tn,
                                        int Cx, int Cm, int H, int W, long long total) {
  const int C = Cx + Cm;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int xx = (int)(r % W); r /= W;
    const int yy = (int)(r % H); r /= H;
    const int b = (int)(r % B);
    const int t = (int)(r / B);
    const float v = c < Cx ? x[((((long long)b * Tn + t) * Cx + c) * H + yy) * W + xx]
                           : m[(((long long)b * Cm + (c - Cx)) * H + yy) * W + xx];
    st1(y + i, v);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ y, float* __restrict__ out, int N, int HW, int C, long long total) {
  // destination order (n, c, p): coalesced writes
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    long long r = i / HW;
    const int c = (int)(r % C);
    const long long n = r / C;
    out[i] = ld1(y + (n * HW + p) * C + c);
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ g, T* __restrict__ out, int N, int HW, int C, long long total) {
  // destination order (n, p, c)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int p = (int)(r % HW);
    const long long n = r / HW;
    st1(out + i, g[(n * C + c) * HW + p]);
  }
}

// im2col for small-channel convolutions (7x7 stem with Cin=1, UNet first conv): out[m][(ky,kx,ci)] zero-padded to Kpad,
// so the convolution becomes a K = Kpad GEMM the tcgen05 family can take.  One thread = 8 consecutive k (16 B store).
// KC / CINC > 0: compile-time filter size / channel count (the divisions become multiplies; the runtime form spent its
// time in integer div/mod: 0.51 ms for the 7x7 stem at 128 x 256^2 against 0.05 ms of HBM time).
template <int KC, int CINC>
__global__ void __launch_bounds__(256) im2col_small_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           int N, int H, int W, int Cin_, int Ho, int Wo, int k_, int stride,
                                                           int pad, int Kpad, long long total) {
  const int k = KC > 0 ? KC : k_;
  const int Cin = CINC > 0 ? CINC : Cin_;
  const int chunks = Kpad / 8;
  const int Ktot = k * k * Cin;
  const unsigned short* __restrict__ xs = reinterpret_cast<const unsigned short*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long r = i / chunks;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
    const unsigned short* __restrict__ xn = xs + (long long)n * H * W * Cin;
    unsigned short v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = ch * 8 + j;
      const int tap = kk / Cin, ci = kk - tap * Cin;
      const int ky = tap / k, kx = tap - ky * k;
      const int iy = iy0 + ky, ix = ix0 + kx;
      const bool ok = kk < Ktot && iy >= 0 && iy < H && ix >= 0 && ix < W;
      v[j] = ok ? __ldg(xn + ((long long)iy * W + ix) * Cin + ci) : (unsigned short)0;
    }
    uint4 pk;
    pk.x = v[0] | ((unsigned)v[1] << 16); pk.y = v[2] | ((unsigned)v[3] << 16);
    pk.z = v[4] | ((unsigned)v[5] << 16); pk.w = v[6] | ((unsigned)v[7] << 16);
    *reinterpret_cast<uint4*>(out + i * 8) = pk;
  }
}

// Tiled form for the compile-time geometries (K padded to the next multiple of 64): a CTA of 128 threads owns IM_TR x IM_TP
// output pixels of one image.  It stages the input patch they touch ((IM_TR-1)*S + K rows x (IM_TP-1)*S + K columns x CIN
// channels, zero padded) in shared memory with coalesced loads; thread p then assembles pixel p's K-padded row 64 values
// at a time -- every (ky, kx, ci) offset is a compile-time immediate -- into a padded staging tile, and the tile leaves
// with 16-byte stores, 8 lanes per 128-byte row segment.  (The flat form above ran ~480 instructions per 16-byte chunk:
// 199 us for the stem against 44 us of HBM time; a first tiled form with run-time chunk indices still took 160 us.)
constexpr int IM_TP = 32, IM_TR = 4, IM_PIX = IM_TP * IM_TR;
template <int KC, int CINC, int S>
__global__ void __launch_bounds__(IM_PIX) im2col_tiled_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                              int N, int H, int W, int Ho, int Wo, int pad) {
  constexpr int PW = (IM_TP - 1) * S + KC, PR = (IM_TR - 1) * S + KC, ROW = PW * CINC, KTOT = KC * KC * CINC;
  constexpr int NSEG = (KTOT + 63) / 64, KPAD = NSEG * 64;
  constexpr int SROW = 36;                                 // staging row: 32 words of payload + 4 of padding (conflict free)
  __shared__ unsigned short patch[PR * ROW + 2];
  __shared__ __align__(16) uint32_t stage[IM_PIX * SROW];
  const unsigned short* __restrict__ xs = reinterpret_cast<const unsigned short*>(x);
  const int ox0 = blockIdx.x * IM_TP, oy0 = blockIdx.y * IM_TR;
  const int tid = threadIdx.x;
  const int lx = tid % IM_TP, ly = tid / IM_TP;
  for (int n = blockIdx.z; n < N; n += gridDim.z) {
    const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
    const unsigned short* __restrict__ xn = xs + (long long)n * H * W * CINC;
    for (int e = tid; e < PR * ROW; e += IM_PIX) {
      const int py = e / ROW, rem = e - py * ROW;          // rem = px * CINC + ci: contiguous in global memory
      const int px = rem / CINC;
      const int iy = iy0 + py, ix = ix0 + px;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
      patch[e] = ok ? __ldg(xn + ((long long)iy * W + ix0) * CINC + rem) : (unsigned short)0;
    }
    __syncthreads();
    const unsigned short* __restrict__ pb = patch + (ly * S) * ROW + (lx * S) * CINC;
#pragma unroll
    for (int seg = 0; seg < NSEG; ++seg) {
      uint32_t wv[32];
#pragma unroll
      for (int wd = 0; wd < 32; ++wd) {
        uint32_t lo = 0u, hi = 0u;
        const int k0 = seg * 64 + 2 * wd, k1 = k0 + 1;          // compile-time after unrolling: every offset is an immediate
        if (k0 < KTOT) { const int tap = k0 / CINC, ci = k0 % CINC; lo = pb[(tap / KC) * ROW + (tap % KC) * CINC + ci]; }
        if (k1 < KTOT) { const int tap = k1 / CINC, ci = k1 % CINC; hi = pb[(tap / KC) * ROW + (tap % KC) * CINC + ci]; }
        wv[wd] = lo | (hi << 16);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(stage + tid * SROW + q * 4) = make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
      __syncthreads();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int q = it * IM_PIX + tid;
        const int pix = q >> 3, c16 = q & 7;
        const int oy = oy0 + pix / IM_TP, ox = ox0 + pix % IM_TP;
        if (oy < Ho && ox < Wo)
          *reinterpret_cast<uint4*>(out + ((((long long)n * Ho + oy) * Wo + ox) * KPAD + seg * 64 + c16 * 8)) =
              *reinterpret_cast<const uint4*>(stage + pix * SROW + c16 * 4);
      }
      __syncthreads();
    }
  }
}

// dW[r][ci][tap] += src[r][tap*Cin + ci]: folds the K-padded, (tap, ci)-ordered im2col weight gradient back into the
// parameter layout [Cout][Cin][kh][kw]
__global__ void unpad_wgrad_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int Cin, int khw, int ld_src) {
  const long long total = (long long)rows * Cin * khw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % khw);
    long long r = i / khw;
    const int ci = (int)(r % Cin);
    r /= Cin;
    dst[i] += src[r * ld_src + (long long)tap * Cin + ci];
  }
}

template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ d, const T* __restrict__ s, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st1(d + i, ld1(d + i) + ld1(s + i));
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st1(d + i, ld1(s + i));
}

}  // namespace stfb

using namespace stfb;

#define DT_OK(d) ((d) == STFB_F32 || (d) == STFB_BF16)
#define DISPATCH_T(dtype, ...)                                   \
  do {                                                           \
    if ((dtype) == STFB_F32) { using T = float; __VA_ARGS__; }   \
    else { using T = __nv_bfloat16; __VA_ARGS__; }               \
  } while (0)

// widest vector the channel count / alignment allows: 8 (bf16, 16 B) > 4 > 1
static int pick_vec(int C, int dtype, std::initializer_list<const void*> ptrs) {
  auto ok = [&](int v) {
    if (C % v != 0) return false;
    const int b = v * (dtype == STFB_BF16 ? 2 : 4);
    for (const void* p : ptrs)
      if (p && !aligned_to(p, b > 16 ? 16 : b)) return false;
    return true;
  };
  if (dtype == STFB_BF16 && ok(8)) return 8;
  if (ok(4)) return 4;
  return 1;
}

static bool vec4_ok(int C, int dtype, std::initializer_list<const void*> ptrs) {
  if (C % 4 != 0) return false;
  const int b = dtype == STFB_BF16 ? 8 : 16;
  for (const void* p : ptrs)
    if (p && !aligned_to(p, b)) return false;
  return true;
}

// launch geometry of the row-tiled kernels: tpr threads span a row's vectors, ~8 CTAs per SM, >= 8 row passes per CTA
struct RowTile { int tpr; long long rows_per_block; dim3 grid; };
static RowTile row_tile(int G, long long R, int C, int vec, int unroll) {
  RowTile t;
  const int CVn = C / vec;
  int tpr = 1;
  while (tpr * 2 <= CVn && tpr * 2 <= 256) tpr *= 2;
  t.tpr = tpr;
  const long long lanes = 256 / tpr;
  // CTAs per SM the grid is sized for: 4 = one resident wave (the kernels hold 4 CTAs of 256 threads per SM), so the
  // per-CTA coefficient preamble is paid once per SM slot; measured on the graph step: 4 -> 9.85 ms, 8 -> 9.93, 16 -> 9.97
  static int per_sm = 0;
  if (per_sm == 0) { const char* e = getenv("STFB_ROWTILE_CTAS_PER_SM"); per_sm = e ? atoi(e) : 4; if (per_sm < 1 || per_sm > 64) per_sm = 4; }
  long long want = ((long long)per_sm * num_sms() + G - 1) / G;
  long long cap = (R + lanes * unroll * 2 - 1) / (lanes * unroll * 2);
  long long nblk = want < cap ? want : cap;
  if (nblk < 1) nblk = 1;
  if (nblk > 65535) nblk = 65535;
  t.rows_per_block = (R + nblk - 1) / nblk;
  nblk = (R + t.rows_per_block - 1) / t.rows_per_block;
  t.grid = dim3((unsigned)nblk, (unsigned)G);
  return t;
}

extern "C" int stfb_bn_partial_blocks(int G, long long R) {
  if (G <= 0 || R <= 0) return 1;
  return colreduce_blocks(G, R);
}

extern "C" int stfb_bn_stats(const void* x, float* partial, int nblk, int G, long long R, int C, int dtype, void* stream) {
  STFB_REQUIRE(x && partial && G > 0 && R >= 0 && C > 0 && DT_OK(dtype), "bn_stats: bad arguments");
  STFB_DEVICE_OR_RETURN();
  STFB_REQUIRE(nblk == colreduce_blocks(G, R), "bn_stats: nblk must come from stfb_bn_partial_blocks");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ColRedArgs A{};
  A.a = x; A.partial = partial; A.G = G; A.R = R; A.C = C;
  return launch_colreduce<0>(A, dtype, pick_vec(C, dtype, {x}), s, "bn_stats", nblk);
}

extern "C" int stfb_bn_finalize_train(const float* partial, int nblk, const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, long long* nbt, float* scale, float* shift, float* mean,
                                      float* invstd, int G, long long R, int C, float eps, float momentum, void* stream) {
  STFB_REQUIRE(partial && nblk > 0 && gamma && beta && scale && shift && mean && invstd && G > 0 && R > 0 && C > 0,
               "bn_finalize_train: bad arguments");
  STFB_DEVICE_OR_RETURN();
  const int fin_g = G < FIN_MAXG ? G : FIN_MAXG;
  if (nblk > 64 && G < 4)
    bn_finalize_train_kernel<FIN_WIDE_SL><<<ceil_div(C, FIN_CH), FIN_CH * FIN_WIDE_SL, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        partial, nblk, gamma, beta, running_mean, running_var, nbt, scale, shift, mean, invstd, G, R, C, eps, momentum);
  else
    bn_finalize_train_kernel<1><<<ceil_div(C, FIN_CH), FIN_CH * fin_g, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        partial, nblk, gamma, beta, running_mean, running_var, nbt, scale, shift, mean, invstd, G, R, C, eps, momentum);
  return post_launch("bn_finalize_train");
}

extern "C" int stfb_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv, float* scale,
                                 float* shift, int C, float eps, void* stream) {
  STFB_REQUIRE(gamma && beta && rm && rv && scale && shift && C > 0, "bn_fold_eval: bad arguments");
  STFB_DEVICE_OR_RETURN();
  bn_fold_eval_kernel<<<ceil_div(C, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gamma, beta, rm, rv, scale, shift, C, eps);
  return post_launch("bn_fold_eval");
}

static int bn_apply_launch(const void* x, const float* scale, const float* shift, const void* residual, void* y, int G, long long R,
                           int C, int relu, int dtype, const float* partial, int nblk, const float* gamma, const float* beta,
                           float eps, cudaStream_t s) {
  const long long rows = (long long)G * R;
  if (rows == 0) return STFB_OK;
  const int v = pick_vec(C, dtype, {x, residual, y});
  const RowTile rt = row_tile(G, R, C, v, 2);
  const size_t sm = (size_t)2 * C * sizeof(float);
  STFB_REQUIRE(sm <= 48 * 1024, "bn_apply: %d channels exceed the shared-memory coefficient table", C);
#define BN_APPLY(V) launch_ex<2>(bn_apply_kernel<T, V>, rt.grid, dim3(256), sm, s, (const T*)x, scale, shift, (const T*)residual, (T*)y, R, C, rt.rows_per_block, rt.tpr, relu, partial, nblk, gamma, beta, eps)
  DISPATCH_T(dtype, {
    if (v == 8) BN_APPLY(8);
    else if (v == 4) BN_APPLY(4);
    else BN_APPLY(1);
  });
#undef BN_APPLY
  return post_launch("bn_apply");
}

extern "C" int stfb_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, void* y, int G,
                             long long R, int C, int relu, int dtype, void* stream) {
  STFB_REQUIRE(x && scale && shift && y && G > 0 && R >= 0 && C > 0 && DT_OK(dtype), "bn_apply: bad arguments");
  STFB_DEVICE_OR_RETURN();
  return bn_apply_launch(x, scale, shift, residual, y, G, R, C, relu, dtype, nullptr, 0, nullptr, nullptr, 0.f,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stfb_bn_apply_from_stats(const void* x, const float* partial, int nblk, const float* gamma, const float* beta,
                                        const void* residual, void* y, int G, long long R, int C, float eps, int relu, int dtype,
                                        void* stream) {
  STFB_REQUIRE(x && partial && gamma && beta && y && G > 0 && G <= 65535 && R > 0 && C > 0 && DT_OK(dtype), "bn_apply_from_stats: bad arguments");
  STFB_REQUIRE(nblk >= 1 && nblk <= 8, "bn_apply_from_stats: 1..8 statistics slots (got %d); finalize first for more", nblk);
  STFB_DEVICE_OR_RETURN();
  return bn_apply_launch(x, nullptr, nullptr, residual, y, G, R, C, relu, dtype, partial, nblk, gamma, beta, eps,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stfb_bn_bwd_reduce(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                                  const float* scale, const float* shift, float* partial, int nblk, int G, long long R, int C,
                                  int relu, int dtype, void* stream) {
  STFB_REQUIRE(dy && x && mean && invstd && partial && G > 0 && R >= 0 && C > 0 && DT_OK(dtype), "bn_bwd_reduce: bad arguments");
  STFB_REQUIRE(!relu || y || (scale && shift), "bn_bwd_reduce: relu needs y, or scale and shift to recompute the mask from x");
  STFB_DEVICE_OR_RETURN();
  STFB_REQUIRE(nblk == colreduce_blocks(G, R), "bn_bwd_reduce: nblk must come from stfb_bn_partial_blocks");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ColRedArgs A{};
  A.a = dy; A.b = y; A.c = x; A.mean = mean; A.invstd = invstd; A.partial = partial; A.G = G; A.R = R; A.C = C; A.relu = relu;
  A.scale = scale; A.shift = shift;
  return launch_colreduce<1>(A, dtype, pick_vec(C, dtype, {dy, y, x}), s, "bn_bwd_reduce", nblk);
}

extern "C" int stfb_bn_bwd_finalize(const float* partial, int nblk, const float* gamma, const float* invstd, float* dgamma,
                                    float* dbeta, float* coef, int G, long long R, int C, void* stream) {
  STFB_REQUIRE(partial && nblk > 0 && gamma && invstd && coef && G > 0 && R > 0 && C > 0, "bn_bwd_finalize: bad arguments");
  STFB_DEVICE_OR_RETURN();
  if (nblk > 64 && G < 4)
    bn_bwd_finalize_kernel<32, 1><<<ceil_div(C, FIN_CH), 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        partial, nblk, gamma, invstd, dgamma, dbeta, coef, G, R, C);
  else
    bn_bwd_finalize_kernel<4, 8><<<ceil_div(C, FIN_CH), 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        partial, nblk, gamma, invstd, dgamma, dbeta, coef, G, R, C);
  return post_launch("bn_bwd_finalize");
}

extern "C" int stfb_bn_bwd_apply(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                                 const float* coef, const float* shift, void* dx, void* dres, int accum_dres, int G, long long R,
                                 int C, int relu, int dtype, void* stream) {
  STFB_REQUIRE(dy && x && mean && invstd && coef && dx && G > 0 && R >= 0 && C > 0 && DT_OK(dtype), "bn_bwd_apply: bad arguments");
  STFB_REQUIRE(!relu || y || shift, "bn_bwd_apply: relu needs y, or shift to recompute the mask from x");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)G * R;
  if (rows == 0) return STFB_OK;
  const int v = pick_vec(C, dtype, {dy, y, x, dx, dres});
  const RowTile rt = row_tile(G, R, C, v, 1);
  DISPATCH_T(dtype, {
    if (v == 8) bn_bwd_apply_kernel<T, 8><<<rt.grid, 256, 0, s>>>((const T*)dy, (const T*)y, (const T*)x, mean, invstd, coef, (T*)dx, (T*)dres, R, C, rt.rows_per_block, rt.tpr, relu, accum_dres, shift);
    else if (v == 4) bn_bwd_apply_kernel<T, 4><<<rt.grid, 256, 0, s>>>((const T*)dy, (const T*)y, (const T*)x, mean, invstd, coef, (T*)dx, (T*)dres, R, C, rt.rows_per_block, rt.tpr, relu, accum_dres, shift);
    else bn_bwd_apply_kernel<T, 1><<<rt.grid, 256, 0, s>>>((const T*)dy, (const T*)y, (const T*)x, mean, invstd, coef, (T*)dx, (T*)dres, R, C, rt.rows_per_block, rt.tpr, relu, accum_dres, shift);
  });
  return post_launch("bn_bwd_apply");
}

extern "C" int stfb_colsum(const void* x, float* out, long long R, int C, int dtype, void* stream) {
  STFB_REQUIRE(x && out && R >= 0 && C > 0 && DT_OK(dtype), "colsum: bad arguments");
  STFB_DEVICE_OR_RETURN();
  ColRedArgs A{};
  A.a = x; A.outf = out; A.G = 1; A.R = R; A.C = C;
  return launch_colreduce<2>(A, dtype, pick_vec(C, dtype, {x}), reinterpret_cast<cudaStream_t>(stream), "colsum");
}

// co-resident capacity of the fused kernel (CTAs per SM from the occupancy API, capped at the 2 its register budget is
// sized for), per instantiation
template <typename T, int VEC>
static int bn_bwd_fused_capacity() {
  static int cap = 0;
  if (cap == 0) {
    int per_sm = 0, per_sm_lean = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lean, bn_bwd_fused_kernel<T, VEC, true>, 256, 0) != cudaSuccess) per_sm_lean = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel<T, VEC, false>, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm_lean < per_sm) per_sm = per_sm_lean;
    if (per_sm > 2) per_sm = 2;
    cap = per_sm * num_sms();
  }
  return cap;
}

extern "C" size_t stfb_bn_bwd_fused_scratch_floats(int G, int C) {
  return (size_t)2 * G * C + (size_t)((G + 15) / 16 * 16);      // acc[G][2][C] + one arrival counter per group
}

extern "C" int stfb_bn_bwd_fused(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                                 const float* gamma, const float* shift, float* scratch, float* dgamma, float* dbeta, void* dx,
                                 void* dres, int accum_dres, int G, long long R, int C, int relu, int dtype, void* stream) {
  STFB_REQUIRE(dy && x && mean && invstd && gamma && scratch && dx && G > 0 && G <= 65535 && R >= 0 && C > 0 && DT_OK(dtype),
               "bn_bwd_fused: bad arguments");
  STFB_REQUIRE(!relu || y || shift, "bn_bwd_fused: relu needs y, or shift to recompute the mask from x");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if ((long long)G * R == 0) return STFB_OK;
  const int v = pick_vec(C, dtype, {dy, y, x, dx, dres});
  int cap = 0;
  DISPATCH_T(dtype, { cap = v == 8 ? bn_bwd_fused_capacity<T, 8>() : (v == 4 ? bn_bwd_fused_capacity<T, 4>() : bn_bwd_fused_capacity<T, 1>()); });
  STFB_REQUIRE(G <= cap, "bn_bwd_fused: %d groups exceed the co-resident capacity (%d CTAs); use the three-launch chain", G, cap);
  const int CVn = C / v;
  int tpr = 1;
  while (tpr * 2 <= CVn && tpr * 2 <= 256) tpr *= 2;
  const long long lanes = 256 / tpr;
  long long nblk = cap / G;                                   // single wave: G * nblk <= capacity
  const long long by_rows = (R + lanes * 4 - 1) / (lanes * 4); // at least four row passes per CTA
  if (nblk > by_rows) nblk = by_rows;
  if (nblk < 1) nblk = 1;
  const long long rpb = (R + nblk - 1) / nblk;
  nblk = (R + rpb - 1) / rpb;
  float* acc = scratch;
  unsigned* counters = reinterpret_cast<unsigned*>(scratch + (size_t)2 * G * C);
  dim3 grid((unsigned)nblk, (unsigned)G);
  const bool lean = !dres && (!relu || !y);                   // nothing but dy and x is read
#define BN_FUSED(V, L) launch_ex<2>(bn_bwd_fused_kernel<T, V, L>, grid, dim3(256), 0, s, (const T*)dy, (const T*)y, (const T*)x, mean, invstd, gamma, shift, acc, counters, dgamma, dbeta, (T*)dx, (T*)dres, R, C, rpb, tpr, relu, accum_dres)
  DISPATCH_T(dtype, {
    if (lean) { if (v == 8) BN_FUSED(8, true); else if (v == 4) BN_FUSED(4, true); else BN_FUSED(1, true); }
    else { if (v == 8) BN_FUSED(8, false); else if (v == 4) BN_FUSED(4, false); else BN_FUSED(1, false); }
  });
#undef BN_FUSED
  return post_launch("bn_bwd_fused");
}

extern "C" int stfb_maxpool_fwd(const void* x, void* y, int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad,
                                int dtype, void* stream) {
  STFB_REQUIRE(x && y && N >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && 2 * pad <= k && DT_OK(dtype), "maxpool_fwd: bad arguments");
  STFB_REQUIRE(Ho == (H + 2 * pad - k) / stride + 1 && Wo == (W + 2 * pad - k) / stride + 1, "maxpool_fwd: bad output size");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == STFB_BF16 && C % 8 == 0 && (k == 2 || k == 3) && aligned_to(x, 16) && aligned_to(y, 16) && N > 0) {
    const PixTile ptile = make_pix_tile(C / 8, Ho, Wo);
    const unsigned gy = (unsigned)((ptile.tiles_h + MP_TPB - 1) / MP_TPB);
    if (gy <= 65535) {
      const dim3 grid((unsigned)ptile.tiles_w, gy, (unsigned)(N < 65535 ? N : 65535));
      if (k == 3) maxpool_fwd_bf16_kernel<3><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C, Ho, Wo, stride, pad);
      else maxpool_fwd_bf16_kernel<2><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C, Ho, Wo, stride, pad);
      return post_launch("maxpool_fwd");
    }
  }
  const bool v = vec4_ok(C, dtype, {x, y});
  const long long tv = (long long)N * Ho * Wo * (C / (v ? 4 : 1));
  if (tv == 0) return STFB_OK;
  DISPATCH_T(dtype, {
    if (v) maxpool_fwd_kernel<T, 4><<<grid_for(tv), 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo, k, stride, pad, tv);
    else maxpool_fwd_kernel<T, 1><<<grid_for(tv), 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo, k, stride, pad, tv);
  });
  return post_launch("maxpool_fwd");
}

extern "C" int stfb_maxpool_fwd_idx(const void* x, void* y, unsigned char* idx, int N, int H, int W, int C, int Ho, int Wo, int k,
                                    int stride, int pad, int dtype, void* stream) {
  STFB_REQUIRE(x && y && idx && N >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && k <= 15 && stride > 0 && pad >= 0 && 2 * pad <= k && DT_OK(dtype),
               "maxpool_fwd_idx: bad arguments");
  STFB_REQUIRE(Ho == (H + 2 * pad - k) / stride + 1 && Wo == (W + 2 * pad - k) / stride + 1, "maxpool_fwd_idx: bad output size");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool v = (C % 8 == 0) && aligned_to(x, 16) && aligned_to(y, 16) && aligned_to(idx, 8);
  const PixTile ptile = make_pix_tile(C / (v ? 8 : 1), Ho, Wo);
  if (N == 0) return STFB_OK;
  STFB_REQUIRE(ptile.tiles_h <= 65535, "maxpool_fwd_idx: output too tall (%d rows)", Ho);
  const dim3 grid((unsigned)ptile.tiles_w, (unsigned)((ptile.tiles_h + MP_TPB - 1) / MP_TPB), (unsigned)(N < 65535 ? N : 65535));
#define MP_FWD(V, KC) maxpool_fwd_idx_kernel<T, V, KC><<<grid, 256, 0, s>>>((const T*)x, (T*)y, idx, N, H, W, C, Ho, Wo, k, stride, pad)
  DISPATCH_T(dtype, {
    if (v) { if (k == 3) MP_FWD(8, 3); else if (k == 2) MP_FWD(8, 2); else MP_FWD(8, 0); }
    else { if (k == 3) MP_FWD(1, 3); else if (k == 2) MP_FWD(1, 2); else MP_FWD(1, 0); }
  });
#undef MP_FWD
  return post_launch("maxpool_fwd_idx");
}

extern "C" int stfb_bn_relu_maxpool_from_stats(const void* x, const float* partial, int nblk, const float* gamma, const float* beta,
                                              void* y, unsigned char* idx, int G, int N, int H, int W, int C, int Ho, int Wo, int k,
                                              int stride, int pad, float eps, int dtype, void* stream) {
  STFB_REQUIRE(x && partial && gamma && beta && y && idx && G > 0 && N >= 0 && H > 0 && W > 0 && C > 0 && stride > 0 && pad >= 0,
               "bn_relu_maxpool_from_stats: bad arguments");
  STFB_REQUIRE(dtype == STFB_BF16 && C % 8 == 0 && (k == 2 || k == 3) && 2 * pad <= k, "bn_relu_maxpool_from_stats: bf16, C %% 8 == 0, k in {2, 3}");
  STFB_REQUIRE(nblk >= 1 && nblk <= 8, "bn_relu_maxpool_from_stats: 1..8 statistics slots (got %d)", nblk);
  STFB_REQUIRE(N % G == 0 && N <= 65535, "bn_relu_maxpool_from_stats: N (%d) must be a multiple of G (%d) and <= 65535", N, G);
  STFB_REQUIRE(Ho == (H + 2 * pad - k) / stride + 1 && Wo == (W + 2 * pad - k) / stride + 1, "bn_relu_maxpool_from_stats: bad output size");
  STFB_REQUIRE(aligned_to(x, 16) && aligned_to(y, 16) && aligned_to(idx, 8), "bn_relu_maxpool_from_stats: unaligned buffers");
  STFB_REQUIRE((size_t)2 * C * sizeof(float) <= 48 * 1024, "bn_relu_maxpool_from_stats: too many channels (%d)", C);
  STFB_DEVICE_OR_RETURN();
  if (N == 0) return STFB_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const PixTile ptile = make_pix_tile(C / 8, Ho, Wo);
  const unsigned gy = (unsigned)((ptile.tiles_h + MP_TPB - 1) / MP_TPB);
  STFB_REQUIRE(gy <= 65535, "bn_relu_maxpool_from_stats: output too tall (%d rows)", Ho);
  const dim3 grid((unsigned)ptile.tiles_w, gy > 1 ? (gy + 1) / 2 : 1, (unsigned)N);
  const size_t sm = (size_t)2 * C * sizeof(float);
  const long long R = (long long)(N / G) * H * W;
  const int row_groups = (int)gy;
  if (k == 3)
    bn_relu_maxpool_idx_kernel<3><<<grid, 256, sm, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, partial, nblk, gamma, beta, eps, G, R, N, H, W, C, Ho, Wo, stride, pad, row_groups);
  else
    bn_relu_maxpool_idx_kernel<2><<<grid, 256, sm, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, partial, nblk, gamma, beta, eps, G, R, N, H, W, C, Ho, Wo, stride, pad, row_groups);
  return post_launch("bn_relu_maxpool_from_stats");
}

extern "C" int stfb_maxpool_bwd_idx(const unsigned char* idx, const void* dy, void* dx, int N, int H, int W, int C, int Ho, int Wo,
                                    int k, int stride, int pad, int dtype, void* stream) {
  STFB_REQUIRE(idx && dy && dx && N >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && DT_OK(dtype), "maxpool_bwd_idx: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool v = (C % 8 == 0) && aligned_to(dy, 16) && aligned_to(dx, 16) && aligned_to(idx, 8);
  const PixTile ptile = make_pix_tile(C / (v ? 8 : 1), H, W);
  if (N == 0) return STFB_OK;
  const int wmax = (k + stride - 1) / stride;        // windows that can contain one input element, per axis
  STFB_REQUIRE(wmax <= 3, "maxpool_bwd_idx: k (%d) > 3 * stride (%d) is not supported", k, stride);
  if (v && dtype == STFB_BF16 && k == 3 && stride == 2 && pad == 1 && H % 2 == 0 && W % 2 == 0 && Ho == H / 2 && Wo == W / 2) {
    const int cslices = (C / 8 + 7) / 8;
    const unsigned gy = (unsigned)((H / 2 + MP3_BH * MP3_ROWS - 1) / (MP3_BH * MP3_ROWS));
    if (gy <= 65535) {
      const dim3 g3((unsigned)(((W / 2 + MP3_BW - 1) / MP3_BW) * cslices), gy, (unsigned)(N < 65535 ? N : 65535));
      maxpool3s2_bwd_idx_kernel<<<g3, 256, 0, s>>>(idx, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo);
      return post_launch("maxpool_bwd_idx");
    }
  }
  STFB_REQUIRE(ptile.tiles_h <= 65535, "maxpool_bwd_idx: input too tall (%d rows)", H);
  const dim3 grid((unsigned)ptile.tiles_w, (unsigned)((ptile.tiles_h + MP_TPB - 1) / MP_TPB), (unsigned)(N < 65535 ? N : 65535));
#define MP_BWD(V, WM) maxpool_bwd_idx_kernel<T, V, WM><<<grid, 256, 0, s>>>(idx, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, k, stride, pad)
  DISPATCH_T(dtype, {
    if (v) { if (wmax == 1) MP_BWD(8, 1); else if (wmax == 2) MP_BWD(8, 2); else MP_BWD(8, 3); }
    else { if (wmax == 1) MP_BWD(1, 1); else if (wmax == 2) MP_BWD(1, 2); else MP_BWD(1, 3); }
  });
#undef MP_BWD
  return post_launch("maxpool_bwd_idx");
}

extern "C" int stfb_maxpool_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int Ho, int Wo, int k,
                                int stride, int pad, int dtype, void* stream) {
  STFB_REQUIRE(x && dy && dx && N >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && DT_OK(dtype), "maxpool_bwd: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool v = vec4_ok(C, dtype, {x, dy, dx});
  const long long tv = (long long)N * H * W * (C / (v ? 4 : 1));
  if (tv == 0) return STFB_OK;
  DISPATCH_T(dtype, {
    if (v) maxpool_bwd_kernel<T, 4><<<grid_for(tv), 256, 0, s>>>((const T*)x, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, k, stride, pad, tv);
    else maxpool_bwd_kernel<T, 1><<<grid_for(tv), 256, 0, s>>>((const T*)x, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, k, stride, pad, tv);
  });
  return post_launch("maxpool_bwd");
}

extern "C" int stfb_bilinear_fwd(const void* x, void* y, int N, int H, int W, int C, int Ho, int Wo, int dtype, void* stream) {
  STFB_REQUIRE(x && y && N >= 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && DT_OK(dtype), "bilinear_fwd: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)N * Ho * Wo * C;
  if (total == 0) return STFB_OK;
  DISPATCH_T(dtype, { bilinear_fwd_kernel<T><<<grid_for(total), 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C, Ho, Wo, sy, sx, total); });
  return post_launch("bilinear_fwd");
}

extern "C" int stfb_bilinear_bwd(const void* dy, float* dx, int N, int H, int W, int C, int Ho, int Wo, int dtype, void* stream) {
  STFB_REQUIRE(dy && dx && N >= 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && DT_OK(dtype), "bilinear_bwd: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)N * Ho * Wo * C;
  if (total == 0) return STFB_OK;
  DISPATCH_T(dtype, { bilinear_bwd_kernel<T><<<grid_for(total), 256, 0, s>>>((const T*)dy, dx, N, H, W, C, Ho, Wo, sy, sx, total); });
  return post_launch("bilinear_bwd");
}

extern "C" int stfb_lstm_cell_fwd(const float* gates, const float* c_prev, void* acts, float* c_out, void* h_out, long long R,
                                  int C, int dtype, void* stream) {
  STFB_REQUIRE(gates && c_out && h_out && R >= 0 && C > 0 && C % 4 == 0 && DT_OK(dtype), "lstm_cell_fwd: bad arguments (C %% 4 == 0 required)");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long tv = R * (C / 4);
  if (tv == 0) return STFB_OK;
  DISPATCH_T(dtype, { lstm_cell_fwd_kernel<T><<<grid_for(tv), 256, 0, s>>>(gates, c_prev, (T*)acts, c_out, (T*)h_out, R, C, tv); });
  return post_launch("lstm_cell_fwd");
}

extern "C" int stfb_lstm_cell_bwd(const float* dh, float* dc, const void* acts, const float* c_prev, const float* c_cur,
                                  void* dgates, long long R, int C, int acts_il, int dtype, void* stream) {
  STFB_REQUIRE(dh && dc && acts && c_cur && dgates && R >= 0 && C > 0 && C % 4 == 0 && DT_OK(dtype), "lstm_cell_bwd: bad arguments");
  STFB_REQUIRE(!acts_il || C % 64 == 0, "lstm_cell_bwd: interleaved activations need C %% 64 == 0");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long tv = R * (C / 4);
  if (tv == 0) return STFB_OK;
  DISPATCH_T(dtype, { lstm_cell_bwd_kernel<T><<<grid_for(tv), 256, 0, s>>>(dh, dc, (const T*)acts, c_prev, c_cur, (T*)dgates, R, C, tv, acts_il); });
  return post_launch("lstm_cell_bwd");
}

extern "C" int stfb_pack_series(const float* x, void* y, int B, int T_, int C, int H, int W, int dtype, void* stream) {
  STFB_REQUIRE(x && y && B >= 0 && T_ > 0 && C > 0 && H > 0 && W > 0 && DT_OK(dtype), "pack_series: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * T_ * C * H * W;
  if (total == 0) return STFB_OK;
  DISPATCH_T(dtype, { pack_series_kernel<T><<<grid_for(total), 256, 0, s>>>(x, (T*)y, B, T_, C, H, W, total); });
  return post_launch("pack_series");
}

extern "C" int stfb_nhwc_to_nchw(const void* y, float* out, int N, int H, int W, int C, int dtype, void* stream) {
  STFB_REQUIRE(y && out && N >= 0 && H > 0 && W > 0 && C > 0 && DT_OK(dtype), "nhwc_to_nchw: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)N * H * W * C;
  if (total == 0) return STFB_OK;
  DISPATCH_T(dtype, { nhwc_to_nchw_kernel<T><<<grid_for(total), 256, 0, s>>>((const T*)y, out, N, H * W, C, total); });
  return post_launch("nhwc_to_nchw");
}

extern "C" int stfb_nchw_to_nhwc(const float* g, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  STFB_REQUIRE(g && out && N >= 0 && H > 0 && W > 0 && C > 0 && DT_OK(dtype), "nchw_to_nhwc: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)N * H * W * C;
  if (total == 0) return STFB_OK;
  DISPATCH_T(dtype, { nchw_to_nhwc_kernel<T><<<grid_for(total), 256, 0, s>>>(g, (T*)out, N, H * W, C, total); });
  return post_launch("nchw_to_nhwc");
}

extern "C" int stfb_cast(const void* src, int sd, void* dst, int dd, long long n, void* stream) {
  STFB_REQUIRE(src && dst && n >= 0 && DT_OK(sd) && DT_OK(dd), "cast: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) return STFB_OK;
  const int g = grid_for(n);
  if (sd == STFB_F32 && dd == STFB_BF16) cast_kernel<float, __nv_bfloat16><<<g, 256, 0, s>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == STFB_BF16 && dd == STFB_F32) cast_kernel<__nv_bfloat16, float><<<g, 256, 0, s>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == STFB_F32) cast_kernel<float, float><<<g, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  return post_launch("cast");
}

extern "C" int stfb_add_inplace(void* dst, const void* src, long long n, int dtype, void* stream) {
  STFB_REQUIRE(dst && src && n >= 0 && DT_OK(dtype), "add_inplace: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) return STFB_OK;
  DISPATCH_T(dtype, { add_inplace_kernel<T><<<grid_for(n), 256, 0, s>>>((T*)dst, (const T*)src, n); });
  return post_launch("add_inplace");
}

extern "C" int stfb_im2col_small(const void* x, void* out, int N, int H, int W, int Cin, int Ho, int Wo, int k, int stride, int pad,
                                 int Kpad, void* stream) {
  STFB_REQUIRE(x && out && N >= 0 && H > 0 && W > 0 && Cin > 0 && k > 0 && stride > 0 && pad >= 0, "im2col_small: bad arguments");
  STFB_REQUIRE(Kpad % 8 == 0 && Kpad >= k * k * Cin, "im2col_small: Kpad must be a multiple of 8 covering k*k*Cin");
  STFB_REQUIRE(Ho == (H + 2 * pad - k) / stride + 1 && Wo == (W + 2 * pad - k) / stride + 1, "im2col_small: bad output size");
  STFB_REQUIRE(aligned_to(out, 16), "im2col_small: out must be 16-byte aligned");
  STFB_DEVICE_OR_RETURN();
  const long long total = (long long)N * Ho * Wo * (Kpad / 8);
  if (total == 0) return STFB_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(total);
  const dim3 tgrid((unsigned)((Wo + IM_TP - 1) / IM_TP), (unsigned)((Ho + IM_TR - 1) / IM_TR), (unsigned)(N < 65535 ? N : 65535));
  const bool tiled_ok = tgrid.y <= 65535 && Kpad == (k * k * Cin + 63) / 64 * 64;
#define IM2COL_TILED(KC, CC, S_) \
  im2col_tiled_kernel<KC, CC, S_><<<tgrid, IM_PIX, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, N, H, W, Ho, Wo, pad)
#define IM2COL_LAUNCH(KC, CC)                                                                                         \
  im2col_small_kernel<KC, CC><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, N, H, W, Cin, Ho, Wo, k, \
                                                   stride, pad, Kpad, total)
  if (tiled_ok && k == 7 && Cin == 1 && stride == 2) IM2COL_TILED(7, 1, 2);        // STF stem (src/stf_lstm_unet.py:105)
  else if (tiled_ok && k == 7 && Cin == 4 && stride == 2) IM2COL_TILED(7, 4, 2);   // STF stem with PK maps
  else if (tiled_ok && k == 3 && Cin == 1 && stride == 1) IM2COL_TILED(3, 1, 1);   // UNet enc1.0 with one input channel
  else if (tiled_ok && k == 3 && Cin == 8 && stride == 1) IM2COL_TILED(3, 8, 1);   // UNet enc1.0, 8 DCE phases as channels
  else if (k == 7 && Cin == 1) IM2COL_LAUNCH(7, 1);
  else if (k == 7 && Cin == 4) IM2COL_LAUNCH(7, 4);
  else if (k == 3 && Cin == 1) IM2COL_LAUNCH(3, 1);
  else if (k == 3 && Cin == 8) IM2COL_LAUNCH(3, 8);
  else IM2COL_LAUNCH(0, 0);
#undef IM2COL_LAUNCH
#undef IM2COL_TILED
  return post_launch("im2col_small");
}

extern "C" int stfb_unpad_wgrad(float* dW, const float* src, int Cout, int Cin, int kh, int kw, int ld_src, void* stream) {
  STFB_REQUIRE(dW && src && Cout > 0 && Cin > 0 && kh > 0 && kw > 0 && ld_src >= Cin * kh * kw, "unpad_wgrad: bad arguments");
  STFB_DEVICE_OR_RETURN();
  const long long total = (long long)Cout * Cin * kh * kw;
  unpad_wgrad_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dW, src, Cout, Cin, kh * kw, ld_src);
  return post_launch("unpad_wgrad");
}

extern "C" int stfb_pack_series_maps(const float* x, const float* maps, void* y, int B, int T_, int Cx, int Cm, int H, int W,
                                     int dtype, void* stream) {
  STFB_REQUIRE(x && maps && y && B >= 0 && T_ > 0 && Cx > 0 && Cm > 0 && H > 0 && W > 0 && DT_OK(dtype), "pack_series_maps: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * T_ * (Cx + Cm) * H * W;
  if (total == 0) return STFB_OK;
  DISPATCH_T(dtype, { pack_series_maps_kernel<T><<<grid_for(total), 256, 0, s>>>(x, maps, (T*)y, B, T_, Cx, Cm, H, W, total); });
  return post_launch("pack_series_maps");
}

extern "C" int stfb_repeat(const void* src, void* dst, size_t bytes, int times, void* stream) {
  STFB_REQUIRE(src && dst && times >= 0, "repeat: bad arguments");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (int t = 0; t < times; ++t) {
    if (cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)t * bytes, src, bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
      set_error("repeat: cudaMemcpyAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
      return STFB_ECUDA;
    }
  }
  return STFB_OK;
}
