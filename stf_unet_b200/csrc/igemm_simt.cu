// fp32-FFMA implicit-GEMM convolution family (forward gather, transposed gather, weight gradient).
//
// This is the fp32-accurate kernel family (north_star: fp32 logits rel-err <= 1e-4 cannot ride bf16
// tensor cores) and the catch-all for shapes the tcgen05 family does not take (7x7 stem with Cin=1,
// stride-2 convs, transposed convs, Cout=2 head).  Activations NHWC, fp32 accumulate, fused epilogue.
#include "common.cuh"
#include "split.cuh"

namespace stfb {

constexpr int BM = 128;   // output pixels per CTA
constexpr int BK = 16;    // reduction slice
constexpr int NTHREADS = 256;

struct ConvArgs {
  stfb_conv_params p;
  long long M;   // N*Ho*Wo
  int Cin;       // C1 + C2
  int Ktot;      // kh*kw*Cin
  int fastA;     // 16-channel slices never straddle a tap / source and are 16B-aligned
  int vecB;      // weight rows can be read 4 at a time
  int vecY;      // outputs can be written 4 at a time
};

template <typename TI, typename TO, int TN>
__global__ void __launch_bounds__(NTHREADS) igemm_simt_kernel(const ConvArgs a) {
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const stfb_conv_params& p = a.p;
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tn = tid % 16, tm = tid / 16;

  // ---- A loader: row r of the tile, 8 consecutive k ----
  const int ar = tid % BM, ah = tid / BM;
  const long long am = m0 + ar;
  const bool arow_ok = am < a.M;
  int an = 0, aoy = 0, aox = 0;
  if (arow_ok) {
    const int hw = p.Ho * p.Wo;
    an = (int)(am / hw);
    const int rem = (int)(am - (long long)an * hw);
    aoy = rem / p.Wo;
    aox = rem - aoy * p.Wo;
  }
  const TI* __restrict__ x1 = reinterpret_cast<const TI*>(p.x);
  const TI* __restrict__ x2 = reinterpret_cast<const TI*>(p.x2);
  const TI* __restrict__ wp = reinterpret_cast<const TI*>(p.w);

  // ---- B loader: row kk, TN consecutive columns ----
  const int bk = tid / 16, bn = (tid % 16) * TN;

  float areg[8];
  float breg[TN];

  auto gather_coords = [&](int ky, int kx, int& iy, int& ix) -> bool {
    if (p.mode == STFB_CONV_FWD) {
      iy = aoy * p.stride - p.pad + ky;
      ix = aox * p.stride - p.pad + kx;
      return iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    } else {
      const int ty = aoy + p.pad - ky, tx = aox + p.pad - kx;
      if (ty < 0 || tx < 0) return false;
      iy = ty / p.stride;
      ix = tx / p.stride;
      return (iy * p.stride == ty) && (ix * p.stride == tx) && iy < p.H && ix < p.W;
    }
  };

  auto load_tiles = [&](int kt) {
    const int k0 = kt * BK;
    // A
    if (a.fastA) {
      const int tap = k0 / a.Cin, ci0 = k0 - tap * a.Cin + ah * 8;
      const int ky = tap / p.kw, kx = tap - ky * p.kw;
      int iy, ix;
      bool ok = arow_ok && gather_coords(ky, kx, iy, ix);
      if (ok) {
        const long long pix = ((long long)an * p.H + iy) * p.W + ix;
        f8 v = (ci0 < p.C1) ? ld8(x1 + pix * p.C1 + ci0) : ld8(x2 + pix * p.C2 + (ci0 - p.C1));
#pragma unroll
        for (int j = 0; j < 8; ++j) areg[j] = v.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) areg[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + ah * 8 + j;
        float v = 0.f;
        if (arow_ok && k < a.Ktot) {
          const int tap = k / a.Cin, ci = k - tap * a.Cin;
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          int iy, ix;
          if (gather_coords(ky, kx, iy, ix)) {
            const long long pix = ((long long)an * p.H + iy) * p.W + ix;
            v = (ci < p.C1) ? ld1(x1 + pix * p.C1 + ci) : ld1(x2 + pix * p.C2 + (ci - p.C1));
          }
        }
        areg[j] = v;
      }
    }
    // B
    const int k = k0 + bk;
    if (k < a.Ktot) {
      const TI* row = wp + (long long)k * p.ldw + n0 + bn;
      if (a.vecB && n0 + bn + TN <= p.Cout) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          f4 v = ld4(row + j);
          breg[j] = v.v[0]; breg[j + 1] = v.v[1]; breg[j + 2] = v.v[2]; breg[j + 3] = v.v[3];
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) breg[j] = (n0 + bn + j < p.Cout) ? ld1(row + j) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) breg[j] = 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][ah * 8 + j][ar] = areg[j];
#pragma unroll
    for (int j = 0; j < TN; ++j) Bs[buf][bk][bn + j] = breg[j];
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (a.Ktot + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[8], bv[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 8 + 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn * TN + j]);
        bv[j] = b0.x; bv[j + 1] = b0.y; bv[j + 2] = b0.z; bv[j + 3] = b0.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue ----
  TO* __restrict__ y = reinterpret_cast<TO*>(p.y);
  const TO* res = reinterpret_cast<const TO*>(p.residual);
  const int c0 = n0 + tn * TN;
  float cb[TN], cs[TN], ct[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int c = c0 + j;
    const bool ok = c < p.Cout;
    cb[j] = 0.f;
    if (ok && p.bias) cb[j] += p.bias[c];
    if (ok && p.bias2) cb[j] += p.bias2[c];
    cs[j] = (ok && p.scale) ? p.scale[c] : 1.f;
    ct[j] = (ok && p.shift) ? p.shift[c] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + tm * 8 + i;
    if (m >= a.M) continue;
    const long long off = m * p.Cout + c0;
    if (a.vecY && c0 + TN <= p.Cout) {
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        f4 v;
#pragma unroll
        for (int q = 0; q < 4; ++q) v.v[q] = (acc[i][j + q] + cb[j + q]) * cs[j + q] + ct[j + q];
        if (res) {
          f4 r = ld4(res + off + j);
#pragma unroll
          for (int q = 0; q < 4; ++q) v.v[q] += r.v[q];
        }
        if (p.relu) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v.v[q] = fmaxf(v.v[q], 0.f);
        }
        st4(y + off + j, v);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        if (c0 + j < p.Cout) {
          float v = (acc[i][j] + cb[j]) * cs[j] + ct[j];
          if (res) v += ld1(res + off + j);
          if (p.relu) v = fmaxf(v, 0.f);
          st1(y + off + j, v);
        }
      }
    }
  }
}

template <typename TI, typename TO>
static int launch_igemm(const ConvArgs& a, cudaStream_t st) {
  const stfb_conv_params& p = a.p;
  const long long mtiles = (a.M + BM - 1) / BM;
  if (mtiles > 2147483647LL) { set_error("conv2d: too many output pixels"); return STFB_EINVAL; }
  // wide tile only when it still fills the machine
  const bool wide = p.Cout >= 128 && mtiles * ((p.Cout + 127) / 128) >= 2LL * num_sms();
  if (wide) {
    dim3 grid((unsigned)mtiles, (unsigned)((p.Cout + 127) / 128));
    igemm_simt_kernel<TI, TO, 8><<<grid, NTHREADS, 0, st>>>(a);
  } else {
    dim3 grid((unsigned)mtiles, (unsigned)((p.Cout + 63) / 64));
    igemm_simt_kernel<TI, TO, 4><<<grid, NTHREADS, 0, st>>>(a);
  }
  return post_launch("conv2d(simt)");
}

// ---- pointwise convolutions with a handful of channels on one side (the 32 -> 2 segmentation head, forward and dgrad) ---------
// The implicit-GEMM tile above is 128 pixels x 64 channels: with 2 output channels 97 % of it is padding and the launch takes
// 51 us for a 17 MB map (profiles/r02_timeline_summary.txt).  These are plain streaming kernels: one pass over the pixels,
// weights (<= 8 x C floats) in shared memory.  Weight packing is the SIMT family's [(k)][n] with row stride ldw.
//   narrow OUT (Cout <= 8): one thread per pixel reads its C channels in 16-byte pieces and keeps Cout accumulators;
//   narrow IN  (Cin  <= 8): one thread per (pixel, 8 output channels).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) pointwise_narrow_out_kernel(const TI* __restrict__ x, const TI* __restrict__ w, const float* __restrict__ bias,
                                                                   TO* __restrict__ y, long long P, int Cin, int Cout, int ldw, int relu) {
  extern __shared__ float s_w[];                       // [Cin][8]
  for (int i = threadIdx.x; i < Cin * 8; i += blockDim.x) {
    const int k = i >> 3, n = i & 7;
    s_w[i] = n < Cout ? ld1(w + (long long)k * ldw + n) : 0.f;
  }
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = (bias != nullptr && n < Cout) ? __ldg(bias + n) : 0.f;
    const TI* xp = x + p * Cin;
    for (int k = 0; k < Cin; k += 8) {
      const f8 v = ld8(xp + k);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + (k + e) * 8), w1 = *reinterpret_cast<const float4*>(s_w + (k + e) * 8 + 4);
        acc[0] = fmaf(v.v[e], w0.x, acc[0]); acc[1] = fmaf(v.v[e], w0.y, acc[1]); acc[2] = fmaf(v.v[e], w0.z, acc[2]); acc[3] = fmaf(v.v[e], w0.w, acc[3]);
        acc[4] = fmaf(v.v[e], w1.x, acc[4]); acc[5] = fmaf(v.v[e], w1.y, acc[5]); acc[6] = fmaf(v.v[e], w1.z, acc[6]); acc[7] = fmaf(v.v[e], w1.w, acc[7]);
      }
    }
    for (int n = 0; n < Cout; ++n) st1(y + p * Cout + n, relu ? fmaxf(acc[n], 0.f) : acc[n]);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) pointwise_narrow_in_kernel(const TI* __restrict__ x, const TI* __restrict__ w, const float* __restrict__ bias,
                                                                  TO* __restrict__ y, long long P, int Cin, int Cout, int ldw, int relu) {
  extern __shared__ float s_w[];                       // [Cin][Cout]
  for (int i = threadIdx.x; i < Cin * Cout; i += blockDim.x) s_w[i] = ld1(w + (long long)(i / Cout) * ldw + (i % Cout));
  __syncthreads();
  const int chunks = Cout / 8;
  const long long total = P * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / chunks;
    const int c0 = (int)(i - p * chunks) * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? __ldg(bias + c0 + e) : 0.f;
    for (int k = 0; k < Cin; ++k) {
      const float xv = ld1(x + p * Cin + k);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv, s_w[k * Cout + c0 + e], acc[e]);
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.f);
    }
    st8(y + p * Cout + c0, acc);
  }
}

// -> STFB_OK when the launch was taken, 1 when the shape is not one of the narrow pointwise cases
template <typename TI, typename TO>
static int launch_pointwise_narrow(const stfb_conv_params* p, cudaStream_t st) {
  const int Cin = p->C1;
  const long long P = (long long)p->N * p->Ho * p->Wo;
  auto aligned = [](const void* q, int b) { return (reinterpret_cast<uintptr_t>(q) % b) == 0; };
  if (p->kh != 1 || p->kw != 1 || p->stride != 1 || p->pad != 0 || p->C2 != 0 || p->x2 != nullptr || p->scale != nullptr ||
      p->residual != nullptr || p->bias2 != nullptr || p->Ho != p->H || p->Wo != p->W)
    return 1;
  long long blocks = (P + 255) / 256;
  if (p->Cout <= 8 && Cin % 8 == 0 && Cin <= 512 && aligned(p->x, 16)) {
    if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
    pointwise_narrow_out_kernel<TI, TO><<<(unsigned)blocks, 256, (size_t)Cin * 8 * sizeof(float), st>>>(
        reinterpret_cast<const TI*>(p->x), reinterpret_cast<const TI*>(p->w), p->bias, reinterpret_cast<TO*>(p->y), P, Cin, p->Cout, p->ldw, p->relu);
    return post_launch("conv2d(pointwise, narrow out)");
  }
  if (Cin <= 8 && p->Cout % 8 == 0 && p->Cout <= 512 && aligned(p->y, 16)) {
    blocks = (P * (p->Cout / 8) + 255) / 256;
    if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
    pointwise_narrow_in_kernel<TI, TO><<<(unsigned)blocks, 256, (size_t)Cin * p->Cout * sizeof(float), st>>>(
        reinterpret_cast<const TI*>(p->x), reinterpret_cast<const TI*>(p->w), p->bias, reinterpret_cast<TO*>(p->y), P, Cin, p->Cout, p->ldw, p->relu);
    return post_launch("conv2d(pointwise, narrow in)");
  }
  return 1;
}

int conv2d_simt(const stfb_conv_params* p, cudaStream_t st) {
  if ((long long)p->N * p->Ho * p->Wo > 0 && (p->Cout <= 8 || p->C1 + p->C2 <= 8)) {
    int rc = 1;
    if (p->x_dtype == STFB_F32 && p->y_dtype == STFB_F32) rc = launch_pointwise_narrow<float, float>(p, st);
    else if (p->x_dtype == STFB_BF16 && p->y_dtype == STFB_BF16) rc = launch_pointwise_narrow<__nv_bfloat16, __nv_bfloat16>(p, st);
    else if (p->x_dtype == STFB_BF16 && p->y_dtype == STFB_F32) rc = launch_pointwise_narrow<__nv_bfloat16, float>(p, st);
    if (rc != 1) return rc;
  }
  ConvArgs a;
  a.p = *p;
  a.M = (long long)p->N * p->Ho * p->Wo;
  a.Cin = p->C1 + p->C2;
  a.Ktot = p->kh * p->kw * a.Cin;
  const int esz = p->x_dtype == STFB_BF16 ? 2 : 4;
  auto aligned = [](const void* q, int b) { return (reinterpret_cast<uintptr_t>(q) % b) == 0; };
  a.fastA = (a.Cin % 16 == 0) && (p->C1 % 16 == 0) && (p->C2 % 8 == 0) && aligned(p->x, 16) &&
            (p->x2 == nullptr || aligned(p->x2, 16));
  a.vecB = (p->ldw % 4 == 0) && aligned(p->w, 4 * esz);
  const int ysz = p->y_dtype == STFB_BF16 ? 2 : 4;
  a.vecY = (p->Cout % 4 == 0) && aligned(p->y, 4 * ysz) && (p->residual == nullptr || aligned(p->residual, 4 * ysz));
  if (a.M == 0) return STFB_OK;
  if (p->x_dtype == STFB_F32 && p->y_dtype == STFB_F32) return launch_igemm<float, float>(a, st);
  if (p->x_dtype == STFB_BF16 && p->y_dtype == STFB_BF16) return launch_igemm<__nv_bfloat16, __nv_bfloat16>(a, st);
  if (p->x_dtype == STFB_BF16 && p->y_dtype == STFB_F32) return launch_igemm<__nv_bfloat16, float>(a, st);
  set_error("conv2d: unsupported dtype pair (%d,%d)", p->x_dtype, p->y_dtype);
  return STFB_EINVAL;
}

// =================================================================================================
// weight gradient: dW[cp][cg_off+cg][ky][kx] += sum_pix P[pix,cp] * G[gather(pix,ky,kx), cg]
// GEMM view: rows = kg = (ky,kx,cg) (BM=128), cols = cp (BN=64), reduction over pixels (BK=16), split over
// gridDim.z with fp32 atomics into the reference-layout gradient.
// =================================================================================================
struct WgradArgs {
  const void* P;
  const void* G;
  float* dW;
  int N, Hp, Wp, Cp, Hg, Wg, Cg, cg_off, cg_total, kh, kw, stride, pad;
  long long npix;        // N*Hp*Wp
  long long chunk;       // pixels per z-slice (multiple of BK)
  int Kg;                // kh*kw*Cg
  int fastA;             // Cg % 8 == 0 and aligned
  int vecB;              // Cp % 4 == 0 and aligned
};

template <typename T>
__global__ void __launch_bounds__(NTHREADS) wgrad_simt_kernel(const WgradArgs a) {
  constexpr int BN = 64, TN = 4;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int kg0 = blockIdx.x * BM, cp0 = blockIdx.y * BN;
  const long long pbeg = (long long)blockIdx.z * a.chunk;
  const long long pend = min(a.npix, pbeg + a.chunk);
  if (pbeg >= pend) return;
  const int tn = tid % 16, tm = tid / 16;
  const T* __restrict__ Pp = reinterpret_cast<const T*>(a.P);
  const T* __restrict__ Gp = reinterpret_cast<const T*>(a.G);

  const int pl = tid / 16;            // pixel lane within the slice (both loaders)
  const int akg = kg0 + (tid % 16) * 8;
  const int bcp = cp0 + (tid % 16) * 4;
  // fast path: the 8 kg of this thread share a tap
  int f_ky = 0, f_kx = 0, f_cg = 0;
  const bool a_in = akg < a.Kg;
  if (a.fastA && a_in) {
    const int tap = akg / a.Cg;
    f_cg = akg - tap * a.Cg;
    f_ky = tap / a.kw;
    f_kx = tap - f_ky * a.kw;
  }
  float areg[8], breg[4];
  const int hw = a.Hp * a.Wp;

  auto load_tiles = [&](long long p0) {
    const long long pix = p0 + pl;
    const bool pok = pix < pend;
    int n = 0, py = 0, px = 0;
    if (pok) {
      n = (int)(pix / hw);
      const int rem = (int)(pix - (long long)n * hw);
      py = rem / a.Wp;
      px = rem - py * a.Wp;
    }
    if (a.fastA) {
      bool ok = pok && a_in;
      int iy = 0, ix = 0;
      if (ok) {
        iy = py * a.stride - a.pad + f_ky;
        ix = px * a.stride - a.pad + f_kx;
        ok = iy >= 0 && iy < a.Hg && ix >= 0 && ix < a.Wg;
      }
      if (ok) {
        f8 v = ld8(Gp + (((long long)n * a.Hg + iy) * a.Wg + ix) * a.Cg + f_cg);
#pragma unroll
        for (int j = 0; j < 8; ++j) areg[j] = v.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) areg[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = 0.f;
        const int kg = akg + j;
        if (pok && kg < a.Kg) {
          const int tap = kg / a.Cg, cg = kg - tap * a.Cg;
          const int ky = tap / a.kw, kx = tap - ky * a.kw;
          const int iy = py * a.stride - a.pad + ky, ix = px * a.stride - a.pad + kx;
          if (iy >= 0 && iy < a.Hg && ix >= 0 && ix < a.Wg)
            v = ld1(Gp + (((long long)n * a.Hg + iy) * a.Wg + ix) * a.Cg + cg);
        }
        areg[j] = v;
      }
    }
    if (pok) {
      const T* row = Pp + pix * a.Cp + bcp;
      if (a.vecB && bcp + 4 <= a.Cp) {
        f4 v = ld4(row);
        breg[0] = v.v[0]; breg[1] = v.v[1]; breg[2] = v.v[2]; breg[3] = v.v[3];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) breg[j] = (bcp + j < a.Cp) ? ld1(row + j) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) breg[j] = 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][pl][(tid % 16) * 8]) = make_float4(areg[0], areg[1], areg[2], areg[3]);
    *reinterpret_cast<float4*>(&As[buf][pl][(tid % 16) * 8 + 4]) = make_float4(areg[4], areg[5], areg[6], areg[7]);
    *reinterpret_cast<float4*>(&Bs[buf][pl][(tid % 16) * 4]) = make_float4(breg[0], breg[1], breg[2], breg[3]);
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (int)((pend - pbeg + BK - 1) / BK);
  load_tiles(pbeg);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(pbeg + (long long)(kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  const int khw = a.kh * a.kw;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kg = kg0 + tm * 8 + i;
    if (kg >= a.Kg) continue;
    const int tap = kg / a.Cg, cg = kg - tap * a.Cg;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int cp = cp0 + tn * 4 + j;
      if (cp < a.Cp) atomicAdd(a.dW + ((long long)cp * a.cg_total + a.cg_off + cg) * khw + tap, acc[i][j]);
    }
  }
}

int conv2d_wgrad_simt(const WgradArgs& a0, int dtype, cudaStream_t st) {
  WgradArgs a = a0;
  a.npix = (long long)a.N * a.Hp * a.Wp;
  a.Kg = a.kh * a.kw * a.Cg;
  if (a.npix == 0) return STFB_OK;
  const int esz = dtype == STFB_BF16 ? 2 : 4;
  auto aligned = [](const void* q, int b) { return (reinterpret_cast<uintptr_t>(q) % b) == 0; };
  a.fastA = (a.Cg % 8 == 0) && aligned(a.G, 16);
  a.vecB = (a.Cp % 4 == 0) && aligned(a.P, 4 * esz);
  const int gx = (a.Kg + BM - 1) / BM, gy = (a.Cp + 63) / 64;
  // split the pixel reduction until the grid covers ~4 waves, keeping >= 8 slices of work per CTA
  long long want = (4LL * num_sms() + (long long)gx * gy - 1) / ((long long)gx * gy);
  long long maxsplit = (a.npix + 8 * BK - 1) / (8 * BK);
  long long splits = want < 1 ? 1 : want;
  if (splits > maxsplit) splits = maxsplit;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  long long chunk = (a.npix + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  splits = (a.npix + chunk - 1) / chunk;
  a.chunk = chunk;
  dim3 grid(gx, gy, (unsigned)splits);
  if (dtype == STFB_F32) wgrad_simt_kernel<float><<<grid, NTHREADS, 0, st>>>(a);
  else wgrad_simt_kernel<__nv_bfloat16><<<grid, NTHREADS, 0, st>>>(a);
  return post_launch("conv2d_wgrad(simt)");
}

// 1x1 weight gradient with very few output channels (the 32 -> 2 segmentation head, src/stf_lstm_unet.py:137):
// dW[cp][cg] += sum_rows P[row][cp] * G[row][cg] is a pure streaming reduction over G (Cp <= 4 accumulators per channel),
// not a GEMM: the tiled FFMA kernel spent 150 us on it (one 128 x 64 tile, 2 useful columns).
template <typename T, int CPMAX>
__global__ void __launch_bounds__(256) wgrad_head_kernel(const T* __restrict__ P, const T* __restrict__ G, float* dW, long long rows,
                                                         int Cp, int Cg, int cg_off, int cg_total, long long rows_per_block) {
  __shared__ float red[256 * 8];
  const int tpr = Cg / 8, lanes = 256 / tpr;
  const int cl = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[CPMAX][8];
#pragma unroll
  for (int p = 0; p < CPMAX; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
  if (rl < lanes) {
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const f8 g = ld8(G + r * Cg + cl * 8);
      float pv[CPMAX];
#pragma unroll
      for (int p = 0; p < CPMAX; ++p) pv[p] = p < Cp ? ld1(P + r * Cp + p) : 0.f;
#pragma unroll
      for (int p = 0; p < CPMAX; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(pv[p], g.v[j], acc[p][j]);
    }
  }
  for (int p = 0; p < Cp; ++p) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = (rl < lanes) ? acc[p][j] : 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < Cg; t += 256) {        // column t = (cl = t / 8, j = t % 8)
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[(l * tpr + t / 8) * 8 + (t & 7)];
      atomicAdd(dW + (long long)p * cg_total + cg_off + t, s);
    }
    __syncthreads();
  }
}

static bool wgrad_head_ok(const void* P, const void* G, int Cp, int Cg, int kh, int kw, int stride, int pad, int Hp, int Wp, int Hg,
                          int Wg) {
  return kh == 1 && kw == 1 && stride == 1 && pad == 0 && Hp == Hg && Wp == Wg && Cp <= 4 && Cg % 8 == 0 && Cg <= 2048 &&
         256 % (Cg / 8) == 0 && (reinterpret_cast<uintptr_t>(G) % 16) == 0 && P != nullptr;
}

static int wgrad_head(const void* P, const void* G, float* dW, long long rows, int Cp, int Cg, int cg_off, int cg_total, int dtype,
                      cudaStream_t st) {
  if (rows == 0) return STFB_OK;
  long long nblk = (rows + 511) / 512;
  if (nblk > 4LL * num_sms()) nblk = 4LL * num_sms();
  const long long rpb = (rows + nblk - 1) / nblk;
  nblk = (rows + rpb - 1) / rpb;
  if (dtype == STFB_F32)
    wgrad_head_kernel<float, 4><<<(unsigned)nblk, 256, 0, st>>>((const float*)P, (const float*)G, dW, rows, Cp, Cg, cg_off, cg_total, rpb);
  else
    wgrad_head_kernel<__nv_bfloat16, 4><<<(unsigned)nblk, 256, 0, st>>>((const __nv_bfloat16*)P, (const __nv_bfloat16*)G, dW, rows, Cp,
                                                                      Cg, cg_off, cg_total, rpb);
  return post_launch("conv2d_wgrad(head)");
}

// =================================================================================================
// weight packing
// =================================================================================================
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, int D0, int D1, int khw,
                                   int k_is_dim1, int n_major, int flip, int ld, int gate_c) {
  const long long total = (long long)D0 * D1 * khw;
  const int Kc = k_is_dim1 ? D1 : D0, Nc = k_is_dim1 ? D0 : D1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n, k, tap;
    if (!n_major) {   // destination index i = (tap*Kc + k)*Nc + n
      n = (int)(i % Nc);
      const long long r = i / Nc;
      k = (int)(r % Kc);
      tap = (int)(r / Kc);
    } else {          // destination index i = (n*khw + tap)*Kc + k     ([N][(tap,k)], K-major rows for tcgen05)
      k = (int)(i % Kc);
      const long long r = i / Kc;
      tap = (int)(r % khw);
      n = (int)(r / khw);
    }
    const int d0 = k_is_dim1 ? n : k, d1 = k_is_dim1 ? k : n;
    const int stap = flip ? (khw - 1 - tap) : tap;   // spatial flip (ky,kx) -> (kh-1-ky, kw-1-kx)
    // n_major rows may be padded to `ld` elements (K padded to the 64-wide k-block; the caller zero-fills the pad).
    // gate_c > 0: LSTM gate interleave -- source row gate*C + u lands at (u/16)*64 + gate*16 + u%16, so one 256-column
    // GEMM tile holds i,f,g,o of the same 64 hidden units and every 64 columns the four gates of 16 units.
    int nd = n;
    if (gate_c > 0) { const int gate = n / gate_c, u = n - gate * gate_c; nd = (u / 64) * 256 + ((u % 64) / 16) * 64 + gate * 16 + (u % 16); }
    const long long di = n_major ? (long long)nd * ld + (long long)tap * Kc + k : i;
    st1(wp + di, w[((long long)d0 * D1 + d1) * khw + stap]);
  }
}

// One vector work item of the batched pack: EIGHT consecutive k of one n -> one 16-byte store per tap (and split segment) instead of
// eight 2-byte ones, the job search and the index arithmetic paid once per eight elements.  All taps of the eight source positions
// are read first (each position's khw floats are contiguous: one or two sectors, fetched once -- a tap-outer order re-touched 256
// scattered chunks per warp nine times and thrashed L1 on the dgrad packs).  KHW = compile-time tap count (registers), 0 = runtime
// count <= 9 with the tap-outer order.
template <typename T, int KHW>
__device__ __forceinline__ void pack_vec_item(const stfb_pack_job& jb, long long i, int Kc) {
  const int kc8 = Kc >> 3;
  const int n = (int)(i / kc8), k0 = (int)(i - (long long)n * kc8) * 8;
  int nd = n;
  if (jb.pad_ > 0) { const int gate = n / jb.pad_, u = n - gate * jb.pad_; nd = (u / 64) * 256 + ((u % 64) / 16) * 64 + gate * 16 + (u % 16); }
  // source of element e: [d0][d1][taps]; with k on dim 1 the eight elements are 8 * khw contiguous floats
  const long long sbase = jb.k_is_dim1 ? ((long long)n * jb.D1 + k0) * jb.khw : ((long long)k0 * jb.D1 + n) * jb.khw;
  const long long sstep = jb.k_is_dim1 ? (long long)jb.khw : (long long)jb.D1 * jb.khw;
  constexpr int NT = KHW > 0 ? KHW : 1;
  float vv[8][NT];
  if constexpr (KHW > 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int tp = 0; tp < KHW; ++tp) vv[e][tp] = __ldg(jb.src + sbase + e * sstep + tp);
  }
  const int ntap = KHW > 0 ? KHW : jb.khw;
#pragma unroll
  for (int tap = 0; tap < ntap; ++tap) {
    float v[8];
    if constexpr (KHW > 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (jb.flip & 1) ? vv[e][KHW - 1 - tap] : vv[e][tap];
    } else {
      const int stap = (jb.flip & 1) ? (jb.khw - 1 - tap) : tap;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(jb.src + sbase + e * sstep + stap);
    }
    if (jb.flip & 2) {
      // split-precision operand (STFB_BF16X3, csrc/split.cu): bf16 [n][tap][6 segments][k] whatever T is
      __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(jb.dst) + (long long)n * jb.ld + (long long)tap * 6 * Kc + k0;
#pragma unroll
      for (int seg = 0; seg < 6; ++seg) {
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = weight_plane(v[e], seg);
        st8(d16 + (long long)seg * Kc, w);
      }
    } else {
      st8(reinterpret_cast<T*>(jb.dst) + (long long)nd * jb.ld + (long long)tap * Kc + k0, v);
    }
  }
}

// all weight packs of a step in ONE launch: job table in device memory, binary search on the work-item offset.
// One work item = one (d0, d1) position of a weight = its khw filter taps: the taps are contiguous in the source
// ([D0][D1][kh][kw]), so a warp reads a contiguous run and writes khw coalesced rows of the packed operand.
template <typename T>
__global__ void pack_weights_batched_kernel(const stfb_pack_job* __restrict__ jobs, int njobs, long long total) {
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(&jobs[mid].start) <= gi) lo = mid; else hi = mid - 1;
    }
    const stfb_pack_job jb = jobs[lo];
    const long long i = gi - jb.start;
    const int Kc = jb.k_is_dim1 ? jb.D1 : jb.D0, Nc = jb.k_is_dim1 ? jb.D0 : jb.D1;
    if (jb.flip & 4) {
      // vector item: EIGHT consecutive k of one n (n_major rows, Kc % 8 == 0, khw <= 9)
      switch (jb.khw) {
        case 9: pack_vec_item<T, 9>(jb, i, Kc); break;
        case 4: pack_vec_item<T, 4>(jb, i, Kc); break;
        case 1: pack_vec_item<T, 1>(jb, i, Kc); break;
        default: pack_vec_item<T, 0>(jb, i, Kc); break;
      }
      continue;
    }
    int n, k;
    if (!jb.n_major) { n = (int)(i % Nc); k = (int)(i / Nc); }      // destination (tap, k, n): n fastest
    else { k = (int)(i % Kc); n = (int)(i / Kc); }                  // destination (n, tap, k): k fastest
    const int d0 = jb.k_is_dim1 ? n : k, d1 = jb.k_is_dim1 ? k : n;
    int nd = n;
    if (jb.pad_ > 0) { const int gate = n / jb.pad_, u = n - gate * jb.pad_; nd = (u / 64) * 256 + ((u % 64) / 16) * 64 + gate * 16 + (u % 16); }
    const float* sp = jb.src + ((long long)d0 * jb.D1 + d1) * jb.khw;
    if (jb.flip & 2) {
      // split-precision operand (STFB_BF16X3, csrc/split.cu): bf16 [n][tap][6 segments][k] whatever T is
      __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(jb.dst) + (long long)n * jb.ld + k;
      for (int tap = 0; tap < jb.khw; ++tap) {
        const float v = sp[(jb.flip & 1) ? (jb.khw - 1 - tap) : tap];
#pragma unroll
        for (int seg = 0; seg < 6; ++seg) d16[((long long)tap * 6 + seg) * Kc] = __float2bfloat16_rn(weight_plane(v, seg));
      }
      continue;
    }
    T* dp = reinterpret_cast<T*>(jb.dst);
    for (int tap = 0; tap < jb.khw; ++tap) {
      const int stap = (jb.flip & 1) ? (jb.khw - 1 - tap) : tap;
      const long long di = jb.n_major ? (long long)nd * jb.ld + (long long)tap * Kc + k : ((long long)tap * Kc + k) * Nc + n;
      st1(dp + di, sp[stap]);
    }
  }
}

}  // namespace stfb

using namespace stfb;

extern "C" int stfb_pack_weights_batched(const stfb_pack_job* jobs_dev, int njobs, long long total, int dtype, void* stream) {
  STFB_REQUIRE(jobs_dev && njobs > 0 && total > 0, "pack_weights_batched: bad arguments");
  STFB_REQUIRE(dtype == STFB_F32 || dtype == STFB_BF16, "pack_weights_batched: bad dtype");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int blocks = ceil_div(total, 256);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  if (dtype == STFB_F32) pack_weights_batched_kernel<float><<<blocks, 256, 0, s>>>(jobs_dev, njobs, total);
  else pack_weights_batched_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(jobs_dev, njobs, total);
  return post_launch("pack_weights_batched");
}

namespace stfb {
int conv2d_tcgen05(const stfb_conv_params* p, cudaStream_t st);
int conv2d_tcgen05_supported(const stfb_conv_params* p);
int conv2d_stats_fusable(const stfb_conv_params* p, int groups);
int wgrad_tcgen05_supported(int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg, int kh, int kw, int stride, int pad,
                            int dtype, const void* P, const void* G);
int wgrad_tcgen05(const void* P, const void* G, const void* G2, float* dW, int N, int H, int W, int Cp, int Hg, int Wg,
                  int C1, int C2, int cg_off, int cg_total, int kh, int kw, int stride, int pad, float* ws, size_t ws_bytes,
                  cudaStream_t st, int split = 0);
size_t wgrad_tcgen05_workspace_bytes(int N, int H, int W, int Cp, int cg_total, int kh, int kw);
int wgrad_pairs_supported(int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg, int cg_off, int cg_total, int kh, int kw,
                          int stride, int pad, int dtype, const void* P, const void* G);
size_t wgrad_pairs_workspace_bytes(int Cp, int Cg);
void wgrad_set_scratch(void* p, size_t bytes);
size_t wgrad_scratch_bytes();
int wgrad_pairs(const void* P, const void* G, float* dW, int N, int H, int W, int Cp, int Cg, float* ws, size_t ws_bytes,
                cudaStream_t st);
int wgrad_scatter_batched(const stfb_scatter_job* jobs_dev, int njobs, long long total, const float* acc, float* grad,
                          cudaStream_t st);
int lstm_step_tcgen05(const void* x_t, const void* h_prev, const void* w_xh_il, const float* b_ih, const float* b_hh,
                      const float* c_prev, float* c_out, void* h_out, void* acts, int N, int H, int W, int C, cudaStream_t st);
}

static int validate_conv(const stfb_conv_params* p) {
  STFB_REQUIRE(p != nullptr, "conv2d: null params");
  STFB_REQUIRE(p->x && p->w && p->y, "conv2d: null x/w/y");
  STFB_REQUIRE(p->N >= 0 && p->H > 0 && p->W > 0 && p->C1 > 0 && p->C2 >= 0 && p->Ho > 0 && p->Wo > 0 && p->Cout > 0,
               "conv2d: bad dims N=%d H=%d W=%d C1=%d C2=%d Ho=%d Wo=%d Cout=%d", p->N, p->H, p->W, p->C1, p->C2, p->Ho,
               p->Wo, p->Cout);
  STFB_REQUIRE(p->C2 == 0 || p->x2 != nullptr, "conv2d: C2 > 0 needs x2");
  STFB_REQUIRE(p->kh > 0 && p->kw > 0 && p->stride > 0 && p->pad >= 0, "conv2d: bad kernel geometry");
  if (p->impl == STFB_IMPL_TCGEN05) {
    const int kmul = p->x_dtype == STFB_BF16X3 ? 6 : 1;      // split-precision operands walk six K segments per channel
    STFB_REQUIRE(p->ldw >= p->kh * p->kw * kmul * (p->C1 + p->C2) && p->ldw % 8 == 0,
                 "conv2d(tcgen05): weights are [Cout][ldw] K-major, ldw (%d) must be >= kh*kw*Cin (x6 for bf16x3) and a multiple of 8", p->ldw);
  } else {
    STFB_REQUIRE(p->x_dtype != STFB_BF16X3, "conv2d: bf16x3 operands need the tcgen05 family (impl = STFB_IMPL_TCGEN05)");
    STFB_REQUIRE(p->ldw >= p->Cout, "conv2d: ldw (%d) < Cout (%d)", p->ldw, p->Cout);
  }
  STFB_REQUIRE(p->mode == STFB_CONV_FWD || p->mode == STFB_CONV_TRANSPOSED, "conv2d: bad mode %d", p->mode);
  STFB_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "conv2d: scale and shift come together");
  if (p->mode == STFB_CONV_FWD) {
    STFB_REQUIRE(p->Ho == (p->H + 2 * p->pad - p->kh) / p->stride + 1 && p->Wo == (p->W + 2 * p->pad - p->kw) / p->stride + 1,
                 "conv2d: output size %dx%d does not match floor((H+2p-k)/s)+1", p->Ho, p->Wo);
  } else {
    const int lo_h = (p->H - 1) * p->stride - 2 * p->pad + p->kh, lo_w = (p->W - 1) * p->stride - 2 * p->pad + p->kw;
    STFB_REQUIRE(p->Ho >= lo_h && p->Ho < lo_h + p->stride && p->Wo >= lo_w && p->Wo < lo_w + p->stride,
                 "conv2d(transposed): output size %dx%d outside [(H-1)s-2p+k, +stride)", p->Ho, p->Wo);
  }
  return STFB_OK;
}

extern "C" int stfb_conv2d_tcgen05_supported(const stfb_conv_params* p) {
  if (p == nullptr) return 0;
  return stfb::conv2d_tcgen05_supported(p);   // pure shape / dtype / alignment check; w and ldw are not inspected
}

extern "C" int stfb_conv2d_stats_fusable(const stfb_conv_params* p, int groups) {
  if (p == nullptr) return 0;
  return stfb::conv2d_stats_fusable(p, groups);
}

extern "C" int stfb_conv2d(const stfb_conv_params* p, void* stream) {
  int st = validate_conv(p);
  if (st != STFB_OK) return st;
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (p->impl == STFB_IMPL_TCGEN05) {
    if (!stfb::conv2d_tcgen05_supported(p)) {
      set_error("conv2d: shape not supported by the tcgen05 family");
      return STFB_ENOTSUP;
    }
    return stfb::conv2d_tcgen05(p, s);
  }
  STFB_REQUIRE(p->stat_partial == nullptr, "conv2d: fused BatchNorm statistics need the tcgen05 family");
  return conv2d_simt(p, s);   // STFB_IMPL_AUTO == SIMT: the two families take different weight packings
}

extern "C" int stfb_conv2d_wgrad_tcgen05_supported(const void* P, const void* G, int N, int Hp, int Wp, int Cp, int Hg, int Wg,
                                                   int Cg, int kh, int kw, int stride, int pad, int dtype) {
  return stfb::wgrad_tcgen05_supported(N, Hp, Wp, Cp, Hg, Wg, Cg, kh, kw, stride, pad, dtype, P, G);
}

extern "C" size_t stfb_conv2d_wgrad_workspace_bytes(const void* P, const void* G, int N, int Hp, int Wp, int Cp, int Hg, int Wg,
                                                    int Cg, int cg_total, int kh, int kw, int stride, int pad, int dtype,
                                                    int impl) {
  if (impl == STFB_IMPL_SIMT) return 0;
  if (stfb::wgrad_pairs_supported(N, Hp, Wp, Cp, Hg, Wg, Cg, 0, cg_total, kh, kw, stride, pad, dtype, P, G))
    return stfb::wgrad_pairs_workspace_bytes(Cp, Cg);
  if (!stfb::wgrad_tcgen05_supported(N, Hp, Wp, Cp, Hg, Wg, Cg, kh, kw, stride, pad, dtype, P, G)) return 0;
  return stfb::wgrad_tcgen05_workspace_bytes(N, Hp, Wp, Cp, cg_total, kh, kw);
}

extern "C" size_t stfb_wgrad_scratch_bytes() {
  if (stfb::check_device() != STFB_OK) return 0;
  return stfb::wgrad_scratch_bytes();
}

extern "C" int stfb_set_wgrad_scratch(void* scratch, size_t bytes) {
  STFB_REQUIRE(scratch == nullptr || (reinterpret_cast<uintptr_t>(scratch) % 16) == 0, "set_wgrad_scratch: 16-byte alignment required");
  stfb::wgrad_set_scratch(scratch, scratch ? bytes : 0);
  return STFB_OK;
}

extern "C" int stfb_wgrad_scatter_batched(const stfb_scatter_job* jobs_dev, int njobs, long long total, const float* acc_flat,
                                          float* grad_flat, void* stream) {
  STFB_REQUIRE(jobs_dev && njobs > 0 && total > 0 && acc_flat && grad_flat, "wgrad_scatter_batched: bad arguments");
  STFB_DEVICE_OR_RETURN();
  return stfb::wgrad_scatter_batched(jobs_dev, njobs, total, acc_flat, grad_flat, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stfb_conv2d_wgrad(const void* P, const void* G, float* dW, int N, int Hp, int Wp, int Cp, int Hg, int Wg,
                                 int Cg, int cg_off, int cg_total, int kh, int kw, int stride, int pad, int dtype, int impl,
                                 void* workspace, size_t ws_bytes, void* stream) {
  STFB_REQUIRE(P && G && (dW || workspace), "conv2d_wgrad: null pointer");
  STFB_REQUIRE(N >= 0 && Hp > 0 && Wp > 0 && Cp > 0 && Hg > 0 && Wg > 0 && Cg > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0,
               "conv2d_wgrad: bad dims");
  STFB_REQUIRE(cg_off >= 0 && cg_off + Cg <= cg_total, "conv2d_wgrad: channel window [%d,%d) outside %d", cg_off, cg_off + Cg, cg_total);
  STFB_REQUIRE(dtype == STFB_F32 || dtype == STFB_BF16 || dtype == STFB_BF16X3, "conv2d_wgrad: bad dtype");
  STFB_REQUIRE(dtype != STFB_BF16X3 || impl != STFB_IMPL_SIMT, "conv2d_wgrad: bf16x3 operands need the tcgen05 family");
  STFB_DEVICE_OR_RETURN();
  if (dtype == STFB_BF16X3) {
    if (!stfb::wgrad_tcgen05_supported(N, Hp, Wp, Cp, Hg, Wg, Cg, kh, kw, stride, pad, dtype, P, G)) {
      set_error("conv2d_wgrad: bf16x3 shape not supported by the tcgen05 family");
      return STFB_ENOTSUP;
    }
    return stfb::wgrad_tcgen05(P, G, nullptr, dW, N, Hp, Wp, Cp, Hg, Wg, Cg, 0, cg_off, cg_total, kh, kw, stride, pad,
                               reinterpret_cast<float*>(workspace), ws_bytes, reinterpret_cast<cudaStream_t>(stream), 1);
  }
  if (impl != STFB_IMPL_SIMT) {
    const int ok = stfb::wgrad_tcgen05_supported(N, Hp, Wp, Cp, Hg, Wg, Cg, kh, kw, stride, pad, dtype, P, G);
    if (ok) return stfb::wgrad_tcgen05(P, G, nullptr, dW, N, Hp, Wp, Cp, Hg, Wg, Cg, 0, cg_off, cg_total, kh, kw, stride, pad,
                                       reinterpret_cast<float*>(workspace), ws_bytes, reinterpret_cast<cudaStream_t>(stream));
    if (dW != nullptr && workspace != nullptr &&
        stfb::wgrad_pairs_supported(N, Hp, Wp, Cp, Hg, Wg, Cg, cg_off, cg_total, kh, kw, stride, pad, dtype, P, G))
      return stfb::wgrad_pairs(P, G, dW, N, Hp, Wp, Cp, Cg, reinterpret_cast<float*>(workspace), ws_bytes,
                               reinterpret_cast<cudaStream_t>(stream));
    if (impl == STFB_IMPL_TCGEN05) {
      set_error("conv2d_wgrad: shape not supported by the tcgen05 family");
      return STFB_ENOTSUP;
    }
  }
  STFB_REQUIRE(dW != nullptr, "conv2d_wgrad: deferred mode (dW = NULL) needs the tcgen05 family");
  if (wgrad_head_ok(P, G, Cp, Cg, kh, kw, stride, pad, Hp, Wp, Hg, Wg))
    return wgrad_head(P, G, dW, (long long)N * Hp * Wp, Cp, Cg, cg_off, cg_total, dtype, reinterpret_cast<cudaStream_t>(stream));
  WgradArgs a{};
  a.P = P; a.G = G; a.dW = dW; a.N = N; a.Hp = Hp; a.Wp = Wp; a.Cp = Cp; a.Hg = Hg; a.Wg = Wg; a.Cg = Cg;
  a.cg_off = cg_off; a.cg_total = cg_total; a.kh = kh; a.kw = kw; a.stride = stride; a.pad = pad;
  return conv2d_wgrad_simt(a, dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stfb_pack_weight_ex(const float* w, void* wp, int D0, int D1, int kh, int kw, int k_is_dim1, int n_major,
                                   int flip, int ld, int gate_c, int dtype, void* stream) {
  STFB_REQUIRE(w && wp && D0 > 0 && D1 > 0 && kh > 0 && kw > 0, "pack_weight: bad arguments");
  {
    const int Kc_ = k_is_dim1 ? D1 : D0;
    if (ld <= 0) ld = kh * kw * Kc_;
    STFB_REQUIRE(!n_major || ld >= kh * kw * Kc_, "pack_weight: ld (%d) smaller than the packed row (%d)", ld, kh * kw * Kc_);
    STFB_REQUIRE(gate_c == 0 || (n_major && k_is_dim1 && gate_c % 64 == 0 && D0 == 4 * gate_c),
                 "pack_weight: gate interleave needs an n_major [4C][C] LSTM matrix with C %% 64 == 0");
  }
  STFB_REQUIRE(dtype == STFB_F32 || dtype == STFB_BF16, "pack_weight: bad dtype");
  STFB_DEVICE_OR_RETURN();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)D0 * D1 * kh * kw;
  int blocks = ceil_div(total, 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  if (dtype == STFB_F32)
    pack_weight_kernel<float><<<blocks, 256, 0, s>>>(w, reinterpret_cast<float*>(wp), D0, D1, kh * kw, k_is_dim1, n_major, flip, ld, gate_c);
  else
    pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(w, reinterpret_cast<__nv_bfloat16*>(wp), D0, D1, kh * kw, k_is_dim1, n_major, flip, ld, gate_c);
  return post_launch("pack_weight");
}

extern "C" int stfb_pack_weight(const float* w, void* wp, int D0, int D1, int kh, int kw, int k_is_dim1, int dtype,
                                void* stream) {
  return stfb_pack_weight_ex(w, wp, D0, D1, kh, kw, k_is_dim1, 0, 0, 0, 0, dtype, stream);
}

namespace stfb {
int lstm_seq_supported(int T, int B, int H, int W, int C);
int lstm_seq_tcgen05(const void* x_seq, const void* w_xh_il, const float* b_ih, const float* b_hh, float* c_all, void* h_all,
                     void* acts_all, int T, int B, int H, int W, int C, int keep, cudaStream_t st);
}

namespace stfb {
int lstm_bwd_step_tcgen05(const void* dg_next, const void* w_hh_d, const void* acts, const float* c_prev, const float* c_cur,
                          float* dc, void* dg_out, int N, int H, int W, int C, cudaStream_t st);
}

extern "C" int stfb_lstm_bwd_step_fused(const void* dg_next, const void* w_hh_d, const void* acts, const float* c_prev,
                                        const float* c_cur, float* dc, void* dg_out, int N, int H, int W, int C, void* stream) {
  STFB_REQUIRE(dg_next && w_hh_d && acts && c_cur && dc && dg_out && N >= 0 && H > 0 && W > 0, "lstm_bwd_step_fused: bad arguments");
  STFB_REQUIRE(C > 0 && C % 64 == 0, "lstm_bwd_step_fused: hidden size must be a multiple of 64 (got %d)", C);
  STFB_REQUIRE(dg_next != dg_out, "lstm_bwd_step_fused: dg_out must not alias dg_next (other tiles still read it)");
  STFB_REQUIRE((long long)N * H * W < 2000000000LL, "lstm_bwd_step_fused: too many rows");
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0; };
  STFB_REQUIRE(al(dg_next) && al(w_hh_d) && al(acts) && al(c_prev) && al(c_cur) && al(dc) && al(dg_out),
               "lstm_bwd_step_fused: pointers must be 16-byte aligned");
  STFB_DEVICE_OR_RETURN();
  return stfb::lstm_bwd_step_tcgen05(dg_next, w_hh_d, acts, c_prev, c_cur, dc, dg_out, N, H, W, C,
                                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stfb_lstm_seq_supported(int T, int B, int H, int W, int C) { return stfb::lstm_seq_supported(T, B, H, W, C); }

extern "C" int stfb_lstm_seq_fused(const void* x_seq, const void* w_xh_il, const float* b_ih, const float* b_hh, float* c_all,
                                   void* h_all, void* acts_all, int T, int B, int H, int W, int C, int keep, void* stream) {
  STFB_REQUIRE(x_seq && w_xh_il && b_ih && b_hh && h_all && T > 0 && B >= 0 && H > 0 && W > 0, "lstm_seq_fused: bad arguments");
  STFB_REQUIRE(!keep || (c_all && acts_all), "lstm_seq_fused: training (keep) needs c_all and acts_all");
  STFB_REQUIRE((long long)T * B * H * W < 2000000000LL, "lstm_seq_fused: too many rows");
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0; };
  STFB_REQUIRE(al(x_seq) && al(w_xh_il) && al(b_ih) && al(b_hh) && al(c_all) && al(h_all) && al(acts_all),
               "lstm_seq_fused: pointers must be 16-byte aligned");
  STFB_DEVICE_OR_RETURN();
  return stfb::lstm_seq_tcgen05(x_seq, w_xh_il, b_ih, b_hh, c_all, h_all, acts_all, T, B, H, W, C, keep,
                                reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stfb_lstm_step_fused(const void* x_t, const void* h_prev, const void* w_xh_il, const float* b_ih,
                                    const float* b_hh, const float* c_prev, float* c_out, void* h_out, void* acts, int N, int H,
                                    int W, int C, void* stream) {
  STFB_REQUIRE(x_t && w_xh_il && b_ih && b_hh && c_out && h_out && N >= 0 && H > 0 && W > 0, "lstm_step_fused: bad arguments");
  STFB_REQUIRE(C > 0 && C % 64 == 0, "lstm_step_fused: hidden size must be a multiple of 64 (got %d)", C);
  STFB_REQUIRE(h_prev == nullptr || c_prev != nullptr, "lstm_step_fused: h_prev needs c_prev");
  STFB_REQUIRE(h_prev != h_out && x_t != h_out, "lstm_step_fused: h_out must not alias an input (other tiles still read it)");
  STFB_REQUIRE((long long)N * H * W < 2000000000LL, "lstm_step_fused: too many rows");
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0; };
  STFB_REQUIRE(al(x_t) && al(h_prev) && al(w_xh_il) && al(b_ih) && al(b_hh) && al(c_prev) && al(c_out) && al(h_out) && al(acts),
               "lstm_step_fused: pointers must be 16-byte aligned");
  STFB_DEVICE_OR_RETURN();
  return stfb::lstm_step_tcgen05(x_t, h_prev, w_xh_il, b_ih, b_hh, c_prev, c_out, h_out, acts, N, H, W, C,
                                 reinterpret_cast<cudaStream_t>(stream));
}
