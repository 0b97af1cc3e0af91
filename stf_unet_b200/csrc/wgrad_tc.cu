// tcgen05 / TMEM weight-gradient kernel for stride-1 "same" convolutions (and 1x1 GEMMs: LSTM matrices).
//
//   dW[co][ci][ky][kx] += sum_pix dy[pix, co] * x[pix + (ky - pad, kx - pad), ci]
//
// GEMM view: D[m = (tap, ci)][n = co] = sum_k A[m][k = pix] * B[n][k = pix].  Both operands are read straight out of
// the NHWC activations, whose contiguous axis is the channel axis, i.e. the M / N axis of this GEMM: they are
// "MN-major" operands.  A 4-D TMA box {64 ch, TW, TH, TN} (64 pixels) lands in smem as 64 rows x 128 B with the
// 128-byte swizzle = one MN-major SWIZZLE_128B column block; tcgen05.mma consumes it with a_major = b_major = MN.
//   * A tile (M = 128) = two independent 64-channel column blocks: they may belong to different filter taps, so
//     64-channel layers still fill the 128-row datapath (block j of the tile is the x patch shifted by ITS tap).
//   * B tile (N = BN <= 256) = BN/64 column blocks of the dy patch (unshifted).
//   * The pixel reduction is split across CTAs (gridDim.z); each CTA accumulates its pixel range in TMEM and writes its
//     fp32 tile to a workspace slice [split][(tap,ci)][co] with coalesced 128-byte row stores; a second small kernel sums
//     the slices in a fixed order and adds the result into dW in the reference layout.  No atomics: the first version
//     used red.global.add.f32 straight into [co][ci][ky][kx] and was bound by the L2 atomic units (10 M scattered
//     sector updates per launch, tensor pipe 12-20 % busy -- profiles/r01_ncu_wgrad_atomics.txt); it is also
//     bitwise deterministic now.
#include "tc_common.cuh"
#include <stdlib.h>

namespace stfb {

struct WgTcArgs {
  float* ws;                // [splits][Kg][Cp] fp32 partial tiles
  int N, H, W, Cp;          // P (dy) is [N, H, W, Cp]; G / G2 are [N, Hg, Wg, C1 / C2] with H = (Hg + 2 pad - k)/stride + 1
  int C1, C2;
  int cg_off, cg_total;     // dW channel window (ci axis) of this launch inside the full weight
  int kh, kw, pad;
  int g_scale;              // conv stride: G-map coordinate = patch pixel * g_scale - pad + tap (TMA elementStrides)
  int TW, TH, TN;           // 64-pixel patch
  int tiles_w, tiles_h, n_patches;
  int patches_per_split;
  int m_chunks;             // kh*kw*(C1+C2)/64 column blocks on the M axis
  int debug;                // STFB_WG_DEBUG: 1 = no MMAs (TMA pipeline only), 2 = no TMA (MMA issue only); timing experiments
};

constexpr int WG_PIX = 64;                       // K per stage
constexpr int WG_BLK_BYTES = WG_PIX * 128;       // one 64-ch x 64-pixel column block = 8 KB
constexpr int WG_THREADS = 192;

template <int BN, int STAGES>
constexpr int wg_smem_bytes() {
  return STAGES * (2 + BN / 64) * WG_BLK_BYTES + 256 + 1024;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG,
                                                               const __grid_constant__ CUtensorMap tmG2,
                                                               const __grid_constant__ CUtensorMap tmP, const WgTcArgs a) {
  constexpr int NB = BN / 64;
  constexpr int STAGE_BYTES = (2 + NB) * WG_BLK_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int mtile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int p_beg = blockIdx.z * a.patches_per_split;
  const int p_end = min(a.n_patches, p_beg + a.patches_per_split);
  const int Cin = a.C1 + a.C2;
  const int cpt = Cin / 64;                       // column blocks per tap
  // the two column blocks of this M tile (the second falls back to the first when the tile is half empty)
  const int chunk0 = mtile * 2;
  const bool has1 = chunk0 + 1 < a.m_chunks;
  const int chunk1 = has1 ? chunk0 + 1 : chunk0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmG);
    prefetch_tensormap(&tmP);
    if (a.C2 > 0) prefetch_tensormap(&tmG2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  const int n_iter = p_end - p_beg;

  if (warp == 0) {
    if (lane == 0 && n_iter > 0) {
      int stage = 0;
      uint32_t phase = 0;
      int tapc[2], cc[2];
      tapc[0] = chunk0 / cpt; cc[0] = (chunk0 - tapc[0] * cpt) * 64;
      tapc[1] = chunk1 / cpt; cc[1] = (chunk1 - tapc[1] * cpt) * 64;
      for (int p = p_beg; p < p_end; ++p) {
        int t = p;
        const int wb = t % a.tiles_w; t /= a.tiles_w;
        const int hb = t % a.tiles_h;
        const int nb = t / a.tiles_h;
        const int w0 = wb * a.TW, h0 = hb * a.TH, i0 = nb * a.TN;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        if (a.debug == 2) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = tapc[j] / a.kw, s = tapc[j] - r * a.kw;
          const int gw = w0 * a.g_scale - a.pad + s, gh = h0 * a.g_scale - a.pad + r;
          if (cc[j] < a.C1) tma_load_4d(sa + j * WG_BLK_BYTES, &tmG, &full_bar[stage], cc[j], gw, gh, i0);
          else tma_load_4d(sa + j * WG_BLK_BYTES, &tmG2, &full_bar[stage], cc[j] - a.C1, gw, gh, i0);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i)
          tma_load_4d(sa + (2 + i) * WG_BLK_BYTES, &tmP, &full_bar[stage], n0 + i * 64, w0, h0, i0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_mnmajor(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_mnmajor_sw128_desc(sa, WG_BLK_BYTES);
        const uint64_t bdesc = make_mnmajor_sw128_desc(sa + 2 * WG_BLK_BYTES, WG_BLK_BYTES);
        if (a.debug != 1) {
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k) {
            // 16 pixels = 16 rows of 128 B = 2048 B further down each column block
            umma_bf16(tmem_acc, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (it == n_iter - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (n_iter > 0) {
    // epilogue: row m of the accumulator = (column block m / 64, channel m % 64) = row kg of the [Kg][Cp] slice
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool valid = (m < 64) || has1;
    const long long kg = (long long)chunk0 * 64 + m;
    float* rowp = a.ws + ((long long)blockIdx.z * a.m_chunks * 64 + kg) * a.Cp + n0;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr + c0, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(rowp + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN);
  }
}

// dW[co][cg_off + ci][tap] += sum_s ws[s][(tap, ci)][co]   (fixed summation order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dW, int splits, int Kg, int Cp, int Cg,
                                    int khw, int cg_off, int cg_total) {
  const long long total = (long long)Kg * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cp);
    const int kg = (int)(i / Cp);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * total + i];
    const int tap = kg / Cg, ci = kg - tap * Cg;
    dW[((long long)co * cg_total + cg_off + ci) * khw + tap] += acc;
  }
}

static bool g_wg_strided = true;   // TMA elementStrides path (stride-2 convolutions / transposed convolutions)
static bool wgrad_strided_ok() { return g_wg_strided; }

static int wg_pick_bn(int Cp) {
  if (Cp % 256 == 0) return 256;
  if (Cp % 128 == 0) return 128;
  if (Cp % 64 == 0) return 64;
  return 0;
}

int wgrad_tcgen05_supported(int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg, int kh, int kw, int stride, int pad,
                            int dtype, const void* P, const void* G) {
  if (dtype != STFB_BF16 || kh != kw || kh > 3) return 0;
  if (stride == 1) {
    if (2 * pad != kh - 1 || Hp != Hg || Wp != Wg) return 0;
  } else if (stride == 2) {
    if (!wgrad_strided_ok() || Hp != (Hg + 2 * pad - kh) / 2 + 1 || Wp != (Wg + 2 * pad - kw) / 2 + 1) return 0;
  } else {
    return 0;
  }
  if (Cg % 64 != 0 || wg_pick_bn(Cp) == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(P) % 16) || (reinterpret_cast<uintptr_t>(G) % 16)) return 0;
  if ((long long)N * Hp * Wp > 2000000000LL) return 0;
  return 1;
}

template <int BN, int STAGES>
static int launch_wg(const CUtensorMap& tG, const CUtensorMap& tG2, const CUtensorMap& tP, const WgTcArgs& a, dim3 grid,
                     cudaStream_t st) {
  constexpr int smem = wg_smem_bytes<BN, STAGES>();
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv2d_wgrad(tcgen05): cannot reserve %d bytes of shared memory", smem);
      cudaGetLastError();
      return STFB_ECUDA;
    }
    configured = true;
  }
  wgrad_tc_kernel<BN, STAGES><<<grid, WG_THREADS, smem, st>>>(tG, tG2, tP, a);
  return post_launch("conv2d_wgrad(tcgen05)");
}

// G (and optional G2, concatenated on the channel axis after G) are the gathered activations; P = dy.
struct WgPlan { int TW, TH, TN, tiles_w, tiles_h, n_patches, m_chunks, BN, m_tiles, n_tiles, pps, splits; };

static WgPlan wg_plan(int N, int H, int W, int Cp, int Ctot, int kh, int kw) {
  WgPlan p{};
  p.TW = pow2_floor(W); if (p.TW > 8) p.TW = 8;
  p.TH = pow2_floor(H); if (p.TH > WG_PIX / p.TW) p.TH = WG_PIX / p.TW;
  p.TN = WG_PIX / (p.TW * p.TH);
  p.tiles_w = (W + p.TW - 1) / p.TW;
  p.tiles_h = (H + p.TH - 1) / p.TH;
  p.n_patches = ((N + p.TN - 1) / p.TN) * p.tiles_h * p.tiles_w;
  p.m_chunks = kh * kw * Ctot / 64;
  p.BN = wg_pick_bn(Cp);
  p.m_tiles = (p.m_chunks + 1) / 2;
  p.n_tiles = Cp / p.BN;
  // split the pixel reduction so the grid covers ~2 waves, keeping >= 4 patches per CTA
  long long want = (2LL * num_sms() + (long long)p.m_tiles * p.n_tiles - 1) / ((long long)p.m_tiles * p.n_tiles);
  long long maxsplit = (p.n_patches + 3) / 4;
  long long splits = want < 1 ? 1 : want;
  if (splits > maxsplit) splits = maxsplit;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.pps = (int)((p.n_patches + splits - 1) / splits);
  p.splits = (p.n_patches + p.pps - 1) / p.pps;
  return p;
}

size_t wgrad_tcgen05_workspace_bytes(int N, int H, int W, int Cp, int Cg, int kh, int kw) {
  const WgPlan p = wg_plan(N, H, W, Cp, Cg, kh, kw);
  return (size_t)p.splits * p.m_chunks * 64 * Cp * sizeof(float);
}

int wgrad_tcgen05(const void* P, const void* G, const void* G2, float* dW, int N, int H, int W, int Cp, int Hg, int Wg,
                  int C1, int C2, int cg_off, int cg_total, int kh, int kw, int stride, int pad, float* ws, size_t ws_bytes,
                  cudaStream_t st) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) { set_error("conv2d_wgrad(tcgen05): cuTensorMapEncodeTiled not available"); return STFB_ECUDA; }
  if ((long long)N * H * W == 0) return STFB_OK;
  const WgPlan pl = wg_plan(N, H, W, Cp, C1 + C2, kh, kw);
  const size_t need = (size_t)pl.splits * pl.m_chunks * 64 * Cp * sizeof(float);
  if (ws == nullptr || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) % 16) != 0) {
    set_error("conv2d_wgrad(tcgen05): workspace of %zu bytes (16-byte aligned) required, got %zu", need, ws_bytes);
    return STFB_EINVAL;
  }
  WgTcArgs a{};
  a.ws = ws; a.N = N; a.H = H; a.W = W; a.Cp = Cp; a.C1 = C1; a.C2 = C2; a.cg_off = cg_off; a.cg_total = cg_total;
  a.kh = kh; a.kw = kw; a.pad = pad; a.g_scale = stride;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("STFB_WG_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.debug = dbg;
  }
  a.TW = pl.TW; a.TH = pl.TH; a.TN = pl.TN; a.tiles_w = pl.tiles_w; a.tiles_h = pl.tiles_h; a.n_patches = pl.n_patches;
  a.m_chunks = pl.m_chunks; a.patches_per_split = pl.pps;
  const int BN = pl.BN, m_tiles = pl.m_tiles, n_tiles = pl.n_tiles;
  const long long splits = pl.splits;

  CUtensorMap tG, tG2, tP;
  if (!encode_nhwc_map_strided(enc, &tG, G, N, Hg, Wg, C1, a.TW, a.TH, a.TN, stride, 64)) { set_error("conv2d_wgrad(tcgen05): tensor map (G) failed"); return STFB_ECUDA; }
  tG2 = tG;
  if (C2 > 0 && !encode_nhwc_map_strided(enc, &tG2, G2, N, Hg, Wg, C2, a.TW, a.TH, a.TN, stride, 64)) { set_error("conv2d_wgrad(tcgen05): tensor map (G2) failed"); return STFB_ECUDA; }
  if (!encode_nhwc_map(enc, &tP, P, N, H, W, Cp, a.TW, a.TH, a.TN)) { set_error("conv2d_wgrad(tcgen05): tensor map (P) failed"); return STFB_ECUDA; }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
  int rc = STFB_ENOTSUP;
  switch (BN) {
    case 256: rc = launch_wg<256, 4>(tG, tG2, tP, a, grid, st); break;
    case 128: rc = launch_wg<128, 5>(tG, tG2, tP, a, grid, st); break;
    case 64: rc = launch_wg<64, 6>(tG, tG2, tP, a, grid, st); break;
    default: set_error("conv2d_wgrad(tcgen05): no tile for Cp=%d", Cp); return STFB_ENOTSUP;
  }
  if (rc != STFB_OK) return rc;
  const int Kg = pl.m_chunks * 64;
  const long long total = (long long)Kg * Cp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(ws, dW, (int)splits, Kg, Cp, C1 + C2, kh * kw, cg_off, cg_total);
  return post_launch("conv2d_wgrad(reduce)");
}

}  // namespace stfb
