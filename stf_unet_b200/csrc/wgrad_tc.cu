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
//   * The pixel reduction is split across CTAs (gridDim.z, one wave); each CTA accumulates its pixel range in TMEM and
//     adds its fp32 tile into an L2-resident accumulation buffer [(tap,ci)][co] with 16-byte vector reductions
//     (red.global.add.v4.f32: a 128-byte row per thread); a small second kernel transposes that buffer into dW in the
//     reference layout.  (v1 issued scalar red.add straight into [co][ci][ky][kx]: 10 M scattered sector updates per
//     launch -- profiles/r01_ncu_wgrad_atomics.txt; v2 wrote per-split partial tiles: 38 MB out + 38 MB back per launch.)
#include "tc_common.cuh"
#include <stdlib.h>
#include <mutex>

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
  float* partial;           // halo kernel: per-split partial tiles [split][pair][576][64] fp32 (NULL: red.add into ws)
  int debug;                // STFB_WG_DEBUG: 1 = no MMAs (TMA pipeline only), 2 = no TMA (MMA issue only); timing experiments
  int split;                // STFB_BF16X3: P / G hold three planes [hi | mid | lo] of Cp / C1 channels; a CTA walks its patch
                            // range six times, once per plane pair (wg_plane_g / wg_plane_p)
};

// pass v over the patch range (0..5): gathered-operand plane lo, hi, mid, mid, hi, hi against per-pixel plane hi, lo, mid, hi,
// mid, hi -- the five correction terms first, the hi*hi chain last (the TMEM accumulator rounds toward zero once per MMA: the
// fewer roundings happen at full magnitude the better, see csrc/split.cu)
__device__ __forceinline__ int wg_plane_g(int v) { return (int)((0x001102u >> (4 * v)) & 0xFu); }
__device__ __forceinline__ int wg_plane_p(int v) { return (int)((0x010120u >> (4 * v)) & 0xFu); }

int wgrad_tcgen05_supported(int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg, int kh, int kw, int stride, int pad,
                            int dtype, const void* P, const void* G);
int wgrad_tcgen05(const void* P, const void* G, const void* G2, float* dW, int N, int H, int W, int Cp, int Hg, int Wg,
                  int C1, int C2, int cg_off, int cg_total, int kh, int kw, int stride, int pad, float* ws, size_t ws_bytes,
                  cudaStream_t st, int split = 0);

constexpr int WG_PIX = 64;                       // K per stage
constexpr int WG_BLK_BYTES = WG_PIX * 128;       // one 64-ch x 64-pixel column block = 8 KB
constexpr int WG_THREADS = 192;

// NA accumulators (M = 128 each = two 64-channel column blocks of the gathered operand) share ONE dy tile per stage:
// the kernel is bound by the L2 -> shared-memory fill rate (~12 TB/s chip-wide, profiles/), so every byte of dy that is
// fetched should feed as many (tap, ci) rows as TMEM can hold (NA * BN <= 512 columns).
template <int BN, int NA, int STAGES>
constexpr int wg_smem_bytes() {
  return STAGES * (2 * NA + BN / 64) * WG_BLK_BYTES + 256 + 1024;
}

template <int BN, int NA, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG,
                                                                  const __grid_constant__ CUtensorMap tmG2,
                                                                  const __grid_constant__ CUtensorMap tmP, const WgTcArgs a) {
  constexpr int NB = BN / 64;
  constexpr int STAGE_BYTES = (2 * NA + NB) * WG_BLK_BYTES;
  constexpr int TMEM_COLS = (NA * BN <= 32) ? 32 : (NA * BN <= 64) ? 64 : (NA * BN <= 128) ? 128 : (NA * BN <= 256) ? 256 : 512;
  static_assert(NA * BN <= 512, "accumulators exceed tensor memory");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.y * BN;
  const int p_beg = blockIdx.z * a.patches_per_split;
  const int p_end = min(a.n_patches, p_beg + a.patches_per_split);
  const int Cin = a.C1 + a.C2;
  const int cpt = Cin / 64;                       // column blocks per tap
  // this CTA owns column blocks [chunk_lo, chunk_hi) of the (tap, ci) axis: up to 2*NA of them
  const int chunk_lo = blockIdx.x * 2 * NA;
  const int chunk_hi = min(a.m_chunks, chunk_lo + 2 * NA);
  const int nchunks = chunk_hi - chunk_lo;
  const int nacc = (nchunks + 1) / 2;             // accumulators in use

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmG);
    prefetch_tensormap(&tmP);
    if (a.C2 > 0) prefetch_tensormap(&tmG2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_real = p_end - p_beg;
  const int n_iter = a.split ? 6 * n_real : n_real;

  if (warp == 0) {
    if (n_iter > 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // an odd block count leaves the upper half of the last accumulator unused: it re-reads the previous block
      // (its rows are never written out), so every stage moves the same number of bytes
      const int nload = 2 * nacc;
      const uint32_t bytes = (uint32_t)(nload + NB) * WG_BLK_BYTES;
      // per-block constants of this CTA (tap offsets, channel coordinate, which operand): decoded once, not per patch
      int jcc[2 * NA], jdw[2 * NA], jdh[2 * NA];
#pragma unroll
      for (int j = 0; j < 2 * NA; ++j) {
        const int chunk = min(chunk_lo + j, chunk_hi - 1);
        const int tap = chunk / cpt, r = tap / a.kw;
        jcc[j] = (chunk - tap * cpt) * 64;
        jdw[j] = tap - r * a.kw - a.pad;
        jdh[j] = r - a.pad;
      }
      const int visits = a.split ? 6 : 1;
      for (int v = 0; v < visits; ++v) {
      const int gpl = a.split ? wg_plane_g(v) : 0, ppl = a.split ? wg_plane_p(v) * a.Cp : 0;
      // patch coordinates by counting: the div/mod chain per patch sat on the producer's issue path
      int wb = p_beg % a.tiles_w, hb = (p_beg / a.tiles_w) % a.tiles_h, nb = p_beg / (a.tiles_w * a.tiles_h);
      for (int p = p_beg; p < p_end; ++p) {
        const int w0 = wb * a.TW, h0 = hb * a.TH, i0 = nb * a.TN;
        if (++wb == a.tiles_w) { wb = 0; if (++hb == a.tiles_h) { hb = 0; ++nb; } }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        if (a.debug == 2) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[stage], bytes);
#pragma unroll
        for (int j = 0; j < 2 * NA; ++j) {
          if (j >= nload) break;
          const int cc = jcc[j];
          const int gw = w0 * a.g_scale + jdw[j], gh = h0 * a.g_scale + jdh[j];
          if (cc < a.C1) tma_load_4d(sa + j * WG_BLK_BYTES, &tmG, &full_bar[stage], gpl * a.C1 + cc, gw, gh, i0);
          else tma_load_4d(sa + j * WG_BLK_BYTES, &tmG2, &full_bar[stage], gpl * a.C2 + cc - a.C1, gw, gh, i0);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i)
          tma_load_4d(sa + (2 * NA + i) * WG_BLK_BYTES, &tmP, &full_bar[stage], ppl + n0 + i * 64, w0, h0, i0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_mnmajor(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t bdesc = make_mnmajor_sw128_desc(sa + 2 * NA * WG_BLK_BYTES, WG_BLK_BYTES);
        if (a.debug != 1) {
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k) {
            // 16 pixels = 16 rows of 128 B = 2048 B further down each column block
            for (int j = 0; j < nacc; ++j) {
              const uint64_t adesc = make_mnmajor_sw128_desc(sa + (uint32_t)(2 * j) * WG_BLK_BYTES, WG_BLK_BYTES);
              umma_bf16(tmem_base + (uint32_t)(j * BN), adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc,
                        (it | k) != 0);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (it == n_iter - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (n_iter > 0) {
    // epilogue: accumulator j, row m = (column block chunk_lo + 2j + m/64, channel m % 64) = row kg of the [Kg][Cp] slice
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    for (int j = 0; j < nacc; ++j) {
      const int chunk = chunk_lo + 2 * j + (m >> 6);
      const bool valid = chunk < chunk_hi;
      // accumulation buffer is [(tap, ci over the FULL weight)][co]: this launch owns the ci window [cg_off, cg_off + Cin)
      const int tapq = chunk / cpt;
      const long long kg = (long long)tapq * a.cg_total + a.cg_off + (chunk - tapq * cpt) * 64 + (m & 63);
      float* rowp = a.ws + kg * a.Cp + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(taddr + c0, r);
        tmem_ld_wait();
        if (valid && a.debug != 3) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + c0 + e), "f"(__uint_as_float(r[e])),
                         "f"(__uint_as_float(r[e + 1])), "f"(__uint_as_float(r[e + 2])), "f"(__uint_as_float(r[e + 3]))
                         : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Halo variant for 3x3 / stride-1 / pad-1 weight gradients.  The kernel above fetches the gathered operand once per
// filter tap (nine shifted 8 KB boxes per 64 pixels x 64 channels) and is bound by the L2 -> SM fill rate (ncu: 6-7 TB/s
// of xbar traffic, tensor pipe 17-41 % active).  Here a CTA owns ONE (64 input channels, 64 output channels) pair and all
// nine taps of it: per 8 x 8 pixel patch it loads one (8+2) x (8+2) halo box of x (12.5 KB) and the 8 x 8 box of dy
// (8 KB); tap (r, s) is the same shared-memory block read through an MN-major descriptor that starts (r*10 + s) rows
// further down with 1280 B between 8-pixel row groups.  An M = 128 accumulator holds two taps (its two 64-channel
// column blocks are `LBO` = the distance between the two tap offsets apart), so nine taps fill 4.5 accumulators of
// 64 TMEM columns each.  The pixel reduction is split over gridDim.z as before.
constexpr int WH_PATCH = 8;
constexpr int WH_PITCH = (WH_PATCH + 2) * 128;                         // bytes between patch rows inside the halo block
constexpr int WH_X_TX = (WH_PATCH + 2) * (WH_PATCH + 2) * 128;          // 12800 B per TMA box
constexpr int WH_X_BYTES = 13312;                                      // rounded up to the 1 KB swizzle period
constexpr int WH_STAGE_BYTES = WH_X_BYTES + WG_BLK_BYTES;              // + one 64-pixel x 64-channel block of dy
constexpr int WH_STAGES = 6;
constexpr int WH_NACC = 5;
constexpr int wh_smem_bytes() { return WH_STAGES * WH_STAGE_BYTES + 256 + 1024; }

__device__ __forceinline__ uint64_t make_mnmajor_halo_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(WH_PITCH >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr int wh_tap_off(int tap) { return (tap / 3) * WH_PITCH + (tap % 3) * 128; }

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmG,
                                                                    const __grid_constant__ CUtensorMap tmG2,
                                                                    const __grid_constant__ CUtensorMap tmP, const WgTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WH_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WH_STAGES;
  uint64_t* accum_bar = empty_bar + WH_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int cblk = blockIdx.x;                    // 64-channel block of the gathered operand (ci axis of dW)
  const int n0 = blockIdx.y * 64;                 // output-channel block
  const int p_beg = blockIdx.z * a.patches_per_split;
  const int p_end = min(a.n_patches, p_beg + a.patches_per_split);
  const int n_real = p_end - p_beg;
  const int n_iter = a.split ? 6 * n_real : n_real;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WH_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmG);
    prefetch_tensormap(&tmP);
    if (a.C2 > 0) prefetch_tensormap(&tmG2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (n_iter > 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int cc = cblk * 64;
      const int visits = a.split ? 6 : 1;
      for (int v = 0; v < visits; ++v) {
      const int gpl = a.split ? wg_plane_g(v) : 0, ppl = a.split ? wg_plane_p(v) * a.Cp : 0;
      // patch coordinates by counting: the div/mod chain per patch sat on the producer's issue path
      int wb = p_beg % a.tiles_w, hb = (p_beg / a.tiles_w) % a.tiles_h, nb_ = p_beg / (a.tiles_w * a.tiles_h);
      for (int p = p_beg; p < p_end; ++p) {
        const int w0 = wb * WH_PATCH, h0 = hb * WH_PATCH, nb = nb_;
        if (++wb == a.tiles_w) { wb = 0; if (++hb == a.tiles_h) { hb = 0; ++nb_; } }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sx = smem + stage * WH_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], WH_X_TX + WG_BLK_BYTES);
        if (cc < a.C1) tma_load_4d(sx, &tmG, &full_bar[stage], gpl * a.C1 + cc, w0 - 1, h0 - 1, nb);
        else tma_load_4d(sx, &tmG2, &full_bar[stage], gpl * a.C2 + cc - a.C1, w0 - 1, h0 - 1, nb);
        tma_load_4d(sx + WH_X_BYTES, &tmP, &full_bar[stage], ppl + n0, w0, h0, nb);
        if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
      }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_mnmajor(128, 64);
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t s0 = smem_u32(smem);
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sx = s0 + stage * WH_STAGE_BYTES;
        const uint64_t bdesc = make_mnmajor_sw128_desc(sx + WH_X_BYTES, WG_BLK_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // 16 pixels = two patch rows per MMA
#pragma unroll
          for (int j = 0; j < WH_NACC; ++j) {
            // accumulator j: taps 2j (rows 0..63) and 2j+1 (rows 64..127); the last one repeats tap 8 (never written out)
            constexpr int dummy = 0;
            const int t0 = 2 * j, t1 = (2 * j + 1 < 9) ? 2 * j + 1 : 2 * j;
            const uint32_t lbo = (uint32_t)(wh_tap_off(t1) - wh_tap_off(t0)) + dummy;
            const uint64_t adesc = make_mnmajor_halo_desc(sx + wh_tap_off(t0) + k * 2 * WH_PITCH, lbo);
            umma_bf16(tmem_base + (uint32_t)(j * 64), adesc, bdesc + (uint64_t)(128 * k), idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (it == n_iter - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (n_iter > 0) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int j = 0; j < WH_NACC; ++j) {
      const int tap = 2 * j + (m >> 6);
      const bool valid = tap < 9;
      const long long kg = (long long)tap * a.cg_total + a.cg_off + cblk * 64 + (m & 63);
      float* rowp = a.ws + kg * a.Cp + n0;
      // with a scratch buffer every CTA stores its tile as a plain partial (summed by wgrad_halo_reduce_kernel): the
      // red.add epilogue (41 k fp32 atomics per CTA) took longer than the 55-patch main loop
      float* prow = a.partial ? a.partial + ((((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 576 +
                                             tap * 64 + (m & 63)) * 64
                              : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 64);
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(taddr + c0, r);
        tmem_ld_wait();
        if (valid && a.debug != 3) {
          if (prow) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<uint4*>(prow + c0 + e) = make_uint4(r[e], r[e + 1], r[e + 2], r[e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + c0 + e), "f"(__uint_as_float(r[e])),
                           "f"(__uint_as_float(r[e + 1])), "f"(__uint_as_float(r[e + 2])), "f"(__uint_as_float(r[e + 3]))
                           : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// acc[(tap, cg_off + c*64 + i)][n*64 + j] += sum over splits of partial[split][n][c][tap*64 + i][j]
// grid (36, pairs, zchunks): one thread = 4 consecutive output channels of one (tap, ci) row over ONE chunk of the splits
// (a 64 -> 64 layer has a single pair and 148 splits: without the z axis 36 CTAs would walk all of them); the partials
// are L2 resident, the chunk sums meet in acc through 16-byte red.adds
__global__ void __launch_bounds__(256) wgrad_halo_reduce_kernel(const float* __restrict__ partial, float* __restrict__ acc,
                                                               int splits, int cblocks, int nblocks, int Cp, int cg_off,
                                                               int cg_total) {
  const int pair = blockIdx.y;                       // = n * cblocks + c  (blockIdx.y * gridDim.x + blockIdx.x of the main grid)
  const int c = pair % cblocks, n = pair / cblocks;
  const int e = blockIdx.x * 256 + threadIdx.x;      // float4 index inside the 576 x 64 tile
  const int row = e >> 4, j = (e & 15) * 4;
  const int pairs = cblocks * nblocks;
  const int per = (splits + gridDim.z - 1) / gridDim.z;
  const int sp0 = blockIdx.z * per, sp1 = min(splits, sp0 + per);
  if (sp0 >= sp1) return;
  const float4* src = reinterpret_cast<const float4*>(partial + ((long long)pair * 576 + row) * 64 + j);
  const long long stride4 = (long long)pairs * 576 * 64 / 4;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
  int sp = sp0;
  for (; sp + 3 < sp1; sp += 4) {
    const float4 u = src[(long long)sp * stride4], v = src[(long long)(sp + 1) * stride4];
    const float4 w = src[(long long)(sp + 2) * stride4], z = src[(long long)(sp + 3) * stride4];
    s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    s2.x += w.x; s2.y += w.y; s2.z += w.z; s2.w += w.w;
    s3.x += z.x; s3.y += z.y; s3.z += z.z; s3.w += z.w;
  }
  for (; sp < sp1; ++sp) {
    const float4 u = src[(long long)sp * stride4];
    s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
  }
  const int tap = row >> 6, i = row & 63;
  float* dst = acc + ((long long)tap * cg_total + cg_off + c * 64 + i) * Cp + n * 64 + j;
  const float ox = (s0.x + s1.x) + (s2.x + s3.x), oy = (s0.y + s1.y) + (s2.y + s3.y);
  const float oz = (s0.z + s1.z) + (s2.z + s3.z), ow = (s0.w + s1.w) + (s2.w + s3.w);
  if (gridDim.z == 1) {
    float4 o = *reinterpret_cast<float4*>(dst);
    o.x += ox; o.y += oy; o.z += oz; o.w += ow;
    *reinterpret_cast<float4*>(dst) = o;
  } else {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(ox), "f"(oy), "f"(oz), "f"(ow) : "memory");
  }
}

// caller-owned scratch for the per-split partial tiles (registered once per process: one process drives one GPU).
// The buffer is cut into SLOTS of wgrad_scratch_bytes() each and every stream that launches a halo wgrad gets its own slot
// (first come, first served): the partials of a launch are read back by the reduce kernel that follows it on the SAME
// stream, so launches on different streams -- the engine rotates weight gradients over several side streams, and they
// become parallel branches of a captured graph -- must not share partial tiles.  A stream that finds no free slot falls
// back to the red.add epilogue (slower, still correct).
static float* g_wg_scratch = nullptr;
static size_t g_wg_scratch_bytes = 0;
constexpr int WG_MAX_SLOTS = 16;
static cudaStream_t g_wg_slot_stream[WG_MAX_SLOTS];
static int g_wg_slots_used = 0;
static std::mutex g_wg_mu;
void wgrad_set_scratch(void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_wg_mu);
  g_wg_scratch = reinterpret_cast<float*>(p);
  g_wg_scratch_bytes = bytes;
  g_wg_slots_used = 0;
}
size_t wgrad_scratch_bytes() { return (size_t)num_sms() * 576 * 64 * sizeof(float); }
// the slot of stream `st` (NULL: none registered / all taken / a slot cannot hold `need` bytes)
static float* wgrad_scratch_slot(cudaStream_t st, size_t need) {
  std::lock_guard<std::mutex> lk(g_wg_mu);
  const size_t per = wgrad_scratch_bytes();
  if (g_wg_scratch == nullptr || need > per || (reinterpret_cast<uintptr_t>(g_wg_scratch) % 16) != 0) return nullptr;
  int nslots = (int)(g_wg_scratch_bytes / per);
  if (nslots > WG_MAX_SLOTS) nslots = WG_MAX_SLOTS;
  for (int i = 0; i < g_wg_slots_used; ++i)
    if (g_wg_slot_stream[i] == st) return g_wg_scratch + (size_t)i * (per / sizeof(float));
  if (g_wg_slots_used >= nslots) return nullptr;
  g_wg_slot_stream[g_wg_slots_used] = st;
  return g_wg_scratch + (size_t)(g_wg_slots_used++) * (per / sizeof(float));
}

// dW[co][cg_off + ci][tap] += acc[(tap, ci)][co]: 32 ci x 32 co tiles transposed through shared memory so that both the
// reads (along co) and the writes (along (ci, tap), contiguous in the reference layout) are coalesced
__global__ void __launch_bounds__(256) wgrad_scatter_kernel(const float* __restrict__ acc, float* __restrict__ dW, int Cp, int Cg,
                                                           int khw, int cg_off, int cg_total) {
  __shared__ float tile[9][32 * 41 + 1];   // ci stride 41 = 9 (mod 32), tap stride 1313 = 1 (mod 32): the (ci, tap) walk of
                                            // the write phase maps element e to bank (e + r) % 32 -- conflict free
  Cg = cg_total;   // the accumulation buffer spans the full ci axis
  (void)cg_off;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;       // 32 x 8
  for (int tap = 0; tap < khw; ++tap)
    for (int r = ty; r < 32; r += 8) {
      const int ci = ci0 + r, co = co0 + tx;
      tile[tap][r * 41 + tx] = (ci < Cg && co < Cp) ? acc[((long long)tap * Cg + ci) * Cp + co] : 0.f;
    }
  __syncthreads();
  // one co row at a time: 32 ci x khw taps = a contiguous run of 32*khw floats in dW
  const int run = 32 * khw;
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r;
    if (co >= Cp) continue;
    float* dst = dW + ((long long)co * cg_total + ci0) * khw;
    for (int e = tx; e < run; e += 32) {
      const int ci = e / khw, tap = e - ci * khw;
      if (ci0 + ci < Cg) dst[e] += tile[tap][ci * 41 + r];
    }
  }
}

// ---- 32-channel layers (decoder tail) through the 64-channel kernel: pixel pairing -----------------------------------
// [N,H,W,32] viewed as [N,H,W/2,64] puts two neighbouring pixels (parity a / b) in one 64-wide "channel" vector.  The
// 3x3 weight gradient of the paired tensors, acc[(r, kx'), (a, ci)][(b, co)], contains every product the 32-channel
// gradient needs: dW[(r,kx),ci][co] = sum over b of the entry whose gathered pixel 2px'+b+kx-1 falls in pair px'+kx'-1
// with parity a.  Half of acc is unused (2x the FLOPs, still ~30x faster than the FFMA kernel on these layers).
__global__ void wgrad_fold_pairs_kernel(const float* __restrict__ acc, float* __restrict__ dW, int Cp, int Cg) {
  const int total = Cp * Cg * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kx = i % 3, r = (i / 3) % 3;
    const int ci = (i / 9) % Cg, co = i / (9 * Cg);
    auto A = [&](int kxp, int a, int b) {
      return acc[((long long)((r * 3 + kxp) * 2 * Cg + a * Cg + ci)) * (2 * Cp) + b * Cp + co];
    };
    float v;
    if (kx == 1) v = A(1, 0, 0) + A(1, 1, 1);
    else if (kx == 2) v = A(1, 1, 0) + A(2, 0, 1);
    else v = A(0, 1, 0) + A(1, 0, 1);
    dW[i] += v;
  }
}

int wgrad_pairs_supported(int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg, int cg_off, int cg_total, int kh, int kw,
                          int stride, int pad, int dtype, const void* P, const void* G) {
  if (dtype != STFB_BF16 || kh != 3 || kw != 3 || stride != 1 || pad != 1) return 0;
  if (Hp != Hg || Wp != Wg || (Wp & 1)) return 0;
  if (Cp % 32 != 0 || Cg % 32 != 0 || (Cp % 64 == 0 && Cg % 64 == 0)) return 0;
  if (cg_off != 0 || cg_total != Cg) return 0;
  if (2 * Cp > 256) return 0;
  return wgrad_tcgen05_supported(N, Hp, Wp / 2, 2 * Cp, Hg, Wg / 2, 2 * Cg, 3, 3, 1, 1, dtype, P, G);
}

size_t wgrad_pairs_workspace_bytes(int Cp, int Cg) { return (size_t)9 * 2 * Cg * 2 * Cp * sizeof(float); }

int wgrad_pairs(const void* P, const void* G, float* dW, int N, int H, int W, int Cp, int Cg, float* ws, size_t ws_bytes,
                cudaStream_t st) {
  const size_t need = wgrad_pairs_workspace_bytes(Cp, Cg);
  if (ws == nullptr || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) % 16) != 0) {
    set_error("conv2d_wgrad(pairs): workspace of %zu bytes (16-byte aligned) required, got %zu", need, ws_bytes);
    return STFB_EINVAL;
  }
  cudaMemsetAsync(ws, 0, need, st);
  int rc = wgrad_tcgen05(P, G, nullptr, nullptr, N, H, W / 2, 2 * Cp, H, W / 2, 2 * Cg, 0, 0, 2 * Cg, 3, 3, 1, 1, ws, need, st);
  if (rc != STFB_OK) return rc;
  const int total = Cp * Cg * 9;
  wgrad_fold_pairs_kernel<<<(total + 255) / 256, 256, 0, st>>>(ws, dW, Cp, Cg);
  return post_launch("conv2d_wgrad(fold pairs)");
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
  if ((dtype != STFB_BF16 && dtype != STFB_BF16X3) || kh != kw || kh > 3) return 0;
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

template <int BN, int NA, int STAGES>
static int launch_wg(const CUtensorMap& tG, const CUtensorMap& tG2, const CUtensorMap& tP, const WgTcArgs& a, dim3 grid,
                     cudaStream_t st) {
  constexpr int smem = wg_smem_bytes<BN, NA, STAGES>();
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel<BN, NA, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv2d_wgrad(tcgen05): cannot reserve %d bytes of shared memory", smem);
      cudaGetLastError();
      return STFB_ECUDA;
    }
    configured = true;
  }
  wgrad_tc_kernel<BN, NA, STAGES><<<grid, WG_THREADS, smem, st>>>(tG, tG2, tP, a);
  return post_launch("conv2d_wgrad(tcgen05)");
}

// G (and optional G2, concatenated on the channel axis after G) are the gathered activations; P = dy.
struct WgPlan { int TW, TH, TN, tiles_w, tiles_h, n_patches, m_chunks, BN, NA, m_tiles, n_tiles, pps, splits; };

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
  // accumulators per CTA: as many as TMEM holds (BN=256: 2, 128: 3, 64: 5) but no more than the problem has
  const int na_max = p.BN == 256 ? 2 : (p.BN == 128 ? 3 : 5);
  const int na_need = (p.m_chunks + 1) / 2;
  p.NA = na_need >= na_max ? na_max : (na_need >= 2 ? 2 : 1);
  if (p.BN == 64 && p.NA == 2 && na_need > 2) p.NA = na_max;
  p.m_tiles = (p.m_chunks + 2 * p.NA - 1) / (2 * p.NA);
  p.n_tiles = Cp / p.BN;
  // ONE wave: m_tiles * n_tiles * splits <= #SMs (a 2.03-wave grid costs a whole third round), >= 2 patches per CTA
  long long tiles = (long long)p.m_tiles * p.n_tiles;
  long long splits = num_sms() / tiles;
  long long maxsplit = (p.n_patches + 1) / 2;
  if (splits > maxsplit) splits = maxsplit;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.pps = (int)((p.n_patches + splits - 1) / splits);
  p.splits = (p.n_patches + p.pps - 1) / p.pps;
  return p;
}

size_t wgrad_tcgen05_workspace_bytes(int N, int H, int W, int Cp, int cg_total, int kh, int kw) {
  return (size_t)kh * kw * cg_total * Cp * sizeof(float);     // fp32 accumulation buffer [(tap, ci)][co] of the full weight
}

// all deferred weight-gradient transposes of a backward pass in one launch:
//   grad_flat[off + (co*Cg + ci)*khw + tap] += acc_flat[off + (tap*Cg + ci)*Cp + co]
// one CTA = one 32 (ci) x 32 (co) tile of one weight, all taps, transposed through shared memory (coalesced both ways)
__global__ void __launch_bounds__(256) wgrad_scatter_batched_kernel(const stfb_scatter_job* __restrict__ jobs, int njobs,
                                                                   const float* __restrict__ acc, float* __restrict__ grad) {
  __shared__ float tile[9][32 * 41 + 1];   // ci stride 41 = 9 (mod 32), tap stride 1313 = 1 (mod 32): the (ci, tap) walk of
                                            // the write phase maps element e to bank (e + r) % 32 -- conflict free
  __shared__ stfb_scatter_job jb;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;
    const long long b = blockIdx.x;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].start <= b) lo = mid; else hi = mid - 1;
    }
    jb = jobs[lo];
  }
  __syncthreads();
  const int Cp = jb.Cp, Cg = jb.Cg, khw = jb.khw;
  const int local = (int)(blockIdx.x - jb.start);
  const int tiles_ci = (Cg + 31) / 32;
  const int ci0 = (local % tiles_ci) * 32, co0 = (local / tiles_ci) * 32;
  const float* a = acc + jb.off;
  float* g = grad + jb.off;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int tap = 0; tap < khw; ++tap)
    for (int r = ty; r < 32; r += 8) {
      const int ci = ci0 + r, co = co0 + tx;
      tile[tap][r * 41 + tx] = (ci < Cg && co < Cp) ? a[((long long)tap * Cg + ci) * Cp + co] : 0.f;
    }
  __syncthreads();
  const int run = 32 * khw;
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r;
    if (co >= Cp) continue;
    float* dst = g + ((long long)co * Cg + ci0) * khw;
    for (int e = tx; e < run; e += 32) {
      const int ci = e / khw, tap = e - ci * khw;
      if (ci0 + ci < Cg) dst[e] += tile[tap][ci * 41 + r];
    }
  }
}

int wgrad_scatter_batched(const stfb_scatter_job* jobs_dev, int njobs, long long total_tiles, const float* acc, float* grad,
                          cudaStream_t st) {
  wgrad_scatter_batched_kernel<<<(unsigned)total_tiles, 256, 0, st>>>(jobs_dev, njobs, acc, grad);
  return post_launch("wgrad_scatter_batched");
}

int wgrad_tcgen05(const void* P, const void* G, const void* G2, float* dW, int N, int H, int W, int Cp, int Hg, int Wg,
                  int C1, int C2, int cg_off, int cg_total, int kh, int kw, int stride, int pad, float* ws, size_t ws_bytes,
                  cudaStream_t st, int split) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) { set_error("conv2d_wgrad(tcgen05): cuTensorMapEncodeTiled not available"); return STFB_ECUDA; }
  if ((long long)N * H * W == 0) return STFB_OK;
  const WgPlan pl = wg_plan(N, H, W, Cp, C1 + C2, kh, kw);
  const size_t need = (size_t)kh * kw * cg_total * Cp * sizeof(float);
  if (ws == nullptr || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) % 16) != 0) {
    set_error("conv2d_wgrad(tcgen05): workspace of %zu bytes (16-byte aligned) required, got %zu", need, ws_bytes);
    return STFB_EINVAL;
  }
  WgTcArgs a{};
  a.ws = ws; a.N = N; a.H = H; a.W = W; a.Cp = Cp; a.C1 = C1; a.C2 = C2; a.cg_off = cg_off; a.cg_total = cg_total;
  a.kh = kh; a.kw = kw; a.pad = pad; a.g_scale = stride;
  a.split = split ? 1 : 0;
  const int pmul = split ? 3 : 1;       // bf16x3: three planes per tensor (the kernels walk their patch range six times)
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("STFB_WG_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.debug = dbg;
  }
  {
    static int no_halo = -1;
    if (no_halo < 0) { const char* e = getenv("STFB_NO_HALO"); no_halo = (e && atoi(e)) ? 1 : 0; }
    if (!no_halo && kh == 3 && kw == 3 && stride == 1 && pad == 1 && H == Hg && W == Wg && H >= 8 && W >= 8 && C1 % 64 == 0 &&
        C2 % 64 == 0 && Cp % 64 == 0) {
      a.TW = WH_PATCH; a.TH = WH_PATCH; a.TN = 1;
      a.tiles_w = (W + WH_PATCH - 1) / WH_PATCH;
      a.tiles_h = (H + WH_PATCH - 1) / WH_PATCH;
      a.n_patches = N * a.tiles_h * a.tiles_w;
      const int pairs = ((C1 + C2) / 64) * (Cp / 64);
      long long splits = num_sms() / pairs;
      const long long maxsplit = (a.n_patches + 1) / 2;
      if (splits > maxsplit) splits = maxsplit;
      if (splits < 1) splits = 1;
      if (splits > 65535) splits = 65535;
      a.patches_per_split = (int)((a.n_patches + splits - 1) / splits);
      splits = (a.n_patches + a.patches_per_split - 1) / a.patches_per_split;
      CUtensorMap tG, tG2, tP;
      if (!encode_nhwc_map(enc, &tG, G, N, Hg, Wg, pmul * C1, WH_PATCH + 2, WH_PATCH + 2, 1)) { set_error("conv2d_wgrad(halo): tensor map (G) failed"); return STFB_ECUDA; }
      tG2 = tG;
      if (C2 > 0 && !encode_nhwc_map(enc, &tG2, G2, N, Hg, Wg, pmul * C2, WH_PATCH + 2, WH_PATCH + 2, 1)) { set_error("conv2d_wgrad(halo): tensor map (G2) failed"); return STFB_ECUDA; }
      if (!encode_nhwc_map(enc, &tP, P, N, H, W, pmul * Cp, WH_PATCH, WH_PATCH, 1)) { set_error("conv2d_wgrad(halo): tensor map (P) failed"); return STFB_ECUDA; }
      static bool configured = false;
      if (!configured) {
        if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wh_smem_bytes()) != cudaSuccess) {
          set_error("conv2d_wgrad(halo): cannot reserve %d bytes of shared memory", wh_smem_bytes());
          cudaGetLastError();
          return STFB_ECUDA;
        }
        configured = true;
      }
      if (dW != nullptr) cudaMemsetAsync(ws, 0, need, st);
      dim3 hgrid((unsigned)((C1 + C2) / 64), (unsigned)(Cp / 64), (unsigned)splits);
      const size_t part_bytes = (size_t)pairs * splits * 576 * 64 * sizeof(float);
      float* slot = Cp % 4 == 0 ? wgrad_scratch_slot(st, part_bytes) : nullptr;
      const bool use_partials = slot != nullptr;
      a.partial = slot;
      wgrad_halo_kernel<<<hgrid, WG_THREADS, wh_smem_bytes(), st>>>(tG, tG2, tP, a);
      int rc = post_launch("conv2d_wgrad(tcgen05 halo)");
      if (rc == STFB_OK && use_partials) {
        // ~4 CTAs per SM: split the walk over the partials when there are few (ci, co) pairs
        int zc = (4 * num_sms() + 36 * pairs - 1) / (36 * pairs);
        if (zc > (int)splits) zc = (int)splits;
        if (zc > 16) zc = 16;
        if (zc < 1) zc = 1;
        dim3 rgrid(576 * 64 / 4 / 256, (unsigned)pairs, (unsigned)zc);
        wgrad_halo_reduce_kernel<<<rgrid, 256, 0, st>>>(slot, ws, (int)splits, (C1 + C2) / 64, Cp / 64, Cp, cg_off, cg_total);
        rc = post_launch("conv2d_wgrad(halo reduce)");
      }
      if (rc != STFB_OK || dW == nullptr) return rc;
      dim3 sgrid((unsigned)((cg_total + 31) / 32), (unsigned)((Cp + 31) / 32));
      wgrad_scatter_kernel<<<sgrid, 256, 0, st>>>(ws, dW, Cp, C1 + C2, kh * kw, cg_off, cg_total);
      return post_launch("conv2d_wgrad(scatter)");
    }
  }
  a.TW = pl.TW; a.TH = pl.TH; a.TN = pl.TN; a.tiles_w = pl.tiles_w; a.tiles_h = pl.tiles_h; a.n_patches = pl.n_patches;
  a.m_chunks = pl.m_chunks; a.patches_per_split = pl.pps;

  const int BN = pl.BN, m_tiles = pl.m_tiles, n_tiles = pl.n_tiles;
  const long long splits = pl.splits;

  CUtensorMap tG, tG2, tP;
  if (!encode_nhwc_map_strided(enc, &tG, G, N, Hg, Wg, pmul * C1, a.TW, a.TH, a.TN, stride, 64)) { set_error("conv2d_wgrad(tcgen05): tensor map (G) failed"); return STFB_ECUDA; }
  tG2 = tG;
  if (C2 > 0 && !encode_nhwc_map_strided(enc, &tG2, G2, N, Hg, Wg, pmul * C2, a.TW, a.TH, a.TN, stride, 64)) { set_error("conv2d_wgrad(tcgen05): tensor map (G2) failed"); return STFB_ECUDA; }
  if (!encode_nhwc_map(enc, &tP, P, N, H, W, pmul * Cp, a.TW, a.TH, a.TN)) { set_error("conv2d_wgrad(tcgen05): tensor map (P) failed"); return STFB_ECUDA; }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
  if (dW != nullptr) cudaMemsetAsync(ws, 0, need, st);   // immediate mode owns the buffer; deferred mode: the caller zeroed it
  int rc = STFB_ENOTSUP;
  const int key = BN * 10 + pl.NA;
  switch (key) {
    case 2562: rc = launch_wg<256, 2, 3>(tG, tG2, tP, a, grid, st); break;   // stage 64 KB
    case 2561: rc = launch_wg<256, 1, 4>(tG, tG2, tP, a, grid, st); break;   // stage 48 KB
    case 1283: rc = launch_wg<128, 3, 3>(tG, tG2, tP, a, grid, st); break;   // stage 64 KB
    case 1282: rc = launch_wg<128, 2, 4>(tG, tG2, tP, a, grid, st); break;   // stage 48 KB
    case 1281: rc = launch_wg<128, 1, 6>(tG, tG2, tP, a, grid, st); break;   // stage 32 KB
    case 645: rc = launch_wg<64, 5, 2>(tG, tG2, tP, a, grid, st); break;     // stage 88 KB
    case 642: rc = launch_wg<64, 2, 5>(tG, tG2, tP, a, grid, st); break;     // stage 40 KB
    case 641: rc = launch_wg<64, 1, 8>(tG, tG2, tP, a, grid, st); break;     // stage 24 KB
    default: set_error("conv2d_wgrad(tcgen05): no kernel for Cp=%d NA=%d", Cp, pl.NA); return STFB_ENOTSUP;
  }
  if (rc != STFB_OK) return rc;
  (void)splits;
  if (dW == nullptr) return STFB_OK;                    // deferred: stfb_wgrad_scatter_batched folds the buffer into dW later
  dim3 sgrid((unsigned)((cg_total + 31) / 32), (unsigned)((Cp + 31) / 32));
  wgrad_scatter_kernel<<<sgrid, 256, 0, st>>>(ws, dW, Cp, C1 + C2, kh * kw, cg_off, cg_total);
  return post_launch("conv2d_wgrad(scatter)");
}

}  // namespace stfb
