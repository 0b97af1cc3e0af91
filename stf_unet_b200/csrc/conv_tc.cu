// tcgen05 / TMEM implicit-GEMM convolution family (bf16 in, fp32 accumulate in tensor memory).
//
//   D[128 pixels x BN channels] (TMEM) += A[128 x 64] (smem, TMA) * B[BN x 64]^T (smem, TMA)   per k-block
//
// * GEMM rows are a TN x TH x TW patch of output pixels (TN*TH*TW = 128).  For filter tap (r, s) the A tile is
//   the SAME patch of the NHWC input shifted by (r - pad, s - pad): one 4-D tiled TMA load with box
//   {64 ch, TW, TH, TN}; out-of-image coordinates (the conv padding, ragged edges) are zero-filled by TMA.
//   The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, which is exactly the K-major
//   SWIZZLE_128B operand layout tcgen05.mma consumes -- no im2col buffer, no register staging.
// * B is the weight matrix packed [Cout][(ky,kx,ci)] (K-major), box {64, BN}.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> bias / folded BN / residual / ReLU -> bf16|fp32 NHWC stores).  smem ring of STAGES k-blocks
//   synchronised with mbarriers (TMA complete_tx -> MMA, tcgen05.commit -> producer / epilogue).
// * Concat inputs (decoder fusion, UNet skip) are two tensor maps walking one K loop: no materialised torch.cat.
//
// * Geometry is table driven (TcArgs): a launch is 1..4 "phases" (gridDim.z), each with its own list of filter taps
//   (A-load offset + weight K block).  That one mechanism covers
//     - stride-1 "same" convolutions and 1x1 GEMMs (1 phase, kh*kw taps),
//     - stride-2 forward convolutions (A map built with TMA elementStrides = 2: the box walks every other pixel),
//     - transposed gathers = ConvTranspose2d forward and Conv2d dgrad at stride 1 or 2: the s*s output parities are
//       separate phases with 1/2/2/4 dense taps each (no multiplication by inserted zeros), outputs scattered at
//       stride s.
#include "tc_common.cuh"
#include <stdlib.h>

namespace stfb {

// ------------------------------------------------------------------------------------------------
struct TcArgs {
  void* y;
  const void* residual;
  const float* bias;
  const float* bias2;
  const float* scale;
  const float* shift;
  int N;
  int Hout, Wout, Cout;  // output tensor
  int C1, C2;
  int a_scale;           // A-map coordinate = logical pixel * a_scale + tap offset (2 for stride-2 forward)
  int o_scale;           // output pixel = logical pixel * o_scale + phase offset   (2 for stride-2 transposed)
  int nphase_w;          // phase z -> (z / nphase_w, z % nphase_w) output parity
  int TW, TH, TN;        // logical pixel patch of one tile, TN*TH*TW == 128
  int tiles_w, tiles_h, tiles_n;
  int num_tiles;         // phases * tiles_n * tiles_h * tiles_w * (Cout / BN)
  int relu;
  // fused LSTM cell epilogue (EPI == 1): accumulator = [x_t, h_{t-1}] [W_ih | W_hh]^T with chunk-interleaved columns
  // (column j of a 256-wide tile: 16-unit chunk j / 64, gate (j % 64) / 16, unit u_base + 16 * chunk + j % 16)
  const float* c_prev;   // [rows][Chid] fp32 or NULL (t = 0)
  const float* c_cur;    // EPI == 2 (LSTM backward step): cell state of the step being differentiated, [rows][Chid] fp32
  float* c_out;          // [rows][Chid] fp32 (EPI == 2: the running d(cell) gradient, read and rewritten in place)
  void* acts;            // [rows][4*Chid] bf16 post-activation gates in accumulator column order (NULL in eval)
  int Chid;
  float* stat_partial;   // fused BatchNorm statistics: [stat_slots][2][stat_groups][Cout] fp32 (NULL = off)
  int stat_slots, stat_groups, stat_imgs;   // stat_imgs = images per group
  int w_resident;        // halo kernel: the whole weight matrix sits in the B ring (loaded once, never released)
  int split_c1, split_c2;  // STFB_BF16X3 (0 = off): logical channels of x / x2; C1 / C2 above are then the virtual 6x counts
  int debug;             // STFB_TC_DEBUG: 1 = no MMAs (TMA pipeline only), 2 = no TMA (MMA issue only); timing experiments
  int ntaps[4];
  signed char dh[4][9], dw[4][9], ktap[4][9];
};

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 192;
// BK = channels per k-block = one swizzle row: 64 bf16 (128 B, SWIZZLE_128B) or 32 bf16 (64 B, SWIZZLE_64B) for the
// 32-channel layers at the end of the decoder.

// Epilogue staging: every epilogue warp owns TC_STG_SEGS segments of 32 rows x 128 B in shared memory.  Accumulator rows
// live one per thread (TMEM lane = pixel), but a pixel's channels are contiguous in memory: results are written to the
// segment row by row (16-byte chunks XOR-swizzled by row, conflict free) and then moved to global memory with 8 lanes per
// 128-byte row segment, so a warp store instruction touches 4 full lines instead of 32 partial ones (and likewise for
// the residual / previous-cell-state loads).
constexpr int TC_SEG_BYTES = 32 * 128;
template <int EPI>
__host__ __device__ constexpr int tc_stg_segs() { return EPI == 1 ? 5 : (EPI == 2 ? 2 : 1); }   // LSTM: c x2, h, acts, (bias in the 5th); LSTM backward: dh x2
template <int BN, int STAGES, int BK, int EPI = 0>
constexpr int tc_smem_bytes() {
  return STAGES * (TC_BM * BK * 2 + BN * BK * 2) + 4 * tc_stg_segs<EPI>() * TC_SEG_BYTES + 256 + 1024;
}

// K-major descriptor for a [rows][BK] bf16 tile whose rows are BK*2 bytes with the matching swizzle
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
  if constexpr (BK == 64) return make_kmajor_sw128_desc(saddr);
  // SWIZZLE_64B: layout code 4, 8-row groups are 8 * 64 B = 512 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Channel coordinate of K position c (a multiple of the k-block) in the activation operand: returns true when it lies in the
// second concat operand (x2), `chan` = its channel there.  STFB_BF16X3: the K axis is six segments of all C1 + C2 logical
// channels; segment s reads plane {lo, hi, mid, mid, hi, hi}[s] of the three-plane tensors [hi | mid | lo] (csrc/split.cu packs
// the weight blocks hi, lo, mid, hi, mid, hi against them: correction terms first, the hi*hi chain last).
__device__ __forceinline__ bool a_coord(const TcArgs& a, int c, int& chan) {
  if (a.split_c1 == 0) {
    if (c < a.C1) { chan = c; return false; }
    chan = c - a.C1;
    return true;
  }
  const int cin = a.split_c1 + a.split_c2;
  const int seg = c / cin, cr = c - seg * cin;
  const int pl = (int)((0x001102u >> (4 * seg)) & 0xFu);
  if (cr < a.split_c1) { chan = pl * a.split_c1 + cr; return false; }
  chan = pl * a.split_c2 + cr - a.split_c1;
  return true;
}

struct TileCoord {
  int ph, ph_h, ph_w, wb, hb, nb, n0, num_kb;
};

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const TcArgs& a, int tile, int n_tiles, int cpt) {
  // order: N tile fastest (CTAs that run concurrently share the A patch in L2), then pixel tile, then phase
  TileCoord t;
  const int nt = tile % n_tiles;
  int r = tile / n_tiles;
  t.wb = r % a.tiles_w; r /= a.tiles_w;
  t.hb = r % a.tiles_h; r /= a.tiles_h;
  t.nb = r % a.tiles_n;
  t.ph = r / a.tiles_n;
  t.ph_h = t.ph / a.nphase_w;
  t.ph_w = t.ph - t.ph_h * a.nphase_w;
  t.n0 = nt * BN;
  t.num_kb = a.ntaps[t.ph] * cpt;
  return t;
}

// k-blocks of a tile without the full decode (the MMA issuer needs nothing else: five div/mod per tile sat on its issue path,
// and an LSTM step tile is only two k-blocks long)
__device__ __forceinline__ int tile_num_kb(const TcArgs& a, int tile, int n_tiles, int cpt) {
  if (a.nphase_w == 1) return a.ntaps[0] * cpt;
  return a.ntaps[tile / (n_tiles * a.tiles_w * a.tiles_h * a.tiles_n)] * cpt;
}

// Persistent: one CTA per SM walks tiles `blockIdx.x, +gridDim.x, ...`.  The TMA ring keeps streaming across tile
// boundaries and the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the
// main loop of tile i+1 and the per-tile prologue (barrier init, TMEM alloc, first TMA round trip) is paid once.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: one MUFU instead of EX2 + RCP (tanh.approx: 2^-11 relative, below bf16 resolution)
__device__ __forceinline__ float tanh_sigmoid(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(0.5f * x));
  return fmaf(0.5f, y, 0.5f);
}
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- epilogue staging helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stg_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
// explicit shared-space accesses: the staging pointer comes from an integer round-up of the dynamic smem base, which
// hides the address space from the compiler (generic ST.E / LD.E to shared memory are several times slower than STS / LDS)
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  sts128(addr, make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d)));
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  const uint4 v = lds128(addr);
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 t;
  t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  t.z = *reinterpret_cast<uint32_t*>(&c); t.w = *reinterpret_cast<uint32_t*>(&d);
  return t;
}
__device__ __forceinline__ void unpack8_bf16(const uint4& t, float* f) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f[2 * i] = p.x;
    f[2 * i + 1] = p.y;
  }
}
__device__ __forceinline__ float4 ldg_add4(const float* p, const float* q) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(q));
  return make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}
// global -> staging segment: row r of the segment is the 128 bytes at base + cpix[r] * pitch (8 lanes x 16 B per row)
__device__ __forceinline__ void seg_load(uint32_t seg, const uint8_t* base, const int* cpix, unsigned cmask, long long pitch, int lane) {
  const int crow = lane >> 3, cch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if ((cmask >> i) & 1u)
      sts128(seg + stg_off(i * 4 + crow, cch), *reinterpret_cast<const uint4*>(base + (long long)cpix[i] * pitch + cch * 16));
  __syncwarp();
}
// staging segment -> global; the leading barrier orders the row-per-thread writes before the cooperative reads, the
// trailing one lets the segment be rewritten
__device__ __forceinline__ void seg_store(uint32_t seg, uint8_t* base, const int* cpix, unsigned cmask, long long pitch, int lane) {
  const int crow = lane >> 3, cch = lane & 7;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if ((cmask >> i) & 1u)
      *reinterpret_cast<uint4*>(base + (long long)cpix[i] * pitch + cch * 16) = lds128(seg + stg_off(i * 4 + crow, cch));
  __syncwarp();
}


// Generic epilogue of one output tile (bias / folded BN / residual / ReLU -> TO), shared by the streaming and the halo
// kernels.  `taddr` = this warp's TMEM lane quadrant + accumulator buffer, `stg` = its staging segment (shared space).
template <int BN, typename TO>
__device__ __forceinline__ void epilogue_tile(const TcArgs& a, const TileCoord& t, uint32_t taddr, uint32_t stg, const int* cpix,
                                              unsigned cmask, bool valid, int on, int oh, int ow, uint64_t* tempty, int lane,
                                              uint32_t tempty_cluster = 0) {
  if constexpr (BN * (int)sizeof(TO) >= 128) {
    // ===== staged epilogue: one 128-byte row segment (64 bf16 / 32 fp32 channels) at a time
    constexpr int ESZ = (int)sizeof(TO);
    constexpr int SEGC = 128 / ESZ;
    constexpr int NSEG = BN / SEGC;
    const long long pitch = (long long)a.Cout * ESZ;
    uint8_t* ybase = reinterpret_cast<uint8_t*>(a.y) + (long long)t.n0 * ESZ;
    const uint8_t* rbase = a.residual ? reinterpret_cast<const uint8_t*>(a.residual) + (long long)t.n0 * ESZ : nullptr;
#pragma unroll 1
    for (int seg = 0; seg < NSEG; ++seg) {
      if (rbase) seg_load(stg, rbase + seg * 128, cpix, cmask, pitch, lane);
#pragma unroll
      for (int h = 0; h < SEGC / 32; ++h) {
        const int c0 = seg * SEGC + h * 32;
        uint32_t r[32];
        if (t.num_kb > 0) {
          tmem_ld_x32(taddr + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (c0 + 32 >= BN) {         // last chunk is in registers: hand the accumulator buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (tempty) mbar_arrive(tempty); else if (tempty_cluster) mbar_arrive_cluster(tempty_cluster); }
        }
        float v[32];
        const int cg = t.n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + cg + j);
        }
        if (a.bias2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias2 + cg + j);
        }
        if (a.scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], __ldg(a.scale + cg + j), __ldg(a.shift + cg + j));
        }
        if (rbase) {
          if constexpr (ESZ == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 rv = lds128(stg + stg_off(lane, h * 4 + k));
              float f[8];
              unpack8_bf16(rv, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[k * 8 + e] += f[e];
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 rv = lds128f(stg + stg_off(lane, k));
              v[k * 4] += rv.x; v[k * 4 + 1] += rv.y; v[k * 4 + 2] += rv.z; v[k * 4 + 3] += rv.w;
            }
          }
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if constexpr (ESZ == 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) sts128(stg + stg_off(lane, h * 4 + k), pack8_bf16(v + k * 8));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            sts128f(stg + stg_off(lane, k), v[k * 4], v[k * 4 + 1], v[k * 4 + 2], v[k * 4 + 3]);
        }
      }
      if constexpr (ESZ == 2) {
        if (a.stat_partial) {
          // train-mode BatchNorm statistics of this tile, from the bf16 values that are about to be stored: lane l owns
          // channels 2l, 2l+1 of the segment and walks the warp's 32 rows in shared memory (conflict free)
          __syncwarp();
          const unsigned vmask = __ballot_sync(0xffffffffu, valid);
          // all 32 loads first (independent), then four interleaved accumulation chains: the serial LDS -> unpack -> add
          // form left the epilogue warps waiting on shared-memory latency (profiles/r01_ncu_halo_conv_l1_instep.txt)
          uint32_t w[32];
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            w[r] = 0u;
            if ((vmask >> r) & 1u)
              asm volatile("ld.shared.b32 %0, [%1];"
                           : "=r"(w[r])
                           : "r"(stg + (uint32_t)(r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4)));
          }
          float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f}, qb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[r]));
            sa[r & 3] += f.x; sb[r & 3] += f.y;
            qa[r & 3] = fmaf(f.x, f.x, qa[r & 3]); qb[r & 3] = fmaf(f.y, f.y, qb[r & 3]);
          }
          const float s0 = (sa[0] + sa[1]) + (sa[2] + sa[3]), s1 = (sb[0] + sb[1]) + (sb[2] + sb[3]);
          const float q0 = (qa[0] + qa[1]) + (qa[2] + qa[3]), q1 = (qb[0] + qb[1]) + (qb[2] + qb[3]);
          const int g = (t.nb * a.TN) / a.stat_imgs;
          float* dst = a.stat_partial + ((long long)(blockIdx.x % a.stat_slots) * 2 * a.stat_groups + g) * a.Cout + t.n0 +
                       seg * SEGC + 2 * lane;
          const long long qoff = (long long)a.stat_groups * a.Cout;
          if (vmask) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(s0) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + 1), "f"(s1) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + qoff), "f"(q0) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + qoff + 1), "f"(q1) : "memory");
          }
        }
      }
      seg_store(stg, ybase + seg * 128, cpix, cmask, pitch, lane);
    }
  } else {
    // ===== direct epilogue (rows narrower than 128 B: the 32-channel decoder tail)
    const long long row = (((long long)on * a.Hout + oh) * a.Wout + ow) * a.Cout + t.n0;
    TO* __restrict__ yp = reinterpret_cast<TO*>(a.y) + row;
    const TO* rp = a.residual ? reinterpret_cast<const TO*>(a.residual) + row : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (t.num_kb > 0) {
        tmem_ld_x32(taddr + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (c0 + 32 >= BN) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (tempty) mbar_arrive(tempty); else if (tempty_cluster) mbar_arrive_cluster(tempty_cluster); }
      }
      if (valid && a.debug != 3) {
        float v[32];
        const int cg = t.n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + cg + j);
        }
        if (a.bias2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias2 + cg + j);
        }
        if (a.scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], __ldg(a.scale + cg + j), __ldg(a.shift + cg + j));
        }
        if (rp) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            f8 tt = ld8(rp + c0 + j);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j + e] += tt.v[e];
          }
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) st8(yp + c0 + j, v + j);
      }
    }
  }
}

template <int BN, int STAGES, typename TO, int BK, int EPI, int MINB = 1>
__global__ void __launch_bounds__(TC_THREADS, MINB) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmA2,
                                                                 const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  constexpr int TC_BK = BK;
  constexpr int TC_A_BYTES = TC_BM * BK * 2;
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();       // the next kernel of the stream may start its prologue while this one runs (it waits before reading)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // layout: [STAGES x (A, B)] [4 epilogue warps x staging segments] [mbarriers, TMEM slot]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 4 * tc_stg_segs<EPI>() * TC_SEG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready  (MMA -> epilogue)
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int Cin = a.C1 + a.C2;
  const int cpt = Cin / TC_BK;                 // k-blocks per filter tap
  const int n_tiles = a.Cout / BN;
  const int num_tiles = a.num_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (a.C2 > 0) prefetch_tensormap(&tmA2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // prologue done (barriers, TMEM, tensor maps): only now anything the previous kernel wrote may be touched

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<BN>(a, tile, n_tiles, cpt);
        const int wbase = t.wb * a.TW * a.a_scale, hbase = t.hb * a.TH * a.a_scale, i0 = t.nb * a.TN;
        // k-blocks walk taps outermost; split-precision operands channel blocks outermost (the hi*hi segment comes last)
        const int ntp = a.ntaps[t.ph];
        int tp = 0, chunk = 0;
        for (int kb = 0; kb < t.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int c = chunk * TC_BK;
          const int tp_now = tp;
          if (a.split_c1) { if (++tp == ntp) { tp = 0; ++chunk; } }
          else if (++chunk == cpt) { chunk = 0; ++tp; }
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + TC_A_BYTES;
          if (a.debug == 2) {
            mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int w0 = wbase + a.dw[t.ph][tp_now], h0 = hbase + a.dh[t.ph][tp_now];
          int ach;
          if (!a_coord(a, c, ach)) tma_load_4d(sa, &tmA, &full_bar[stage], ach, w0, h0, i0);
          else tma_load_4d(sa, &tmA2, &full_bar[stage], ach, w0, h0, i0);
          tma_load_2d(sb, &tmB, &full_bar[stage], (int)a.ktap[t.ph][tp_now] * Cin + c, t.n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      struct { int num_kb; } t{tile_num_kb(a, tile, n_tiles, cpt)};
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);       // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < t.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_kmajor_desc<BK>(sa);
          const uint64_t bdesc = make_kmajor_desc<BK>(sa + TC_A_BYTES);
          if (a.debug != 1) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
              umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);                    // frees the smem slot when these MMAs have read it
          if (kb == t.num_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (t.num_kb == 0 && lane == 0) mbar_arrive(&tfull_bar[acc]);   // a parity no tap reaches: nothing to wait for
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =================
    constexpr int STG_SEGS = tc_stg_segs<EPI>();
    const int q = warp & 3;
    const uint32_t stg = smem_u32(smem) + STAGES * STAGE_BYTES + q * (STG_SEGS * TC_SEG_BYTES);
    const int m = q * 32 + lane;                       // accumulator row = TMEM lane = pixel in the patch
    const int tw = m % a.TW, th = (m / a.TW) % a.TH, tn = m / (a.TW * a.TH);
    // cooperative mapping of the staged moves: instruction i handles rows i*4 + lane/8, 16-byte chunk lane%8
    int cpk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mm = q * 32 + i * 4 + (lane >> 3);
      cpk[i] = (mm % a.TW) | (((mm / a.TW) % a.TH) << 8) | ((mm / (a.TW * a.TH)) << 16);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile<BN>(a, tile, n_tiles, cpt);
      const int ow = (t.wb * a.TW + tw) * a.o_scale + t.ph_w, oh = (t.hb * a.TH + th) * a.o_scale + t.ph_h;
      const int on = t.nb * a.TN + tn;
      const bool valid = ow < a.Wout && oh < a.Hout && on < a.N;
      // pixel index of the rows this lane moves cooperatively (N*Ho*Wo < 2^31 is checked on the host)
      int cpix[8];
      unsigned cmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cw = (t.wb * a.TW + (cpk[i] & 0xFF)) * a.o_scale + t.ph_w;
        const int chh = (t.hb * a.TH + ((cpk[i] >> 8) & 0xFF)) * a.o_scale + t.ph_h;
        const int cn = t.nb * a.TN + (cpk[i] >> 16);
        const bool ok = cw < a.Wout && chh < a.Hout && cn < a.N;
        cpix[i] = ok ? (cn * a.Hout + chh) * a.Wout + cw : 0;
        cmask |= (ok ? 1u : 0u) << i;
      }
      if (a.debug == 3) cmask = 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (EPI != 1) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      }
      if constexpr (EPI == 1) {
        // ===== fused LSTM cell.  Tile = 64 hidden units [u_base, u_base + 64); TMEM column 64*k + 16*gate + e holds
        // gate (i,f,g,o) of unit u_base + 16*k + e.  c (fp32) is staged 32 units at a time, h (bf16) once per tile, the
        // saved activations (bf16, stored in accumulator column order) once per 16-unit chunk.
        static_assert(EPI != 1 || BN == 256, "LSTM epilogue needs the 256-column gate tile");
        const int C = a.Chid;
        const int u_base = t.n0 / 4;
        const uint32_t seg_c0 = stg, seg_h = stg + 2 * TC_SEG_BYTES, seg_a = stg + 3 * TC_SEG_BYTES;
        const uint32_t bias_s = stg + 4 * TC_SEG_BYTES;                  // 256 floats: b_ih + b_hh in accumulator column order
        const long long pitch_c = (long long)C * 4, pitch_h = (long long)C * 2, pitch_a = (long long)C * 8;
        const uint8_t* cprev_base = a.c_prev ? reinterpret_cast<const uint8_t*>(a.c_prev) + (long long)u_base * 4 : nullptr;
        uint8_t* cout_base = reinterpret_cast<uint8_t*>(a.c_out) + (long long)u_base * 4;
        uint8_t* h_base = reinterpret_cast<uint8_t*>(a.y) + (long long)u_base * 2;
        uint8_t* acts_base = a.acts ? reinterpret_cast<uint8_t*>(a.acts) + (long long)t.n0 * 2 : nullptr;
        // both halves of c_prev and the tile's biases are requested before the first accumulator column is touched
        if (cprev_base) {
          seg_load(seg_c0, cprev_base, cpix, cmask, pitch_c, lane);
          seg_load(seg_c0 + TC_SEG_BYTES, cprev_base + 128, cpix, cmask, pitch_c, lane);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = j * 32 + lane;                                  // column (k, gate, e) -> bias row gate*C + unit
          const int uu = u_base + (col >> 6) * 16 + (col & 15), gate = (col >> 4) & 3;
          const float b = __ldg(a.bias + gate * C + uu) + __ldg(a.bias2 + gate * C + uu);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + (uint32_t)col * 4), "f"(b) : "memory");
        }
        __syncwarp();
        mbar_wait(&tfull_bar[acc], acc_phase);       // the loads above were in flight while the MMAs of this tile ran
        tc_fence_after();
#pragma unroll 1
        for (int pair = 0; pair < 2; ++pair) {
          const uint32_t seg_c = seg_c0 + pair * TC_SEG_BYTES;
#pragma unroll 1
          for (int kk = 0; kk < 2; ++kk) {
            const int k = pair * 2 + kk;
            uint32_t ri[16], rf[16], rg[16], ro[16];
            if (t.num_kb > 0) {
              tmem_ld_x16(taddr + k * 64, ri);
              tmem_ld_x16(taddr + k * 64 + 16, rf);
              tmem_ld_x16(taddr + k * 64 + 32, rg);
              tmem_ld_x16(taddr + k * 64 + 48, ro);
              tmem_ld_wait();
            }
            if (k == 3) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            float hv[16], cv[16], ai[16], af[16], ag[16], ao[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
              if (cprev_base) cc = lds128f(seg_c + stg_off(lane, kk * 4 + (e >> 2)));
              const float pc[4] = {cc.x, cc.y, cc.z, cc.w};
              const float4 bi4 = lds128f(bias_s + (uint32_t)(k * 64 + e) * 4), bf4 = lds128f(bias_s + (uint32_t)(k * 64 + 16 + e) * 4);
              const float4 bg4 = lds128f(bias_s + (uint32_t)(k * 64 + 32 + e) * 4), bo4 = lds128f(bias_s + (uint32_t)(k * 64 + 48 + e) * 4);
              const float pbi[4] = {bi4.x, bi4.y, bi4.z, bi4.w}, pbf[4] = {bf4.x, bf4.y, bf4.z, bf4.w};
              const float pbg[4] = {bg4.x, bg4.y, bg4.z, bg4.w}, pbo[4] = {bo4.x, bo4.y, bo4.z, bo4.w};
#pragma unroll
              for (int z = 0; z < 4; ++z) {
                ai[e + z] = tanh_sigmoid(__uint_as_float(ri[e + z]) + pbi[z]);
                af[e + z] = tanh_sigmoid(__uint_as_float(rf[e + z]) + pbf[z]);
                ag[e + z] = fast_tanh(__uint_as_float(rg[e + z]) + pbg[z]);
                ao[e + z] = tanh_sigmoid(__uint_as_float(ro[e + z]) + pbo[z]);
                cv[e + z] = af[e + z] * pc[z] + ai[e + z] * ag[e + z];
                hv[e + z] = ao[e + z] * fast_tanh(cv[e + z]);
              }
              sts128f(seg_c + stg_off(lane, kk * 4 + (e >> 2)), cv[e], cv[e + 1], cv[e + 2], cv[e + 3]);
            }
            sts128(seg_h + stg_off(lane, k * 2), pack8_bf16(hv));
            sts128(seg_h + stg_off(lane, k * 2 + 1), pack8_bf16(hv + 8));
            if (acts_base) {
              sts128(seg_a + stg_off(lane, 0), pack8_bf16(ai));
              sts128(seg_a + stg_off(lane, 1), pack8_bf16(ai + 8));
              sts128(seg_a + stg_off(lane, 2), pack8_bf16(af));
              sts128(seg_a + stg_off(lane, 3), pack8_bf16(af + 8));
              sts128(seg_a + stg_off(lane, 4), pack8_bf16(ag));
              sts128(seg_a + stg_off(lane, 5), pack8_bf16(ag + 8));
              sts128(seg_a + stg_off(lane, 6), pack8_bf16(ao));
              sts128(seg_a + stg_off(lane, 7), pack8_bf16(ao + 8));
              seg_store(seg_a, acts_base + k * 128, cpix, cmask, pitch_a, lane);
            }
          }
          seg_store(seg_c, cout_base + pair * 128, cpix, cmask, pitch_c, lane);
        }
        seg_store(seg_h, h_base, cpix, cmask, pitch_h, lane);
      } else if constexpr (EPI == 2) {
        // ===== fused LSTM backward step.  The accumulator is dh_{t-1} = dG_t W_hh for hidden units [n0, n0 + BN); the
        // epilogue differentiates the cell of step t-1 right there (the arithmetic of lstm_cell_bwd_kernel, operation for
        // operation) and writes dG_{t-1} (bf16, gate-major [rows][4C]) and the running d(cell): dh never exists in memory
        // and the step is one launch instead of two.
        // Accumulator rows live one per lane, but the saved state is row-major in memory: 64 units of dh at a time go through
        // the warp's staging segments (two 32-unit halves, fp32) and are then processed in the elementwise kernel's mapping
        // -- 16 lanes x 4 units per row, two rows per instruction -- so every global access is a 256-byte (c, dc, dh) or
        // 128-byte (dG) contiguous run.  (The first version read row-per-lane: 32 lines per load instruction, and the step
        // got SLOWER than the two-launch route.)
        const int C = a.Chid;
        const long long row_l = valid ? ((long long)on * a.Hout + oh) * a.Wout + ow : -1;
        const uint32_t ridx_s = stg + 2 * TC_SEG_BYTES - 0;                 // (segments: [0] units 0..31, [1] units 32..63)
        (void)ridx_s;
#pragma unroll 1
        for (int c64 = 0; c64 < BN; c64 += 64) {
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            uint32_t r[32];
            tmem_ld_x32(taddr + c64 + hseg * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k)
              sts128(stg + hseg * TC_SEG_BYTES + stg_off(lane, k), make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]));
          }
          if (c64 + 64 >= BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          __syncwarp();
          const int u_base = t.n0 + c64;
#pragma unroll 2
          for (int it = 0; it < 16; ++it) {
            const int rl = it * 2 + (lane >> 4);                          // row of this warp's 32
            const int c = (lane & 15) * 4;                                // 4 units of the 64
            const long long row = __shfl_sync(0xffffffffu, row_l, rl);
            if (row < 0) continue;
            const float4 vdh4 = lds128f(stg + (c >> 5) * TC_SEG_BYTES + stg_off(rl, (c & 31) >> 2));
            const int u0 = u_base + c;
            const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(a.acts) + row * 4 * C + (u0 >> 4) * 64 + (u0 & 15);
            const f4 ai = ld4(ap), af = ld4(ap + 16), ag = ld4(ap + 32), ao = ld4(ap + 48);
            const f4 cc = ld4(a.c_cur + row * C + u0), vdc = ld4(a.c_out + row * C + u0);
            f4 cp{{0.f, 0.f, 0.f, 0.f}};
            if (a.c_prev) cp = ld4(a.c_prev + row * C + u0);
            const float vdh[4] = {vdh4.x, vdh4.y, vdh4.z, vdh4.w};
            f4 di, df, dg, dO, dcp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float tc = tanhf(cc.v[j]);
              dO.v[j] = vdh[j] * tc * ao.v[j] * (1.f - ao.v[j]);
              const float dct = vdc.v[j] + vdh[j] * ao.v[j] * (1.f - tc * tc);
              di.v[j] = dct * ag.v[j] * ai.v[j] * (1.f - ai.v[j]);
              df.v[j] = dct * cp.v[j] * af.v[j] * (1.f - af.v[j]);
              dg.v[j] = dct * ai.v[j] * (1.f - ag.v[j] * ag.v[j]);
              dcp.v[j] = dct * af.v[j];
            }
            __nv_bfloat16* gp = reinterpret_cast<__nv_bfloat16*>(a.y) + row * 4 * C + u0;
            st4(gp, di); st4(gp + C, df); st4(gp + 2 * C, dg); st4(gp + 3 * C, dO);
            st4(a.c_out + row * C + u0, dcp);
          }
          __syncwarp();
        }
      } else {
        epilogue_tile<BN, TO>(a, t, taddr, stg, cpix, cmask, valid, on, oh, ow, &tempty_bar[acc], lane);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
// Halo variant for 3x3 / stride-1 "same" convolutions (forward and dgrad): the streaming kernel above fetches the
// input patch once per filter tap (9 shifted 16 KB boxes per 64 channels) and is bound by the L2 -> SM fill rate
// (profiles/: 8.5 TB/s, tensor pipe 18 % on the 64-channel layer).  Here a tile is an 8 x 16 pixel patch of ONE image and
// the A operand of all nine taps is ONE (8+2) x (16+2) pixel box per 64 channels (23 KB instead of 144 KB): tap (r, s) is
// the same shared-memory block read through a descriptor that starts (r*10 + s) rows further down, with 1280 B (= one
// halo row of 10 pixels) between 8-pixel row groups instead of the dense 1024 B.  The 128-byte swizzle is a function of
// the shared-memory address bits, so TMA's write pattern and the shifted MMA reads agree.
// Weights stream through their own ring of (tap, 64-channel) blocks; when the whole matrix fits (64 -> 64 channels:
// 72 KB) it is loaded once per CTA and stays resident.
constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_PITCH = (HALO_TW + 2) * 128;                      // bytes between patch rows in the halo block
constexpr int HALO_TX_BYTES = (HALO_TH + 2) * (HALO_TW + 2) * 128;   // 23040 B per TMA box
constexpr int HALO_BLK_BYTES = 23552;                                // rounded up to the 1 KB swizzle period

template <int BN, int SA, int SB>
constexpr int halo_smem_bytes() {
  return SA * HALO_BLK_BYTES + SB * BN * 128 + 4 * TC_SEG_BYTES + 512 + 1024;
}

__device__ __forceinline__ uint64_t make_halo_a_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(HALO_PITCH >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// byte offset of filter tap tp = (dh + 1) * 3 + (dw + 1) inside the halo block (taps are issued in this canonical order;
// the host permutes the weight k-blocks to match: forward ktap = tp, dgrad ktap = 8 - tp)
__host__ __device__ constexpr uint32_t halo_tap_off(int tp) { return (uint32_t)((tp / 3) * HALO_PITCH + (tp % 3) * 128); }

// MT = pixel tiles per weight pass: with MT = 2 every weight block that reaches shared memory feeds two 128-pixel
// accumulators (half the L2 -> SM weight traffic per output, twice the MMAs per barrier round trip).  Tiles 2p and 2p+1
// of the same output-channel block form a "super tile"; TMEM holds MT * BN columns per buffer, double buffered when
// 2 * MT * BN <= 512.  SB divides 9, so the weight slot of a tap is a compile-time constant of the unrolled tap loop.
template <int BN, int MT, int SA, int SB, typename TO, int MINB = 1>
__global__ void __launch_bounds__(TC_THREADS, MINB) conv_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmA2,
                                                                   const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  static_assert(9 % SB == 0, "the weight ring must divide the nine taps");
  static_assert(SA % MT == 0, "the A ring holds whole super-tile channel blocks");
  constexpr int B_BYTES = BN * 128;
  constexpr int NBUF = (2 * MT * BN <= 512) ? 2 : 1;
  constexpr int TMEM_COLS = (NBUF * MT * BN < 32) ? 32 : NBUF * MT * BN;
  constexpr int WRAPS = 9 / SB;                       // ring wraps per channel block
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();       // the next kernel of the stream may start its prologue while this one runs (it waits before reading)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + SA * HALO_BLK_BYTES;
  const uint32_t sStg = sB + SB * B_BYTES;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + SA * HALO_BLK_BYTES + SB * B_BYTES + 4 * TC_SEG_BYTES);
  uint64_t* aempty = afull + SA;
  uint64_t* bfull = aempty + SA;
  uint64_t* bempty = bfull + SB;
  uint64_t* tfull_bar = bempty + SB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int Cin = a.C1 + a.C2;
  const int cpt = Cin / 64;
  const int n_tiles = a.Cout / BN;
  const int pix_tiles = a.num_tiles / n_tiles;
  const int n_super = ((pix_tiles + MT - 1) / MT) * n_tiles;       // super tile = n-tile fastest, then pixel-tile pairs

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (a.C2 > 0) prefetch_tensormap(&tmA2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // prologue done (barriers, TMEM, tensor maps): only now anything the previous kernel wrote may be touched

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int sa = 0;
      uint32_t pa = 0, nblk = 0;
      bool first = true;
      for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
        const int nt = st % n_tiles, pp = st / n_tiles;
        for (int c = 0; c < cpt; ++c, ++nblk) {
          const int cc = c * 64;
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const int ptile = pp * MT + m;
            mbar_wait(&aempty[sa], pa ^ 1);
            if (ptile < pix_tiles) {
              const TileCoord t = decode_tile<BN>(a, ptile * n_tiles + nt, n_tiles, cpt);
              mbar_arrive_expect_tx(&afull[sa], HALO_TX_BYTES);
              uint8_t* dst = smem + sa * HALO_BLK_BYTES;
              const int w0 = t.wb * HALO_TW - 1, h0 = t.hb * HALO_TH - 1;
              int ach;
              if (!a_coord(a, cc, ach)) tma_load_4d(dst, &tmA, &afull[sa], ach, w0, h0, t.nb);
              else tma_load_4d(dst, &tmA2, &afull[sa], ach, w0, h0, t.nb);
            } else {
              mbar_arrive(&afull[sa]);              // odd tail: the second accumulator is computed on stale data, never stored
            }
            if (++sa == SA) { sa = 0; pa ^= 1; }
          }
          if (first || !a.w_resident) {
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
              const int sb = tp % SB;
              mbar_wait(&bempty[sb], ((nblk * WRAPS + tp / SB) & 1) ^ 1);
              mbar_arrive_expect_tx(&bfull[sb], B_BYTES);
              tma_load_2d(smem + SA * HALO_BLK_BYTES + sb * B_BYTES, &tmB, &bfull[sb], (int)a.ktap[0][tp] * Cin + cc, nt * BN);
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
    int sa = 0, acc = 0;
    uint32_t pa = 0, acc_phase = 0, nblk = 0;
    bool first = true;
    for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * MT * BN);
      for (int c = 0; c < cpt; ++c, ++nblk) {
        uint32_t ablk[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          mbar_wait(&afull[sa], pa);
          ablk[m] = sA + sa * HALO_BLK_BYTES;
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        tc_fence_after();
        const bool streamed = first || !a.w_resident;
        if (elect_one()) {
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const int sb = tp % SB;
            if (streamed) {
              mbar_wait(&bfull[sb], (nblk * WRAPS + tp / SB) & 1);
              tc_fence_after();
            }
            const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb * B_BYTES);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              const uint64_t adesc = make_halo_a_desc(ablk[m] + halo_tap_off(tp));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_acc + (uint32_t)(m * BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                          (c | tp | k) != 0);
            }
            if (!a.w_resident) umma_commit(&bempty[sb]);
          }
          // the A blocks of this channel step are free once the MMAs above have read them
          {
            int s2 = sa;
#pragma unroll
            for (int m = MT - 1; m >= 0; --m) {
              s2 = (s2 == 0) ? SA - 1 : s2 - 1;
              umma_commit(&aempty[s2]);
            }
          }
          if (c == cpt - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
      first = false;
      if (NBUF == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      else acc_phase ^= 1;
    }
  } else {
    // ================= epilogue (same as the streaming kernel), one pixel tile of the super tile after the other ========
    const int q = warp & 3;
    const uint32_t stg = sStg + q * TC_SEG_BYTES;
    const int m_ = q * 32 + lane;
    const int tw = m_ % HALO_TW, th = m_ / HALO_TW;
    int cpk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mm = q * 32 + i * 4 + (lane >> 3);
      cpk[i] = (mm % HALO_TW) | ((mm / HALO_TW) << 8);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
      const int nt = st % n_tiles, pp = st / n_tiles;
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int ptile = pp * MT + m;
        const bool live = ptile < pix_tiles;
        const TileCoord t = decode_tile<BN>(a, (live ? ptile : pix_tiles - 1) * n_tiles + nt, n_tiles, cpt);
        const int ow = t.wb * HALO_TW + tw, oh = t.hb * HALO_TH + th, on = t.nb;
        const bool valid = live && ow < a.Wout && oh < a.Hout && on < a.N;
        int cpix[8];
        unsigned cmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int cw = t.wb * HALO_TW + (cpk[i] & 0xFF), chh = t.hb * HALO_TH + (cpk[i] >> 8);
          const bool ok = live && cw < a.Wout && chh < a.Hout && on < a.N;
          cpix[i] = ok ? (on * a.Hout + chh) * a.Wout + cw : 0;
          cmask |= (ok ? 1u : 0u) << i;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + m) * BN);
        if (m == 0) {          // after the coordinate arithmetic: it overlaps the main loop instead of delaying the drain
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
        }
        epilogue_tile<BN, TO>(a, t, taddr, stg, cpix, cmask, valid, on, oh, ow, m == MT - 1 ? &tempty_bar[acc] : nullptr, lane);
      }
      if (NBUF == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      else acc_phase ^= 1;
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
// CTA-pair variant of the halo kernel (tcgen05 cta_group::2): the two CTAs of a (2,1,1) cluster -- the two SMs of a TPC --
// compute two neighbouring 128-pixel tiles against the SAME output-channel block as ONE M = 256 MMA.  Each CTA stages its
// own halo block (A is split along M) and only HALF of every weight block (B is split along N: rows
// [rank * BN/2, +BN/2) of the tap's K-major tile), so a weight byte crosses the L2 -> SM fabric once per 256 pixels instead
// of once per 128 and the shared-memory reads per MMA drop from A + B to A + B/2 per SM: the 1-CTA kernel is bound by
// exactly these two (layers 2/3: ~310 KB of L2 reads per 4.6 k MMA cycles per SM, 1.6x the chip's L2 bandwidth; layer 1:
// 6 KB of operand reads per 32-cycle N = 64 MMA against 128 B/clk of shared-memory bandwidth).
//   * barriers: the "full" barriers live in the leader (rank 0): it arms them with the bytes of BOTH CTAs and both CTAs'
//     TMA loads complete_tx on them; the MMA warp of the leader frees ring slots and
//     publishes accumulators with multicast commits that arrive in both CTAs; the epilogue warps of both CTAs hand the
//     accumulator back by arriving on the leader's "tmem empty" barrier.
//   * 8 epilogue warps in two groups: group g drains accumulator buffer g (tiles alternate between the two TMEM buffers),
//     so every group has two MMA periods per tile -- the narrow layers were bound by their four epilogue warps.
constexpr int HALO2_THREADS = 320;
template <int BN, int SA, int SB>
constexpr int halo2_smem_bytes() {
  return SA * HALO_BLK_BYTES + SB * (BN / 2) * 128 + 8 * TC_SEG_BYTES + 512 + 1024;
}

template <int BN, int SA, int SB, typename TO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO2_THREADS, 1)
    conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  static_assert(9 % SB == 0, "the weight ring must divide the nine taps");
  constexpr int BH_BYTES = (BN / 2) * 128;            // this CTA's half of one (tap, 64-channel) weight block
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr int WRAPS = 9 / SB;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();       // the next kernel of the stream may start its prologue while this one runs (it waits before reading)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + SA * HALO_BLK_BYTES;
  const uint32_t sStg = sB + SB * BH_BYTES;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + SA * HALO_BLK_BYTES + SB * BH_BYTES + 8 * TC_SEG_BYTES);
  uint64_t* aempty = afull + SA;
  uint64_t* bfull = aempty + SA;
  uint64_t* bempty = bfull + SB;
  uint64_t* tfull_bar = bempty + SB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int Cin = a.C1 + a.C2;
  const int cpt = Cin / 64;
  const int n_tiles = a.Cout / BN;
  const int pix_tiles = a.num_tiles / n_tiles;                  // even (checked on the host)
  const int n_super = (pix_tiles / 2) * n_tiles;                // super tile = n-tile fastest, then pixel-tile pairs
  const int cid = (int)cluster_id_x(), ncl = (int)num_clusters_x();

  if (threadIdx.x == 0) {
    // full barriers (used in the leader only): ONE arrival, the leader's arrive.expect_tx with the bytes of both CTAs.  The
    // peer needs no arrival of its own: it cannot issue the loads of a slot's next use before the leader has consumed the
    // current one (its "empty" barrier is released by the same multicast commit), and bytes that land before the leader
    // has armed the barrier only drive the transaction count negative for a moment.
    for (int s = 0; s < SA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8); }   // 4 warps x 2 CTAs
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (a.C2 > 0) prefetch_tensormap(&tmA2);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // prologue done (barriers, TMEM, tensor maps): only now anything the previous kernel wrote may be touched

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own halo block, own half of the weights) =================
    if (elect_one()) {
      int sa = 0;
      uint32_t pa = 0, nblk = 0;
      bool first = true;
      for (int st = cid; st < n_super; st += ncl) {
        const int nt = st % n_tiles, pp = st / n_tiles;
        const TileCoord t = decode_tile<BN>(a, (pp * 2 + (int)rank) * n_tiles + nt, n_tiles, cpt);
        const int w0 = t.wb * HALO_TW - 1, h0 = t.hb * HALO_TH - 1;
        for (int c = 0; c < cpt; ++c, ++nblk) {
          const int cc = c * 64;
          mbar_wait(&aempty[sa], pa ^ 1);
          const uint32_t af = mapa_shared(smem_u32(&afull[sa]), 0);
          if (leader) mbar_arrive_expect_tx(&afull[sa], 2 * HALO_TX_BYTES);
          int ach;
          if (!a_coord(a, cc, ach)) tma_load_4d_pair(sA + sa * HALO_BLK_BYTES, &tmA, af, ach, w0, h0, t.nb);
          else tma_load_4d_pair(sA + sa * HALO_BLK_BYTES, &tmA2, af, ach, w0, h0, t.nb);
          if (++sa == SA) { sa = 0; pa ^= 1; }
          if (!first && a.w_resident) continue;        // the whole (half) weight matrix sits in the 9-slot ring since the first tile
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const int sb = tp % SB;
            mbar_wait(&bempty[sb], ((nblk * WRAPS + tp / SB) & 1) ^ 1);
            const uint32_t bf = mapa_shared(smem_u32(&bfull[sb]), 0);
            if (leader) mbar_arrive_expect_tx(&bfull[sb], 2 * BH_BYTES);
            tma_load_2d_pair(sB + sb * BH_BYTES, &tmB, bf, (int)a.ktap[0][tp] * Cin + cc, nt * BN + (int)rank * (BN / 2));
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader only) =================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      int sa = 0;
      uint32_t pa = 0, nblk = 0, it = 0;
      bool first = true;
      for (int st = cid; st < n_super; st += ncl, ++it) {
        const int acc = (int)(it & 1);
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
        for (int c = 0; c < cpt; ++c, ++nblk) {
          mbar_wait(&afull[sa], pa);
          const uint32_t ablk = sA + sa * HALO_BLK_BYTES;
          const int sa_used = sa;
          if (++sa == SA) { sa = 0; pa ^= 1; }
          tc_fence_after();
          const bool streamed = first || !a.w_resident;
          if (elect_one()) {
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
              const int sb = tp % SB;
              if (streamed) {
                mbar_wait(&bfull[sb], (nblk * WRAPS + tp / SB) & 1);
                tc_fence_after();
              }
              const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb * BH_BYTES);
              const uint64_t adesc = make_halo_a_desc(ablk + halo_tap_off(tp));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_pair(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (c | tp | k) != 0);
              if (!a.w_resident) umma_commit_pair(&bempty[sb]);
            }
            umma_commit_pair(&aempty[sa_used]);
            if (c == cpt - 1) umma_commit_pair(&tfull_bar[acc]);
          }
          __syncwarp();
        }
        first = false;
      }
    }
  } else {
    // ================= epilogue: group g = (warp - 2) / 4 drains accumulator buffer g =================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const uint32_t stg = sStg + (uint32_t)(g * 4 + q) * TC_SEG_BYTES;
    const int m_ = q * 32 + lane;
    const int tw = m_ % HALO_TW, th = m_ / HALO_TW;
    int cpk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mm = q * 32 + i * 4 + (lane >> 3);
      cpk[i] = (mm % HALO_TW) | ((mm / HALO_TW) << 8);
    }
    const uint32_t tempty_leader = mapa_shared(smem_u32(&tempty_bar[g]), 0);
    uint32_t it = 0;
    for (int st = cid; st < n_super; st += ncl, ++it) {
      if ((int)(it & 1) != g) continue;
      const int nt = st % n_tiles, pp = st / n_tiles;
      // tile coordinates and row addresses BEFORE the wait: the div/mod chain overlaps the main loop instead of delaying the drain
      const TileCoord t = decode_tile<BN>(a, (pp * 2 + (int)rank) * n_tiles + nt, n_tiles, cpt);
      const int ow = t.wb * HALO_TW + tw, oh = t.hb * HALO_TH + th, on = t.nb;
      const bool valid = ow < a.Wout && oh < a.Hout && on < a.N;
      int cpix[8];
      unsigned cmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cw = t.wb * HALO_TW + (cpk[i] & 0xFF), chh = t.hb * HALO_TH + (cpk[i] >> 8);
        const bool ok = cw < a.Wout && chh < a.Hout && on < a.N;
        cpix[i] = ok ? (on * a.Hout + chh) * a.Wout + cw : 0;
        cmask |= (ok ? 1u : 0u) << i;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN);
      mbar_wait(&tfull_bar[g], (it >> 1) & 1);
      tc_fence_after();
      epilogue_tile<BN, TO>(a, t, taddr, stg, cpix, cmask, valid, on, oh, ow, nullptr, lane, tempty_leader);
    }
  }
  // neither CTA may leave (or free its tensor memory) while the pair's MMAs can still touch its shared / tensor memory
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// The whole time loop of a per-pixel LSTM level in ONE kernel (hidden size 64: [W_ih | W_hh] = 64 KB of bf16 stays in shared
// memory).  Reference: nn.LSTM(C, C) over [B*h*w, T, C] (/root/reference/src/stf_lstm_unet.py:124-127, :216-221).
// A CTA owns a 128-pixel tile for t = 0 .. T-1:
//   * x_t tiles stream through a TMA ring; h_{t-1} never leaves the SM: the epilogue writes h_t (bf16) straight into a
//     shared-memory tile in the K-major SWIZZLE_128B layout (fence.proxy.async, mbarrier) and the next step's MMAs read it as
//     their A operand; c stays in the epilogue warps' staging segments (fp32) from step to step.
//   * two TMEM accumulators alternate by time step: the x_{t+1} W_ih^T half of step t+1 is issued while the epilogue of step t
//     still runs; only the four h_t W_hh^T MMAs wait for it.
//   * training writes c_t, h_t and the gate activations of every step (the backward pass needs them); inference writes h_T only.
// Same arithmetic, same accumulation order as conv_tc_kernel<EPI = 1> launched once per step: results are bit-identical.
struct LstmSeqArgs {
  const float* bias;        // b_ih [4C]
  const float* bias2;       // b_hh [4C]
  float* c_all;             // [T][rows][64] fp32 (training) or NULL
  void* h_all;              // training: [T][rows][64] bf16; inference: [rows][64] = h_T
  void* acts_all;           // [T][rows][256] bf16 in accumulator column order (training) or NULL
  long long rows;           // B * H * W
  int T, B, H, W;
  int TW, TH, TN, tiles_w, tiles_h, tiles_n, num_tiles;
  int keep;                 // 1: training (every step is stored)
};
constexpr int LSEQ_XS = 3;
constexpr int LSEQ_W_BYTES = 2 * 256 * 128;              // W_ih block + W_hh block, 256 gate rows x 64 k each
constexpr int LSEQ_X_BYTES = TC_BM * 128;
constexpr int LSEQ_WARP_STG = 2 * TC_SEG_BYTES + TC_SEG_BYTES + 1024;   // c (two segments), activations, biases
constexpr int lseq_smem_bytes() { return LSEQ_W_BYTES + LSEQ_XS * LSEQ_X_BYTES + LSEQ_X_BYTES + 4 * LSEQ_WARP_STG + 256 + 1024; }

__global__ void __launch_bounds__(TC_THREADS, 1) lstm_seq64_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                   const __grid_constant__ CUtensorMap tmW, const LstmSeqArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sW = smem_u32(smem);
  const uint32_t sX = sW + LSEQ_W_BYTES;
  const uint32_t sH = sX + LSEQ_XS * LSEQ_X_BYTES;
  const uint32_t sStg = sH + LSEQ_X_BYTES;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + LSEQ_W_BYTES + (LSEQ_XS + 1) * LSEQ_X_BYTES + 4 * LSEQ_WARP_STG);
  uint64_t* xfull = wfull + 1;
  uint64_t* xempty = xfull + LSEQ_XS;
  uint64_t* hfull = xempty + LSEQ_XS;
  uint64_t* tfull_bar = hfull + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int T = a.T;

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < LSEQ_XS; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 1); }
    mbar_init(hfull, 4);
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmW);
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
    // ================= TMA producer: the weights once, then one x_t tile per (tile, step) =================
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, LSEQ_W_BYTES);
      tma_load_2d(smem, &tmW, wfull, 0, 0);                              // W_ih: k in [0, 64)
      tma_load_2d(smem + 256 * 128, &tmW, wfull, 64, 0);                 // W_hh: k in [64, 128)
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        int r = tile;
        const int wb = r % a.tiles_w; r /= a.tiles_w;
        const int hb = r % a.tiles_h;
        const int nb = r / a.tiles_h;
        for (int t = 0; t < T; ++t) {
          mbar_wait(&xempty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&xfull[stage], LSEQ_X_BYTES);
          tma_load_4d(smem + LSEQ_W_BYTES + stage * LSEQ_X_BYTES, &tmX, &xfull[stage], 0, wb * a.TW, hb * a.TH, t * a.B + nb * a.TN);
          if (++stage == LSEQ_XS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, 256);
    mbar_wait(wfull, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0, s = 0, hp = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      for (int t = 0; t < T; ++t, ++s) {
        const int acc = (int)(s & 1);
        mbar_wait(&tempty_bar[acc], ((s >> 1) & 1) ^ 1);
        mbar_wait(&xfull[stage], phase);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 256);
        if (elect_one()) {
          const uint64_t adesc = make_kmajor_sw128_desc(sX + stage * LSEQ_X_BYTES), bdesc = make_kmajor_sw128_desc(sW);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0);
          umma_commit(&xempty[stage]);
          if (t == 0) umma_commit(&tfull_bar[acc]);          // zero initial state: no recurrent half
        }
        __syncwarp();
        if (++stage == LSEQ_XS) { stage = 0; phase ^= 1; }
        if (t > 0) {
          mbar_wait(hfull, hp & 1);                          // h_{t-1} of this tile is in shared memory
          ++hp;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_kmajor_sw128_desc(sH), bdesc = make_kmajor_sw128_desc(sW + 256 * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
            umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogue: gates -> cell update; c and h stay on the SM between steps =================
    const int q = warp & 3;
    const uint32_t stg = sStg + q * LSEQ_WARP_STG;
    const uint32_t seg_c0 = stg, seg_a = stg + 2 * TC_SEG_BYTES, bias_s = stg + 3 * TC_SEG_BYTES;
    const uint32_t seg_h = sH + q * TC_SEG_BYTES;            // rows q*32 .. q*32+31 of the next step's A operand
    const int m = q * 32 + lane;
    const int tw = m % a.TW, th = (m / a.TW) % a.TH, tn = m / (a.TW * a.TH);
    (void)tw; (void)th; (void)tn;
    int cpk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mm = q * 32 + i * 4 + (lane >> 3);
      cpk[i] = (mm % a.TW) | (((mm / a.TW) % a.TH) << 8) | ((mm / (a.TW * a.TH)) << 16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {                            // biases once per CTA (column (k, gate, e) -> bias row gate*64 + unit)
      const int col = j * 32 + lane;
      const int uu = (col >> 6) * 16 + (col & 15), gate = (col >> 4) & 3;
      const float b = __ldg(a.bias + gate * 64 + uu) + __ldg(a.bias2 + gate * 64 + uu);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + (uint32_t)col * 4), "f"(b) : "memory");
    }
    __syncwarp();
    const long long pitch_c = 64 * 4, pitch_h = 64 * 2, pitch_a = 256 * 2;
    uint32_t s = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      int r = tile;
      const int wb = r % a.tiles_w; r /= a.tiles_w;
      const int hb = r % a.tiles_h;
      const int nb = r / a.tiles_h;
      int cpix[8];
      unsigned cmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cw = wb * a.TW + (cpk[i] & 0xFF), chh = hb * a.TH + ((cpk[i] >> 8) & 0xFF), cn = nb * a.TN + (cpk[i] >> 16);
        const bool ok = cw < a.W && chh < a.H && cn < a.B;
        cpix[i] = ok ? (cn * a.H + chh) * a.W + cw : 0;
        cmask |= (ok ? 1u : 0u) << i;
      }
      for (int t = 0; t < T; ++t, ++s) {
        const int acc = (int)(s & 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
        const bool store_h = a.keep || t == T - 1;
        uint8_t* c_base = a.keep ? reinterpret_cast<uint8_t*>(a.c_all) + (long long)t * a.rows * pitch_c : nullptr;
        uint8_t* h_base = reinterpret_cast<uint8_t*>(a.h_all) + (a.keep ? (long long)t * a.rows * pitch_h : 0);
        uint8_t* acts_base = (a.keep && a.acts_all) ? reinterpret_cast<uint8_t*>(a.acts_all) + (long long)t * a.rows * pitch_a : nullptr;
        mbar_wait(&tfull_bar[acc], (s >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int pair = 0; pair < 2; ++pair) {
          const uint32_t seg_c = seg_c0 + pair * TC_SEG_BYTES;
#pragma unroll 1
          for (int kk = 0; kk < 2; ++kk) {
            const int k = pair * 2 + kk;
            uint32_t ri[16], rf[16], rg[16], ro[16];
            tmem_ld_x16(taddr + k * 64, ri);
            tmem_ld_x16(taddr + k * 64 + 16, rf);
            tmem_ld_x16(taddr + k * 64 + 32, rg);
            tmem_ld_x16(taddr + k * 64 + 48, ro);
            tmem_ld_wait();
            if (k == 3) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            float hv[16], cv[16], ai[16], af[16], ag[16], ao[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
              if (t > 0) cc = lds128f(seg_c + stg_off(lane, kk * 4 + (e >> 2)));      // c_{t-1}: left here by the previous step
              const float pc[4] = {cc.x, cc.y, cc.z, cc.w};
              const float4 bi4 = lds128f(bias_s + (uint32_t)(k * 64 + e) * 4), bf4 = lds128f(bias_s + (uint32_t)(k * 64 + 16 + e) * 4);
              const float4 bg4 = lds128f(bias_s + (uint32_t)(k * 64 + 32 + e) * 4), bo4 = lds128f(bias_s + (uint32_t)(k * 64 + 48 + e) * 4);
              const float pbi[4] = {bi4.x, bi4.y, bi4.z, bi4.w}, pbf[4] = {bf4.x, bf4.y, bf4.z, bf4.w};
              const float pbg[4] = {bg4.x, bg4.y, bg4.z, bg4.w}, pbo[4] = {bo4.x, bo4.y, bo4.z, bo4.w};
#pragma unroll
              for (int z = 0; z < 4; ++z) {
                ai[e + z] = tanh_sigmoid(__uint_as_float(ri[e + z]) + pbi[z]);
                af[e + z] = tanh_sigmoid(__uint_as_float(rf[e + z]) + pbf[z]);
                ag[e + z] = fast_tanh(__uint_as_float(rg[e + z]) + pbg[z]);
                ao[e + z] = tanh_sigmoid(__uint_as_float(ro[e + z]) + pbo[z]);
                cv[e + z] = af[e + z] * pc[z] + ai[e + z] * ag[e + z];
                hv[e + z] = ao[e + z] * fast_tanh(cv[e + z]);
              }
              sts128f(seg_c + stg_off(lane, kk * 4 + (e >> 2)), cv[e], cv[e + 1], cv[e + 2], cv[e + 3]);
            }
            sts128(seg_h + stg_off(lane, k * 2), pack8_bf16(hv));
            sts128(seg_h + stg_off(lane, k * 2 + 1), pack8_bf16(hv + 8));
            if (acts_base) {
              sts128(seg_a + stg_off(lane, 0), pack8_bf16(ai));
              sts128(seg_a + stg_off(lane, 1), pack8_bf16(ai + 8));
              sts128(seg_a + stg_off(lane, 2), pack8_bf16(af));
              sts128(seg_a + stg_off(lane, 3), pack8_bf16(af + 8));
              sts128(seg_a + stg_off(lane, 4), pack8_bf16(ag));
              sts128(seg_a + stg_off(lane, 5), pack8_bf16(ag + 8));
              sts128(seg_a + stg_off(lane, 6), pack8_bf16(ao));
              sts128(seg_a + stg_off(lane, 7), pack8_bf16(ao + 8));
              seg_store(seg_a, acts_base + k * 128, cpix, cmask, pitch_a, lane);
            }
          }
          if (c_base) seg_store(seg_c, c_base + pair * 128, cpix, cmask, pitch_c, lane);
        }
        // h_t is complete in shared memory: make it visible to the tensor core (async proxy) and release the next step
        fence_proxy_async();
        __syncwarp();
        if (t < T - 1 && lane == 0) mbar_arrive(hfull);
        if (store_h) seg_store(seg_h, h_base, cpix, cmask, pitch_h, lane);
        else __syncwarp();
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_tensormap_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

static void pick_patch(int H, int W, int& TW, int& TH, int& TN) {
  TW = pow2_floor(W); if (TW > 16) TW = 16;
  TH = pow2_floor(H); if (TH > TC_BM / TW) TH = TC_BM / TW;
  TN = TC_BM / (TW * TH);
}

static int pick_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  if (Cout % 32 == 0) return 32;
  return 0;
}

// Output-channel tile for THIS problem.  The widest tile that divides Cout has the best tensor-pipe efficiency, but small maps
// (UNet bottleneck at batch 4: 8 pixel tiles x 4 channel tiles of 256 = 32 CTAs on 148 SMs) leave most SMs idle: when the
// widest tile fills less than half of the last wave, take the width with the best (wave fill x relative tile rate) instead.
// Large problems (every layer of the headline workload) keep the widest tile.
static int pick_bn_for(const stfb_conv_params* p) {
  const int wide = pick_bn(p->Cout);
  static int occ = -1;
  if (occ < 0) { const char* e = getenv("STFB_BN_OCC"); occ = e ? atoi(e) : 1; }
  if (!occ || wide <= 64) return wide;
  const long long pix = p->mode == STFB_CONV_FWD ? (long long)p->N * p->Ho * p->Wo
                                                 : (long long)p->N * ((p->Ho + p->stride - 1) / p->stride) * ((p->Wo + p->stride - 1) / p->stride) * p->stride * p->stride;
  const long long pix_tiles = (pix + TC_BM - 1) / TC_BM;
  const int sms = num_sms();
  auto fill = [&](int bn) {
    const long long tiles = pix_tiles * (p->Cout / bn);
    return (double)tiles / (double)(((tiles + sms - 1) / sms) * sms);
  };
  if (fill(wide) >= 0.5) return wide;
  int best = wide;
  double best_score = 1.5 * fill(wide);          // a narrower tile has to promise a clear win (latency-bound small layers do not care)
  for (int bn = wide / 2; bn >= 64; bn /= 2) {
    if (p->Cout % bn != 0) continue;
    const double score = fill(bn) * (bn == 128 ? 0.9 : 0.6);
    if (score > best_score) { best_score = score; best = bn; }
  }
  return best;
}

static bool g_tc_strided_fwd = true;   // TMA elementStrides path (stride-2 forward)

int conv2d_tcgen05_supported(const stfb_conv_params* p) {
  if (p->x_dtype != STFB_BF16 && !(p->x_dtype == STFB_BF16X3 && p->y_dtype == STFB_F32)) return 0;
  if (p->kh != p->kw || p->kh > 3) return 0;
  if (p->stride != 1 && p->stride != 2) return 0;
  if (p->mode == STFB_CONV_FWD) {
    if (p->stride == 1 && (2 * p->pad != p->kh - 1 || p->Ho != p->H || p->Wo != p->W)) return 0;
    if (p->stride == 2 && !g_tc_strided_fwd) return 0;
    if (p->Ho != (p->H + 2 * p->pad - p->kh) / p->stride + 1 || p->Wo != (p->W + 2 * p->pad - p->kw) / p->stride + 1) return 0;
  } else if (p->mode != STFB_CONV_TRANSPOSED) {
    return 0;
  }
  if (p->C1 % 32 != 0 || p->C2 % 32 != 0) return 0;
  if (pick_bn(p->Cout) == 0) return 0;
  auto al = [](const void* q, int b) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % b) == 0; };
  if (!al(p->x, 16) || !al(p->x2, 16) || !al(p->y, 16) || !al(p->residual, 16)) return 0;
  if ((long long)p->N * p->Ho * p->Wo > 2000000000LL || (long long)p->N * p->H * p->W > 2000000000LL) return 0;
  return 1;
}

// 4-D NHWC bf16 activation map, box {64 ch, TW, TH, TN} pixels visited with traversal stride `estride` in W and H
bool encode_nhwc_map_strided(EncodeTiledFn enc, CUtensorMap* tm, const void* base, int N, int H, int W, int C, int TW, int TH,
                             int TN, int estride, int cblock) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)cblock, (cuuint32_t)(TW * estride), (cuuint32_t)(TH * estride), (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, cblock == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool encode_nhwc_map(EncodeTiledFn enc, CUtensorMap* tm, const void* base, int N, int H, int W, int C, int TW, int TH, int TN) {
  return encode_nhwc_map_strided(enc, tm, base, N, H, W, C, TW, TH, TN, 1, 64);
}

template <int BN, int STAGES, typename TO, int BK, int EPI = 0, int MINB = 1>
static int launch_tc(const CUtensorMap& tA, const CUtensorMap& tA2, const CUtensorMap& tB, const TcArgs& a, dim3 grid,
                     cudaStream_t st) {
  constexpr int smem = tc_smem_bytes<BN, STAGES, BK, EPI>();
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, TO, BK, EPI, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv2d(tcgen05): cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return STFB_ECUDA;
    }
    configured = true;
  }
  launch_ex<1>(conv_tc_kernel<BN, STAGES, TO, BK, EPI, MINB>, grid, dim3(TC_THREADS), smem, st, tA, tA2, tB, a);
  return post_launch("conv2d(tcgen05)");
}


template <int BN, int MT, int SA, int SB, typename TO, int MINB = 1>
static int launch_halo(const CUtensorMap& tA, const CUtensorMap& tA2, const CUtensorMap& tB, const TcArgs& a, dim3 grid,
                       cudaStream_t st) {
  constexpr int smem = halo_smem_bytes<BN, SA, SB>();
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_halo_kernel<BN, MT, SA, SB, TO, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv2d(tcgen05 halo): cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return STFB_ECUDA;
    }
    configured = true;
  }
  launch_ex<1>(conv_halo_kernel<BN, MT, SA, SB, TO, MINB>, grid, dim3(TC_THREADS), smem, st, tA, tA2, tB, a);
  return post_launch("conv2d(tcgen05 halo)");
}

template <int BN, int SA, int SB, typename TO>
static int launch_halo2(const CUtensorMap& tA, const CUtensorMap& tA2, const CUtensorMap& tB, const TcArgs& a, int n_super,
                        cudaStream_t st) {
  constexpr int smem = halo2_smem_bytes<BN, SA, SB>();
  static_assert(smem <= 232448, "shared memory budget");
  static int max_clusters = -1;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(conv_halo2_kernel<BN, SA, SB, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv2d(tcgen05 pair): cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return STFB_ECUDA;
    }
    // how many CTA pairs fit at once (a TPC with a disabled SM cannot host one)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)num_sms() / 2 * 2);
    cfg.blockDim = dim3(HALO2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute cattr[1];
    cattr[0].id = cudaLaunchAttributeClusterDimension;
    cattr[0].val.clusterDim.x = 2; cattr[0].val.clusterDim.y = 1; cattr[0].val.clusterDim.z = 1;
    cfg.attrs = cattr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, conv_halo2_kernel<BN, SA, SB, TO>, &cfg) != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      nc = num_sms() / 2;
    }
    max_clusters = nc < num_sms() / 2 ? nc : num_sms() / 2;
  }
  const int ncl = n_super < max_clusters ? n_super : max_clusters;
  launch_ex<1>(conv_halo2_kernel<BN, SA, SB, TO>, dim3((unsigned)(2 * ncl)), dim3(HALO2_THREADS), smem, st, tA, tA2, tB, a);
  return post_launch("conv2d(tcgen05 pair)");
}

static bool halo_ok(const stfb_conv_params* p, int BN);

// can this launch also reduce the BatchNorm statistics of its output? (bf16 rows of >= 128 B through the staged epilogue,
// plain forward conv without epilogue extras, every 128-pixel tile inside one image group)
int conv2d_stats_fusable(const stfb_conv_params* p, int groups) {
  if (!conv2d_tcgen05_supported(p)) return 0;
  if (p->mode != STFB_CONV_FWD || p->y_dtype != STFB_BF16) return 0;
  if (p->scale || p->residual || p->relu) return 0;
  if (groups <= 0 || p->N % groups != 0) return 0;
  const int BN = pick_bn(p->Cout);
  if (BN * 2 < 128) return 0;
  if (halo_ok(p, BN)) return 1;              // one image per tile
  int TW, TH, TN;
  pick_patch(p->Ho, p->Wo, TW, TH, TN);
  return ((p->N / groups) % TN == 0) ? 1 : 0;
}

static int g_tc_halo = -1;   // STFB_NO_HALO=1 keeps every 3x3 on the streaming kernel (A/B measurements)

// 3x3 / stride 1 / pad 1 (forward or dgrad), 64-channel multiples, maps at least 12 x 8: the halo kernel
static bool halo_ok(const stfb_conv_params* p, int BN) {
  if (g_tc_halo < 0) { const char* e = getenv("STFB_NO_HALO"); g_tc_halo = (e && atoi(e)) ? 0 : 1; }
  if (!g_tc_halo) return false;
  if (p->kh != 3 || p->kw != 3 || p->stride != 1 || p->pad != 1) return false;
  if (p->C1 % 64 != 0 || p->C2 % 64 != 0) return false;
  if (BN < 64) return false;
  if (p->Ho != p->H || p->Wo != p->W) return false;
  return p->Ho >= 12 && p->Wo >= 8;
}

int conv2d_tcgen05(const stfb_conv_params* p, cudaStream_t st) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) { set_error("conv2d(tcgen05): cuTensorMapEncodeTiled not available from the driver"); return STFB_ECUDA; }
  if ((long long)p->N * p->Ho * p->Wo == 0) return STFB_OK;
  if (reinterpret_cast<uintptr_t>(p->w) % 16 != 0) { set_error("conv2d(tcgen05): weights must be 16-byte aligned"); return STFB_EINVAL; }
  TcArgs a{};
  a.y = p->y; a.residual = p->residual; a.bias = p->bias; a.bias2 = p->bias2; a.scale = p->scale; a.shift = p->shift;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("STFB_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.debug = dbg;
  }
  // STFB_BF16X3: three planes per operand in memory, six segments on the K axis
  const bool split = p->x_dtype == STFB_BF16X3;
  const int kmul = split ? 6 : 1, pmul = split ? 3 : 1;
  a.N = p->N; a.Hout = p->Ho; a.Wout = p->Wo; a.Cout = p->Cout; a.C1 = kmul * p->C1; a.C2 = kmul * p->C2; a.relu = p->relu;
  a.split_c1 = split ? p->C1 : 0; a.split_c2 = split ? p->C2 : 0;
  if (p->stat_partial) {
    if (p->stat_slots <= 0 || !conv2d_stats_fusable(p, p->stat_groups)) {
      set_error("conv2d(tcgen05): fused BatchNorm statistics not available for this launch (stfb_conv2d_stats_fusable)");
      return STFB_ENOTSUP;
    }
    a.stat_partial = p->stat_partial; a.stat_slots = p->stat_slots; a.stat_groups = p->stat_groups;
    a.stat_imgs = p->N / p->stat_groups;
  }
  const int s = p->stride, k = p->kh;
  int Hl, Wl;   // logical pixel grid the tiles cover
  if (p->mode == STFB_CONV_FWD) {
    a.a_scale = s; a.o_scale = 1; a.nphase_w = 1;
    Hl = p->Ho; Wl = p->Wo;
    a.ntaps[0] = k * k;
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) {
        a.dh[0][r * k + c] = (signed char)(r - p->pad);
        a.dw[0][r * k + c] = (signed char)(c - p->pad);
        a.ktap[0][r * k + c] = (signed char)(r * k + c);
      }
  } else {
    // output pixel o = j*s + ph gathers input (o + pad - ky)/s = j + (ph + pad - ky)/s for the ky that divide
    a.a_scale = 1; a.o_scale = s; a.nphase_w = s;
    Hl = (p->Ho + s - 1) / s; Wl = (p->Wo + s - 1) / s;
    for (int phh = 0; phh < s; ++phh)
      for (int phw = 0; phw < s; ++phw) {
        const int z = phh * s + phw;
        int n = 0;
        for (int r = 0; r < k; ++r) {
          if ((phh + p->pad - r) % s != 0) continue;
          for (int c = 0; c < k; ++c) {
            if ((phw + p->pad - c) % s != 0) continue;
            a.dh[z][n] = (signed char)((phh + p->pad - r) / s);
            a.dw[z][n] = (signed char)((phw + p->pad - c) / s);
            a.ktap[z][n] = (signed char)(r * k + c);
            ++n;
          }
        }
        a.ntaps[z] = n;
      }
  }
  pick_patch(Hl, Wl, a.TW, a.TH, a.TN);
  a.tiles_w = (Wl + a.TW - 1) / a.TW;
  a.tiles_h = (Hl + a.TH - 1) / a.TH;
  const int tiles_n = (p->N + a.TN - 1) / a.TN;
  const int BN = pick_bn_for(p);
  const int Ktot = k * k * kmul * (p->C1 + p->C2);
  if (p->ldw < Ktot) { set_error("conv2d(tcgen05): ldw %d < kh*kw*Cin %d", p->ldw, Ktot); return STFB_EINVAL; }

  CUtensorMap tA, tA2, tB;
  const int BK = (p->C1 % 64 == 0 && p->C2 % 64 == 0) ? 64 : 32;
  const bool halo = halo_ok(p, BN);
  if (halo) { a.TW = HALO_TW; a.TH = HALO_TH; a.TN = 1; a.tiles_w = (Wl + a.TW - 1) / a.TW; a.tiles_h = (Hl + a.TH - 1) / a.TH; }
  const int tiles_n_eff = halo ? p->N : tiles_n;
  const int abox_w = halo ? HALO_TW + 2 : a.TW, abox_h = halo ? HALO_TH + 2 : a.TH;
  if (!encode_nhwc_map_strided(enc, &tA, p->x, p->N, p->H, p->W, pmul * p->C1, abox_w, abox_h, a.TN, a.a_scale, BK)) {
    set_error("conv2d(tcgen05): cuTensorMapEncodeTiled failed for x"); return STFB_ECUDA;
  }
  tA2 = tA;
  if (p->C2 > 0 && !encode_nhwc_map_strided(enc, &tA2, p->x2, p->N, p->H, p->W, pmul * p->C2, abox_w, abox_h, a.TN, a.a_scale, BK)) {
    set_error("conv2d(tcgen05): cuTensorMapEncodeTiled failed for x2"); return STFB_ECUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)p->Cout};
    cuuint64_t strides[1] = {(cuuint64_t)p->ldw * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("conv2d(tcgen05): cuTensorMapEncodeTiled failed for the weights"); return STFB_ECUDA;
    }
  }
  a.tiles_n = tiles_n_eff;
  a.num_tiles = a.nphase_w * a.nphase_w * tiles_n_eff * a.tiles_h * a.tiles_w * (p->Cout / BN);
  dim3 grid((unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms()));
  const bool f32out = p->y_dtype == STFB_F32;
  if (halo) {
    const int cpt = kmul * (p->C1 + p->C2) / 64;
    // canonical tap order tp = (dh + 1) * 3 + (dw + 1): permute the weight k-block table to match
    {
      signed char kt[9];
      for (int i = 0; i < 9; ++i) kt[(a.dh[0][i] + 1) * 3 + (a.dw[0][i] + 1)] = a.ktap[0][i];
      for (int tp = 0; tp < 9; ++tp) { a.ktap[0][tp] = kt[tp]; a.dh[0][tp] = (signed char)(tp / 3 - 1); a.dw[0][tp] = (signed char)(tp % 3 - 1); }
    }
    const int n_tiles_ = p->Cout / BN, pix_tiles = a.num_tiles / n_tiles_;
    {
      // CTA pairs (cta_group::2) whenever the pixel tiles pair up and there is at least one pair per two SMs worth of work
      const char* pe = getenv("STFB_HALO_PAIR");           // read per launch: tests switch it
      const int pair_mode = pe ? atoi(pe) : 1;
      if (pair_mode != 0 && pix_tiles % 2 == 0 && (pix_tiles / 2) * n_tiles_ >= 16) {
        CUtensorMap tBh;
        cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)p->Cout};
        cuuint64_t strides[1] = {(cuuint64_t)p->ldw * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(BN / 2)};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&tBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
          set_error("conv2d(tcgen05 pair): cuTensorMapEncodeTiled failed for the weights"); return STFB_ECUDA;
        }
        const int n_super2 = (pix_tiles / 2) * n_tiles_;
        // one channel block, one output-channel block, nine ring slots: the CTA's half of the weight matrix (36 KB at 64 -> 64)
        // is loaded once and stays resident
        a.w_resident = (cpt == 1 && n_tiles_ == 1 && BN <= 128) ? 1 : 0;
#define HALO2_LAUNCH(BN_, SA_, SB_)                                                                            \
        return f32out ? launch_halo2<BN_, SA_, SB_, float>(tA, tA2, tBh, a, n_super2, st)                     \
                      : launch_halo2<BN_, SA_, SB_, __nv_bfloat16>(tA, tA2, tBh, a, n_super2, st)
        switch (BN) {
          case 256: HALO2_LAUNCH(256, 4, 3);      // 94 KB halo ring + 3 x 16 KB weight halves
          case 128: HALO2_LAUNCH(128, 4, 9);      // + the whole tap set of a channel block: 9 x 8 KB
          case 64: HALO2_LAUNCH(64, 4, 9);        // 9 x 4 KB
        }
#undef HALO2_LAUNCH
      }
    }
    // MT = 2 halves the weight traffic per pixel but needs enough super tiles to keep the SMs busy
    // Pairing two pixel tiles per weight pass (MT = 2) halves the weight traffic and wins on L2-warm microbenchmarks
    // (64->64: 70 -> 52 us, 128->128: 47 -> 41 us) but loses inside the training step (13.45 vs 13.85 ms), where the tiles
    // stream from DRAM and the longer super-tile epilogue is exposed: off unless STFB_HALO_MT=2.
    const char* mt_env = getenv("STFB_HALO_MT");       // read per launch: tests switch it
    const int mt_mode = mt_env ? atoi(mt_env) : 0;
    const char* mt_bn = getenv("STFB_HALO_MT_BN");     // optional filter: comma list of the BN values MT = 2 applies to
    bool bn_sel = true;
    if (mt_bn && *mt_bn) {
      bn_sel = false;
      for (const char* q = mt_bn; *q;) {
        if (atoi(q) == BN) bn_sel = true;
        while (*q && *q != ',') ++q;
        if (*q == ',') ++q;
      }
    }
    const bool mt2 = mt_mode == 2 && bn_sel && ((pix_tiles + 1) / 2) * n_tiles_ >= (num_sms() * 3) / 4;
    const int n_super = mt2 ? ((pix_tiles + 1) / 2) * n_tiles_ : a.num_tiles;
    dim3 hgrid((unsigned)(n_super < num_sms() ? n_super : num_sms()));
#define HALO_LAUNCH(BN_, MT_, SA_, SB_)                                                                        \
    return f32out ? launch_halo<BN_, MT_, SA_, SB_, float>(tA, tA2, tB, a, hgrid, st)                         \
                  : launch_halo<BN_, MT_, SA_, SB_, __nv_bfloat16>(tA, tA2, tB, a, hgrid, st)
    // Two CTAs per SM (STFB_HALO_OCC2, default on for BN <= 128): the narrow layers are bound by the latency of their four
    // epilogue warps, not by the tensor pipe; a second resident CTA doubles the warps that drain accumulators.  It needs
    // <= 113 KB of shared memory and <= 168 registers per thread, so the weights stream (3-slot ring) instead of staying
    // resident and the A ring is two blocks deep.
    const char* occ_env = getenv("STFB_HALO_OCC2");
    const bool occ2 = !mt2 && (occ_env ? atoi(occ_env) != 0 : true) && a.num_tiles >= 2 * num_sms();
#define HALO_LAUNCH2(BN_, SA_, SB_)                                                                            \
    return f32out ? launch_halo<BN_, 1, SA_, SB_, float, 2>(tA, tA2, tB, a, hgrid2, st)                       \
                  : launch_halo<BN_, 1, SA_, SB_, __nv_bfloat16, 2>(tA, tA2, tB, a, hgrid2, st)
    dim3 hgrid2((unsigned)(a.num_tiles < 2 * num_sms() ? a.num_tiles : 2 * num_sms()));
    switch (BN) {
      case 256: if (mt2) { HALO_LAUNCH(256, 2, 4, 3); } else { HALO_LAUNCH(256, 1, 3, 3); }
      case 128:
        if (mt2) { HALO_LAUNCH(128, 2, 4, 3); }
        else if (occ2) { HALO_LAUNCH2(128, 2, 3); }
        else { HALO_LAUNCH(128, 1, 4, 3); }
      case 64:
        if (occ2) { a.w_resident = 0; HALO_LAUNCH2(64, 2, 3); }
        a.w_resident = (cpt == 1 && p->Cout == 64) ? 1 : 0;         // the whole 64 x 576 matrix = the 9-slot ring
        if (mt2) { HALO_LAUNCH(64, 2, 4, 9); } else { HALO_LAUNCH(64, 1, 4, 9); }
    }
#undef HALO_LAUNCH2
#undef HALO_LAUNCH
  }
#define TC_LAUNCH(BN_, ST_, BK_)                                                                              \
  return f32out ? launch_tc<BN_, ST_, float, BK_>(tA, tA2, tB, a, grid, st)                                 \
                : launch_tc<BN_, ST_, __nv_bfloat16, BK_>(tA, tA2, tB, a, grid, st)
  // two CTAs per SM for the 64-wide tiles when there is work for them (see the halo kernel): 3-stage ring.  (A 128-wide
  // tile with three 32 KB stages + staging is 256 bytes over half an SM's shared memory.)
  const char* occ_env2 = getenv("STFB_TC_OCC2");
  const bool tc_occ2 = (occ_env2 ? atoi(occ_env2) != 0 : true) && a.num_tiles >= 2 * num_sms();
  dim3 grid2((unsigned)(a.num_tiles < 2 * num_sms() ? a.num_tiles : 2 * num_sms()));
#define TC_LAUNCH2(BN_, ST_)                                                                                  \
  return f32out ? launch_tc<BN_, ST_, float, 64, 0, 2>(tA, tA2, tB, a, grid2, st)                            \
                : launch_tc<BN_, ST_, __nv_bfloat16, 64, 0, 2>(tA, tA2, tB, a, grid2, st)
  if (BK == 64) {
    if (tc_occ2 && BN == 64) { TC_LAUNCH2(64, 3); }        // 3 x 24 KB + staging = 89 KB
    switch (BN) {
      case 256: TC_LAUNCH(256, 4, 64);   // 4 x 48 KB
      case 128: TC_LAUNCH(128, 6, 64);   // 6 x 32 KB
      case 64: TC_LAUNCH(64, 8, 64);     // 8 x 24 KB
      case 32: TC_LAUNCH(32, 8, 64);     // 8 x 20 KB
    }
  } else {
    switch (BN) {                        // 32-channel inputs (decoder tail): small, memory-bound problems
      case 256: TC_LAUNCH(256, 6, 32);
      case 128: TC_LAUNCH(128, 8, 32);
      case 64: TC_LAUNCH(64, 8, 32);
      case 32: TC_LAUNCH(32, 8, 32);
    }
  }
#undef TC_LAUNCH2
#undef TC_LAUNCH
  set_error("conv2d(tcgen05): no tile for Cout=%d", p->Cout);
  return STFB_ENOTSUP;
}

// One LSTM step on the tensor cores, input and recurrent GEMM in ONE K loop with the cell update in the epilogue:
//   gates = [x_t, h_{t-1}] [W_ih | W_hh]^T (TMEM) + b_ih + b_hh ; c = sig(f) c_prev + sig(i) tanh(g) ; h = sig(o) tanh(c)
// w_xh_il: [4C][2C] K-major, rows chunk-interleaved (stfb_pack_weight_ex gate_c = C), columns [0,C) = W_ih, [C,2C) = W_hh.
// h_prev == NULL (t = 0): only the W_ih half of K is walked and c_prev is taken as zero.
int lstm_step_tcgen05(const void* x_t, const void* h_prev, const void* w_xh_il, const float* b_ih, const float* b_hh,
                      const float* c_prev, float* c_out, void* h_out, void* acts, int N, int H, int W, int C, cudaStream_t st) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) { set_error("lstm_step(tcgen05): cuTensorMapEncodeTiled not available"); return STFB_ECUDA; }
  if ((long long)N * H * W == 0) return STFB_OK;
  TcArgs a{};
  a.y = h_out; a.bias = b_ih; a.bias2 = b_hh; a.c_prev = h_prev ? c_prev : nullptr; a.c_out = c_out; a.acts = acts; a.Chid = C;
  a.N = N; a.Hout = H; a.Wout = W; a.Cout = 4 * C; a.C1 = C; a.C2 = h_prev ? C : 0;
  a.a_scale = 1; a.o_scale = 1; a.nphase_w = 1;
  a.ntaps[0] = 1; a.dh[0][0] = 0; a.dw[0][0] = 0; a.ktap[0][0] = 0;
  pick_patch(H, W, a.TW, a.TH, a.TN);
  a.tiles_w = (W + a.TW - 1) / a.TW;
  a.tiles_h = (H + a.TH - 1) / a.TH;
  a.tiles_n = (N + a.TN - 1) / a.TN;
  a.num_tiles = a.tiles_n * a.tiles_h * a.tiles_w * (4 * C / 256);
  CUtensorMap tA, tA2, tB;
  if (!encode_nhwc_map_strided(enc, &tA, x_t, N, H, W, C, a.TW, a.TH, a.TN, 1, 64)) {
    set_error("lstm_step(tcgen05): tensor map (x) failed"); return STFB_ECUDA;
  }
  tA2 = tA;
  if (h_prev && !encode_nhwc_map_strided(enc, &tA2, h_prev, N, H, W, C, a.TW, a.TH, a.TN, 1, 64)) {
    set_error("lstm_step(tcgen05): tensor map (h) failed"); return STFB_ECUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(a.C1 + a.C2), (cuuint64_t)(4 * C)};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * C) * 2};
    cuuint32_t box[2] = {64, 256};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_xh_il), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("lstm_step(tcgen05): tensor map (W) failed"); return STFB_ECUDA;
    }
  }
  dim3 grid((unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms()));
  return launch_tc<256, 3, __nv_bfloat16, 64, 1>(tA, tA2, tB, a, grid, st);
}


// All T steps of a 64-unit per-pixel LSTM in one launch (lstm_seq64_kernel).  x_seq: [T*B, H, W, 64] bf16 time-major;
// h_all: training [T][B,H,W,64], inference [B,H,W,64] (= h_T); c_all [T][rows][64] fp32 and acts_all [T][rows][256] bf16 only in
// training (keep != 0).  -> STFB_ENOTSUP when the geometry does not fit (caller falls back to one launch per step).
int lstm_seq_supported(int T, int B, int H, int W, int C) {
  if (C != 64 || T < 1 || B < 1) return 0;
  int TW, TH, TN;
  pick_patch(H, W, TW, TH, TN);
  return (B % TN == 0) ? 1 : 0;            // a tile must not straddle two time steps
}

int lstm_seq_tcgen05(const void* x_seq, const void* w_xh_il, const float* b_ih, const float* b_hh, float* c_all, void* h_all,
                     void* acts_all, int T, int B, int H, int W, int C, int keep, cudaStream_t st) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) { set_error("lstm_seq(tcgen05): cuTensorMapEncodeTiled not available"); return STFB_ECUDA; }
  if (!lstm_seq_supported(T, B, H, W, C)) { set_error("lstm_seq(tcgen05): unsupported geometry (C=%d, B=%d, %dx%d)", C, B, H, W); return STFB_ENOTSUP; }
  if ((long long)B * H * W == 0) return STFB_OK;
  LstmSeqArgs a{};
  a.bias = b_ih; a.bias2 = b_hh; a.c_all = keep ? c_all : nullptr; a.h_all = h_all; a.acts_all = keep ? acts_all : nullptr;
  a.rows = (long long)B * H * W; a.T = T; a.B = B; a.H = H; a.W = W; a.keep = keep ? 1 : 0;
  pick_patch(H, W, a.TW, a.TH, a.TN);
  a.tiles_w = (W + a.TW - 1) / a.TW;
  a.tiles_h = (H + a.TH - 1) / a.TH;
  a.tiles_n = B / a.TN;
  a.num_tiles = a.tiles_n * a.tiles_h * a.tiles_w;
  CUtensorMap tX, tW;
  if (!encode_nhwc_map_strided(enc, &tX, x_seq, T * B, H, W, C, a.TW, a.TH, a.TN, 1, 64)) {
    set_error("lstm_seq(tcgen05): tensor map (x) failed"); return STFB_ECUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(2 * C), (cuuint64_t)(4 * C)};
    cuuint64_t strides[1] = {(cuuint64_t)(2 * C) * 2};
    cuuint32_t box[2] = {64, 256};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_xh_il), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("lstm_seq(tcgen05): tensor map (W) failed"); return STFB_ECUDA;
    }
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lstm_seq64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lseq_smem_bytes()) != cudaSuccess) {
      set_error("lstm_seq(tcgen05): cannot reserve %d bytes of shared memory", lseq_smem_bytes());
      cudaGetLastError();
      return STFB_ECUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms()));
  lstm_seq64_kernel<<<grid, TC_THREADS, lseq_smem_bytes(), st>>>(tX, tW, a);
  return post_launch("lstm_seq(tcgen05)");
}


// One LSTM BACKWARD step on the tensor cores with the cell backward in the epilogue (conv_tc_kernel<EPI = 2>):
//   dh_{t-1} = dG_t W_hh (TMEM)  ->  dG_{t-1}, dc  (lstm_cell_bwd arithmetic on step t-1's saved activations and cell states)
// dg_next = dG_t [N,H,W,4C] bf16 gate-major; w_hh_d = W_hh packed for dgrad ([C][4C], K-major rows = hidden unit); acts / c_cur /
// c_prev belong to step t-1 (c_prev == NULL when t-1 == 0); dc is read and rewritten in place; dg_out = dG_{t-1}.
int lstm_bwd_step_tcgen05(const void* dg_next, const void* w_hh_d, const void* acts, const float* c_prev, const float* c_cur,
                          float* dc, void* dg_out, int N, int H, int W, int C, cudaStream_t st) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) { set_error("lstm_bwd_step(tcgen05): cuTensorMapEncodeTiled not available"); return STFB_ECUDA; }
  if ((long long)N * H * W == 0) return STFB_OK;
  const int BN = C % 256 == 0 ? 256 : (C % 128 == 0 ? 128 : 64);
  TcArgs a{};
  a.y = dg_out; a.acts = const_cast<void*>(acts); a.c_prev = c_prev; a.c_cur = c_cur; a.c_out = dc; a.Chid = C;
  a.N = N; a.Hout = H; a.Wout = W; a.Cout = C; a.C1 = 4 * C; a.C2 = 0;
  a.a_scale = 1; a.o_scale = 1; a.nphase_w = 1;
  a.ntaps[0] = 1; a.dh[0][0] = 0; a.dw[0][0] = 0; a.ktap[0][0] = 0;
  pick_patch(H, W, a.TW, a.TH, a.TN);
  a.tiles_w = (W + a.TW - 1) / a.TW;
  a.tiles_h = (H + a.TH - 1) / a.TH;
  a.tiles_n = (N + a.TN - 1) / a.TN;
  a.num_tiles = a.tiles_n * a.tiles_h * a.tiles_w * (C / BN);
  CUtensorMap tA, tB;
  if (!encode_nhwc_map_strided(enc, &tA, dg_next, N, H, W, 4 * C, a.TW, a.TH, a.TN, 1, 64)) {
    set_error("lstm_bwd_step(tcgen05): tensor map (dG) failed"); return STFB_ECUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(4 * C), (cuuint64_t)C};
    cuuint64_t strides[1] = {(cuuint64_t)(4 * C) * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_hh_d), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("lstm_bwd_step(tcgen05): tensor map (W) failed"); return STFB_ECUDA;
    }
  }
  // Shallow rings (K = 4C is only 4..32 k-blocks) and, where it fits, two CTAs per SM: the four LSTM levels run their backward
  // chains concurrently on four streams, and a CTA that takes a whole SM's shared memory serialises them.
  dim3 grid((unsigned)(a.num_tiles < num_sms() ? a.num_tiles : num_sms()));
  dim3 grid2((unsigned)(a.num_tiles < 2 * num_sms() ? a.num_tiles : 2 * num_sms()));
  switch (BN) {
    case 256: return launch_tc<256, 2, __nv_bfloat16, 64, 2>(tA, tA, tB, a, grid, st);        // 2 x 48 KB + 32 KB staging
    case 128: return launch_tc<128, 2, __nv_bfloat16, 64, 2, 2>(tA, tA, tB, a, grid2, st);    // 2 x 32 KB + 32 KB: two per SM
    default: return launch_tc<64, 3, __nv_bfloat16, 64, 2, 2>(tA, tA, tB, a, grid2, st);      // 3 x 24 KB + 32 KB: two per SM
  }
}

}  // namespace stfb
