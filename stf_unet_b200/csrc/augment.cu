// Paired train-time augmentation of an 8-bit DCE series and its mask on the device.
//
// Reference: get_transform(train=True) (/root/reference/train.py:51-67) = RandomResize -> RandomHorizontalFlip ->
// RandomVerticalFlip -> RandomRotation -> RandomCrop -> ToTensor -> Normalize (/root/reference/transforms.py:18-157), applied
// on the CPU through PIL, one image at a time, once per DCE phase (/root/reference/my_dataset.py:173-179).
//
// Here ONE launch produces the normalised crops of all T phases and the mask of a whole batch.  Every output pixel walks the
// chain backwards (crop -> rotation -> flips -> resize) to the source pixels it depends on and re-does the reference's
// arithmetic for exactly those: Pillow's two-pass fixed-point triangle-filter resize (22-bit coefficients, 8-bit rounding
// after each pass), its double-precision bilinear rotation with truncation to 8 bits, its 16.16 fixed-point nearest
// rotation and accumulated-scale nearest resize for the mask.  Results are bit-identical to the PIL pipeline
// (tests/golden/augment_6x3x256.npz, generated from the reference's own transforms).  The geometry is drawn once per
// sample on the host (stf_unet_b200/augment.py, same draw order as the reference) and shared by all phases and the mask --
// the reference draws independently per phase, a bug.
//
// Cost: a rotated pixel needs 4 resized neighbours x (<= 5 x 5 source taps); the whole batch of 16 x 8 x 224 x 224 outputs is
// ~0.6 G integer multiply-adds -- the input is 8 MB, the output 26 MB: a memory-light kernel whose point is to take the
// loader's PIL work (~40 ms of CPU per sample) off the critical path.
#include "common.cuh"

namespace stfb {

constexpr int AUG_PREC = 22;

struct AugGeo {
  const unsigned char* src;    // one phase [H][W]
  int W;
  const int* hb; const int* hk; int ksh;   // horizontal pass: bounds [rw][2], coefficients [rw][ksh]; ksh == 0: identity
  const int* vb; const int* vk; int ksv;   // vertical pass
};

__device__ __forceinline__ int aug_clip8(int acc) {
  const int v = acc >> AUG_PREC;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontally resized 8-bit value at (source row, resized column)
__device__ __forceinline__ int aug_hres(const AugGeo& g, int row, int xx) {
  const unsigned char* r = g.src + (long long)row * g.W;
  if (g.ksh == 0) return r[xx];
  const int x0 = g.hb[2 * xx], n = g.hb[2 * xx + 1];
  int acc = 1 << (AUG_PREC - 1);
  for (int i = 0; i < n; ++i) acc += (int)r[x0 + i] * g.hk[xx * g.ksh + i];
  return aug_clip8(acc);
}

// resized 8-bit value at (yy, xx) of the resized image: Pillow runs the horizontal pass over the rows the vertical pass needs
__device__ __forceinline__ int aug_resized(const AugGeo& g, int yy, int xx) {
  if (g.ksv == 0) return aug_hres(g, yy, xx);
  const int y0 = g.vb[2 * yy], n = g.vb[2 * yy + 1];
  int acc = 1 << (AUG_PREC - 1);
  for (int j = 0; j < n; ++j) acc += aug_hres(g, y0 + j, xx) * g.vk[yy * g.ksv + j];
  return aug_clip8(acc);
}

__global__ void __launch_bounds__(128) augment_series_u8_kernel(const unsigned char* __restrict__ series,
                                                                 const unsigned char* __restrict__ masks,
                                                                 const stfb_aug_sample* __restrict__ samples,
                                                                 const int* __restrict__ tables, float* __restrict__ x_out,
                                                                 long long* __restrict__ t_out, int B, int T, int H, int W, int S,
                                                                 int tstride, float mean, float stdv) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= S) return;
  const stfb_aug_sample sm = samples[b];
  const int* tb = tables + sm.tab_off;
  AugGeo g;
  g.W = W; g.ksh = sm.ksize_h; g.ksv = sm.ksize_v;
  g.hb = tb; g.hk = g.hb + 2 * sm.rw;
  g.vb = g.hk + sm.rw * sm.ksize_h; g.vk = g.vb + 2 * sm.rh;
  const int* xtab = g.vk + sm.rh * sm.ksize_v;
  const int* ytab = xtab + sm.rw;
  const int yc = y + sm.h0, xc = x + sm.w0;                  // position in the (zero padded) rotated image
  const bool in_img = yc < sm.rh && xc < sm.rw;

  // ---- image: up to four resized neighbours blended in double (Pillow's bilinear_filter8), or one pixel without rotation
  int py[2] = {yc, yc}, px[2] = {xc, xc};
  double dx = 0.0, dy = 0.0;
  bool live = in_img, y1ok = false, blend = false;
  if (in_img && sm.rot) {
    // Pillow is plain C without fused multiply-adds: a0 * xin + a1 * yin + a2 rounds after every operation.  Spelled with the
    // _rn intrinsics so that nvcc does not contract the expression into FMAs (one ulp in the coordinate moves a pixel whose
    // blend lands exactly on an integer across the truncation below).
    const double xin = (double)xc + 0.5, yin = (double)yc + 0.5;
    const double xi = __dadd_rn(__dadd_rn(__dmul_rn(sm.m[0], xin), __dmul_rn(sm.m[1], yin)), sm.m[2]);
    const double yi = __dadd_rn(__dadd_rn(__dmul_rn(sm.m[3], xin), __dmul_rn(sm.m[4], yin)), sm.m[5]);
    live = xi >= 0.0 && xi < (double)sm.rw && yi >= 0.0 && yi < (double)sm.rh;
    if (live) {
      const double xf = xi - 0.5, yf = yi - 0.5;
      const double x0 = floor(xf), y0 = floor(yf);
      dx = xf - x0; dy = yf - y0;
      const int ix = (int)x0, iy = (int)y0;
      px[0] = min(max(ix, 0), sm.rw - 1); px[1] = min(max(ix + 1, 0), sm.rw - 1);
      py[0] = min(max(iy, 0), sm.rh - 1); py[1] = min(max(iy + 1, 0), sm.rh - 1);
      y1ok = iy + 1 >= 0 && iy + 1 < sm.rh;
      blend = true;
    }
  }
  // flips act on the resized image: position (yy, xx) of the flipped image is (rh-1-yy, rw-1-xx) of the resized one
  int ry[2], rx[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    ry[k] = sm.vflip ? sm.rh - 1 - py[k] : py[k];
    rx[k] = sm.hflip ? sm.rw - 1 - px[k] : px[k];
  }
  for (int t = 0; t < T; ++t) {
    int v = 0;
    if (live) {
      g.src = series + ((long long)b * T + t) * H * W;
      if (!blend) {
        v = aug_resized(g, ry[0], rx[0]);
      } else {
        // BILINEAR(v, a, b, d): v = a + (b - a) * d, again without contraction
        const double a00 = aug_resized(g, ry[0], rx[0]), a01 = aug_resized(g, ry[0], rx[1]);
        double v1 = __dadd_rn(a00, __dmul_rn(a01 - a00, dx)), v2 = v1;
        if (y1ok) {
          const double a10 = aug_resized(g, ry[1], rx[0]), a11 = aug_resized(g, ry[1], rx[1]);
          v2 = __dadd_rn(a10, __dmul_rn(a11 - a10, dx));
        }
        v = (int)__dadd_rn(v1, __dmul_rn(v2 - v1, dy));      // (UINT8) cast: truncation
      }
    }
    // ToTensor: x / 255; Normalize: (x - mean) / std -- IEEE divisions in the loader's order, like stfb_pack_series_u8
    x_out[(((long long)b * T + t) * S + y) * S + x] = (((float)v / 255.0f) - mean) / stdv;
  }

  // ---- mask: nearest neighbour all the way (16.16 fixed-point rotation, accumulated-scale resize tables)
  if (t_out != nullptr && (y % tstride) == 0 && (x % tstride) == 0) {
    int mv = 0;
    if (in_img) {
      int my = yc, mx = xc;
      bool ok = true;
      if (sm.rot) {
        const long long xx = (long long)sm.fix[2] + (long long)sm.fix[1] * yc + (long long)sm.fix[0] * xc;
        const long long yy = (long long)sm.fix[5] + (long long)sm.fix[4] * yc + (long long)sm.fix[3] * xc;
        mx = (int)(xx >> 16); my = (int)(yy >> 16);
        ok = mx >= 0 && mx < sm.rw && my >= 0 && my < sm.rh;
      }
      if (ok) {
        const int fy = sm.vflip ? sm.rh - 1 - my : my, fx = sm.hflip ? sm.rw - 1 - mx : mx;
        const int sy = ytab[fy], sx = xtab[fx];
        if (sy >= 0 && sx >= 0) mv = masks[((long long)b * H + sy) * W + sx];
      }
    }
    const int So = (S + tstride - 1) / tstride;
    t_out[((long long)b * So + y / tstride) * So + x / tstride] = mv;
  }
}

}  // namespace stfb

using namespace stfb;

extern "C" int stfb_augment_series_u8(const unsigned char* series, const unsigned char* masks, const stfb_aug_sample* samples_dev,
                                      const int* tables_dev, float* x_out, long long* target_out, int B, int T, int H, int W, int S,
                                      int target_stride, float mean, float stdv, void* stream) {
  STFB_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && S > 0 && target_stride >= 1, "augment_series_u8: bad sizes");
  if (B == 0) return STFB_OK;
  STFB_REQUIRE(series && samples_dev && tables_dev && x_out, "augment_series_u8: null argument");
  STFB_REQUIRE((target_out == nullptr) || masks != nullptr, "augment_series_u8: a target needs the masks");
  STFB_REQUIRE(B <= 65535 && S <= 65535, "augment_series_u8: batch or crop too large");
  STFB_DEVICE_OR_RETURN();
  dim3 grid((unsigned)((S + 127) / 128), (unsigned)S, (unsigned)B);
  augment_series_u8_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(series, masks, samples_dev, tables_dev, x_out,
                                                                                     target_out, B, T, H, W, S, target_stride, mean, stdv);
  return post_launch("augment_series_u8");
}
