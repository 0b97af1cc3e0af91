"""CPU restatement (numpy, integer / float64 arithmetic) of the reference's paired training augmentation.
TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ and never by the product.

The reference builds its training pipeline in train.py:51-67 from transforms.py:

    RandomResize(0.5*256 .. 1.2*256)  transforms.py:18-33   torchvision F.resize on PIL images: BILINEAR (image), NEAREST (mask)
    RandomHorizontalFlip(0.5)         :36-45
    RandomVerticalFlip(0.5)           :48-57
    RandomRotation(30)                :136-157              PIL Image.rotate, BILINEAR (image) / NEAREST (mask), with probability 0.5
    RandomCrop(224)                   :60-117               zero padding at the bottom / right if smaller, then a random window
    ToTensor, Normalize(0.709, 0.127) :120-133

and applies it once per DCE phase (my_dataset.py:173-179) with INDEPENDENT random draws per phase -- a reference bug
(SURVEY.md section 2 row 8); the B200 pipeline draws ONCE per sample and applies the same geometry to every phase and to
the mask.  This file restates the arithmetic of each step exactly as Pillow 12 / torchvision 0.26 execute it on 8-bit
single-channel images (the libraries are absent from /root/reference; the algorithms are restated from their behaviour
and pinned bit for bit against the live libraries by tests/golden/make_golden_augment.py):

  * Image.resize(BILINEAR): separable triangle filter, support max(scale, 1), coefficients normalised in double and
    quantised to 22-bit fixed point, horizontal pass then vertical pass, each rounded to 8 bits.
  * Image.resize(NEAREST): source index table built by ACCUMULATING the scale in double (xo += a0), truncated.
  * Image.rotate(BILINEAR): output centre (x+0.5, y+0.5) through the double-precision affine matrix (entries rounded to 15
    decimals), inside test on the raw coordinate, bilinear blend in double with clamped neighbours, TRUNCATED to 8 bits.
  * Image.rotate(NEAREST): 16.16 fixed-point affine walk.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def draw_params(rng, in_hw=(256, 256), min_size=128, max_size=307, hflip=0.5, vflip=0.5, degrees=30.0, crop=224):
    """One sample's random draws, in the order the reference's Compose consumes Python's `random` (transforms.py:26, :41, :53,
    :152-153, :100-101).  rng: a random.Random (or the random module)."""
    h, w = in_hw
    size = rng.randint(min_size, max_size)
    # torchvision F.resize(img, int): the SHORTER side becomes `size`, the other keeps the aspect ratio (int truncation)
    if w <= h:
        rw, rh = size, int(size * h / w)
    else:
        rh, rw = size, int(size * w / h)
    hf = rng.random() < hflip
    vf = rng.random() < vflip
    rot = rng.random() < 0.5
    angle = rng.uniform(-degrees, degrees) if rot else 0.0
    ph, pw = max(rh, crop), max(rw, crop)
    h0 = rng.randint(0, ph - crop)
    w0 = rng.randint(0, pw - crop)
    return {"rh": rh, "rw": rw, "hflip": hf, "vflip": vf, "rot": rot, "angle": angle, "h0": h0, "w0": w0, "crop": crop}


def resize_coeffs(in_size, out_size):
    """-> (bounds [out,2] = (first source index, tap count), coefficients [out, ksize] int) of Pillow's 8bpc bilinear resize."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            c = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + c * (1 << PRECISION_BITS)) if c < 0 else int(0.5 + c * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_bilinear_u8(img, out_h, out_w):
    in_h, in_w = img.shape
    a = img.astype(np.int64)
    if out_w != in_w:
        b, kk = resize_coeffs(in_w, out_w)
        tmp = np.zeros((in_h, out_w), dtype=np.int64)
        for xx in range(out_w):
            acc = np.full(in_h, 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for x in range(b[xx, 1]):
                acc += a[:, b[xx, 0] + x] * int(kk[xx, x])
            tmp[:, xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
        a = tmp
    if out_h != in_h:
        b, kk = resize_coeffs(in_h, out_h)
        out = np.zeros((out_h, a.shape[1]), dtype=np.int64)
        for yy in range(out_h):
            acc = np.full(a.shape[1], 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for y in range(b[yy, 1]):
                acc += a[b[yy, 0] + y] * int(kk[yy, y])
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        a = out
    return a.astype(np.uint8)


def nearest_table(n_out, n_in):
    """Source index per output index of Image.resize(NEAREST): ImagingScaleAffine accumulates xo += scale in double."""
    a0 = n_in / n_out
    xo = np.add.accumulate(np.concatenate([[a0 * 0.5], np.full(n_out - 1, a0)]))
    xi = np.where(xo < 0, -1, xo.astype(np.int64))
    return np.where(xi < n_in, xi, -1).astype(np.int32)


def resize_nearest(img, out_h, out_w):
    in_h, in_w = img.shape
    xi, yi = nearest_table(out_w, in_w), nearest_table(out_h, in_h)
    out = np.zeros((out_h, out_w), dtype=img.dtype)
    vx, vy = xi >= 0, yi >= 0
    out[np.ix_(vy, vx)] = img[np.ix_(yi[vy], xi[vx])]
    return out


def rotate_matrix(angle, w, h):
    """The affine matrix Image.rotate(angle, expand=False) hands to Image.transform (output pixel -> input coordinate)."""
    angle = angle % 360.0
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def rotate_bilinear_u8(img, angle):
    h, w = img.shape
    m = rotate_matrix(angle, w, h)
    ys, xs = np.mgrid[0:h, 0:w]
    xin = m[0] * (xs + 0.5) + m[1] * (ys + 0.5) + m[2]
    yin = m[3] * (xs + 0.5) + m[4] * (ys + 0.5) + m[5]
    inside = (xin >= 0.0) & (xin < w) & (yin >= 0.0) & (yin < h)
    xf, yf = xin - 0.5, yin - 0.5
    x0, y0 = np.floor(xf).astype(np.int64), np.floor(yf).astype(np.int64)
    dx, dy = xf - x0, yf - y0
    xc0, xc1 = np.clip(x0, 0, w - 1), np.clip(x0 + 1, 0, w - 1)
    yc0, yc1 = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    a = img.astype(np.float64)
    v1 = a[yc0, xc0] + (a[yc0, xc1] - a[yc0, xc0]) * dx
    v2 = a[yc1, xc0] + (a[yc1, xc1] - a[yc1, xc0]) * dx
    v2 = np.where((y0 + 1 >= 0) & (y0 + 1 < h), v2, v1)
    v = v1 + (v2 - v1) * dy
    return np.where(inside, np.floor(v), 0.0).astype(np.uint8)       # (UINT8) cast of a non-negative double: truncation


def fix16(v):
    return int(math.floor(v * 65536.0 + 0.5))


def rotate_fixed_coeffs(angle, w, h):
    a = rotate_matrix(angle, w, h)
    return [fix16(a[0]), fix16(a[1]), fix16(a[2] + a[0] * 0.5 + a[1] * 0.5), fix16(a[3]), fix16(a[4]),
            fix16(a[5] + a[3] * 0.5 + a[4] * 0.5)]


def rotate_nearest(img, angle):
    h, w = img.shape
    a0, a1, a2, a3, a4, a5 = rotate_fixed_coeffs(angle, w, h)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    xin, yin = (a2 + a1 * ys + a0 * xs) >> 16, (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    return np.where(ok, img[np.clip(yin, 0, h - 1), np.clip(xin, 0, w - 1)], 0).astype(img.dtype)


def apply(series_u8, mask_u8, p, mean=0.709, std=0.127):
    """series_u8 [T, H, W] uint8, mask_u8 [H, W] uint8 in {0, 1} -> (x float32 [T, 1, S, S], target int64 [S, S]): the
    reference's training pipeline with ONE set of draws `p` (draw_params) shared by all phases and the mask."""
    S = p["crop"]

    def geom(img, is_mask):
        r = resize_nearest(img, p["rh"], p["rw"]) if is_mask else resize_bilinear_u8(img, p["rh"], p["rw"])
        if p["hflip"]:
            r = r[:, ::-1]
        if p["vflip"]:
            r = r[::-1, :]
        r = np.ascontiguousarray(r)
        if p["rot"]:
            r = rotate_nearest(r, p["angle"]) if is_mask else rotate_bilinear_u8(r, p["angle"])
        ph, pw = max(r.shape[0], S), max(r.shape[1], S)
        padded = np.zeros((ph, pw), dtype=np.uint8)
        padded[:r.shape[0], :r.shape[1]] = r
        return padded[p["h0"]:p["h0"] + S, p["w0"]:p["w0"] + S]

    imgs = np.stack([geom(series_u8[t], False) for t in range(series_u8.shape[0])])
    x = ((imgs.astype(np.float32) / np.float32(255.0)) - np.float32(mean)) / np.float32(std)     # ToTensor.div(255), Normalize
    return x[:, None].astype(np.float32), geom(mask_u8, True).astype(np.int64)


def fixture_inputs(B=6, T=3, H=256, seed=41):
    """The deterministic 8-bit series + masks the augmentation fixture was generated from (numpy PCG64 streams are stable
    across versions, so the fixture stores digests of the outputs instead of megabytes of inputs)."""
    from stf_unet_b200.synthetic import synthetic_dce_batch_u8       # the product's generator (test-side use only)
    u8, _ = synthetic_dce_batch_u8(B, T, H, H, seed=seed)
    g = np.random.Generator(np.random.PCG64(3))
    u8 = np.clip(u8.numpy().astype(np.int64) + g.integers(-20, 21, size=tuple(u8.shape)), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:H]
    masks = np.zeros((B, H, H), dtype=np.uint8)
    for b in range(B):
        cy, cx, r = g.integers(70, 186), g.integers(70, 186), g.integers(12, 45)
        masks[b] = ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.uint8)
        masks[b][g.integers(0, H, 40), g.integers(0, H, 40)] ^= 1
    return u8, masks


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
