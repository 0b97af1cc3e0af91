"""Paired train-time augmentation on the device -- the step before the hot path.

Mirror of the reference's ``get_transform`` (/root/reference/train.py:51-75) over ``transforms.py`` (:18-157): the training
pipeline RandomResize(128..307) -> RandomHorizontalFlip -> RandomVerticalFlip -> RandomRotation(30) -> RandomCrop(224) ->
ToTensor -> Normalize(0.709, 0.127), and the validation pipeline (resize to 224, ToTensor, Normalize).  The reference runs it
on the CPU through PIL, one image at a time, once per DCE phase (my_dataset.py:173-179); here the raw 8-bit series of a whole
batch goes to the device and ONE kernel (csrc/augment.cu) emits the normalised crops of every phase plus the mask, bit for bit
what PIL produces.

What stays on the host is the control logic: the random draws -- same generator (Python's ``random``), same order, same
distributions as the reference's Compose, ONE set per sample shared by all phases and the mask (the reference re-draws per
phase: a bug) -- and the per-size resize tables (Pillow's coefficient recipe in double precision,
cached per (input, output) size: at most 180 sizes exist).
"""
from __future__ import annotations

import ctypes as C
import math
import random

import numpy as np
import torch

from . import _lib
from ._lib import check

__all__ = ["PairedAugment", "AugSample"]

PRECISION_BITS = 32 - 8 - 2


class AugSample(C.Structure):
    """stfb_aug_sample (include/stfb200.h)."""
    _fields_ = [("rh", C.c_int), ("rw", C.c_int), ("hflip", C.c_int), ("vflip", C.c_int), ("rot", C.c_int), ("h0", C.c_int),
                ("w0", C.c_int), ("tab_off", C.c_int), ("ksize_h", C.c_int), ("ksize_v", C.c_int), ("fix", C.c_int * 6),
                ("pad_", C.c_int), ("m", C.c_double * 6)]


def _bilinear_tables(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the triangle filter: (bounds [out,2], coeffs [out,ksize]) int32;
    (None, None) when the pass is skipped (same size)."""
    if in_size == out_size:
        return None, None
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)             # C (int) cast of a non-negative double
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    j = np.arange(ksize, dtype=np.float64)[None, :]
    v = np.abs((j + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(v < 1.0, 1.0 - v, 0.0)
    w = np.where(j < xmax[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for t in range(ksize):                                                       # sequential sum, like the C loop
        ww = ww + w[:, t]
    c = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kk = (0.5 + c * (1 << PRECISION_BITS)).astype(np.int64)                      # coefficients are non-negative for this filter
    kk = np.where(j < xmax[:, None], kk, 0).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk


def _nearest_table(n_out, n_in):
    """Source index per output index of Image.resize(NEAREST) (ImagingScaleAffine: the scale is ACCUMULATED in double)."""
    a0 = n_in / n_out
    xo = np.add.accumulate(np.concatenate([[a0 * 0.5], np.full(n_out - 1, a0)]))
    xi = np.where(xo < 0, -1, xo.astype(np.int64))
    return np.where(xi < n_in, xi, -1).astype(np.int32)


def _rotate_matrix(angle, w, h):
    """The matrix Image.rotate(angle, expand=False) passes to Image.transform (entries rounded to 15 decimals, like PIL)."""
    angle = angle % 360.0
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def _fix16(v):
    return int(math.floor(v * 65536.0 + 0.5))


class PairedAugment:
    """aug = PairedAugment(train=True); x, target = aug(series_u8, masks_u8)   (both CUDA uint8 tensors)

    series_u8 [B, T, H, W] (or [B, T, 1, H, W]) raw grey levels, masks_u8 [B, H, W] in {0, 1} ->
    x float32 [B, T, 1, S, S] normalised, target int64 [B, S/stride, S/stride] (target_stride=2 gives STF-LSTM-UNet its
    half-resolution target; 1 is the reference's).  `rng`: a random.Random for reproducible draws (default: Python's global
    generator, what the reference's transforms consume)."""

    def __init__(self, train=True, base_size=256, crop_size=224, hflip_prob=0.5, vflip_prob=0.5, degrees=30.0, mean=0.709, std=0.127,
                 target_stride=1, rng=None):
        self.train = train
        self.crop = int(crop_size)
        self.min_size, self.max_size = (int(0.5 * base_size), int(1.2 * base_size)) if train else (self.crop, self.crop)
        self.hflip_prob, self.vflip_prob, self.degrees = hflip_prob, vflip_prob, float(degrees)
        self.mean, self.std = float(mean), float(std)
        self.target_stride = int(target_stride)
        self.rng = rng if rng is not None else random
        self._tab_cache = {}

    # ---- draws: the order in which the reference's Compose consumes `random` (transforms.py:26, :41, :53, :152-153, :100-101)
    def draw(self, in_hw):
        h, w = in_hw
        r = self.rng
        size = r.randint(self.min_size, self.max_size)
        if w <= h:            # torchvision F.resize(img, int): shorter side -> size, the other side keeps the aspect (truncated)
            rw, rh = size, int(size * h / w)
        else:
            rh, rw = size, int(size * w / h)
        if not self.train:    # validation: RandomResize(crop) only; the crop is the whole resized image's top-left S x S
            return {"rh": rh, "rw": rw, "hflip": False, "vflip": False, "rot": False, "angle": 0.0, "h0": 0, "w0": 0}
        hf = r.random() < self.hflip_prob
        vf = r.random() < self.vflip_prob
        rot = r.random() < 0.5
        angle = r.uniform(-self.degrees, self.degrees) if rot else 0.0
        ph, pw = max(rh, self.crop), max(rw, self.crop)
        return {"rh": rh, "rw": rw, "hflip": hf, "vflip": vf, "rot": rot, "angle": angle, "h0": r.randint(0, ph - self.crop),
                "w0": r.randint(0, pw - self.crop)}

    def _tables(self, H, W, rh, rw):
        key = (H, W, rh, rw)
        t = self._tab_cache.get(key)
        if t is None:
            hb, hk = _bilinear_tables(W, rw)
            vb, vk = _bilinear_tables(H, rh)
            ksh = 0 if hk is None else hk.shape[1]
            ksv = 0 if vk is None else vk.shape[1]
            parts = [np.zeros((rw, 2), np.int32) if hb is None else hb, np.zeros((rw, 0), np.int32) if hk is None else hk,
                     np.zeros((rh, 2), np.int32) if vb is None else vb, np.zeros((rh, 0), np.int32) if vk is None else vk,
                     _nearest_table(rw, W), _nearest_table(rh, H)]
            t = (np.concatenate([p.reshape(-1) for p in parts]).astype(np.int32), ksh, ksv)
            self._tab_cache[key] = t
        return t

    def plan(self, B, H, W, params=None):
        """Host side of one batch: -> (list of draws, AugSample array, int32 table vector)."""
        params = params if params is not None else [self.draw((H, W)) for _ in range(B)]
        samples = (AugSample * B)()
        tabs, off = [], 0
        for b, p in enumerate(params):
            tab, ksh, ksv = self._tables(H, W, p["rh"], p["rw"])
            s = samples[b]
            s.rh, s.rw, s.hflip, s.vflip, s.rot = p["rh"], p["rw"], int(p["hflip"]), int(p["vflip"]), int(p["rot"])
            s.h0, s.w0, s.tab_off, s.ksize_h, s.ksize_v = p["h0"], p["w0"], off, ksh, ksv
            if p["rot"]:
                m = _rotate_matrix(p["angle"], p["rw"], p["rh"])
                for i in range(6):
                    s.m[i] = m[i]
                fx = [_fix16(m[0]), _fix16(m[1]), _fix16(m[2] + m[0] * 0.5 + m[1] * 0.5), _fix16(m[3]), _fix16(m[4]),
                      _fix16(m[5] + m[3] * 0.5 + m[4] * 0.5)]
                for i in range(6):
                    s.fix[i] = fx[i]
            tabs.append(tab)
            off += tab.size
        return params, samples, np.concatenate(tabs) if tabs else np.zeros(0, np.int32)

    def __call__(self, series_u8, masks_u8=None, params=None):
        if not series_u8.is_cuda or series_u8.dtype != torch.uint8:
            raise RuntimeError("PairedAugment takes CUDA uint8 tensors (there is no CPU fallback)")
        if series_u8.dim() == 5:
            if series_u8.shape[2] != 1:
                raise ValueError("8-bit input is single-channel ([B, T, 1, H, W])")
            series_u8 = series_u8[:, :, 0]
        series_u8 = series_u8.contiguous()
        B, T, H, W = series_u8.shape
        if masks_u8 is not None:
            if masks_u8.dtype != torch.uint8 or tuple(masks_u8.shape) != (B, H, W) or not masks_u8.is_cuda:
                raise ValueError("masks: CUDA uint8 [B, H, W] expected")
            masks_u8 = masks_u8.contiguous()
        _, samples, tables = self.plan(B, H, W, params)
        dev = series_u8.device
        S = self.crop
        s_dev = torch.frombuffer(bytearray(bytes(samples)), dtype=torch.uint8).to(dev) if B else torch.empty(0, dtype=torch.uint8, device=dev)
        t_dev = torch.from_numpy(tables).to(dev)
        x = torch.empty((B, T, 1, S, S), dtype=torch.float32, device=dev)
        So = (S + self.target_stride - 1) // self.target_stride
        tgt = torch.empty((B, So, So), dtype=torch.int64, device=dev) if masks_u8 is not None else None
        check(_lib.load().stfb_augment_series_u8(series_u8.data_ptr(), None if masks_u8 is None else masks_u8.data_ptr(),
                                                 s_dev.data_ptr() if B else None, t_dev.data_ptr() if B else None, x.data_ptr(),
                                                 None if tgt is None else tgt.data_ptr(), B, T, H, W, S, self.target_stride,
                                                 self.mean, self.std, torch.cuda.current_stream().cuda_stream), "augment_series_u8")
        return x, tgt
