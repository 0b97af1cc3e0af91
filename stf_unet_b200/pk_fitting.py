"""Extended-Tofts PK-map fitting on libstfb200 -- mirror of ``ToftsModelFitter`` (/root/reference/pk_fitting.py:10-602) for the
two GPU-heavy members the reference has: ``extended_tofts_model_batch`` (:193-231) and ``fit_volume_gpu`` (:233-420).

Same constructor (``time_points``, ``device``, ``aif_method``), same AIF models (``population_aif`` :28-46, ``modified_aif``
:48-56), same fitting recipe (initial guess 0.05 / 0.1 / 0.01, Adam lr 0.005, 100 epochs of 1024-pixel batches, clamps
:302-306).  The difference is where the loop runs: the reference launches ~50 000 small kernels per slice (100 epochs x
batches x 8 time points x a dozen ATen ops); here ONE kernel owns each pixel for the whole fit (csrc/tofts.cu).

Out of scope, as in SURVEY.md section 2 row 11: patient / dataset walkers, AIF auto-detection (broken in the reference:
``aif_concentration`` is undefined at :127), plotting.  ``preprocess_images`` (:157-191) keeps its cv2 morphology on the host.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import check

__all__ = ["ToftsModelFitter"]


def _p(t):
    return None if t is None else t.data_ptr()


class ToftsModelFitter:
    CONV_DT = 0.01                                     # pk_fitting.py:205

    def __init__(self, time_points=None, device=None, aif_method="population"):
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        if self.device.type != "cuda":
            raise RuntimeError("stf_unet_b200.pk_fitting runs on CUDA (sm_100a) only; there is no CPU fallback")
        tp = [0, 1, 2, 3, 4, 5, 6, 7] if time_points is None else time_points
        self.time_points = torch.tensor(tp, dtype=torch.float32, device=self.device)
        if aif_method not in ("population", "modified", "auto"):
            raise ValueError(f"unsupported AIF method: {aif_method}")
        self.aif_method = aif_method
        self._tables = {}

    # ---- arterial input functions (host-side table builders; the tables are a few hundred floats) ----------------------
    @staticmethod
    def population_aif(t, dose=0.1):
        a1, a2, m1, m2 = 3.99, 4.78, 0.144, 0.0111
        return dose * (a1 * torch.exp(-m1 * t) + a2 * torch.exp(-m2 * t))

    @staticmethod
    def modified_aif(t):
        a1, a2, m1, m2 = 3.99, 4.78, 0.144, 0.0111
        return a1 * torch.exp(-m1 * t) + a2 * torch.exp(-m2 * t)

    def aif(self, t):
        if self.aif_method == "population":
            return self.population_aif(t)
        # 'auto' without a detected curve falls back to the bi-exponential model, like the reference (:86-88)
        return self.modified_aif(t)

    def _get_tables(self, t):
        """(t, aif(t), t_conv, aif(t_conv), nvalid) on the device for the acquisition times `t` (cached per time vector)."""
        t_host = t.detach().to("cpu", torch.float32).contiguous()
        key = tuple(t_host.tolist())
        tb = self._tables.get(key)
        if tb is None:
            if t_host.numel() == 0 or t_host.numel() > 32:
                raise ValueError("1..32 time points supported")
            max_time = float(t_host[-1])
            t_conv = torch.arange(0, max_time, self.CONV_DT, dtype=torch.float32)          # :207
            nvalid = torch.tensor([int((t_conv < ti).sum()) for ti in t_host], dtype=torch.int32)
            if not bool((t_conv[1:] > t_conv[:-1]).all()) or any(int(n) != int((t_conv[:int(n)] < ti).sum()) for n, ti in zip(nvalid, t_host)):
                raise ValueError("time grid must be ascending")
            dev = self.device
            tb = (t_host.to(dev), self.aif(t_host).to(dev), t_conv.to(dev), self.aif(t_conv).to(dev), nvalid.to(dev))
            self._tables[key] = tb
        return tb

    # ---- the forward model ---------------------------------------------------------------------------------------------
    def extended_tofts_model_batch(self, t, Ktrans, ve, vp):
        """[N] parameters -> [N, T] concentration curves (pk_fitting.py:193-231)."""
        for x in (Ktrans, ve, vp):
            if not x.is_cuda:
                raise RuntimeError("stf_unet_b200.pk_fitting needs CUDA tensors (there is no CPU fallback)")
        tt, at, tc, ac, nv = self._get_tables(t)
        K, E, V = (x.detach().float().contiguous() for x in (Ktrans, ve, vp))
        N = K.numel()
        out = torch.empty((N, tt.numel()), dtype=torch.float32, device=K.device)
        check(_lib.load().stfb_tofts_forward(_p(tt), _p(at), _p(tc), _p(ac), _p(nv), tt.numel(), tc.numel(), self.CONV_DT, _p(K), _p(E),
                                             _p(V), _p(out), N, torch.cuda.current_stream().cuda_stream), "tofts_forward")
        return out

    # ---- the fit -------------------------------------------------------------------------------------------------------
    def fit_pixels(self, valid_pixels, epochs=100, batch_size=1024, lr=0.005, init=(0.05, 0.1, 0.01), betas=(0.9, 0.999), eps=1e-8,
                   clamp_lo=(0.0, 0.001, 0.0), clamp_hi=(1.0, 0.5, 0.2), return_losses=False):
        """valid_pixels [N, T] curves -> (Ktrans, ve, vp) [N] each: the optimisation of fit_volume_gpu (:288-368) in one launch."""
        if not valid_pixels.is_cuda:
            raise RuntimeError("stf_unet_b200.pk_fitting needs CUDA tensors (there is no CPU fallback)")
        px = valid_pixels.detach().float().contiguous()
        N, T = px.shape
        if T != self.time_points.numel():
            raise ValueError(f"curves have {T} time points, the fitter {self.time_points.numel()}")
        tt, at, tc, ac, nv = self._get_tables(self.time_points)
        dev = px.device
        K = torch.full((N,), init[0], dtype=torch.float32, device=dev)
        E = torch.full((N,), init[1], dtype=torch.float32, device=dev)
        V = torch.full((N,), init[2], dtype=torch.float32, device=dev)
        nb = (N + batch_size - 1) // batch_size
        steps = max(1, epochs * nb)
        # bias corrections in double on the host, like torch.optim.Adam (_single_tensor_adam)
        ss = np.array([lr / (1.0 - betas[0] ** s) for s in range(1, steps + 1)], dtype=np.float32)
        bc = np.array([math.sqrt(1.0 - betas[1] ** s) for s in range(1, steps + 1)], dtype=np.float32)
        ss_d, bc_d = torch.from_numpy(ss).to(dev), torch.from_numpy(bc).to(dev)
        losses = torch.zeros((max(epochs, 1),), dtype=torch.float32, device=dev) if return_losses else None
        import ctypes as C
        lo = (C.c_float * 3)(*clamp_lo)
        hi = (C.c_float * 3)(*clamp_hi)
        check(_lib.load().stfb_tofts_fit(_p(px), _p(tt), _p(at), _p(tc), _p(ac), _p(nv), T, tc.numel(), self.CONV_DT, _p(K), _p(E), _p(V),
                                         N, int(batch_size), int(epochs), _p(ss_d), _p(bc_d), float(betas[0]), float(betas[1]),
                                         float(eps), C.cast(lo, C.c_void_p), C.cast(hi, C.c_void_p), _p(losses),
                                         torch.cuda.current_stream().cuda_stream), "tofts_fit")
        return (K, E, V, losses[:epochs]) if return_losses else (K, E, V)

    def preprocess_images(self, images):
        """uint8 [T, H, W] -> ([T, H, W] float in [0, 1] on the device, bool tissue mask) -- pk_fitting.py:157-191 (the 5x5
        open / close of the threshold mask stays on the host with cv2, as in the reference)."""
        import cv2
        images = np.asarray(images)
        images_tensor = torch.tensor(images, dtype=torch.float32, device=self.device) / 255.0
        first = images[0]
        mask = first > np.mean(first) * 0.15
        kernel = np.ones((5, 5), np.uint8)
        mask = cv2.morphologyEx(mask.astype(np.uint8), cv2.MORPH_OPEN, kernel)
        mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, kernel)
        return images_tensor, torch.tensor(mask, dtype=torch.bool, device=self.device)

    def fit_volume_gpu(self, subtraction_images, output_dir=None, debug_output_dir=None):
        """uint8 [T, H, W] subtraction images -> float32 [3, H, W] maps (Ktrans, ve, vp), zeros outside the tissue mask
        (pk_fitting.py:233-420).  output_dir: `{name}_raw.npy` per map, like the reference; no PNG / heat-map rendering."""
        T, H, W = np.asarray(subtraction_images).shape
        images_tensor, tissue = self.preprocess_images(subtraction_images)
        pixels = images_tensor.permute(1, 2, 0).reshape(-1, T)
        mask = tissue.reshape(-1)
        valid = pixels[mask].contiguous()
        K, E, V = self.fit_pixels(valid)
        maps = torch.zeros((3, H * W), dtype=torch.float32, device=self.device)
        maps[0, mask], maps[1, mask], maps[2, mask] = K, E, V
        out = maps.reshape(3, H, W).cpu().numpy()
        if output_dir is not None:
            import os
            os.makedirs(output_dir, exist_ok=True)
            for i, name in enumerate(("ktrans", "ve", "vp")):
                np.save(os.path.join(output_dir, f"{name}_raw.npy"), out[i])
        return out
