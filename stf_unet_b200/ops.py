"""Thin tensor-level wrappers over the C ABI (include/stfb200.h).

torch is used for device memory (``torch.empty``) and the current stream only; all arithmetic happens in
libstfb200.so.  Feature maps are contiguous NHWC tensors ``[N, H, W, C]`` in fp32 or bf16.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import BF16, BF16X3, CONV_FWD, CONV_TRANSPOSED, F32, IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, ConvParams, check

__all__ = ["dt_code", "conv2d", "conv2d_wgrad", "pack_weight", "bn_stats", "bn_finalize_train", "bn_fold_eval",
           "bn_apply", "bn_bwd", "colsum", "maxpool_fwd", "maxpool_bwd", "maxpool_fwd_idx", "maxpool_bwd_idx", "bilinear_fwd", "bilinear_bwd",
           "tcgen05_ok", "conv_stats_fusable", "STAT_SLOTS", "im2col_small", "unpad_wgrad", "lstm_step_fused", "lstm_bwd_step_fused", "lstm_seq_fused", "lstm_seq_supported", "pack_lstm_xh", "lstm_cell_fwd", "lstm_cell_bwd", "bn_apply_from_stats", "bn_relu_maxpool_from_stats", "bn_bwd_scratch_floats", "pack_series", "pack_series_u8", "adamw_flat_", "pack_series_maps", "repeat_batch", "nhwc_to_nchw", "nchw_to_nhwc", "add_", "cast",
           "ce_dice_fwd", "ce_dice_bwd", "split_bf16x3", "pack_weight_split", "CONV_FWD", "CONV_TRANSPOSED", "IMPL_AUTO", "IMPL_SIMT", "IMPL_TCGEN05"]

_DT = {torch.float32: F32, torch.bfloat16: BF16}


class KernelProfiler:
    """Per-launch CUDA-event timing of the GEMM-shaped kernels (bench.py roofline leg).  Events are recorded on
    torch's current stream, which is the stream every libstfb200 launch uses."""

    def __init__(self):
        self.records = []

    def begin(self):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def end(self, e0, family, flops, nbytes=0, tag=""):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((family, float(flops), float(nbytes), e0, e1, tag))

    def summary(self):
        torch.cuda.synchronize()
        fam = {}
        for family, flops, nbytes, e0, e1, _tag in self.records:
            d = fam.setdefault(family, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
            d["n"] += 1
        return fam

    def top(self, n=25):
        """Slowest launches grouped by (family, shape tag): [(ms, count, TFLOP/s, family, tag)]."""
        torch.cuda.synchronize()
        agg = {}
        for family, flops, nbytes, e0, e1, tag in self.records:
            d = agg.setdefault((family, tag), [0.0, 0, 0.0])
            d[0] += e0.elapsed_time(e1)
            d[1] += 1
            d[2] += flops
        rows = [(v[0], v[1], v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0, k[0], k[1]) for k, v in agg.items()]
        return sorted(rows, reverse=True)[:n]


_prof = None


class _timed:
    """with _timed(family, bytes): ...  -- records a CUDA-event pair when a profiler is installed (no-op otherwise)."""
    __slots__ = ("fam", "nbytes", "tag", "e0")

    def __init__(self, fam, nbytes, tag=""):
        self.fam, self.nbytes, self.tag, self.e0 = fam, nbytes, tag, None

    def __enter__(self):
        if _prof is not None:
            self.e0 = _prof.begin()
        return self

    def __exit__(self, *a):
        if self.e0 is not None:
            _prof.end(self.e0, self.fam, 0.0, self.nbytes, self.tag)
        return False


def _nb(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def set_profiler(p):
    global _prof
    _prof = p


def dt_code(dtype):
    try:
        return _DT[dtype]
    except KeyError:
        raise TypeError(f"stf_unet_b200 supports fp32 and bf16 activations, got {dtype}") from None


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stf_unet_b200 kernels need CUDA tensors (there is no CPU fallback)")


def conv_out_hw(H, W, k, stride, pad, transposed=False, out_pad=0):
    if transposed:
        return (H - 1) * stride - 2 * pad + k + out_pad, (W - 1) * stride - 2 * pad + k + out_pad
    return (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1


STAT_SLOTS = 8   # partial-sum slots of the fused BatchNorm statistics (CTA % slots): bounds the red.add contention


def conv2d(x, wp, Cout, k, stride, pad, *, mode=CONV_FWD, out_hw=None, x2=None, bias=None, bias2=None, scale=None,
           shift=None, residual=None, relu=False, y_dtype=None, ldw=None, out=None, w_offset=0, impl=IMPL_AUTO,
           stat_partial=None, stat_groups=0, split=False):
    """Implicit-GEMM convolution with fused epilogue; see stfb_conv2d in include/stfb200.h.

    split: x / x2 are split-precision operands (split_bf16x3: bf16 [N, H, W, 3C]), wp comes from pack_weight_split and the
    result is fp32 (STFB_BF16X3: fp32-accurate convolution on the tensor cores)."""
    _need_cuda(x, wp)
    N, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[3]
    if split:
        assert x.dtype == torch.bfloat16 and C1 % 3 == 0 and C2 % 3 == 0 and impl == IMPL_TCGEN05
        C1, C2 = C1 // 3, C2 // 3
        y_dtype = torch.float32
    kh, kw = (k, k) if isinstance(k, int) else k
    if out_hw is None:
        out_hw = conv_out_hw(H, W, kh, stride, pad, transposed=(mode == CONV_TRANSPOSED))
    Ho, Wo = out_hw
    y_dtype = y_dtype or x.dtype
    y = out if out is not None else torch.empty((N, Ho, Wo, Cout), dtype=y_dtype, device=x.device)
    esz = x.element_size()
    if impl == IMPL_TCGEN05 and ldw is None:
        ldw = wp.shape[1]
    p = ConvParams(x=_p(x), x2=_p(x2), w=wp.data_ptr() + w_offset * esz, y=_p(y), bias=_p(bias), bias2=_p(bias2),
                   scale=_p(scale), shift=_p(shift), residual=_p(residual), N=N, H=H, W=W, C1=C1, C2=C2, Ho=Ho, Wo=Wo,
                   Cout=Cout, kh=kh, kw=kw, stride=stride, pad=pad, ldw=ldw if ldw is not None else Cout, mode=mode,
                   relu=int(bool(relu)), x_dtype=BF16X3 if split else dt_code(x.dtype), y_dtype=dt_code(y.dtype), impl=impl,
                   stat_partial=_p(stat_partial), stat_slots=0 if stat_partial is None else stat_partial.shape[0],
                   stat_groups=stat_groups)
    if _prof is None:
        check(_lib.load().stfb_conv2d(C.byref(p), _stream()), "conv2d")
        return y
    lib = _lib.load()
    tc = impl == IMPL_TCGEN05
    # algorithmic FLOPs: transposed gathers count the taps that exist, not the zeros
    pix = N * Ho * Wo if mode == CONV_FWD else N * H * W
    flops = 2.0 * pix * Cout * (C1 + C2) * kh * kw
    nbytes = (x.numel() + (0 if x2 is None else x2.numel())) * esz + y.numel() * y.element_size() + wp.numel() * esz
    e0 = _prof.begin()
    check(lib.stfb_conv2d(C.byref(p), _stream()), "conv2d")
    # families follow the kernels: 3x3 / stride-1 "same" layers of 64-channel multiples on maps >= 12 x 8 run the halo kernels
    # (conv_halo2_kernel on CTA pairs, conv_halo_kernel otherwise), every other geometry the streaming conv_tc_kernel
    halo = tc and kh == 3 and kw == 3 and stride == 1 and C1 % 64 == 0 and C2 % 64 == 0 and Cout % 64 == 0 and Ho >= 12 and Wo >= 8 \
        and Ho == H and Wo == W
    fam = ("conv3x3_halo_tcgen05" if halo else "conv_other_tcgen05") if tc else "conv_simt_" + ("bf16" if esz == 2 else "f32")
    if split:
        fam += "_bf16x3"
    _prof.end(e0, fam, flops, nbytes, f"x{tuple(x.shape)}+{C2} ->{Cout} k{kh} s{stride} m{mode}")
    return y


_wg_scratch = {}


WGRAD_SCRATCH_SLOTS = 8   # one slot per stream that launches weight gradients (3 side streams, 4 LSTM forks, the main one)


def _ensure_wgrad_scratch(device):
    """Registers (once per device) the scratch buffer the tcgen05 wgrad kernels keep their per-split partial tiles in:
    WGRAD_SCRATCH_SLOTS slots, one per launching stream (launches on different streams must not share partial tiles)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _wg_scratch:
        lib = _lib.load()
        n = lib.stfb_wgrad_scratch_bytes() * WGRAD_SCRATCH_SLOTS
        buf = torch.empty((max(n, 16) // 4,), dtype=torch.float32, device=device)
        check(lib.stfb_set_wgrad_scratch(buf.data_ptr(), buf.numel() * 4), "set_wgrad_scratch")
        _wg_scratch[key] = buf


def conv2d_wgrad(P, G, dW, k, stride, pad, cg_off=0, cg_total=None, impl=IMPL_AUTO, acc=None, split=False):
    """dW[cp][cg_off+cg][ky][kx] += sum_pix P[pix,cp] * G[gather(pix),cg]; dW fp32, reference layout.

    acc: optional pre-zeroed fp32 accumulation buffer of the whole weight, [(ky,kx,cg_total), Cp].  When given and the
    tcgen05 family takes the shape, the launch only accumulates into it (returns True) and the caller folds it into dW
    later with ScatterPlan; otherwise dW is updated immediately (returns False)."""
    _need_cuda(P, G)
    _ensure_wgrad_scratch(P.device)
    N, Hp, Wp, Cp = P.shape
    _, Hg, Wg, Cg = G.shape
    if split:        # split-precision operands (split_bf16x3): three bf16 planes per tensor, fp32-accurate result
        assert P.dtype == torch.bfloat16 and G.dtype == torch.bfloat16 and Cp % 3 == 0 and Cg % 3 == 0
        Cp, Cg = Cp // 3, Cg // 3
    kh, kw = (k, k) if isinstance(k, int) else k
    cgt = cg_total if cg_total is not None else Cg
    lib = _lib.load()
    dt = BF16X3 if split else dt_code(P.dtype)
    tc = impl != IMPL_SIMT and lib.stfb_conv2d_wgrad_tcgen05_supported(_p(P), _p(G), N, Hp, Wp, Cp, Hg, Wg, Cg, kh, kw, stride,
                                                                       pad, dt) == 1
    deferred = tc and acc is not None
    e0 = _prof.begin() if _prof is not None else None
    if deferred:
        check(lib.stfb_conv2d_wgrad(_p(P), _p(G), None, N, Hp, Wp, Cp, Hg, Wg, Cg, cg_off, cgt, kh, kw, stride, pad, dt, impl,
                                    _p(acc), acc.numel() * 4, _stream()), "conv2d_wgrad")
    else:
        # non-zero for the tcgen05 family and for the pixel-paired 32-channel path (then `tc` is False but the launch is
        # still a tcgen05 one)
        ws_bytes = lib.stfb_conv2d_wgrad_workspace_bytes(_p(P), _p(G), N, Hp, Wp, Cp, Hg, Wg, Cg, cgt, kh, kw, stride, pad, dt,
                                                         impl) if (impl != IMPL_SIMT and cg_off == 0 and cgt == Cg) or tc else 0
        ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=P.device) if ws_bytes else None
        tc = tc or ws_bytes > 0
        check(lib.stfb_conv2d_wgrad(_p(P), _p(G), _p(dW), N, Hp, Wp, Cp, Hg, Wg, Cg, cg_off, cgt, kh, kw, stride, pad, dt,
                                    impl, _p(ws), ws_bytes, _stream()), "conv2d_wgrad")
    if e0 is not None:
        fam = ("wgrad_tcgen05_bf16x3" if split else "wgrad_tcgen05") if tc else "wgrad_simt_" + ("bf16" if P.element_size() == 2 else "f32")
        _prof.end(e0, fam, 2.0 * N * Hp * Wp * Cp * Cg * kh * kw, (P.numel() + G.numel()) * P.element_size(),
                  f"P{tuple(P.shape)} G{tuple(G.shape)} k{kh} s{stride}")
    return deferred


class ScatterPlan:
    """One launch that folds every deferred tcgen05 weight-gradient buffer into the flat gradient."""

    def __init__(self, entries, device):
        """entries: [(flat offset, Cp, Cg_total, khw)] of the weights whose wgrad accumulates in the side buffer."""
        import numpy as np
        job_t = np.dtype([("start", "<i8"), ("off", "<i8"), ("Cp", "<i4"), ("Cg", "<i4"), ("khw", "<i4"), ("pad_", "<i4")])
        jobs = np.zeros(len(entries), dtype=job_t)
        start = 0
        nelem = 0
        for j, (off, Cp, Cg, khw) in enumerate(entries):
            assert khw <= 9
            jobs[j] = (start, off, Cp, Cg, khw, 0)
            start += ((Cg + 31) // 32) * ((Cp + 31) // 32)      # tiles of 32 ci x 32 co
            nelem += Cp * Cg * khw
        self.total = start
        self.nelem = nelem
        self.n = len(entries)
        self.key = tuple(entries)
        self.table = torch.from_numpy(jobs.view(np.uint8)).to(device)

    def run(self, acc_flat, grad_flat):
        with _timed("wgrad_scatter", self.nelem * 12):
            check(_lib.load().stfb_wgrad_scatter_batched(_p(self.table), self.n, self.total, _p(acc_flat), _p(grad_flat),
                                                         _stream()), "wgrad_scatter_batched")


def pack_weight(w, k_is_dim1, dtype, n_major=False, flip=False, kpad=None, gate_c=0):
    """[D0, D1, kh, kw] (or [D0, D1]) fp32 parameter -> GEMM operand in `dtype`:
    [(ky,kx,k), n] (SIMT family) or, with n_major, [n, (ky,kx,k)] (tcgen05 family; kpad zero-pads each row)."""
    _need_cuda(w)
    if w.dim() == 2:
        D0, D1, kh, kw = w.shape[0], w.shape[1], 1, 1
    else:
        D0, D1, kh, kw = w.shape
    Kc, Nc = (D1, D0) if k_is_dim1 else (D0, D1)
    ld = 0
    if n_major:
        ld = kpad if kpad else kh * kw * Kc
        wp = (torch.zeros if kpad else torch.empty)((Nc, ld), dtype=dtype, device=w.device)
    else:
        wp = torch.empty((kh * kw * Kc, Nc), dtype=dtype, device=w.device)
    with _timed("pack_weight", _nb(w, wp)):
        check(_lib.load().stfb_pack_weight_ex(_p(w), _p(wp), D0, D1, kh, kw, int(k_is_dim1), int(n_major), int(flip), ld,
                                              int(gate_c), dt_code(dtype), _stream()), "pack_weight")
    return wp


def packed_shape(w, k_is_dim1, n_major, kpad=None):
    if w.dim() == 2:
        D0, D1, kh, kw = w.shape[0], w.shape[1], 1, 1
    else:
        D0, D1, kh, kw = w.shape
    Kc, Nc = (D1, D0) if k_is_dim1 else (D0, D1)
    if n_major:
        return (Nc, kpad if kpad else kh * kw * Kc)
    return (kh * kw * Kc, Nc)


class PackPlan:
    """Every weight pack of one forward(+backward) as a single launch.  The job table lives in device memory and is
    rebuilt only when a parameter's storage moves."""

    def __init__(self, params, keys, dtype):
        import numpy as np
        self.keys = list(keys)
        self.dtype = dtype
        dev = next(iter(params.values())).device
        self.buffers = {}
        job_t = np.dtype([("src", "<u8"), ("dst", "<u8"), ("start", "<i8"), ("D0", "<i4"), ("D1", "<i4"), ("khw", "<i4"),
                          ("k_is_dim1", "<i4"), ("n_major", "<i4"), ("flip", "<i4"), ("ld", "<i4"), ("pad_", "<i4")])
        jobs = np.zeros(len(self.keys), dtype=job_t)
        start = 0
        cats = {}
        for j, key in enumerate(self.keys):
            name, k_is_dim1, n_major, flip, kpad = key[:5]
            gate_c = key[5] if len(key) > 5 else 0
            cat = key[6] if len(key) > 6 else None       # (buffer name, column offset, row stride): K-concatenated operands
            x3 = len(key) > 7 and key[7] == "x3"         # split-precision operand (pack_weight_split layout, always bf16)
            w = params[name]
            shape = packed_shape(w, k_is_dim1, n_major, kpad)
            esz = torch.empty((), dtype=dtype).element_size()
            if x3:
                buf = torch.empty((shape[0], 6 * shape[1]), dtype=torch.bfloat16, device=dev)
                dst, ld = buf.data_ptr(), 6 * shape[1]
            elif cat is not None:
                cat_name, col_off, ld = cat
                buf = cats.get(cat_name)
                if buf is None:
                    buf = cats[cat_name] = torch.empty((shape[0], ld), dtype=dtype, device=dev)
                dst = buf.data_ptr() + col_off * esz
            else:
                buf = (torch.zeros if kpad else torch.empty)(shape, dtype=dtype, device=dev)
                dst, ld = buf.data_ptr(), shape[1]
            self.buffers[key] = buf
            D0, D1 = w.shape[0], w.shape[1]
            khw = 1 if w.dim() == 2 else w.shape[2] * w.shape[3]
            Kc = D1 if k_is_dim1 else D0
            vec = bool(n_major) and Kc % 8 == 0 and ld % 8 == 0 and dst % 16 == 0 and khw <= 9   # eight k per work item, 16-byte stores
            jobs[j] = (w.data_ptr(), dst, start, D0, D1, khw, int(k_is_dim1), int(n_major),
                       int(flip) | (2 if x3 else 0) | (4 if vec else 0), ld if n_major else 0, int(gate_c))
            start += PackPlan._items(jobs[j])     # one work item per (d0, d1) position (or eight of them), all taps
        self.total = start
        self._job_dtype = job_t
        self.table = torch.from_numpy(jobs.view(np.uint8)).to(dev)
        self.signature = tuple(params[k[0]].data_ptr() for k in self.keys)

    @staticmethod
    def _items(job):
        n = int(job["D0"]) * int(job["D1"])
        return n // 8 if int(job["flip"]) & 4 else n

    def valid_for(self, params, dtype):
        try:
            return dtype == self.dtype and self.signature == tuple(params[k[0]].data_ptr() for k in self.keys)
        except KeyError:
            return False

    def run(self, early=None):
        """One launch for every pack -- or, with `early` (a predicate on the parameter name), two: the packs the first layers
        need on the current stream, the rest on a low-priority side stream beside them.  -> (buffers, side stream or None);
        the caller orders the first consumer of a late pack behind the side stream."""
        lib = _lib.load()
        if early is None or not torch.cuda.is_available():
            with _timed("pack_weight", 0):
                check(lib.stfb_pack_weights_batched(_p(self.table), len(self.keys), self.total, dt_code(self.dtype), _stream()),
                      "pack_weights_batched")
            return dict(self.buffers), None
        split = getattr(self, "_split", None)
        if split is None:
            import numpy as np
            raw = self.table.cpu().numpy().view(self._job_dtype)
            idx_e = [j for j, k in enumerate(self.keys) if early(k[0])]
            idx_l = [j for j, k in enumerate(self.keys) if not early(k[0])]
            parts = []
            for idx in (idx_e, idx_l):
                jobs = raw[idx].copy()
                start = 0
                for j in range(len(jobs)):
                    jobs[j]["start"] = start
                    start += PackPlan._items(jobs[j])
                parts.append((torch.from_numpy(jobs.view(np.uint8).copy()).to(self.table.device), len(idx), start))
            split = self._split = parts
        cur = torch.cuda.current_stream()
        side = PackPlan._side.get(cur.device.index)
        if side is None:
            side = PackPlan._side[cur.device.index] = torch.cuda.Stream(device=cur.device)      # default = lowest priority
        (te, ne, tote), (tl, nl, totl) = split
        if nl:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                check(lib.stfb_pack_weights_batched(_p(tl), nl, totl, dt_code(self.dtype), side.cuda_stream), "pack_weights_batched")
        if ne:
            check(lib.stfb_pack_weights_batched(_p(te), ne, tote, dt_code(self.dtype), _stream()), "pack_weights_batched")
        return dict(self.buffers), (side if nl else None)

    _side = {}


def im2col_small(x, k, stride, pad, kpad, out=None):
    """bf16 [N,H,W,Cin] -> [N,Ho,Wo,kpad]: K order (ky,kx,ci), zero padded (small-channel convs on the tensor cores)."""
    N, H, W, Cin = x.shape
    Ho, Wo = conv_out_hw(H, W, k, stride, pad)
    if out is None:
        out = torch.empty((N, Ho, Wo, kpad), dtype=torch.bfloat16, device=x.device)
    with _timed("im2col_small", _nb(x, out)):
        check(_lib.load().stfb_im2col_small(_p(x), _p(out), N, H, W, Cin, Ho, Wo, k, stride, pad, kpad, _stream()),
              "im2col_small")
    return out


def unpad_wgrad(dW, src):
    """dW [Cout,Cin,kh,kw] += src [Cout, kpad] whose columns are ordered (ky,kx,ci)."""
    Cout, Cin, kh, kw = dW.shape
    check(_lib.load().stfb_unpad_wgrad(_p(dW), _p(src), Cout, Cin, kh, kw, src.shape[1], _stream()), "unpad_wgrad")


def tcgen05_ok(x, Cout, k, stride, pad, mode=CONV_FWD, x2=None, out_hw=None, y_dtype=None, as_split=False):
    """Shape/dtype check: can the tcgen05 family run this convolution?  as_split: x / x2 are the fp32 tensors whose
    split-precision form (split_bf16x3) would be the operand."""
    N, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[3]
    Ho, Wo = out_hw if out_hw is not None else conv_out_hw(H, W, k, stride, pad, transposed=(mode == CONV_TRANSPOSED))
    if as_split:
        if C1 % 8 != 0 or C2 % 8 != 0:
            return False
        p = ConvParams(x=None, x2=None, N=N, H=H, W=W, C1=C1, C2=C2, Ho=Ho, Wo=Wo, Cout=Cout, kh=k, kw=k, stride=stride,
                       pad=pad, mode=mode, x_dtype=BF16X3, y_dtype=F32)
    else:
        p = ConvParams(x=_p(x), x2=_p(x2), N=N, H=H, W=W, C1=C1, C2=C2, Ho=Ho, Wo=Wo, Cout=Cout, kh=k, kw=k, stride=stride,
                       pad=pad, mode=mode, x_dtype=dt_code(x.dtype), y_dtype=dt_code(y_dtype or x.dtype))
    return _lib.load().stfb_conv2d_tcgen05_supported(C.byref(p)) == 1


def wgrad_tcgen05_ok(P, G, k, stride, pad, split=False):
    """Can the tcgen05 family compute this weight gradient?  split: P / G are split-precision operands."""
    N, Hp, Wp, Cp = P.shape
    _, Hg, Wg, Cg = G.shape
    if split:
        Cp, Cg = Cp // 3, Cg // 3
    return _lib.load().stfb_conv2d_wgrad_tcgen05_supported(_p(P), _p(G), N, Hp, Wp, Cp, Hg, Wg, Cg, k, k, stride, pad,
                                                           BF16X3 if split else dt_code(P.dtype)) == 1


def split_bf16x3(x, out=None):
    """fp32 [..., C] -> bf16 [..., 3C] = [hi | mid | lo] (STFB_BF16X3, include/stfb200.h): the operand form of the
    fp32-accurate tensor-core convolutions."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    y = out if out is not None else torch.empty(x.shape[:-1] + (3 * Cc,), dtype=torch.bfloat16, device=x.device)
    with _timed("split_bf16x3", _nb(x, y), f"C{Cc}"):
        check(_lib.load().stfb_split_bf16x3(_p(x), _p(y), rows, Cc, _stream()), "split_bf16x3")
    return y


def pack_weight_split(w, k_is_dim1, flip=False):
    """[D0, D1, kh, kw] (or [D0, D1]) fp32 parameter -> [n, kh*kw*6*K] bf16: the B operand of a split-precision convolution
    (stfb_pack_weight_split)."""
    _need_cuda(w)
    if w.dim() == 2:
        D0, D1, kh, kw = w.shape[0], w.shape[1], 1, 1
    else:
        D0, D1, kh, kw = w.shape
    Kc, Nc = (D1, D0) if k_is_dim1 else (D0, D1)
    wp = torch.empty((Nc, kh * kw * 6 * Kc), dtype=torch.bfloat16, device=w.device)
    with _timed("pack_weight", _nb(w, wp)):
        check(_lib.load().stfb_pack_weight_split(_p(w), _p(wp), D0, D1, kh, kw, int(k_is_dim1), int(flip), _stream()),
              "pack_weight_split")
    return wp


def conv_stats_fusable(x, Cout, k, stride, pad, G, x2=None):
    """Can the tcgen05 conv of this shape also reduce the BatchNorm statistics of its (bf16) output over G image groups?"""
    N, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[3]
    Ho, Wo = conv_out_hw(H, W, k, stride, pad)
    p = ConvParams(x=_p(x), x2=_p(x2), N=N, H=H, W=W, C1=C1, C2=C2, Ho=Ho, Wo=Wo, Cout=Cout, kh=k, kw=k, stride=stride,
                   pad=pad, mode=CONV_FWD, x_dtype=dt_code(x.dtype), y_dtype=BF16)
    return _lib.load().stfb_conv2d_stats_fusable(C.byref(p), G) == 1


def bn_stats(x, G, R, C):
    """-> per-CTA partial sums [nblk, 2, G, C] fp32 (combined in fp64 by bn_finalize_train)."""
    lib = _lib.load()
    nblk = lib.stfb_bn_partial_blocks(G, R)
    partial = torch.empty((nblk, 2, G, C), dtype=torch.float32, device=x.device)
    with _timed("bn_stats", _nb(x), f"C{C}"):
        check(lib.stfb_bn_stats(_p(x), _p(partial), nblk, G, R, C, dt_code(x.dtype), _stream()), "bn_stats")
    return partial


def bn_finalize_train(partial, gamma, beta, running_mean, running_var, nbt, G, R, C, eps=1e-5, momentum=0.1, out=None):
    if out is None:
        out = torch.empty((4, G, C), dtype=torch.float32, device=partial.device)  # scale, shift, mean, invstd
    check(_lib.load().stfb_bn_finalize_train(_p(partial), partial.shape[0], _p(gamma), _p(beta), _p(running_mean),
                                             _p(running_var), _p(nbt), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), G, R,
                                             C, eps, momentum, _stream()), "bn_finalize_train")
    return out


def bn_fold_eval(gamma, beta, running_mean, running_var, eps=1e-5):
    C_ = gamma.numel()
    out = torch.empty((2, C_), dtype=torch.float32, device=gamma.device)
    check(_lib.load().stfb_bn_fold_eval(_p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(out[0]), _p(out[1]), C_,
                                        eps, _stream()), "bn_fold_eval")
    return out


def bn_apply_from_stats(x, partial, gamma, beta, G, R, C, relu, residual=None, out=None, eps=1e-5):
    """bn_apply with scale/shift derived in-kernel from <= 8 statistics slots (same arithmetic as bn_finalize_train)."""
    y = out if out is not None else torch.empty_like(x)
    with _timed("bn_apply", _nb(x, residual, y), f"C{C}"):
        check(_lib.load().stfb_bn_apply_from_stats(_p(x), _p(partial), partial.shape[0], _p(gamma), _p(beta), _p(residual), _p(y),
                                                   G, R, C, eps, int(bool(relu)), dt_code(x.dtype), _stream()), "bn_apply_from_stats")
    return y


def bn_relu_maxpool_from_stats(x, partial, gamma, beta, G, k, stride, pad, eps=1e-5, out=None):
    """maxpool(relu(batchnorm(x))) + argmax index in one pass; scale/shift derived in-kernel from <= 8 statistics slots.
    -> (y [N,Ho,Wo,C], idx uint8): exactly what bn_apply_from_stats + maxpool_fwd_idx give, without the full-size map."""
    N, H, W, C_ = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    if out is not None:
        y, idx = out
    else:
        y = torch.empty((N, Ho, Wo, C_), dtype=x.dtype, device=x.device)
        idx = torch.empty((N, Ho, Wo, C_), dtype=torch.uint8, device=x.device)
    with _timed("bn_relu_maxpool", _nb(x, y) + idx.numel(), f"C{C_}"):
        check(_lib.load().stfb_bn_relu_maxpool_from_stats(_p(x), _p(partial), partial.shape[0], _p(gamma), _p(beta), _p(y), _p(idx),
                                                          G, N, H, W, C_, Ho, Wo, k, stride, pad, eps, dt_code(x.dtype), _stream()),
              "bn_relu_maxpool_from_stats")
    return y, idx


def bn_apply(x, scale, shift, G, R, C, relu, residual=None, out=None):
    y = out if out is not None else torch.empty_like(x)
    with _timed("bn_apply", _nb(x, residual, y), f"C{C}"):
        check(_lib.load().stfb_bn_apply(_p(x), _p(scale), _p(shift), _p(residual), _p(y), G, R, C, int(bool(relu)),
                                        dt_code(x.dtype), _stream()), "bn_apply")
    return y


USE_FUSED_BN_BWD = os.environ.get("STFB_NO_FUSED_BN_BWD", "0") != "1"


def bn_bwd_scratch_floats(G, C):
    """Zeroed fp32 scratch the one-launch BatchNorm backward needs (group sums + arrival counters)."""
    return int(_lib.load().stfb_bn_bwd_fused_scratch_floats(G, C))


def bn_bwd(dy, y, x, mean, invstd, gamma, dgamma, dbeta, G, R, C, relu, want_dres, dres_acc=None, scale=None, shift=None,
           scratch=None):
    """Full BatchNorm(+ReLU)(+residual) backward: returns (dx, dres or None); dgamma/dbeta accumulated.
    dres_acc: existing gradient of the residual input, accumulated into in place.
    scale/shift (the forward's, [G][C]): with relu and no residual the mask is recomputed from x and y is not read.
    scratch: ZEROED fp32 buffer of bn_bwd_scratch_floats(G, C) elements -> the whole backward is ONE launch
    (reduce -> per-group barrier -> finalize -> apply); without it, the three-launch chain."""
    lib = _lib.load()
    s = _stream()
    from_x = bool(relu) and not want_dres and dres_acc is None and scale is not None and shift is not None
    ym = None if (from_x or not relu) else y
    dx = torch.empty_like(x)
    dres = dres_acc if dres_acc is not None else (torch.empty_like(x) if want_dres else None)
    # two tensors per pass (no y, no dres) over a map far beyond L2: the chain's deeper grids still stream faster
    # (tools/kernel_probe.py bn2: 278 vs 307 us on the stem's 268 MB map, 81 vs 85 us on layer 1's 67 MB)
    lean_and_large = ym is None and dres is None and x.numel() * x.element_size() > (48 << 20)
    if scratch is not None and USE_FUSED_BN_BWD and not lean_and_large:
        # algorithmic bytes: every input once (dy, x, the ReLU mask source, the residual gradient being
        # accumulated into) + every output once (dx, dres); the kernel's second pass over dy / x is NOT counted
        with _timed("bn_bwd_fused", _nb(dy, x, ym, dres_acc) + _nb(dx, dres), f"C{C}"):
            check(lib.stfb_bn_bwd_fused(_p(dy), _p(ym), _p(x), _p(mean), _p(invstd), _p(gamma), _p(shift) if from_x else None,
                                        _p(scratch), _p(dgamma), _p(dbeta), _p(dx), _p(dres), int(dres_acc is not None), G, R, C,
                                        int(bool(relu)), dt_code(x.dtype), s), "bn_bwd_fused")
        return dx, dres
    nblk = lib.stfb_bn_partial_blocks(G, R)
    red = torch.empty((nblk, 2, G, C), dtype=torch.float32, device=x.device)
    with _timed("bn_bwd_reduce", _nb(dy, x, ym), f"C{C}"):
        check(lib.stfb_bn_bwd_reduce(_p(dy), _p(ym), _p(x), _p(mean), _p(invstd), _p(scale) if from_x else None,
                                     _p(shift) if from_x else None, _p(red), nblk, G, R, C, int(bool(relu)), dt_code(x.dtype), s),
              "bn_bwd_reduce")
    coef = torch.empty((G, C, 3), dtype=torch.float32, device=x.device)
    check(lib.stfb_bn_bwd_finalize(_p(red), nblk, _p(gamma), _p(invstd), _p(dgamma), _p(dbeta), _p(coef), G, R, C, s),
          "bn_bwd_finalize")
    with _timed("bn_bwd_apply", _nb(dy, x, ym, dres_acc) + _nb(dx, dres), f"C{C}"):
        check(lib.stfb_bn_bwd_apply(_p(dy), _p(ym), _p(x), _p(mean), _p(invstd), _p(coef), _p(shift) if from_x else None, _p(dx),
                                    _p(dres), int(dres_acc is not None), G, R, C, int(bool(relu)), dt_code(x.dtype), s),
              "bn_bwd_apply")
    return dx, dres


def colsum(x, out, R, C):
    with _timed("colsum", R * C * x.element_size(), f"C{C}"):
        check(_lib.load().stfb_colsum(_p(x), _p(out), R, C, dt_code(x.dtype), _stream()), "colsum")


def maxpool_fwd(x, k, stride, pad):
    N, H, W, C_ = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty((N, Ho, Wo, C_), dtype=x.dtype, device=x.device)
    check(_lib.load().stfb_maxpool_fwd(_p(x), _p(y), N, H, W, C_, Ho, Wo, k, stride, pad, dt_code(x.dtype), _stream()),
          "maxpool_fwd")
    return y


def maxpool_fwd_idx(x, k, stride, pad):
    """Forward that also returns the uint8 first-max window position per output element (training)."""
    N, H, W, C_ = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty((N, Ho, Wo, C_), dtype=x.dtype, device=x.device)
    idx = torch.empty((N, Ho, Wo, C_), dtype=torch.uint8, device=x.device)
    with _timed("maxpool_fwd", _nb(x, y, idx)):
        check(_lib.load().stfb_maxpool_fwd_idx(_p(x), _p(y), _p(idx), N, H, W, C_, Ho, Wo, k, stride, pad,
                                               dt_code(x.dtype), _stream()), "maxpool_fwd_idx")
    return y, idx


def maxpool_bwd_idx(idx, dy, in_shape, k, stride, pad):
    N, H, W, C_ = in_shape
    _, Ho, Wo, _ = dy.shape
    dx = torch.empty(in_shape, dtype=dy.dtype, device=dy.device)
    with _timed("maxpool_bwd", _nb(idx, dy, dx)):
        check(_lib.load().stfb_maxpool_bwd_idx(_p(idx), _p(dy), _p(dx), N, H, W, C_, Ho, Wo, k, stride, pad,
                                               dt_code(dy.dtype), _stream()), "maxpool_bwd_idx")
    return dx


def maxpool_bwd(x, dy, k, stride, pad):
    N, H, W, C_ = x.shape
    _, Ho, Wo, _ = dy.shape
    dx = torch.empty_like(x)
    check(_lib.load().stfb_maxpool_bwd(_p(x), _p(dy), _p(dx), N, H, W, C_, Ho, Wo, k, stride, pad, dt_code(x.dtype),
                                       _stream()), "maxpool_bwd")
    return dx


def bilinear_fwd(x, Ho, Wo):
    N, H, W, C_ = x.shape
    y = torch.empty((N, Ho, Wo, C_), dtype=x.dtype, device=x.device)
    check(_lib.load().stfb_bilinear_fwd(_p(x), _p(y), N, H, W, C_, Ho, Wo, dt_code(x.dtype), _stream()), "bilinear_fwd")
    return y


def bilinear_bwd(dy, H, W):
    N, Ho, Wo, C_ = dy.shape
    dx = torch.zeros((N, H, W, C_), dtype=torch.float32, device=dy.device)
    check(_lib.load().stfb_bilinear_bwd(_p(dy), _p(dx), N, H, W, C_, Ho, Wo, dt_code(dy.dtype), _stream()), "bilinear_bwd")
    return dx


def lstm_step_fused(x_t, h_prev, w_xh_il, b_ih, b_hh, c_prev, c_out, h_out, acts):
    """One LSTM step: tcgen05 GEMM over [x_t, h_prev] @ [W_ih | W_hh]^T with bias + cell update fused into the epilogue.
    h_prev / c_prev None at t = 0.  acts (optional) receives the post-activation gates in accumulator column order."""
    N, H, W, C_ = x_t.shape
    e0 = _prof.begin() if _prof is not None else None
    check(_lib.load().stfb_lstm_step_fused(_p(x_t), _p(h_prev), _p(w_xh_il), _p(b_ih), _p(b_hh), _p(c_prev), _p(c_out),
                                           _p(h_out), _p(acts), N, H, W, C_, _stream()), "lstm_step_fused")
    if e0 is not None:
        rows = N * H * W
        kc = C_ * (2 if h_prev is not None else 1)
        _prof.end(e0, "lstm_step_tcgen05", 2.0 * rows * kc * 4 * C_,
                  rows * C_ * (2 + 4 + 2 + (2 + 4 if h_prev is not None else 0) + (8 if acts is not None else 0)),
                  f"lstm_step_fused rows{rows} C{C_} K{kc}")


def lstm_bwd_step_fused(dg_next, w_hh_d, acts, c_prev, c_cur, dc, dg_out):
    """dG_{t-1}, dc <- (dG_t W_hh, saved state of step t-1): the recurrent GEMM of the LSTM backward pass with the cell backward
    in its epilogue.  dg_next / dg_out [N,H,W,4C] bf16 (gate-major), acts [N,H,W,4C] in accumulator column order."""
    N, H, W, C4 = dg_next.shape
    C_ = C4 // 4
    e0 = _prof.begin() if _prof is not None else None
    check(_lib.load().stfb_lstm_bwd_step_fused(_p(dg_next), _p(w_hh_d), _p(acts), _p(c_prev), _p(c_cur), _p(dc), _p(dg_out), N, H, W,
                                               C_, _stream()), "lstm_bwd_step_fused")
    if e0 is not None:
        rows = N * H * W
        _prof.end(e0, "lstm_step_tcgen05", 2.0 * rows * C4 * C_, rows * C_ * 40, f"lstm_bwd_step_fused rows{rows} C{C_}")


def lstm_seq_supported(T, B, H, W, C_):
    return _lib.load().stfb_lstm_seq_supported(T, B, H, W, C_) == 1


def lstm_seq_fused(x_seq, w_xh_il, b_ih, b_hh, T, c_all, h_all, acts_all):
    """All T steps of a 64-unit per-pixel LSTM level in one launch.  x_seq [T*B, H, W, C] bf16 time-major.  Training: c_all
    [T,R,C] fp32, h_all [T,B,H,W,C], acts_all [T,R,4C]; inference: c_all = acts_all = None and h_all [B,H,W,C] receives h_T."""
    TB, H, W, C_ = x_seq.shape
    B = TB // T
    keep = c_all is not None
    e0 = _prof.begin() if _prof is not None else None
    check(_lib.load().stfb_lstm_seq_fused(_p(x_seq), _p(w_xh_il), _p(b_ih), _p(b_hh), _p(c_all), _p(h_all), _p(acts_all), T, B, H, W,
                                          C_, int(keep), _stream()), "lstm_seq_fused")
    if e0 is not None:
        rows = B * H * W
        _prof.end(e0, "lstm_step_tcgen05", 2.0 * rows * C_ * 4 * C_ * (2 * T - 1),
                  rows * C_ * (2 * T + ((4 + 2 + 8) * T if keep else 2)), f"lstm_seq_fused rows{rows} C{C_} T{T}")


def pack_lstm_xh(w_ih, w_hh, dtype):
    """[W_ih | W_hh] as one K-major [4C][2C] operand with chunk-interleaved gate rows (for lstm_step_fused)."""
    C4, C_ = w_ih.shape
    wp = torch.empty((C4, 2 * C_), dtype=dtype, device=w_ih.device)
    lib = _lib.load()
    esz = wp.element_size()
    with _timed("pack_weight", _nb(w_ih, w_hh, wp)):
        check(lib.stfb_pack_weight_ex(_p(w_ih), wp.data_ptr(), C4, C_, 1, 1, 1, 1, 0, 2 * C_, C_, dt_code(dtype), _stream()),
              "pack_weight")
        check(lib.stfb_pack_weight_ex(_p(w_hh), wp.data_ptr() + C_ * esz, C4, C_, 1, 1, 1, 1, 0, 2 * C_, C_, dt_code(dtype),
                                      _stream()), "pack_weight")
    return wp


def lstm_cell_fwd(gates, c_prev, acts, c_out, h_out, R, C_):
    with _timed("lstm_cell_fwd", R * C_ * (16 + (4 if c_prev is not None else 0) + 4 + h_out.element_size() * (5 if acts is not None else 1))):
        _lstm_cell_fwd(gates, c_prev, acts, c_out, h_out, R, C_)


def _lstm_cell_fwd(gates, c_prev, acts, c_out, h_out, R, C_):
    check(_lib.load().stfb_lstm_cell_fwd(_p(gates), _p(c_prev), _p(acts), _p(c_out), _p(h_out), R, C_, dt_code(h_out.dtype),
                                         _stream()), "lstm_cell_fwd")


def lstm_cell_bwd(dh, dc, acts, c_prev, c_cur, dgates, R, C_, acts_il=False):
    with _timed("lstm_cell_bwd", R * C_ * (4 + 8 + 4 + (4 if c_prev is not None else 0) + 8 * dgates.element_size())):
        _lstm_cell_bwd(dh, dc, acts, c_prev, c_cur, dgates, R, C_, acts_il)


def _lstm_cell_bwd(dh, dc, acts, c_prev, c_cur, dgates, R, C_, acts_il=False):
    check(_lib.load().stfb_lstm_cell_bwd(_p(dh), _p(dc), _p(acts), _p(c_prev), _p(c_cur), _p(dgates), R, C_, int(acts_il),
                                         dt_code(dgates.dtype), _stream()), "lstm_cell_bwd")


def pack_series(x, dtype):
    """[B, T, C, H, W] fp32 -> [T*B, H, W, C] `dtype`, time-major image order."""
    _need_cuda(x)
    B, T, C_, H, W = x.shape
    y = torch.empty((T * B, H, W, C_), dtype=dtype, device=x.device)
    check(_lib.load().stfb_pack_series(_p(x), _p(y), B, T, C_, H, W, dt_code(dtype), _stream()), "pack_series")
    return y


def pack_series_u8(x, dtype, mean, std):
    """8-bit series [B, T, H, W] (or [B, T, 1, H, W]) -> normalised [T*B, H, W, 1] `dtype`, time-major:
    ((x / 255) - mean) / std, the loader's ToTensor + Normalize fused into the layout pass."""
    _need_cuda(x)
    if x.dtype != torch.uint8:
        raise TypeError("pack_series_u8 takes uint8 grey levels")
    if x.dim() == 5:
        if x.shape[2] != 1:
            raise ValueError("pack_series_u8: 8-bit input is single-channel ([B, T, 1, H, W])")
        x = x[:, :, 0]
    x = x.contiguous()
    B, T, H, W = x.shape
    y = torch.empty((T * B, H, W, 1), dtype=dtype, device=x.device)
    check(_lib.load().stfb_pack_series_u8(_p(x), _p(y), B, T, H, W, float(mean), float(std), dt_code(dtype), _stream()),
          "pack_series_u8")
    return y


def adamw_flat_(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """One AdamW step in place over flat fp32 buffers (ONE launch for the whole model)."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
            raise ValueError("adamw_flat_: four contiguous fp32 buffers of equal length are required")
    check(_lib.load().stfb_adamw_flat(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n, float(lr), float(beta1), float(beta2),
                                      float(eps), float(weight_decay), int(step), float(grad_scale), _stream()), "adamw_flat")


def pack_series_maps(x, maps, dtype):
    """x [B,T,Cx,H,W] + maps [B,Cm,H,W] (fp32) -> [T*B, H, W, Cx+Cm] `dtype`, time-major."""
    _need_cuda(x, maps)
    B, T, Cx, H, W = x.shape
    Cm = maps.shape[1]
    y = torch.empty((T * B, H, W, Cx + Cm), dtype=dtype, device=x.device)
    check(_lib.load().stfb_pack_series_maps(_p(x), _p(maps), _p(y), B, T, Cx, Cm, H, W, dt_code(dtype), _stream()),
          "pack_series_maps")
    return y


def repeat_batch(src, times):
    """[B, ...] -> [times*B, ...]: `times` back-to-back copies (per-sample maps repeated for every time step)."""
    dst = torch.empty((times * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    check(_lib.load().stfb_repeat(_p(src), _p(dst), src.numel() * src.element_size(), times, _stream()), "repeat")
    return dst


def nhwc_to_nchw(y):
    N, H, W, C_ = y.shape
    out = torch.empty((N, C_, H, W), dtype=torch.float32, device=y.device)
    check(_lib.load().stfb_nhwc_to_nchw(_p(y), _p(out), N, H, W, C_, dt_code(y.dtype), _stream()), "nhwc_to_nchw")
    return out


def nchw_to_nhwc(g, dtype):
    _need_cuda(g)
    N, C_, H, W = g.shape
    out = torch.empty((N, H, W, C_), dtype=dtype, device=g.device)
    check(_lib.load().stfb_nchw_to_nhwc(_p(g), _p(out), N, H, W, C_, dt_code(dtype), _stream()), "nchw_to_nhwc")
    return out


def add_(dst, src):
    assert dst.dtype == src.dtype and dst.numel() == src.numel()
    check(_lib.load().stfb_add_inplace(_p(dst), _p(src), dst.numel(), dt_code(dst.dtype), _stream()), "add_inplace")
    return dst


def cast(src, dtype):
    dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    check(_lib.load().stfb_cast(_p(src), dt_code(src.dtype), _p(dst), dt_code(dtype), src.numel(), _stream()), "cast")
    return dst


def ce_dice_fwd(logits, target, eps=1e-6, weight=None, ignore_index=-100, dice=True):
    """logits NCHW fp32, target int64 -> (loss_out[3] = {total, ce, dice}, stats).  weight: fp32 [C] class weights of the
    cross-entropy term or None; pixels labelled ignore_index enter neither term; dice=False drops the Dice term."""
    _need_cuda(logits, target, weight)
    B, C_, H, W = logits.shape
    stats = torch.empty((B * C_ * 3 + 3,), dtype=torch.float64, device=logits.device)
    out = torch.empty((3,), dtype=torch.float32, device=logits.device)
    check(_lib.load().stfb_ce_dice_fwd_ex(_p(logits), _p(target), _p(weight), _p(stats), _p(out), B, C_, H * W, eps,
                                          int(ignore_index), int(bool(dice)), _stream()), "ce_dice_fwd")
    return out, stats


def ce_dice_bwd(logits, target, stats, dloss, eps=1e-6, weight=None, ignore_index=-100, dice=True):
    B, C_, H, W = logits.shape
    dl = torch.empty_like(logits)
    check(_lib.load().stfb_ce_dice_bwd_ex(_p(logits), _p(target), _p(weight), _p(stats), _p(dloss), _p(dl), B, C_, H * W, eps,
                                          int(ignore_index), int(bool(dice)), _stream()), "ce_dice_bwd")
    return dl
