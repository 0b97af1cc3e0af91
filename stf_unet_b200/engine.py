"""Execution engine of the hot path: a forward executor that records a backward tape.

The reference leaves graph building to torch autograd, one node per library operator.  Here the whole model
forward is ONE autograd node (see ``ModelFunction``); inside it every operator is a libstfb200 kernel launch and
the backward pass is this module's own tape walked in reverse.  That keeps gradient accumulation fused into
kernel epilogues (dgrad + residual add), puts the weight gradients straight into one flat fp32 buffer that the
data-parallel all-reduce consumes, and makes the step capturable in a CUDA graph (no host syncs anywhere).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
#: kill switch for A/B measurements (STFB_NO_TCGEN05=1 keeps every conv on the SIMT family)
import os as _os
USE_TCGEN05 = _os.environ.get("STFB_NO_TCGEN05", "0") != "1"
# fp32 mode: convolutions and weight gradients the tcgen05 family has a shape for run there on split-precision operands
# (STFB_BF16X3, csrc/split.cu: three bf16 planes per fp32 tensor, six products per MAC, fp32-accurate) instead of the FFMA family
USE_SPLIT_FP32 = _os.environ.get("STFB_NO_SPLIT_FP32", "0") != "1"
USE_FP32_SIDE_STREAMS = _os.environ.get("STFB_NO_FP32_SIDE_STREAMS", "0") != "1"   # fp32 mode: weight gradients on the side streams too
USE_FUSED_LSTM = _os.environ.get("STFB_NO_FUSED_LSTM", "0") != "1"
# cell backward in the recurrent GEMM's epilogue (stfb_lstm_bwd_step_fused: one launch per backward time step instead of two,
# dh never in memory).  Correct (tests/test_ops_gpu.py) but MEASURED SLOWER in the step, three times: 9.88 ms against 9.59 with a
# row-per-lane epilogue, 9.86 against 9.57 with the coalesced one (dh staged through shared memory, the elementwise kernel's
# 16-lanes-per-row mapping), 9.88 against 9.56 with shallow rings and two CTAs per SM.  The four levels' backward chains run
# concurrently: the separate lstm_cell_bwd launches are light elementwise grids that fill the SMs between the other levels'
# GEMMs, while a fused step moves all of that traffic through four epilogue warps per tensor-core CTA.  OFF by default;
# STFB_FUSED_LSTM_BWD=1 enables it.
USE_FUSED_LSTM_BWD = _os.environ.get("STFB_FUSED_LSTM_BWD", "0") == "1"
USE_LSTM_SEQ = _os.environ.get("STFB_NO_LSTM_SEQ", "0") != "1"      # all-T kernel for 64-unit levels (csrc/conv_tc.cu lstm_seq64_kernel)
USE_FUSED_BN_STATS = _os.environ.get("STFB_NO_FUSED_BN_STATS", "0") != "1"
USE_LSTM_STREAMS = _os.environ.get("STFB_NO_LSTM_STREAMS", "0") != "1"
USE_WGRAD_STREAM = _os.environ.get("STFB_NO_WGRAD_STREAM", "0") != "1"
USE_SIDE_FINALIZE = _os.environ.get("STFB_NO_SIDE_FINALIZE", "0") != "1"
USE_FUSED_STEM = _os.environ.get("STFB_NO_FUSED_STEM", "0") != "1"
# the training step's weight pack (223 us for 27 M parameters) split in two: the stem's weights at once, the rest on a
# low-priority side stream beside the stem's memory-bound kernels.  MEASURED (round 2): 9.547 ms with it, 9.506 ms without --
# the pack competes with the stem for HBM and delays layer 1 by as much as it saves (same finding as round 1, now with the
# mechanism in the tree).  OFF by default; STFB_PACK_OVERLAP=1 enables it.
USE_PACK_OVERLAP = _os.environ.get("STFB_PACK_OVERLAP", "0") == "1"
# Two forward chains: the encoder's time steps are independent until the LSTMs (BatchNorm statistics are per time step), so
# in the training forward every encoder layer is launched as two half-batch launches (time steps [0, T/2) and [T/2, T)) on two
# streams.  Each chain alternates tensor-bound convolutions with HBM-bound BatchNorm passes; the two chains drift out of
# phase and the block scheduler co-runs one chain's BatchNorm pass with the other's convolution (nothing else is
# independent of the main chain in the forward pass).  Tensors stay whole (a half is a leading-dimension view), so the backward
# tape is unchanged.  MEASURED (round 2, graph step at the headline size): 9.87 ms with the split against 9.82 ms without --
# the half-batch launches cost more (109 extra launches, half the tiles per launch) than the overlap returns, because a
# persistent convolution CTA holds 160-200 KB of shared memory and 54 k registers of its SM and leaves the other chain's
# BatchNorm CTAs almost no room to co-reside.  OFF by default; STFB_FWD_SPLIT=1 enables it (kept tested).
USE_FWD_SPLIT = _os.environ.get("STFB_FWD_SPLIT", "0") == "1"
# Stream priorities: the main chain (graph capture stream, LSTM forks, BatchNorm finalize) is HIGH priority and the
# weight-gradient side streams stay at the default (lowest) one, so the block scheduler hands freed SM slots to the critical
# path first and the tensor-bound wgrad CTAs fill what is left.  Measured on the graph step: 10.29 -> 10.08 ms with one wgrad
# stream, 9.89 ms with three (without priorities a second wgrad stream LOSES 0.18 ms).  STFB_NO_STREAM_PRIO=1 disables.
MAIN_PRIO = 0 if _os.environ.get("STFB_NO_STREAM_PRIO", "0") == "1" else -1
WGRAD_STREAMS = int(_os.environ.get("STFB_WGRAD_STREAMS", "3" if MAIN_PRIO else "1"))   # side streams the wgrad launches rotate over


class Var:
    """An NHWC activation plus (during backward) its gradient."""
    __slots__ = ("data", "grad", "needs_grad", "grad_dtype", "bn_partial", "split")

    def __init__(self, data, needs_grad=True, grad_dtype=None):
        self.bn_partial = None     # fused BatchNorm partial sums produced by the conv that wrote `data`
        self.split = None          # split-precision copy of `data` (ops.split_bf16x3), made by the first conv that reads it
        self.data = data
        self.grad = None
        self.needs_grad = needs_grad
        self.grad_dtype = grad_dtype or data.dtype

    def accumulate(self, g):
        if not self.needs_grad:
            return
        if g.dtype != self.grad_dtype:
            g = ops.cast(g, self.grad_dtype)
        if self.grad is None:
            self.grad = g
        else:
            ops.add_(self.grad, g)


def hs_probe(seq, B):
    """A [B, h, w, C] view of the first time slab (shape probe for the recurrent GEMM)."""
    return seq[:B]


class Executor:
    """Runs layers on NHWC activations; when ``record`` is set every layer pushes its backward closure."""

    def __init__(self, params: Dict[str, torch.Tensor], dtype: torch.dtype, train: bool, record: bool,
                 grads: Optional[Dict[str, torch.Tensor]] = None):
        self.params = params
        self.dtype = dtype
        self.train = train
        self.record = record
        self.grads = grads if grads is not None else {}
        self.tape: List = []
        self._packed = {}
        self._new_pack_keys = []      # packs this forward needed that the module's PackPlan did not hold
        # deferred tcgen05 weight gradients: side buffer with the flat gradient's layout + who used it
        self.acc_flat = None
        self.acc: Dict[str, torch.Tensor] = {}
        self.grad_offsets: Dict[str, int] = {}
        self._deferred = {}           # name -> (flat offset, Cp, Cg_total, khw)
        self._pack_join = None        # side stream still packing the late weights (joined by the first consumer)
        self._owner = None
        self.flat = None              # this forward's flat gradient buffer (self.grads are views of it)
        self.segment_hook = None      # callable(flat_slice): gradient range final (data-parallel all-reduce); None = off
        self._seg_end = None
        self._comm_used = False
        self._owner_for_plans = None
        self._wg_keep = []            # (dy, x) pairs the side-stream wgrad launches still read
        self._fin_keep = []           # (slots, coefficients) the side-stream BatchNorm finalize launches still touch

    # ---------------------------------------------------------------------------------------------
    _split_streams = {}

    def begin_split(self, G):
        """Start the two-chain section (see USE_FWD_SPLIT): only in the recording bf16 training forward -- every tensor is
        then kept alive by the tape until the backward pass, so no allocation made on the main stream is recycled while a
        chain stream still reads it."""
        self._split = None
        if not (USE_FWD_SPLIT and self.record and self.train and self.dtype == torch.bfloat16 and G >= 2 and G % 2 == 0):
            return
        cur = torch.cuda.current_stream()
        pool = Executor._split_streams.get(cur.device.index)
        if pool is None:
            pool = Executor._split_streams[cur.device.index] = [torch.cuda.Stream(device=cur.device, priority=MAIN_PRIO) for _ in range(2)]
        if getattr(self, "_stat_arena", None) is None:        # the zero fill of the statistics arena precedes the fork
            self._stat_arena = torch.zeros(self.STAT_ARENA_FLOATS, dtype=torch.float32, device=cur.device)
            self._stat_off = 0
        for st_ in pool:
            st_.wait_stream(cur)
        self._split = pool

    def end_split(self):
        if getattr(self, "_split", None):
            cur = torch.cuda.current_stream()
            for st_ in self._split:
                cur.wait_stream(st_)
        self._split = None

    def _halves(self, N, G):
        """[(stream, n0, n1, g0, g1)] of the two chains, or None when this tensor cannot be halved on a group boundary."""
        sp = getattr(self, "_split", None)
        if not sp or G < 2 or G % 2 or N % G:
            return None
        nh, gh = (N // G) * (G // 2), G // 2
        return [(sp[0], 0, nh, 0, gh), (sp[1], nh, N, gh, G)]

    def _join_late_packs(self, name):
        """The first consumer of a weight that was packed on the side stream (ops.PackPlan.run(early=...)) waits for it; every
        later consumer is ordered behind this point through the streams it forks from."""
        j = self._pack_join
        if j is not None and not (self._owner is not None and self._owner._early_pack is not None and self._owner._early_pack(name)):
            torch.cuda.current_stream().wait_stream(j)
            self._pack_join = None

    def packed(self, name, k_is_dim1, n_major=False, flip=False, kpad=None, gate_c=0):
        key = (name, bool(k_is_dim1), bool(n_major), bool(flip), kpad, int(gate_c))
        self._join_late_packs(name)
        wp = self._packed.get(key)
        if wp is None:
            wp = ops.pack_weight(self.params[name], k_is_dim1, self.dtype, n_major=n_major, flip=flip, kpad=kpad,
                                 gate_c=gate_c)
            self._packed[key] = wp
            self._new_pack_keys.append(key)
            self._split_wait_current()     # a pack launched after the fork (first forward of a mode, before the plan exists)
        return wp

    def packed_lstm_xh(self, wih, whh, C):
        """[W_ih | W_hh] K-concatenated, gate rows chunk-interleaved: the B operand of ops.lstm_step_fused."""
        cat = "xh:" + wih
        k1 = (wih, True, True, False, None, int(C), (cat, 0, 2 * C))
        k2 = (whh, True, True, False, None, int(C), (cat, C, 2 * C))
        self._join_late_packs(wih)
        wp = self._packed.get(k1)
        if wp is None:
            wp = ops.pack_lstm_xh(self.params[wih], self.params[whh], self.dtype)
            self._packed[k1] = self._packed[k2] = wp
            self._new_pack_keys += [k1, k2]
        return wp

    _tc_cache = {}

    def use_tc(self, x, Cout, k, stride, pad, x2=None, mode=ops.CONV_FWD, out_hw=None):
        """Does this convolution go to the tcgen05 family (bf16, 64-channel multiples, stride 1/2, k <= 3)?"""
        if self.dtype != torch.bfloat16 or not USE_TCGEN05:
            return False
        key = (tuple(x.shape), Cout, k, stride, pad, None if x2 is None else x2.shape[3], mode, out_hw)
        r = Executor._tc_cache.get(key)
        if r is None:
            r = ops.tcgen05_ok(x, Cout, k, stride, pad, mode=mode, x2=x2, out_hw=out_hw)
            Executor._tc_cache[key] = r
        return r

    def use_split(self, x, Cout, k, stride, pad, x2=None, mode=ops.CONV_FWD, out_hw=None):
        """fp32 mode: does this convolution run on the tensor cores with split-precision operands?"""
        if self.dtype != torch.float32 or not (USE_SPLIT_FP32 and USE_TCGEN05) or not x.is_cuda:
            return False
        key = ("x3", tuple(x.shape), Cout, k, stride, pad, None if x2 is None else x2.shape[3], mode, out_hw)
        r = Executor._tc_cache.get(key)
        if r is None:
            r = ops.tcgen05_ok(x, Cout, k, stride, pad, mode=mode, x2=x2, out_hw=out_hw, as_split=True)
            Executor._tc_cache[key] = r
        return r

    def split_of(self, v: Var):
        if v.split is None:
            v.split = ops.split_bf16x3(v.data)
        return v.split

    def packed_split(self, name, k_is_dim1):
        """Split-precision B operand of `name` (ops.pack_weight_split); part of the module's batched PackPlan like every pack."""
        key = (name, bool(k_is_dim1), True, False, None, 0, None, "x3")
        self._join_late_packs(name)
        wp = self._packed.get(key)
        if wp is None:
            wp = self._packed[key] = ops.pack_weight_split(self.params[name], k_is_dim1)
            self._new_pack_keys.append(key)
            self._split_wait_current()
        return wp

    _sf_cache = {}

    def stats_fusable(self, x, Cout, k, stride, pad, G, x2=None):
        if not USE_FUSED_BN_STATS:
            return False
        key = (tuple(x.shape), Cout, k, stride, pad, G, None if x2 is None else x2.shape[3])
        r = Executor._sf_cache.get(key)
        if r is None:
            r = ops.conv_stats_fusable(x, Cout, k, stride, pad, G, x2)
            Executor._sf_cache[key] = r
        return r

    STAT_ARENA_FLOATS = 1 << 21        # 8 MB of zeroed fp32 per forward: one fill instead of one per BatchNorm layer

    def _split_wait_current(self):
        """Inside the two-chain section: work just enqueued on the CURRENT stream (a zero fill, a weight pack) that the chain
        streams -- forked earlier -- are about to consume is ordered before them."""
        if getattr(self, "_split", None):
            cur = torch.cuda.current_stream()
            for st_ in self._split:
                if st_ != cur:
                    st_.wait_stream(cur)

    def _fresh_zeros(self, n, device):
        """n zeroed fp32 filled on the current stream (the chain streams accumulate into these buffers with red.add)."""
        z = torch.zeros(n, dtype=torch.float32, device=device)
        self._split_wait_current()
        return z

    def _arena_take(self, n, device):
        arena = getattr(self, "_stat_arena", None)
        if arena is None or self._stat_off + n > arena.numel():
            if n > self.STAT_ARENA_FLOATS:
                return self._fresh_zeros(n, device)
            arena = self._stat_arena = self._fresh_zeros(self.STAT_ARENA_FLOATS, device)
            self._stat_off = 0
        buf = arena[self._stat_off:self._stat_off + n]
        self._stat_off += (n + 63) // 64 * 64
        return buf

    def stat_buffer(self, G, C, device):
        """Zeroed [STAT_SLOTS, 2, G, C] fp32 for the fused BatchNorm statistics, carved from one arena per forward."""
        return self._arena_take(ops.STAT_SLOTS * 2 * G * C, device).view(ops.STAT_SLOTS, 2, G, C)

    def zeroed_scratch(self, n, device):
        """n zeroed fp32 from the per-forward arena (the one-launch BatchNorm backward's group sums + counters)."""
        return self._arena_take(n, device)

    def wants_grad(self, name):
        return name in self.grads

    _wg_streams = {}

    def side_launch(self, keep, fn):
        """Run fn() -- launches whose results only the optimizer reads (weight / bias gradients) -- on one of the
        low-priority side streams forked from the current stream (bf16 path and the split-precision fp32 path; otherwise inline).  `keep`: the tensors the
        launches read, kept alive until the join at the end of the backward pass (their Python owners drop them earlier
        and the caching allocator would hand the memory to the main stream)."""
        if USE_WGRAD_STREAM and keep[0].is_cuda and (self.dtype == torch.bfloat16 or (USE_SPLIT_FP32 and USE_FP32_SIDE_STREAMS)):
            cur = torch.cuda.current_stream()
            pool = Executor._wg_streams.get(cur.device.index)
            if pool is None:
                pool = Executor._wg_streams[cur.device.index] = [torch.cuda.Stream(device=cur.device) for _ in range(max(1, WGRAD_STREAMS))]
            side = pool[len(self._wg_keep) % len(pool)]
            side.wait_stream(cur)
            self._wg_keep.append(keep)
            with torch.cuda.stream(side):
                return fn()
        return fn()

    def wgrad(self, wname, P, G, k, stride, pad, cg_off=0, cg_total=None, split=False):
        """Weight gradient of `wname`; tcgen05 launches accumulate in the side buffer and are folded in once, at the
        end of the backward pass (Executor.backward).

        Nothing downstream of a weight gradient is needed before the optimizer, so (bf16 path) the wgrad launches rotate
        over WGRAD_STREAMS low-priority side streams forked from the current one: the tensor-core bound wgrad kernels (192 threads, 130 KB of shared
        memory per SM) then share the SMs with the memory-bound BatchNorm-backward kernels of the main chain instead of
        alternating with them.  P and G are kept alive until the join at the end of the backward pass."""
        acc = self.acc.get(wname)
        deferred = self.side_launch((P, G), lambda: ops.conv2d_wgrad(P, G, self.grads[wname], k, stride, pad, cg_off, cg_total,
                                                                     acc=acc, split=split))
        if deferred:
            w = self.params[wname]
            khw = 1 if w.dim() == 2 else w.shape[2] * w.shape[3]
            self._deferred[wname] = (self.grad_offsets[wname], w.shape[0], w.shape[1], khw)

    _fin_streams = {}

    def join_finalize(self):
        """End of the forward pass: the side-stream BatchNorm finalize launches (running statistics, saved coefficients)
        are ordered before whatever follows on the current stream."""
        if self._pack_join is not None:           # nobody consumed a late pack (cannot happen in the models here): join anyway
            torch.cuda.current_stream().wait_stream(self._pack_join)
            self._pack_join = None
        if self._fin_keep:
            cur = torch.cuda.current_stream()
            side = Executor._fin_streams.get(cur.device.index)
            if side is not None:
                cur.wait_stream(side)
            self._fin_keep = []

    def join_wgrad(self):
        if self._wg_keep:
            cur = torch.cuda.current_stream()
            for side in Executor._wg_streams.get(cur.device.index) or []:
                cur.wait_stream(side)
            self._wg_keep = []

    # ---------------------------------------------------------------------------------------------
    def conv(self, x: Var, wname: str, *, k: int, stride: int = 1, pad: int = 0, transposed: bool = False,
             out_pad: int = 0, bname: Optional[str] = None, x2: Optional[Var] = None, scale=None, shift=None,
             residual=None, relu: bool = False, y_dtype=None, stats_G: int = 0) -> Var:
        """Conv2d / ConvTranspose2d (+bias)(+folded BN)(+residual)(+ReLU).  The epilogue extras other than the
        bias are inference-only (no backward through them)."""
        w = self.params[wname]
        Cout = w.shape[1] if transposed else w.shape[0]
        N, H, W, C1 = x.data.shape
        C2 = 0 if x2 is None else x2.data.shape[3]
        assert not (transposed and x2 is not None)
        if (self.dtype == torch.bfloat16 and USE_TCGEN05 and not transposed and x2 is None and C1 % 32 != 0
                and C1 * k * k <= 512 and Cout % 32 == 0 and not x.needs_grad and residual is None):
            return self._conv_small_cin(x, wname, k, stride, pad, bname, scale, shift, relu, y_dtype, stats_G)
        mode = ops.CONV_TRANSPOSED if transposed else ops.CONV_FWD
        out_hw = ops.conv_out_hw(H, W, k, stride, pad, transposed, out_pad)
        x2d = None if x2 is None else x2.data
        tc = self.use_tc(x.data, Cout, k, stride, pad, x2d, mode, out_hw)
        x3 = (not tc) and self.use_split(x.data, Cout, k, stride, pad, x2d, mode, out_hw)
        if x3:
            return self._conv_split(x, wname, k, stride, pad, transposed, out_hw, bname, x2, scale, shift, residual, relu)
        wp = self.packed(wname, k_is_dim1=not transposed, n_major=tc)
        bias = self.params[bname] if bname else None
        partial = None
        if stats_G and tc and not transposed and y_dtype in (None, torch.bfloat16) and self.stats_fusable(x.data, Cout, k, stride, pad, stats_G, x2d):
            # train-mode BatchNorm statistics of the output come out of the conv epilogue (no separate pass over y)
            partial = self.stat_buffer(stats_G, Cout, x.data.device)
        halves = self._halves(N, stats_G) if (stats_G and scale is None and residual is None and not relu) else None
        if halves is not None:
            y = torch.empty((N, out_hw[0], out_hw[1], Cout), dtype=y_dtype or x.data.dtype, device=x.data.device)
            plist = None
            if partial is not None:       # one statistics buffer per chain (groups are renumbered inside each half)
                plist = [partial.view(-1)[:partial.numel() // 2].view(ops.STAT_SLOTS, 2, stats_G // 2, Cout),
                         partial.view(-1)[partial.numel() // 2:].view(ops.STAT_SLOTS, 2, stats_G // 2, Cout)]
            for h, (st_, n0, n1, g0, g1) in enumerate(halves):
                with torch.cuda.stream(st_):
                    ops.conv2d(x.data[n0:n1], wp, Cout, k, stride, pad, mode=mode, out_hw=out_hw, x2=None if x2d is None else x2d[n0:n1],
                               bias=bias, y_dtype=y_dtype, out=y[n0:n1], impl=ops.IMPL_TCGEN05 if tc else ops.IMPL_SIMT,
                               stat_partial=None if plist is None else plist[h], stat_groups=(g1 - g0) if plist is not None else 0)
            partial = plist
        else:
            y = ops.conv2d(x.data, wp, Cout, k, stride, pad, mode=mode, out_hw=out_hw, x2=x2d,
                           bias=bias, scale=scale, shift=shift, residual=residual, relu=relu, y_dtype=y_dtype,
                           impl=ops.IMPL_TCGEN05 if tc else ops.IMPL_SIMT, stat_partial=partial, stat_groups=stats_G if partial is not None else 0)
        out = Var(y, grad_dtype=self.dtype)
        out.bn_partial = partial
        if not self.record:
            return out
        assert scale is None and residual is None and not relu, "fused epilogue is inference-only"

        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            rows = dy.shape[0] * dy.shape[1] * dy.shape[2]
            if bname and self.wants_grad(bname):
                self.side_launch((dy,), lambda: ops.colsum(dy, self.grads[bname], rows, Cout))
            if self.wants_grad(wname):
                if not transposed:
                    self.wgrad(wname, dy, x.data, k, stride, pad, 0, C1 + C2)
                    if x2 is not None:
                        self.wgrad(wname, dy, x2.data, k, stride, pad, C1, C1 + C2)
                else:
                    self.wgrad(wname, x.data, dy, k, stride, pad, 0, Cout)
            srcs = [(x, 0, C1)] + ([(x2, C1, C2)] if x2 is not None else [])
            for src, off, csrc in srcs:
                if not src.needs_grad:
                    continue
                mode_d = ops.CONV_FWD if transposed else ops.CONV_TRANSPOSED
                tcd = self.use_tc(dy, csrc, k, stride, pad, None, mode_d, (H, W))
                wpd = self.packed(wname, k_is_dim1=transposed, n_major=tcd)
                # tcgen05 packing is [n = source channel][K]: a source window is a row offset; SIMT: a column offset
                g = ops.conv2d(dy, wpd, csrc, k, stride, pad, mode=mode_d, out_hw=(H, W), residual=src.grad,
                               out=src.grad, y_dtype=src.grad_dtype, ldw=wpd.shape[1] if tcd else C1 + C2,
                               w_offset=off * wpd.shape[1] if tcd else off,
                               impl=ops.IMPL_TCGEN05 if tcd else ops.IMPL_SIMT)
                src.grad = g

        self.tape.append(bwd)
        return out

    def _conv_split(self, x, wname, k, stride, pad, transposed, out_hw, bname, x2, scale, shift, residual, relu):
        """fp32 mode on the tensor cores: forward, dgrad and wgrad of one convolution over split-precision operands
        (include/stfb200.h STFB_BF16X3).  Every fp32 tensor that feeds a GEMM is split once (the activation by its first
        consumer, the output gradient here) and the three-plane copy serves every GEMM that reads it."""
        w = self.params[wname]
        Cout = w.shape[1] if transposed else w.shape[0]
        N, H, W, C1 = x.data.shape
        C2 = 0 if x2 is None else x2.data.shape[3]
        mode = ops.CONV_TRANSPOSED if transposed else ops.CONV_FWD
        xs = self.split_of(x)
        x2s = None if x2 is None else self.split_of(x2)
        wp = self.packed_split(wname, k_is_dim1=not transposed)
        bias = self.params[bname] if bname else None
        y = ops.conv2d(xs, wp, Cout, k, stride, pad, mode=mode, out_hw=out_hw, x2=x2s, bias=bias, scale=scale, shift=shift,
                       residual=residual, relu=relu, impl=ops.IMPL_TCGEN05, split=True)
        out = Var(y, grad_dtype=self.dtype)
        if not self.record:
            return out
        assert scale is None and residual is None and not relu, "fused epilogue is inference-only"
        lib_ok = ops.wgrad_tcgen05_ok

        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            rows = dy.shape[0] * dy.shape[1] * dy.shape[2]
            if bname and self.wants_grad(bname):
                self.side_launch((dy,), lambda: ops.colsum(dy, self.grads[bname], rows, Cout))
            dys = ops.split_bf16x3(dy)
            if self.wants_grad(wname):
                if not transposed:
                    for src, srcs, off in ((x, xs, 0),) + (((x2, x2s, C1),) if x2 is not None else ()):
                        if lib_ok(dys, srcs, k, stride, pad, split=True):
                            self.wgrad(wname, dys, srcs, k, stride, pad, off, C1 + C2, split=True)
                        else:
                            self.wgrad(wname, dy, src.data, k, stride, pad, off, C1 + C2)
                elif lib_ok(xs, dys, k, stride, pad, split=True):
                    self.wgrad(wname, xs, dys, k, stride, pad, 0, Cout, split=True)
                else:
                    self.wgrad(wname, x.data, dy, k, stride, pad, 0, Cout)
            for src, off, csrc in [(x, 0, C1)] + ([(x2, C1, C2)] if x2 is not None else []):
                if not src.needs_grad:
                    continue
                mode_d = ops.CONV_FWD if transposed else ops.CONV_TRANSPOSED
                if self.use_split(dy, csrc, k, stride, pad, None, mode_d, (H, W)):
                    # [n = source channel][K = taps x 6 x Cout]: a source window is a row offset
                    wpd = self.packed_split(wname, k_is_dim1=transposed)
                    g = ops.conv2d(dys, wpd, csrc, k, stride, pad, mode=mode_d, out_hw=(H, W), residual=src.grad, out=src.grad,
                                   w_offset=off * wpd.shape[1], impl=ops.IMPL_TCGEN05, split=True)
                else:
                    wpd = self.packed(wname, k_is_dim1=transposed, n_major=False)
                    g = ops.conv2d(dy, wpd, csrc, k, stride, pad, mode=mode_d, out_hw=(H, W), residual=src.grad, out=src.grad,
                                   y_dtype=src.grad_dtype, ldw=C1 + C2, w_offset=off, impl=ops.IMPL_SIMT)
                src.grad = g

        self.tape.append(bwd)
        return out

    def _conv_small_cin(self, x, wname, k, stride, pad, bname, scale, shift, relu, y_dtype, stats_G=0):
        """Few-input-channel conv (7x7 stem, UNet enc1.0) as im2col (K padded to 64) + 1x1 tcgen05 GEMM; its weight
        gradient is the 1x1 tcgen05 wgrad over the same im2col buffer.  The input never needs a gradient."""
        w = self.params[wname]
        Cout, Cin = w.shape[0], w.shape[1]
        kpad = (Cin * k * k + 63) // 64 * 64
        wp = self.packed(wname, True, n_major=True, kpad=kpad)
        bias = self.params[bname] if bname else None
        N = x.data.shape[0]
        halves = self._halves(N, stats_G) if (stats_G and scale is None and not relu) else None
        if halves is not None:
            Ho, Wo = ops.conv_out_hw(x.data.shape[1], x.data.shape[2], k, stride, pad)
            col = torch.empty((N, Ho, Wo, kpad), dtype=torch.bfloat16, device=x.data.device)
            y = torch.empty((N, Ho, Wo, Cout), dtype=y_dtype or torch.bfloat16, device=x.data.device)
            partial = None
            if y_dtype in (None, torch.bfloat16) and self.stats_fusable(col[halves[0][1]:halves[0][2]], Cout, 1, 1, 0, stats_G // 2):
                whole = self.stat_buffer(stats_G, Cout, col.device)
                partial = [whole.view(-1)[:whole.numel() // 2].view(ops.STAT_SLOTS, 2, stats_G // 2, Cout),
                           whole.view(-1)[whole.numel() // 2:].view(ops.STAT_SLOTS, 2, stats_G // 2, Cout)]
            for h, (st_, n0, n1, g0, g1) in enumerate(halves):
                with torch.cuda.stream(st_):
                    ops.im2col_small(x.data[n0:n1], k, stride, pad, kpad, out=col[n0:n1])
                    ops.conv2d(col[n0:n1], wp, Cout, 1, 1, 0, bias=bias, y_dtype=y_dtype, out=y[n0:n1], impl=ops.IMPL_TCGEN05,
                               stat_partial=None if partial is None else partial[h], stat_groups=(g1 - g0) if partial is not None else 0)
        else:
            col = ops.im2col_small(x.data, k, stride, pad, kpad)
            partial = None
            if stats_G and y_dtype in (None, torch.bfloat16) and self.stats_fusable(col, Cout, 1, 1, 0, stats_G):
                partial = self.stat_buffer(stats_G, Cout, col.device)
            y = ops.conv2d(col, wp, Cout, 1, 1, 0, bias=bias, scale=scale, shift=shift, relu=relu, y_dtype=y_dtype,
                           impl=ops.IMPL_TCGEN05, stat_partial=partial, stat_groups=stats_G if partial is not None else 0)
        out = Var(y, grad_dtype=self.dtype)
        out.bn_partial = partial
        if not self.record:
            return out
        assert scale is None and not relu, "fused epilogue is inference-only"

        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            if bname and self.wants_grad(bname):
                self.side_launch((dy,), lambda: ops.colsum(dy, self.grads[bname], dy.shape[0] * dy.shape[1] * dy.shape[2], Cout))
            if self.wants_grad(wname):
                def small_wgrad():
                    scratch = torch.zeros((Cout, kpad), dtype=torch.float32, device=dy.device)
                    ops.conv2d_wgrad(dy, col, scratch, 1, 1, 0, 0, kpad)
                    ops.unpad_wgrad(self.grads[wname], scratch)
                    self._wg_keep.append((scratch,))
                self.side_launch((dy, col), small_wgrad)

        self.tape.append(bwd)
        return out

    # ---------------------------------------------------------------------------------------------
    def _bn_forward(self, xd, sums, prefix, G, R, C, relu, resd, y_out, st_out):
        """Statistics (unless the conv epilogue produced them) -> finalize -> apply on the CURRENT stream; -> (y, st) with
        st = [scale, shift, mean, invstd] x [G, C].  y_out / st_out: views to fill (two-chain section)."""
        P = self.params
        if sums is None:
            sums = ops.bn_stats(xd, G, R, C)
        st = st_out if st_out is not None else torch.empty((4, G, C), dtype=torch.float32, device=xd.device)
        if USE_SIDE_FINALIZE and sums.shape[0] <= 8 and xd.is_cuda and 2 * C * 4 <= 48 * 1024:
            # few slots (the conv epilogue's): the apply kernel derives scale/shift itself and the finalize launch -- now
            # only the running statistics and the coefficients the backward pass reads -- runs beside the main chain
            cur = torch.cuda.current_stream()
            side = Executor._fin_streams.get(cur.device.index)
            if side is None:
                side = Executor._fin_streams[cur.device.index] = torch.cuda.Stream(device=cur.device, priority=MAIN_PRIO)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ops.bn_finalize_train(sums, P[prefix + ".weight"], P[prefix + ".bias"], P[prefix + ".running_mean"],
                                      P[prefix + ".running_var"], P[prefix + ".num_batches_tracked"], G, R, C, BN_EPS,
                                      BN_MOMENTUM, out=st)
            self._fin_keep.append((sums, st))
            y = ops.bn_apply_from_stats(xd, sums, P[prefix + ".weight"], P[prefix + ".bias"], G, R, C, relu, resd, out=y_out, eps=BN_EPS)
        else:
            ops.bn_finalize_train(sums, P[prefix + ".weight"], P[prefix + ".bias"], P[prefix + ".running_mean"],
                                  P[prefix + ".running_var"], P[prefix + ".num_batches_tracked"], G, R, C, BN_EPS,
                                  BN_MOMENTUM, out=st)
            y = ops.bn_apply(xd, st[0], st[1], G, R, C, relu, resd, out=y_out)
        return y, st

    def bn(self, x: Var, prefix: str, G: int, relu: bool, residual: Optional[Var] = None) -> Var:
        """Train-mode BatchNorm2d over G row groups (+residual)(+ReLU)."""
        N, H, W, C = x.data.shape
        assert N % G == 0
        R = (N // G) * H * W
        P = self.params
        resd = None if residual is None else residual.data
        halves = self._halves(N, G)
        if halves is None:
            assert not isinstance(x.bn_partial, list), "a tensor of the two-chain section reached a whole-batch BatchNorm"
            y, st = self._bn_forward(x.data, x.bn_partial, prefix, G, R, C, relu, resd, None, None)
        else:
            # two chains: each half normalises its own time steps; the running statistics still see the groups in order
            # (the finalize launches of both halves go to ONE side stream, first half first)
            y = torch.empty_like(x.data)
            st = torch.empty((4, G, C), dtype=torch.float32, device=x.data.device)
            plist = x.bn_partial if isinstance(x.bn_partial, list) else [None, None]
            assert x.bn_partial is None or isinstance(x.bn_partial, list)
            for h, (st_, n0, n1, g0, g1) in enumerate(halves):
                with torch.cuda.stream(st_):
                    self._bn_forward(x.data[n0:n1], plist[h], prefix, g1 - g0, R, C, relu, None if resd is None else resd[n0:n1],
                                     y[n0:n1], st[:, g0:g1])
        x.bn_partial = None
        out = Var(y, grad_dtype=self.dtype)
        if not self.record:
            return out
        mean, invstd = st[2], st[3]
        # zeroed now (one arena fill per forward), consumed by the one-launch backward
        # (bf16 path only: the fp32-accurate path keeps the three-launch chain with its fp64 slot sums -- the one-launch
        # kernel meets its group sums through fp32 atomics, which tiny-batch BatchNorm amplifies beyond the fp32 bars)
        scratch = (self.zeroed_scratch(ops.bn_bwd_scratch_floats(G, C), x.data.device)
                   if ops.USE_FUSED_BN_BWD and self.dtype == torch.bfloat16 else None)

        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            gname, bname = prefix + ".weight", prefix + ".bias"
            dgamma = self.grads.get(gname)
            dbeta = self.grads.get(bname)
            want_dres = residual is not None and residual.needs_grad
            dx, dres = ops.bn_bwd(dy, y, x.data, mean, invstd, P[gname], dgamma, dbeta, G, R, C, relu, want_dres,
                                  dres_acc=residual.grad if want_dres else None,
                                  scale=st[0] if residual is None else None, shift=st[1] if residual is None else None,
                                  scratch=scratch)
            x.grad = dx
            if want_dres:
                residual.grad = dres

        self.tape.append(bwd)
        return out

    def bn_relu_pool(self, x: Var, prefix: str, G: int, k: int, stride: int, pad: int) -> Var:
        """Train-mode BatchNorm2d -> ReLU -> MaxPool2d(k, stride, pad) of a raw conv output (the stem).  bf16 with the
        statistics in the conv epilogue's slots: ONE pass, the full-resolution post-BN map is never written (the backward pass
        recomputes the ReLU mask from x and routes the pool gradient by index); otherwise the two-launch route."""
        N, H, W, C = x.data.shape
        sums = x.bn_partial
        split = isinstance(sums, list)
        s0 = sums[0] if split else sums
        fusable = (USE_FUSED_STEM and self.train and s0 is not None and s0.shape[0] <= 8 and x.data.dtype == torch.bfloat16
                   and C % 8 == 0 and N % G == 0 and N <= 65535 and k in (2, 3) and 2 * C * 4 <= 48 * 1024)
        halves = self._halves(N, G) if split else None
        assert not split or halves is not None
        if not fusable:
            return self.maxpool(self.bn(x, prefix, G, True), k, stride, pad)
        x.bn_partial = None
        R = (N // G) * H * W
        P = self.params
        st = torch.empty((4, G, C), dtype=torch.float32, device=x.data.device)
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty((N, Ho, Wo, C), dtype=x.data.dtype, device=x.data.device)
        idx = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=x.data.device)

        def one(xd, sm, Gh, st_view, y_view, idx_view):
            cur = torch.cuda.current_stream()
            side = Executor._fin_streams.get(cur.device.index)
            if side is None:
                side = Executor._fin_streams[cur.device.index] = torch.cuda.Stream(device=cur.device, priority=MAIN_PRIO)
            side.wait_stream(cur)
            with torch.cuda.stream(side):        # running statistics + the coefficients the backward pass reads
                ops.bn_finalize_train(sm, P[prefix + ".weight"], P[prefix + ".bias"], P[prefix + ".running_mean"],
                                      P[prefix + ".running_var"], P[prefix + ".num_batches_tracked"], Gh, R, C, BN_EPS, BN_MOMENTUM,
                                      out=st_view)
            self._fin_keep.append((sm, st))
            ops.bn_relu_maxpool_from_stats(xd, sm, P[prefix + ".weight"], P[prefix + ".bias"], Gh, k, stride, pad, eps=BN_EPS,
                                           out=(y_view, idx_view))

        if halves is None:
            one(x.data, sums, G, st, y, idx)
        else:
            for h, (st_, n0, n1, g0, g1) in enumerate(halves):
                with torch.cuda.stream(st_):
                    one(x.data[n0:n1], sums[h], g1 - g0, st[:, g0:g1], y[n0:n1], idx[n0:n1])
        out = Var(y, grad_dtype=self.dtype)
        if not self.record:
            return out
        scratch = self.zeroed_scratch(ops.bn_bwd_scratch_floats(G, C), x.data.device) if ops.USE_FUSED_BN_BWD else None
        in_shape = (N, H, W, C)

        def bwd():
            dy = out.grad
            out.grad = None
            if dy is None:
                return
            dfull = ops.maxpool_bwd_idx(idx, dy, in_shape, k, stride, pad)      # gradient of the (never stored) post-BN map
            gname, bname = prefix + ".weight", prefix + ".bias"
            dx, _ = ops.bn_bwd(dfull, None, x.data, st[2], st[3], P[gname], self.grads.get(gname), self.grads.get(bname), G, R, C,
                               True, False, scale=st[0], shift=st[1], scratch=scratch)
            x.grad = dx

        self.tape.append(bwd)
        return out

    def conv_bn(self, x: Var, wname: str, bnprefix: str, *, k: int, stride: int = 1, pad: int = 0, relu: bool = True,
                residual: Optional[Var] = None, bname: Optional[str] = None, x2: Optional[Var] = None, G: int = 1) -> Var:
        """conv -> BatchNorm -> (+residual) -> ReLU.  Eval: one fused launch with the BN folded into the epilogue."""
        if self.train:
            raw = self.conv(x, wname, k=k, stride=stride, pad=pad, bname=bname, x2=x2, stats_G=G)
            return self.bn(raw, bnprefix, G, relu, residual)
        P = self.params
        fold = ops.bn_fold_eval(P[bnprefix + ".weight"], P[bnprefix + ".bias"], P[bnprefix + ".running_mean"],
                                P[bnprefix + ".running_var"], BN_EPS)
        assert not self.record, "eval-mode BatchNorm has no backward in this engine"
        return self.conv(x, wname, k=k, stride=stride, pad=pad, bname=bname, x2=x2, scale=fold[0], shift=fold[1],
                         residual=None if residual is None else residual.data, relu=relu)

    # ---------------------------------------------------------------------------------------------
    def maxpool(self, x: Var, k: int, stride: int, pad: int) -> Var:
        if not self.record:
            return Var(ops.maxpool_fwd(x.data, k, stride, pad), grad_dtype=self.dtype)
        y, idx = ops.maxpool_fwd_idx(x.data, k, stride, pad)
        out = Var(y, grad_dtype=self.dtype)
        in_shape = tuple(x.data.shape)

        def bwd():
            dy = out.grad
            out.grad = None
            if dy is not None and x.needs_grad:
                x.accumulate(ops.maxpool_bwd_idx(idx, dy, in_shape, k, stride, pad))
        self.tape.append(bwd)
        return out

    def resize(self, x: Var, Ho: int, Wo: int) -> Var:
        """Bilinear, align_corners=True."""
        N, H, W, C = x.data.shape
        y = ops.bilinear_fwd(x.data, Ho, Wo)
        out = Var(y, grad_dtype=self.dtype)
        if self.record:
            def bwd():
                dy = out.grad
                out.grad = None
                if dy is not None and x.needs_grad:
                    x.accumulate(ops.bilinear_bwd(dy, H, W))
            self.tape.append(bwd)
        return out

    # ---------------------------------------------------------------------------------------------
    def lstm_last(self, seq: Var, prefix: str, T: int) -> Var:
        """Per-pixel nn.LSTM(C, C) over T time-major slabs of `seq` [T*B, h, w, C]; returns h_T [B, h, w, C]."""
        TB, h, w, C = seq.data.shape
        assert TB % T == 0
        B = TB // T
        R = B * h * w
        P = self.params
        dev = seq.data.device
        wih, whh = prefix + ".weight_ih_l0", prefix + ".weight_hh_l0"
        bih, bhh = prefix + ".bias_ih_l0", prefix + ".bias_hh_l0"
        tc = self.use_tc(seq.data, 4 * C, 1, 1, 0) and self.use_tc(hs_probe(seq.data, B), 4 * C, 1, 1, 0)
        impl = ops.IMPL_TCGEN05 if tc else ops.IMPL_SIMT
        keep = self.record
        acts = torch.empty((T, R, 4 * C), dtype=self.dtype, device=dev) if keep else None
        cs = torch.empty((T if keep else 2, R, C), dtype=torch.float32, device=dev)
        hs = torch.empty((T if keep else 2, B, h, w, C), dtype=self.dtype, device=dev)
        fused = tc and C % 64 == 0 and USE_FUSED_LSTM
        if fused and USE_LSTM_SEQ and ops.lstm_seq_supported(T, B, h, w, C):
            # hidden size 64: the whole time loop in ONE kernel -- weights, c and h_{t-1} stay in shared memory, x_t streams
            # through a TMA ring; inference writes nothing but h_T
            wxh = self.packed_lstm_xh(wih, whh, C)
            if keep:
                ops.lstm_seq_fused(seq.data, wxh, P[bih], P[bhh], T, cs, hs, acts)
            else:
                ops.lstm_seq_fused(seq.data, wxh, P[bih], P[bhh], T, None, hs[(T - 1) % 2], None)
        elif fused:
            # one kernel per time step: implicit GEMM over [x_t, h_{t-1}] -> four gates in TMEM -> cell update in the
            # epilogue; the gate pre-activations never exist in memory
            wxh = self.packed_lstm_xh(wih, whh, C)
            xs = seq.data.view(T, B, h, w, C)
            for t in range(T):
                cur = t if keep else t % 2
                prev = (t - 1) if keep else (t - 1) % 2
                ops.lstm_step_fused(xs[t], hs[prev] if t > 0 else None, wxh, P[bih], P[bhh], cs[prev] if t > 0 else None,
                                    cs[cur], hs[cur], acts[t].view(B, h, w, 4 * C) if keep else None)
        else:
            # hoisted input GEMM for all T: gates_x = X W_ih^T + b_ih + b_hh  (fp32 pre-activations)
            gates = ops.conv2d(seq.data, self.packed(wih, True, n_major=tc), 4 * C, 1, 1, 0, bias=P[bih], bias2=P[bhh],
                               y_dtype=torch.float32, impl=impl)
            gates = gates.view(T, B, h, w, 4 * C)
            whh_p = self.packed(whh, True, n_major=tc)
            for t in range(T):
                cur = t if keep else t % 2
                prev = (t - 1) if keep else (t - 1) % 2
                if t > 0:  # gates_t += h_{t-1} W_hh^T   (in place through the residual epilogue)
                    ops.conv2d(hs[prev], whh_p, 4 * C, 1, 1, 0, residual=gates[t], out=gates[t], y_dtype=torch.float32,
                               impl=impl)
                ops.lstm_cell_fwd(gates[t], cs[prev] if t > 0 else None, acts[t] if keep else None, cs[cur], hs[cur], R, C)
        last = (T - 1) if keep else (T - 1) % 2
        out = Var(hs[last], grad_dtype=torch.float32)
        if not self.record:
            return out

        def bwd():
            dh = out.grad
            out.grad = None
            if dh is None:
                return
            dG = torch.empty((T, B, h, w, 4 * C), dtype=self.dtype, device=dev)
            dc = torch.zeros((R, C), dtype=torch.float32, device=dev)
            tcb = self.use_tc(dG[0], C, 1, 1, 0)
            implb = ops.IMPL_TCGEN05 if tcb else ops.IMPL_SIMT
            whh_d = self.packed(whh, False, n_major=tcb)
            if fused and tcb and USE_FUSED_LSTM_BWD:
                # t = T-1 takes the upstream gradient; every earlier step is ONE launch: dG_t W_hh in TMEM, the cell backward of
                # step t-1 in the epilogue (dh never exists in memory)
                ops.lstm_cell_bwd(dh, dc, acts[T - 1], cs[T - 2] if T > 1 else None, cs[T - 1], dG[T - 1], R, C, acts_il=True)
                for t in range(T - 1, 0, -1):
                    ops.lstm_bwd_step_fused(dG[t], whh_d, acts[t - 1].view(B, h, w, 4 * C), cs[t - 2] if t > 1 else None, cs[t - 1], dc,
                                            dG[t - 1])
            else:
                for t in range(T - 1, -1, -1):
                    ops.lstm_cell_bwd(dh, dc, acts[t], cs[t - 1] if t > 0 else None, cs[t], dG[t], R, C, acts_il=fused)
                    if t > 0:
                        dh = ops.conv2d(dG[t], whh_d, C, 1, 1, 0, y_dtype=torch.float32, impl=implb)
            dG_all = dG.view(T * B, h, w, 4 * C)
            if self.wants_grad(wih):
                self.wgrad(wih, dG_all, seq.data, 1, 1, 0)
            if self.wants_grad(whh) and T > 1:
                self.wgrad(whh, dG[1:].reshape((T - 1) * B, h, w, 4 * C), hs[:T - 1].reshape((T - 1) * B, h, w, C), 1, 1, 0)
            # d b_ih == d b_hh == column sums of dG: reduce once, add the (tiny) result into the second bias
            def bias_grads():
                if self.wants_grad(bih):
                    ops.colsum(dG_all, self.grads[bih], T * R, 4 * C)
                    if self.wants_grad(bhh):
                        ops.add_(self.grads[bhh], self.grads[bih])
                elif self.wants_grad(bhh):
                    ops.colsum(dG_all, self.grads[bhh], T * R, 4 * C)
            self.side_launch((dG_all,), bias_grads)        # only the optimizer reads them: off the LSTM chain
            if seq.needs_grad:
                tca = self.use_tc(dG_all, C, 1, 1, 0)
                seq.grad = ops.conv2d(dG_all, self.packed(wih, False, n_major=tca), C, 1, 1, 0, residual=seq.grad,
                                      out=seq.grad, y_dtype=seq.grad_dtype,
                                      impl=ops.IMPL_TCGEN05 if tca else ops.IMPL_SIMT)

        self.tape.append(bwd)
        return out

    # ---------------------------------------------------------------------------------------------
    _side_streams = {}

    def fork_join(self, thunks):
        """Run independent launch sequences on side streams (forked from / joined to the current stream; capturable).
        The four per-pixel LSTMs are sequential chains of small launches (rows = 65536 .. 1024): run one after the other
        the deep levels leave most SMs idle for 16 of the 32 steps."""
        if not USE_LSTM_STREAMS or len(thunks) < 2:
            return [th() for th in thunks]
        cur = torch.cuda.current_stream()
        key = (cur.device.index, len(thunks))
        pool = Executor._side_streams.get(key)
        if pool is None:
            pool = Executor._side_streams[key] = [torch.cuda.Stream(device=cur.device, priority=MAIN_PRIO) for _ in thunks]
        results = []
        for st_, th in zip(pool, thunks):
            st_.wait_stream(cur)
            with torch.cuda.stream(st_):
                results.append(th())
        for st_ in pool:
            cur.wait_stream(st_)
        return results

    def lstm_levels(self, seqs, prefixes, T):
        """lstm_last for every encoder level, concurrently; ONE tape entry runs the four backward chains concurrently."""
        mark = len(self.tape)
        outs = self.fork_join([(lambda sq=sq, pf=pf: self.lstm_last(sq, pf, T)) for sq, pf in zip(seqs, prefixes)])
        if self.record:
            bwds = self.tape[mark:]
            del self.tape[mark:]
            self.tape.append(lambda: self.fork_join(list(reversed(bwds))))
        return outs

    # ---------------------------------------------------------------------------------------------
    # Gradient segments.  Parameters sit in the flat gradient buffer in registration order, which is also the order of
    # their first use in the forward pass; the backward pass walks the tape in reverse, so when it passes the point where
    # parameter `name` was first used, every gradient at flat offset >= offset(name) is final.  A forward pass marks such
    # points (mark_grad_segment); with a segment hook installed (data parallelism) the backward pass then
    # folds the deferred weight gradients of the finished range into the flat buffer and hands the range to the hook ON A
    # COMMUNICATION STREAM, ordered after the main chain and the weight-gradient side streams as of that moment -- the
    # all-reduce of the decoder / LSTM / layer-4 gradients runs under the backward pass of layers 3..1 instead of after it.
    _comm_streams = {}

    def mark_grad_segment(self, first_param):
        # also without a hook (one GPU): the fold of the segment's deferred weight gradients into the flat buffer then runs
        # beside the backward pass instead of as one 175 us launch after it (the step's tail is the stem's serial chain)
        if not self.record or first_param not in self.grad_offsets or (self.segment_hook is None and self.acc_flat is None):
            return
        start = self.grad_offsets[first_param]
        self.tape.append(lambda: self._segment_ready(start))

    def _scatter_range(self, lo, hi, flat_grad, owner):
        """Fold the deferred tcgen05 weight gradients whose flat offset lies in [lo, hi) into the flat gradient."""
        entries = sorted(e for e in self._deferred.values() if lo <= e[0] < hi)
        if not entries or flat_grad is None:
            return
        plans = owner.__dict__.setdefault("_scatter_plans", {}) if owner is not None else {}
        plan = plans.get(tuple(entries))
        if plan is None:
            plan = plans[tuple(entries)] = ops.ScatterPlan(entries, flat_grad.device)
        plan.run(self.acc_flat, flat_grad)
        for k in [k for k, e in self._deferred.items() if lo <= e[0] < hi]:
            del self._deferred[k]

    def _segment_ready(self, start):
        end = self._seg_end if self._seg_end is not None else self.flat.numel()
        if start >= end:
            return
        cur = torch.cuda.current_stream()
        comm = Executor._comm_streams.get(cur.device.index)
        if comm is None:
            comm = Executor._comm_streams[cur.device.index] = torch.cuda.Stream(device=cur.device, priority=MAIN_PRIO)
        comm.wait_stream(cur)
        for side in Executor._wg_streams.get(cur.device.index) or []:
            comm.wait_stream(side)
        with torch.cuda.stream(comm):
            self._scatter_range(start, end, self.flat, self._owner_for_plans)
            if self.segment_hook is not None:
                self.segment_hook(self.flat[start:end])
        self._seg_end = start
        self._comm_used = True

    def backward(self, out: Var, dout, flat_grad=None, owner=None):
        out.grad = dout
        self._owner_for_plans = owner
        self._seg_end = None
        self._comm_used = False
        for fn in reversed(self.tape):
            fn()
        self.join_wgrad()
        end = self._seg_end if self._seg_end is not None else (flat_grad.numel() if flat_grad is not None else 0)
        self._scatter_range(0, end, flat_grad, owner)
        if self.segment_hook is not None and flat_grad is not None and end > 0:
            self.segment_hook(flat_grad[:end])           # what no segment mark covered (the stem and the first layers)
        if self._comm_used:
            cur = torch.cuda.current_stream()
            cur.wait_stream(Executor._comm_streams[cur.device.index])
        self.tape = []
        self._packed = {}


def flat_grads(named_params, with_offsets=False):
    """One flat fp32 buffer with a view per trainable parameter (what the DP all-reduce walks)."""
    named = [(n, p) for n, p in named_params if p.requires_grad]
    if not named:
        return (None, {}, {}) if with_offsets else (None, {})
    total = sum(p.numel() for _, p in named)
    flat = torch.zeros(total, dtype=torch.float32, device=named[0][1].device)
    views, offsets, off = {}, {}, 0
    for n, p in named:
        views[n] = flat[off:off + p.numel()].view(p.shape)
        offsets[n] = off
        off += p.numel()
    return (flat, views, offsets) if with_offsets else (flat, views)


class ModelFunction(torch.autograd.Function):
    """The whole model forward as one autograd node; backward = the executor's tape."""

    @staticmethod
    def forward(ctx, module, x, *params):
        ex, out_var, logits = module._run(x, record=True)
        ctx.ex, ctx.out_var, ctx.module = ex, out_var, module
        ctx.names = module._trainable_names
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ex, module = ctx.ex, ctx.module
        if ex is None:
            raise RuntimeError("stf_unet_b200: backward through the same forward twice (the tape is consumed by the first one)")
        flat = ex.flat                     # THIS forward's gradient buffer (another forward may have run since)
        d = ops.nchw_to_nhwc(dlogits.contiguous().float(), ex.dtype)
        ex.backward(ctx.out_var, d, flat_grad=flat, owner=module)
        # Gradients live in ONE flat fp32 buffer; each parameter's .grad is a view of it, assigned directly so that autograd
        # does not clone 110 MB per step and the data-parallel all-reduce / the flat optimizer work on the flat buffer.
        # module._last_flat_grad is always the buffer the parameters' .grad VIEW: a backward that finds gradients already
        # there (gradient accumulation, zero_grad(set_to_none=False), a second forward/backward before the optimizer step)
        # adds into that buffer, as autograd would, instead of re-pointing the optimizer at a fresh one.
        params = module._param_cache()
        names = ctx.names
        cur = getattr(module, "_last_flat_grad", None)
        have = [params[n].grad is not None for n in names]
        if not any(have):
            for n in names:
                params[n].grad = ex.grads[n]
            module._last_flat_grad = flat
        elif all(have) and cur is not None and cur is not flat and _views_of(cur, [params[n].grad for n in names], ex.grad_offsets, names):
            ops.add_(cur, flat)            # one launch: both buffers share the layout
        else:
            for n in names:
                p = params[n]
                if p.grad is None:
                    p.grad = ex.grads[n].clone()
                elif p.grad.is_contiguous() and p.grad.dtype == torch.float32:
                    ops.add_(p.grad, ex.grads[n])
                else:
                    p.grad.add_(ex.grads[n])
            # the parameters' gradients no longer form one flat buffer this engine knows: flat consumers must say so
            if cur is None or not _views_of(cur, [params[n].grad for n in names], ex.grad_offsets, names):
                module._last_flat_grad = None
        hook = getattr(module, "_grad_ready_hook", None)
        if hook is not None and ex.segment_hook is None:
            hook(module._last_flat_grad)
        elif ex.segment_hook is not None and module._last_flat_grad is not flat:
            # gradient accumulation into an older buffer: the segments reduced THIS step's buffer before it was added in;
            # averaging is linear, so the accumulated buffer is consistent across ranks without another exchange
            pass
        module._learn_pack_plan(ex)
        ctx.ex = ctx.out_var = None
        return (None, None) + (None,) * len(ctx.names)


def _views_of(flat, grads, offsets, names):
    """Are `grads` exactly the per-parameter views of `flat` (same storage, the engine's offsets)?"""
    base = flat.data_ptr()
    for n, g in zip(names, grads):
        if g is None or g.dtype != torch.float32 or g.data_ptr() != base + 4 * offsets[n]:
            return False
    return True
