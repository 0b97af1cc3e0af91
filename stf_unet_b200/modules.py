"""Parameter containers + the shared nn.Module base of the B200 models.

The classes only *hold* parameters under the reference's ``state_dict`` names; no torch
operator is ever called on them -- the forward pass is ``engine.Executor`` launching libstfb200 kernels.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn

from . import engine, ops


class ConvParams(nn.Module):
    """weight [Cout, Cin, k, k] (+ bias) -- nn.Conv2d's parameters and default init."""

    def __init__(self, cin, cout, k, bias, transposed=False, resnet_init=False):
        super().__init__()
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        if resnet_init:   # torchvision resnet: kaiming_normal_(mode="fan_out", nonlinearity="relu")
            nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")
        else:             # nn.Conv2d / nn.ConvTranspose2d default
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            fan_in = shape[1] * k * k
            bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)


class BNParams(nn.Module):
    """nn.BatchNorm2d's parameters and buffers."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class LSTMParams(nn.Module):
    """nn.LSTM(C, C, batch_first=True) single-layer parameters, gate order i,f,g,o."""

    def __init__(self, c):
        super().__init__()
        k = 1.0 / math.sqrt(c)
        self.weight_ih_l0 = nn.Parameter(torch.empty(4 * c, c).uniform_(-k, k))
        self.weight_hh_l0 = nn.Parameter(torch.empty(4 * c, c).uniform_(-k, k))
        self.bias_ih_l0 = nn.Parameter(torch.empty(4 * c).uniform_(-k, k))
        self.bias_hh_l0 = nn.Parameter(torch.empty(4 * c).uniform_(-k, k))


def seq(**children):
    """ModuleDict whose keys are the reference's nn.Sequential indices (gaps where ReLU sat)."""
    return nn.ModuleDict(children)


class B200Module(nn.Module):
    """Shared forward plumbing: precision selection, tape recording, the single autograd node."""

    #: None = follow torch autocast (fp32 unless autocast is on); or force torch.float32 / torch.bfloat16
    compute_dtype = None
    _warned_fp16 = False
    _warned_eval_grad = False

    def _select_dtype(self):
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            if dt == torch.float16 and not B200Module._warned_fp16:
                warnings.warn("stf_unet_b200: fp16 autocast requested; the B200 kernels run bf16 (same tensor-core "
                              "rate, no loss scaling needed). Call torch.set_autocast_dtype('cuda', torch.bfloat16) "
                              "to silence this.")
                B200Module._warned_fp16 = True
            return torch.bfloat16
        return torch.float32

    def _state(self):
        params = {n: p.data for n, p in self.named_parameters()}
        params.update({n: b for n, b in self.named_buffers()})
        return params

    def _make_executor(self, record):
        named = list(self.named_parameters())
        for n, p in named:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"parameter {n} must be a contiguous fp32 CUDA tensor (stf_unet_b200 has no CPU "
                                   "fallback; call model.to('cuda'))")
        grads, offsets, flat = {}, {}, None
        if record:
            for n, p in named:
                if p.requires_grad and (p._backward_hooks or getattr(p, "_post_accumulate_grad_hooks", None)):
                    raise RuntimeError(
                        f"parameter {n} carries autograd hooks (torch DistributedDataParallel / register_hook): the whole "
                        "model is ONE autograd node here and writes .grad itself, so per-parameter hooks never fire. Use "
                        "stf_unet_b200.parallel.DataParallel for data parallelism.")
            flat, grads, offsets = engine.flat_grads(named, with_offsets=True)
            self._trainable_names = [n for n, p in named if p.requires_grad]
        ex = engine.Executor(self._state(), self._select_dtype(), self.training, record, grads)
        ex.flat = flat
        ex.grad_offsets = offsets
        if record:
            # data parallelism with overlapped exchange: a per-range hook (parallel.DataParallel installs it)
            ex.segment_hook = self.__dict__.get("_grad_segment_hook")
        if record and flat is not None and (ex.dtype == torch.bfloat16 or (engine.USE_SPLIT_FP32 and engine.USE_TCGEN05)):
            # side buffer (same offsets as the flat gradient) the tcgen05 weight-gradient kernels accumulate into
            ex.acc_flat = torch.zeros_like(flat)
            ex.acc = {n: ex.acc_flat[o:o + self._numel(n)] for n, o in offsets.items() if self._is_matrix(n)}
        # one batched launch re-packs every weight this mode needs (plan learned on the first forward of the mode)
        plans = self.__dict__.setdefault("_pack_plans", {})
        plan = plans.get((ex.dtype, self.training, record))
        if plan is not None and plan.valid_for(ex.params, ex.dtype):
            # inference (eval mode, no tape): the fp32 masters do not move between forwards, so the packed operands are
            # reused until a parameter changes -- torch-side updates bump Parameter._version, FlatAdamW (a raw-pointer
            # kernel) bumps the module's _pack_epoch.  0.23 ms of a 2.7 ms inference step.  A CUDA graph captured in this
            # mode therefore bakes the weights of capture time in (a training-mode graph re-packs on every replay).
            stamp = None
            if not self.training and not record:
                stamp = (self.__dict__.get("_pack_epoch", 0), tuple(p._version for _, p in named))
            if stamp is not None and getattr(plan, "stamp", None) == stamp:
                ex._packed = dict(plan.buffers)
            else:
                # training step: only the stem's weights are needed at once; everything else is packed on a low-priority side
                # stream beside the stem's memory-bound kernels and joined at the first layer that needs it (engine.packed)
                early = self._early_pack if (self.training and record and engine.USE_PACK_OVERLAP) else None
                ex._packed, ex._pack_join = plan.run(early)
                plan.stamp = stamp
        else:
            plans.pop((ex.dtype, self.training, record), None)
        return ex

    #: parameters whose packed copies the forward pass needs first (subclasses override); None = no early / late split
    _early_pack = None

    def _numel(self, name):
        return self._param_cache()[name].numel()

    def _is_matrix(self, name):
        return self._param_cache()[name].dim() in (2, 4)

    def _param_cache(self):
        pc = self.__dict__.get("_pcache")
        if pc is None:
            pc = dict(self.named_parameters())
            self.__dict__["_pcache"] = pc
        return pc

    def _forward_impl(self, ex, x):
        raise NotImplementedError

    def _run(self, x, record):
        ex = self._make_executor(record)
        ex._mode_key = (ex.dtype, self.training, record)
        ex._owner = self
        head = self._forward_impl(ex, x)
        ex.join_finalize()
        logits = ops.nhwc_to_nchw(head.data)
        if not record:
            self._learn_pack_plan(ex)
        return ex, head, logits

    def _learn_pack_plan(self, ex):
        """After a forward (and backward, when recording) that had to pack weights one by one: build the batched plan."""
        if ex._new_pack_keys:
            plans = self.__dict__.setdefault("_pack_plans", {})
            old = plans.get(ex._mode_key)
            keys = (old.keys if old is not None else []) + ex._new_pack_keys
            plans[ex._mode_key] = ops.PackPlan(ex.params, keys, ex.dtype)
            ex._new_pack_keys = []

    def _call(self, x):
        if not x.is_cuda:
            raise RuntimeError("stf_unet_b200 models run on CUDA (sm_100a) only; there is no CPU fallback")
        x = x.contiguous() if x.dtype == torch.uint8 else x.contiguous().float()   # uint8: raw grey levels (STFLSTMUNet)
        trainable = [p for p in self.parameters() if p.requires_grad]
        if torch.is_grad_enabled() and self.training and trainable:
            logits = engine.ModelFunction.apply(self, x, *trainable)
        else:
            if torch.is_grad_enabled() and trainable and not self.training and not B200Module._warned_eval_grad:
                # eval mode folds every BatchNorm into the conv epilogues and records no tape: the result cannot be
                # differentiated.  The reference's evaluate() runs under torch.no_grad() (train_and_eval.py:322); say so once
                # instead of handing back a tensor that silently carries no graph.
                warnings.warn("stf_unet_b200: eval-mode forward with gradients enabled returns a NON-differentiable result "
                              "(inference path: folded BatchNorm, no tape). Use torch.no_grad() for evaluation or model.train() "
                              "to differentiate.")
                B200Module._warned_eval_grad = True
            _, _, logits = self._run(x, record=False)
        return {"out": logits}
