"""AdamW over flat buffers -- the step after backward.

The reference builds ``torch.optim.AdamW(params, lr, weight_decay, fused=True)`` and steps a ``LambdaLR`` schedule
every iteration (/root/reference/train.py:227-247, train_utils/train_and_eval.py:404-407).  ``FlatAdamW`` keeps that
surface -- it IS a ``torch.optim.Optimizer`` (``param_groups[0]['lr']`` is what the scheduler writes, ``state_dict`` /
``load_state_dict`` round-trip) -- but re-homes every trainable parameter of the model as a view of ONE flat fp32
buffer, in the order of the engine's flat gradient buffer, so that a step is one kernel launch
(``stfb_adamw_flat``: 28 bytes per parameter) fed directly by the (all-reduced) flat gradient, instead of torch's
multi-tensor-apply over ~190 tensors.

Flatten BEFORE capturing a CUDA graph of the step: the graph bakes parameter addresses in.
"""
from __future__ import annotations

import torch

from . import ops


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("FlatAdamW: the model has no trainable parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW runs on CUDA parameters only (no CPU fallback); call model.to('cuda') first")
        for n, p in named:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError(f"FlatAdamW: parameter {n} must be fp32 on {dev}")
        total = sum(p.numel() for _, p in named)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for _, p in named:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                      # the module now reads its weights straight from the flat buffer
                off += p.numel()
        model.__dict__.pop("_pack_plans", None)    # cached packing plans hold the old addresses
        model.__dict__.pop("_pcache", None)
        self.model = model
        self.flat_param = flat
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.steps = 0
        self.grad_scale = 1.0                      # set to 1/world_size when the all-reduce sums instead of averaging
        super().__init__([p for _, p in named], dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _flat_grad(self):
        g = getattr(self.model, "_last_flat_grad", None)
        if g is None or g.numel() != self.flat_param.numel():
            raise RuntimeError("FlatAdamW.step(): no flat gradient buffer -- run a forward/backward of the model first (or the "
                               "parameters' .grad tensors were replaced by something other than the engine's flat views)")
        # the flat buffer must be what the parameters' .grad view (engine.ModelFunction keeps that invariant; a foreign
        # optimizer / user code that re-assigns .grad breaks it): checked on the first and last parameter, every step
        first, last = self.param_groups[0]["params"][0], self.param_groups[0]["params"][-1]
        if (first.grad is None or first.grad.data_ptr() != g.data_ptr()
                or last.grad is None or last.grad.data_ptr() != g.data_ptr() + 4 * (g.numel() - last.numel())):
            raise RuntimeError("FlatAdamW.step(): the parameters' .grad are not views of the model's flat gradient buffer "
                               "(zero_grad(set_to_none=True) without a following backward, or gradients assigned by hand)")
        return g

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grp = self.param_groups[0]
        self.steps += 1
        ops.adamw_flat_(self.flat_param, self._flat_grad(), self.exp_avg, self.exp_avg_sq, grp["lr"], grp["betas"][0],
                        grp["betas"][1], grp["eps"], grp["weight_decay"], self.steps, self.grad_scale)
        # the kernel writes through raw pointers (no Parameter._version bump): tell the module its cached eval-mode weight
        # packs are stale
        self.model.__dict__["_pack_epoch"] = self.model.__dict__.get("_pack_epoch", 0) + 1
        return loss

    # ---- checkpointing: one flat state instead of per-parameter dicts ------------------------------------------
    def state_dict(self):
        return {"steps": self.steps, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update(saved)
