"""criterion = cross-entropy + softmax-Dice as one fused kernel pair.

Drop-in for ``criterion`` (/root/reference/train_utils/train_and_eval.py:299-313) and ``dice_loss`` /
``build_target`` (/root/reference/train_utils/dice_coefficient_loss.py:5-55), full signature: class weights
(``loss_weight`` -> ``F.cross_entropy(weight=...)``), ``ignore_index`` (pixels with that label are left out of the
cross-entropy mean and of every per-image Dice sum; the reference's ``collate_fn`` pads targets with 255,
my_dataset.py:243) and ``dice=False``.  A label outside ``[0, C)`` that is not ``ignore_index`` makes the reference raise
(``IndexError`` in ``one_hot`` / a device assert in ``cross_entropy``); here the loss comes out NaN without a host
synchronisation -- set ``STFB_CHECK_TARGETS=1`` to validate on the host (one sync per call) and raise instead.
"""
from __future__ import annotations

import os

import torch

from . import ops


class _CEDice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, eps, weight, ignore_index, dice):
        logits = logits.contiguous().float()
        target = target.contiguous()
        out, stats = ops.ce_dice_fwd(logits, target, eps, weight, ignore_index, dice)
        ctx.save_for_backward(logits, target, stats)
        ctx.eps, ctx.weight, ctx.ignore_index, ctx.dice = eps, weight, ignore_index, dice
        return out  # [total, ce, dice_loss]

    @staticmethod
    def backward(ctx, dout):
        logits, target, stats = ctx.saved_tensors
        # only d(total) is supported as an upstream gradient: out[1:] are reporting values
        dloss = dout[0:1].contiguous().float()
        dl = ops.ce_dice_bwd(logits, target, stats, dloss, ctx.eps, ctx.weight, ctx.ignore_index, ctx.dice)
        return dl, None, None, None, None, None


def ce_dice(logits, target, eps=1e-6, weight=None, ignore_index=-100, dice=True):
    """-> tensor [3] = {CE (+ Dice), CE, Dice loss}; differentiable through element 0."""
    if target.dtype != torch.int64:
        raise TypeError("target must be int64 class indices")
    if logits.dim() != 4 or target.shape != (logits.shape[0], logits.shape[2], logits.shape[3]):
        raise ValueError(f"criterion: logits {tuple(logits.shape)} vs target {tuple(target.shape)} size mismatch")
    if weight is not None:
        if weight.numel() != logits.shape[1]:
            raise ValueError(f"criterion: loss_weight has {weight.numel()} entries for {logits.shape[1]} classes")
        weight = weight.detach().to(device=logits.device, dtype=torch.float32).contiguous()
    if os.environ.get("STFB_CHECK_TARGETS", "0") == "1":
        bad = (target != ignore_index) & ((target < 0) | (target >= logits.shape[1]))
        if bool(bad.any()):
            raise IndexError(f"criterion: target holds labels outside [0, {logits.shape[1]}) other than ignore_index={ignore_index}")
    return _CEDice.apply(logits, target, eps, weight, int(ignore_index), bool(dice))


def criterion(inputs, target, loss_weight=None, num_classes: int = 2, dice: bool = True, ignore_index: int = -100):
    """Same signature and semantics as the reference's criterion; ``inputs`` is the model's ``{'out': logits}`` dict."""
    losses = {}
    for name, x in inputs.items():
        if x.shape[1] != num_classes and dice:
            raise ValueError(f"criterion: {x.shape[1]} logit channels for num_classes={num_classes}")
        losses[name] = ce_dice(x, target, weight=loss_weight, ignore_index=ignore_index, dice=dice)[0]
    if len(losses) == 1:
        return losses["out"]
    return losses["out"] + 0.5 * losses["aux"]


def build_target(target, num_classes: int = 2, ignore_index: int = -100):
    """One-hot [N, C, H, W] float target with ignore_index kept on ignored pixels (dice_coefficient_loss.py:5-17).  The fused
    criterion never materialises it (4x the target bytes); provided for callers that use the reference's helper directly."""
    t = target.clone()
    if ignore_index >= 0:
        m = target == ignore_index
        t[m] = 0
        oh = torch.nn.functional.one_hot(t, num_classes).float()
        oh[m] = ignore_index
    else:
        oh = torch.nn.functional.one_hot(t, num_classes).float()
    return oh.permute(0, 3, 1, 2)


def dice_loss(x, target_onehot_or_index, multiclass=True, ignore_index=-100):
    """Dice part alone (reference dice_coefficient_loss.py:51-55); accepts index or one-hot [N,C,H,W] targets (a one-hot
    target built by build_target carries ignore_index on ignored pixels in every channel)."""
    t = target_onehot_or_index
    if t.dim() == 4:
        ign = (t[:, 0] == ignore_index) if ignore_index >= 0 else None
        t = t.argmax(dim=1)
        if ign is not None:
            t = torch.where(ign, torch.full_like(t, ignore_index), t)
    return ce_dice(x, t.long(), ignore_index=ignore_index)[2]
