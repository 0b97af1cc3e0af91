"""criterion = cross-entropy + softmax-Dice as one fused kernel pair.

Drop-in for ``criterion`` (/root/reference/train_utils/train_and_eval.py:299-313) and ``dice_loss`` /
``build_target`` (/root/reference/train_utils/dice_coefficient_loss.py:5-55) on the configuration the reference
trains with: ``loss_weight=None``, ``dice=True``, ``ignore_index=-100`` (so no pixel is ever ignored).
"""
from __future__ import annotations

import torch

from . import ops


class _CEDice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, eps):
        logits = logits.contiguous().float()
        target = target.contiguous()
        out, stats = ops.ce_dice_fwd(logits, target, eps)
        ctx.save_for_backward(logits, target, stats)
        ctx.eps = eps
        return out  # [total, ce, dice_loss]

    @staticmethod
    def backward(ctx, dout):
        logits, target, stats = ctx.saved_tensors
        # only d(total) is supported as an upstream gradient: out[1:] are reporting values
        dloss = dout[0:1].contiguous().float()
        return ops.ce_dice_bwd(logits, target, stats, dloss, ctx.eps), None, None


def ce_dice(logits, target, eps=1e-6):
    """-> tensor [3] = {CE + Dice, CE, Dice loss}; differentiable through element 0."""
    if target.dtype != torch.int64:
        raise TypeError("target must be int64 class indices")
    if logits.dim() != 4 or target.shape != (logits.shape[0], logits.shape[2], logits.shape[3]):
        raise ValueError(f"criterion: logits {tuple(logits.shape)} vs target {tuple(target.shape)} size mismatch")
    return _CEDice.apply(logits, target, eps)


def criterion(inputs, target, loss_weight=None, num_classes: int = 2, dice: bool = True, ignore_index: int = -100):
    """Same signature as the reference's criterion; ``inputs`` is the model's ``{'out': logits}`` dict."""
    if loss_weight is not None or not dice or ignore_index >= 0:
        raise NotImplementedError("stf_unet_b200.criterion implements the reference's training configuration only "
                                  "(loss_weight=None, dice=True, ignore_index<0)")
    losses = {name: ce_dice(x, target)[0] for name, x in inputs.items()}
    if len(losses) == 1:
        return losses["out"]
    return losses["out"] + 0.5 * losses["aux"]


def dice_loss(x, target_onehot_or_index, multiclass=True, ignore_index=-100):
    """Dice part alone (reference dice_coefficient_loss.py:51-55); accepts index or one-hot [N,C,H,W] targets."""
    t = target_onehot_or_index
    if t.dim() == 4:
        t = t.argmax(dim=1)
    return ce_dice(x, t.long())[2]
