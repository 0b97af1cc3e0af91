"""Device-side evaluation metrics -- drop-in for ``ConfusionMatrix`` and ``DiceCoefficient``
(/root/reference/train_utils/train_and_eval.py:25-70, :73-132) as used by ``evaluate`` (:322-336).

The reference does ``output.argmax(1)``, a masked ``bincount`` and a Python loop over classes per batch (with host syncs
at ``if union > 0``).  Here both metrics are updated by ONE kernel pass over the logits (``stfb_eval_metrics``), which also
emits the uint8 tumour mask of the batch (BASELINE.json configs[4], whole-volume inference); nothing leaves the device
until ``compute()``.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check


def _stream():
    return torch.cuda.current_stream().cuda_stream


class EvalMetrics:
    """Confusion matrix + Dice coefficient of a stream of batches, updated together."""

    def __init__(self, num_classes=2, ignore_index=255, device="cuda"):
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        dev = torch.device(device)
        self.mat = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=dev)
        self._counts = torch.zeros((3 * num_classes,), dtype=torch.int64, device=dev)
        self.cumulative_dice = torch.zeros((num_classes,), dtype=torch.float32, device=dev)
        self._updates = torch.zeros((), dtype=torch.int64, device=dev)

    def update(self, output, target, want_mask=False):
        """output: logits [B,C,H,W] (or the model's {'out': logits}); target int64 [B,H,W].  Returns the uint8 argmax mask
        [B,H,W] when want_mask is set."""
        if isinstance(output, dict):
            output = output["out"]
        if not output.is_cuda:
            raise RuntimeError("stf_unet_b200.metrics runs on CUDA tensors only (no CPU fallback)")
        logits = output.contiguous().float()
        target = target.contiguous()
        if target.dtype != torch.int64:
            raise TypeError("target must be int64 class indices")
        B, C, H, W = logits.shape
        if C != self.num_classes or target.shape != (B, H, W):
            raise ValueError(f"metrics: logits {tuple(logits.shape)} vs target {tuple(target.shape)} / {self.num_classes} classes")
        mask = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device) if want_mask else None
        check(_lib.load().stfb_eval_metrics(logits.data_ptr(), target.data_ptr(), None if mask is None else mask.data_ptr(),
                                            self.mat.data_ptr(), self._counts.data_ptr(), self.cumulative_dice.data_ptr(),
                                            self._updates.data_ptr(), B, C, H * W,
                                            -1 if self.ignore_index is None else int(self.ignore_index),
                                            int(self.ignore_index is not None), _stream()), "eval_metrics")
        return mask

    # ---- ConfusionMatrix.compute (:44-49) ----
    def compute_confusion(self):
        h = self.mat.float()
        acc_global = torch.diag(h).sum() / h.sum()
        acc = torch.diag(h) / h.sum(1)
        iu = torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))
        return acc_global, acc, iu

    # ---- DiceCoefficient.compute (:120-123) ----
    def compute_dice(self):
        n = int(self._updates.item())
        if n == 0:
            return torch.tensor(0.0)
        return self.cumulative_dice / n

    def reset(self):
        self.mat.zero_()
        self._counts.zero_()
        self.cumulative_dice.zero_()
        self._updates.zero_()

    def reduce_from_all_processes(self):
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        torch.distributed.barrier()
        torch.distributed.all_reduce(self.mat)
        torch.distributed.all_reduce(self.cumulative_dice)
        torch.distributed.all_reduce(self._updates)


class ConfusionMatrix:
    """Reference-shaped wrapper: ``update(target.flatten(), pred.flatten())`` is replaced by ``update_logits``; the
    index form is kept for callers that already hold predictions."""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.mat = None
        self._m = None

    def update_logits(self, output, target):
        if self._m is None:
            self._m = EvalMetrics(self.num_classes, ignore_index=None, device=target.device)
            self.mat = self._m.mat
        self._m.update(output, target)

    def compute(self):
        return self._m.compute_confusion()

    def reset(self):
        if self._m is not None:
            self._m.reset()


def argmax_mask(output):
    """uint8 [B,H,W] argmax of logits [B,C,H,W] (whole-volume tumour masks, configs[4])."""
    if isinstance(output, dict):
        output = output["out"]
    B, C, H, W = output.shape
    m = EvalMetrics(C, ignore_index=None, device=output.device)
    dummy = torch.zeros((B, H, W), dtype=torch.int64, device=output.device)
    return m.update(output, dummy, want_mask=True)
