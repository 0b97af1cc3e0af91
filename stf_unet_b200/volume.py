"""Whole-volume tumour-mask inference (BASELINE.json configs[4]; SURVEY.md section 8(e) "inference partitioning").

The reference walks the test set one slice per ``model(inputs)`` call (``DataLoader(batch_size=1)``,
/root/reference/test.py:156-175; its overlay images threshold ``sigmoid`` of channel 0) and scores it with the argmax
masks of ``evaluate`` (train_utils/train_and_eval.py:322-336).  The product here is the argmax mask of every slice as
uint8.  A case is a ``[S, T, 1, H, W]`` series in (pinned) host memory; its
slices are independent, so rank r of N takes one contiguous range of slices (no collective on the data path) and
pushes it through the eval-mode forward in fixed-size batches:

    host series --H2D (copy stream, double buffered)--> CUDA-graph replay (forward + fused argmax-mask kernel)
                --D2H--> uint8 masks [S, H/2, W/2] in pinned host memory

The last batch of a range is padded with its own first slices (their masks are discarded), so ONE captured graph
serves every batch.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .metrics import EvalMetrics


def slice_range(num_slices: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of rank `rank`: sizes differ by at most one, earlier ranks take the larger shares."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"slice_range: bad rank {rank} of {world}")
    base, extra = divmod(num_slices, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class VolumePredictor:
    """masks = predictor(series): argmax masks (uint8) of every slice in this rank's share of a case."""

    def __init__(self, model, slice_shape, batch: int = 16, autocast_dtype: Optional[torch.dtype] = torch.bfloat16,
                 num_classes: int = 2, device=None):
        T, C, H, W = slice_shape
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("VolumePredictor: the model must live on a CUDA device (no CPU fallback)")
        self.model, self.batch, self.dev = model.eval(), int(batch), dev
        self.slice_shape = (T, C, H, W)
        self.xbuf = [torch.zeros((batch, T, C, H, W), device=dev), torch.zeros((batch, T, C, H, W), device=dev)]
        self.x = torch.zeros((batch, T, C, H, W), device=dev)                 # the captured graph's input
        self._dummy_t = None                                                  # all-zero target of the metrics kernel (unused counts)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._met = EvalMetrics(num_classes, ignore_index=255, device=dev)
        self._ac = autocast_dtype

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None):
                out = self.model(self.x)["out"]
            if self._dummy_t is None:
                self._dummy_t = torch.zeros((out.shape[0], out.shape[2], out.shape[3]), dtype=torch.long, device=dev)
            return self._met.update(out, self._dummy_t, want_mask=True)     # argmax fused with the (ignored) metric counts

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                fwd()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.mask = fwd()                                                  # [batch, h, w] uint8, static buffer
        self.launches_per_replay = _lib.launch_count() - n0
        self.replays = 0

    def __call__(self, series: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """series: [S, T, C, H, W] float32 on the host (pinned for asynchronous copies) -> uint8 [S, h, w] on the host."""
        if series.is_cuda:
            raise ValueError("VolumePredictor: pass the host series; batches are staged to the device here")
        if tuple(series.shape[1:]) != self.slice_shape:
            raise ValueError(f"VolumePredictor: slices are {tuple(series.shape[1:])}, the graph was captured for {self.slice_shape}")
        S, B = series.shape[0], self.batch
        h, w = self.mask.shape[1], self.mask.shape[2]
        if out is None:
            out = torch.empty((S, h, w), dtype=torch.uint8).pin_memory()
        if S == 0:
            return out
        cur = torch.cuda.current_stream(self.dev)
        nb = (S + B - 1) // B
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def stage(i):
            slot = i & 1
            lo, hi = i * B, min(S, (i + 1) * B)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(consumed[slot])                   # the replay that read this slot last is done
                self.xbuf[slot][:hi - lo].copy_(series[lo:hi], non_blocking=True)
                if hi - lo < B:                                                # pad the ragged tail with its own slices
                    self.xbuf[slot][hi - lo:].copy_(self.xbuf[slot][:1].expand(B - (hi - lo), *self.slice_shape))
                ready[slot].record(self.copy_stream)

        consumed[0].record(cur)
        consumed[1].record(cur)
        stage(0)
        for i in range(nb):
            slot = i & 1
            if i + 1 < nb:
                stage(i + 1)                                                   # overlaps with this batch's forward
            cur.wait_event(ready[slot])
            self.x.copy_(self.xbuf[slot], non_blocking=True)
            consumed[slot].record(cur)
            self.graph.replay()
            self.replays += 1
            lo, hi = i * B, min(S, (i + 1) * B)
            out[lo:hi].copy_(self.mask[:hi - lo], non_blocking=True)
        cur.synchronize()
        return out


def predict_volume(model, series: torch.Tensor, batch: int = 16, rank: int = 0, world: int = 1,
                   autocast_dtype: Optional[torch.dtype] = torch.bfloat16, predictor: Optional[VolumePredictor] = None):
    """This rank's masks of one case: returns ((lo, hi), uint8 [hi-lo, h, w]).  No collective: a caller that wants the
    whole case on one rank gathers the 8-bit masks itself (2.6 MB for 160 slices)."""
    lo, hi = slice_range(series.shape[0], rank, world)
    if predictor is None:
        predictor = VolumePredictor(model, tuple(series.shape[1:]), batch=batch, autocast_dtype=autocast_dtype)
    return (lo, hi), predictor(series[lo:hi])
