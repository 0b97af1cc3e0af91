"""Data parallelism for the hot path: one process per GPU, batch-sharded, replicated weights.

The reference is single-device; the only exchange step data parallelism adds is the
gradient all-reduce.  Because the engine writes every weight gradient into one flat fp32 buffer, the exchange is
a handful of large NCCL all-reduces over NVLink/NVSwitch (bucketed so the first buckets overlap the optimizer's
launch latency) instead of ~150 small ones.  BatchNorm statistics stay per replica, like the reference (no SyncBN).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bucket_ranges(numel, bucket_elems):
    """[(start, end)] covering [0, numel) in buckets of at most bucket_elems elements."""
    if numel <= 0:
        return []
    bucket_elems = max(1, int(bucket_elems))
    return [(s, min(numel, s + bucket_elems)) for s in range(0, numel, bucket_elems)]


def allreduce_mean_(flat, group=None, bucket_elems=16 * 1024 * 1024):
    """In-place mean over ranks of a flat gradient buffer."""
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    works = []
    avg = dist.ReduceOp.AVG if flat.is_cuda else dist.ReduceOp.SUM   # gloo has no AVG
    for s, e in bucket_ranges(flat.numel(), bucket_elems):
        works.append(dist.all_reduce(flat[s:e], op=avg, group=group, async_op=True))
    for w in works:
        w.wait()
    if not flat.is_cuda:
        flat.div_(world)
    return flat


def broadcast_state_(module, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class DataParallel(torch.nn.Module):
    """Thin wrapper: broadcasts the initial state and averages the flat gradient of every backward.

    overlap=True (default on CUDA): the exchange is issued per gradient SEGMENT from inside the backward pass
    (engine.Executor.mark_grad_segment): head + decoder + LSTMs as soon as the LSTM backward is done, layer 4, layer 3 as
    soon as their blocks are done, the rest at the end -- each as one NCCL all-reduce on a communication stream ordered
    behind the kernels that produced the range, so ~95 % of the 110 MB travel under the backward pass of layers 3..1.  The
    calls are capture-safe: graph.GraphedStep records them inside the step's CUDA graph.  overlap=False: one bucketed
    exchange after the backward pass (what round 1 did)."""

    def __init__(self, module, group=None, bucket_elems=16 * 1024 * 1024, overlap=None):
        super().__init__()
        self.module = module
        self.group = group
        self.bucket_elems = bucket_elems
        if overlap is None:
            import os
            overlap = os.environ.get("STFB_DP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        if dist.is_available() and dist.is_initialized():
            broadcast_state_(module, 0, group)
            module._grad_ready_hook = self._on_grads
            if self.overlap and next(module.parameters()).is_cuda and dist.get_world_size(group) > 1:
                module._grad_segment_hook = self._on_segment
        # reference drivers key on this attribute (train_utils/train_and_eval.py:10)
        if hasattr(module, "input_format"):
            self.input_format = module.input_format

    def _on_grads(self, flat):
        if flat is not None:
            allreduce_mean_(flat, self.group, self.bucket_elems)
        else:       # the parameters' gradients no longer form one flat buffer (assigned by hand): one exchange per tensor
            for p in self.module.parameters():
                if p.grad is not None:
                    allreduce_mean_(p.grad.view(-1), self.group, self.bucket_elems)

    def _on_segment(self, flat_slice):
        """One gradient range is final: average it across the ranks, on the CURRENT (communication) stream."""
        if flat_slice.numel():
            dist.all_reduce(flat_slice, op=dist.ReduceOp.AVG, group=self.group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
