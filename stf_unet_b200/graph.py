"""CUDA-graph capture of the hot path: forward + criterion + backward as ONE graph launch.

A training step of STF-LSTM-UNet is ~480 kernel launches; enqueueing them from Python costs ~14 ms of host time per
step, which is more than the GPU needs once the kernels are fast.  Every libstfb200 launch is capture-safe (no host
syncs, no allocation, tensor maps are by-value kernel parameters), so the whole step is captured once and replayed.
The optimizer and the data-parallel all-reduce stay outside the graph (a handful of launches) so the same object
serves 1..8 GPUs.
"""
from __future__ import annotations

import torch


_CAPTURE_STREAMS = {}


class GraphedStep:
    """loss = step(x, target): copies the batch into static buffers, replays fwd + loss + bwd; gradients land in the
    parameters' persistent ``.grad`` views (one flat buffer), ready for all-reduce / optimizer.step()."""

    def __init__(self, model, criterion, example_x, example_target, autocast_dtype=torch.bfloat16, warmup=3,
                 capture_collectives=True):
        self.model = model
        self.x = example_x.clone()
        self.t = example_target.clone()
        # Data parallelism: with a per-segment hook (parallel.DataParallel(overlap=True)) the NCCL all-reduces are part of the
        # backward pass and are captured INSIDE the graph, on their own branch, overlapping the rest of the backward pass;
        # without one the exchange runs after every replay (self._hook).
        self._in_graph_comm = capture_collectives and model.__dict__.get("_grad_segment_hook") is not None
        self._seg_hook = None if self._in_graph_comm else model.__dict__.pop("_grad_segment_hook", None)
        self._hook = model.__dict__.pop("_grad_ready_hook", None)   # the after-replay exchange (unused with in-graph collectives)

        def fwd_bwd():
            with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
                loss = criterion(model(self.x), self.t)
            loss.backward()
            return loss

        model.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # learns the pack plan, configures kernels, warms the allocator
                for p in model.parameters():
                    p.grad = None
                fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in model.parameters():
            p.grad = None
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # captured on a HIGH-priority stream: kernel nodes keep the priority of the stream they were captured on, and the
        # weight-gradient side streams (default = lowest priority) then only fill the SM slots the main chain leaves free
        from . import engine
        # (ONE capture stream per device, shared by every GraphedStep like torch's own default capture stream: a second
        # capture on a fresh stream fails with cudaErrorStreamCaptureIsolation in the autograd engine's stream hand-off)
        cap_stream = None
        if engine.MAIN_PRIO:
            dev_index = self.x.device.index if self.x.device.index is not None else torch.cuda.current_device()
            cap_stream = _CAPTURE_STREAMS.get(dev_index)
            if cap_stream is None:
                cap_stream = _CAPTURE_STREAMS[dev_index] = torch.cuda.Stream(device=self.x.device, priority=engine.MAIN_PRIO)
        # thread_local: NCCL's watchdog thread polls events while the collectives of this step are being captured
        try:
            with torch.cuda.graph(self.graph, stream=cap_stream, capture_error_mode="thread_local" if self._in_graph_comm else "global"):
                self.loss = fwd_bwd()
        except Exception as e:
            if not self._in_graph_comm:
                raise
            # the collectives could not be captured on this system: keep the step graph, run the exchange after each replay
            import warnings
            warnings.warn(f"stf_unet_b200: capturing the gradient all-reduce inside the CUDA graph failed ({type(e).__name__}: {e}); "
                          "falling back to the exchange after the replay")
            torch.cuda.synchronize()
            self._in_graph_comm = False
            self._seg_hook = model.__dict__.pop("_grad_segment_hook", None)
            for p in model.parameters():
                p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph, stream=cap_stream):
                self.loss = fwd_bwd()
        self.launches_per_replay = _lib.launch_count() - n0      # libstfb200 kernels inside the captured graph
        self.flat_grad = model._last_flat_grad
        # replay runs no Python: the .grad views assigned during capture are re-bound after every replay, so a caller that
        # did zero_grad() (set_to_none=True is torch's default) in between still sees gradients -- with ANY optimizer
        self._grad_views = [(p, p.grad) for p in model.parameters() if p.requires_grad and p.grad is not None]
        if self._hook is not None:
            model._grad_ready_hook = self._hook
        if self._seg_hook is not None:
            model._grad_segment_hook = self._seg_hook

    def __call__(self, x, target):
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(target, non_blocking=True)
        self.graph.replay()
        self.model._last_flat_grad = self.flat_grad      # what a flat optimizer steps on (another graph / an eager step may have moved it)
        for p, g in self._grad_views:                    # the graph OVERWRITES its flat buffer: replay == zero_grad + backward
            if p.grad is not g:
                p.grad = g
        if self._hook is not None and not self._in_graph_comm:
            self._hook(self.flat_grad)
        return self.loss
