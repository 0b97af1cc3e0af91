#!/usr/bin/env python
"""bench.py -- headline benchmark of the STF-Unet hot path on B200.

Metric (BASELINE.json): train slices/s of STF-LSTM-UNet, bf16, T=8 DCE phases x 1x256x256, batch 16 per GPU
(configs[2] sharded 16/GPU: weak scaling).  One "step" = forward + CE/Dice loss + backward (+ gradient all-reduce
for N>1) + fused AdamW over one synthetic batch.

    python bench.py --gpus 1 --steps 10 --warmup 3          # own arm (libstfb200 kernels)
    python bench.py --impl reference ...                     # reference arm: the oracle port on the host CPU cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = same step fed from pinned host buffers
with the H2D copies and a D2H read of the loss inside the timed region.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_PHASES, HW, BATCH_PER_GPU = 8, 256, 16
TRAIN_GFLOP_PER_SLICE = 259.35        # BASELINE.md section 3 (fwd + dgrad + wgrad, conv + LSTM GEMMs)


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    version banner there at communicator creation, more with NCCL_DEBUG=INFO): keep a private duplicate of the real stdout
    for the JSON line and point descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, k), int):
                names[getattr(nv, k)] = k[len("nvmlClocksThrottleReason"):]
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (mask & bit) == bit and nm not in ("None", "All"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        rs = sorted(r for r in self.reasons if r not in ("GpuIdle", "ApplicationsClocksSetting"))
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": rs}


def make_batch(rank, batch=None, T=None, hw=None):
    from stf_unet_b200.synthetic import synthetic_dce_batch      # product-side generator (the oracle has its own twin)
    batch, T, hw = batch or BATCH_PER_GPU, T or T_PHASES, hw or HW   # the workload's globals, read at call time
    return synthetic_dce_batch(batch, T, hw, hw, seed=1234 + rank, half_res_target=True)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference modules (baseline/_ref, installed by tools/install_ref.sh; it
# travels to the GPU box) on the host cores; the oracle port only when that install is missing
# ------------------------------------------------------------------------------------------------
def reference_root():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "src")) and os.path.isdir(os.path.join(p, "train_utils")):
            return p
    return None


_REF = None


def import_reference():
    """-> namespace with the reference's STFLSTMUNet, UNet, criterion, train_one_epoch, evaluate (or None)."""
    global _REF
    if _REF is None:
        root = reference_root()
        if root is None:
            _REF = False
        else:
            sys.dont_write_bytecode = True
            if root not in sys.path:
                sys.path.insert(0, root)
            import types
            from src import STFLSTMUNet, UNet                                  # noqa: E402  (the reference's own package)
            from train_utils import train_and_eval as te                        # noqa: E402
            _REF = types.SimpleNamespace(root=root, STFLSTMUNet=STFLSTMUNet, UNet=UNet, criterion=te.criterion,
                                         train_one_epoch=te.train_one_epoch, evaluate=te.evaluate,
                                         create_lr_scheduler=te.create_lr_scheduler)
    return _REF or None


def cpu_train_step_time(batch, iters, warmup, threads=None, workload="train"):
    """Seconds per training step (fwd + criterion + bwd + AdamW) of the reference on the host cores -> (times, threads, kind)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ref = import_reference()
    if workload == "unet":
        from stf_unet_b200.synthetic import synthetic_dce_batch
        x, t = synthetic_dce_batch(batch, 1, HW, HW, seed=1234, half_res_target=False)
        x = x[:, 0]                                                             # [B, 1, H, W]
    else:
        x, t = make_batch(0, batch=batch)
    times = []
    if ref is not None:
        torch.manual_seed(0)
        model = (ref.UNet(1, 2, 64) if workload == "unet" else ref.STFLSTMUNet(1, 2, T_PHASES)).train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            loss = ref.criterion(model(x), t)
            opt.zero_grad()
            loss.backward()
            opt.step()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        return times, threads, "reference"
    from oracle import stf_oracle as O
    from oracle import weights as W
    sd = W.make_state_dict(W.unet_param_spec(1, 2, 64) if workload == "unet" else W.stf_param_spec(1, 2), seed=0)
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.loss_and_grads(sd, x, t, model="unet" if workload == "unet" else "stf", train=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, threads, "port"


def run_reference(args, rank):
    if rank != 0:
        return
    unet = args.workload == "unet"
    batch = 4 if unet else (2 if HW <= 256 else 1)
    times, threads, kind = cpu_train_step_time(batch, args.steps, args.warmup, workload="unet" if unet else "train")
    ms = 1000.0 * sum(times) / len(times)
    val = batch / (ms / 1000.0)
    what = "the unmodified reference modules (baseline/_ref)" if kind == "reference" else "the oracle port of the reference"
    if unet:
        sample = f"fwd+CE/Dice+bwd+AdamW of B={batch} images (1x{HW}x{HW}) per step, fp32, {what}: BASELINE.json configs[0] as written"
        cfg = unet_config(1)
        metric = "train images/s UNet"
    else:
        sample = (f"fwd+CE/Dice+bwd+AdamW of B={batch} slices (T={T_PHASES}, {HW}x{HW}) per step -- a bounded sample of the "
                  f"{BATCH_PER_GPU}-slice step --, fp32, {what}")
        cfg = workload_config(args.gpus)
        metric = "train slices/s STF-LSTM-UNet"
    cfg = dict(cfg)
    cfg["launch"] = f"reference arm: PyTorch CPU backend (oneDNN/MKL), {threads} host threads, eager"
    cfg["l2"] = "n/a (host CPU)"
    line = {"impl": "reference", "metric": metric, "value": round(val, 4), "unit": "images/s" if unet else "slices/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": round(val, 4), "unit": "images/s" if unet else "slices/s", "cores": threads, "kind": kind,
                             "sample": sample},
            "e2e": {"value": round(val, 4), "unit": "images/s" if unet else "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def gpu_eager_baseline(dev, x_dev, t_dev, steps=3, warmup=2, unet=False):
    """The practical bar (SURVEY.md section 8(d) last row): the UNMODIFIED reference module on this
    same B200 through PyTorch eager / cuDNN, same batch, fwd + criterion + bwd + AdamW(fused), CUDA events after warm-up;
    fp32 with TF32 off (the parity oracle's arithmetic) and torch.autocast(bf16) (the like-for-like precision)."""
    ref = import_reference()
    if ref is None:
        return {"unavailable": "baseline/_ref not installed (tools/install_ref.sh)"}
    saved = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {"what": "reference nn.Module (baseline/_ref) on the same GPU, PyTorch eager + cuDNN, cudnn.benchmark=True, "
                   "fwd+criterion+bwd+AdamW(fused), same batch", "batch": int(x_dev.shape[0])}
    try:
        for tag, dt in (("fp32_tf32_off", None), ("autocast_bf16", torch.bfloat16)):
            torch.manual_seed(0)
            model = (ref.UNet(1, 2, 64) if unet else ref.STFLSTMUNet(1, 2, int(x_dev.shape[1]))).to(dev).train()
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)

            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt is not None):
                    loss = ref.criterion(model(x_dev), t_dev)
                opt.zero_grad()
                loss.backward()
                opt.step()
                return loss

            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = {"ms_per_step": round(ms, 3), "per_s": round(x_dev.shape[0] / (ms / 1e3), 2)}
            del model, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


CONFIG_NOTE = "BASELINE.json configs[2] = global batch 128 on 8 GPUs"


def workload_config(n, exchange=None):
    return {"workload": f"STF-LSTM-UNet train fwd+CE/Dice+bwd+AdamW, T={T_PHASES} x 1x{HW}x{HW}, batch {BATCH_PER_GPU}/GPU "
                        f"({CONFIG_NOTE})",
            "global_batch": BATCH_PER_GPU * n, "T": T_PHASES, "hw": HW, "parallelism": f"dp{n}", "launch": "CUDA graph (fwd+loss+bwd" + (exchange or "") + ") + one-launch flat AdamW",
            "l2": "per-step working set (activations of several GB) far exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_own(args, rank, world, local_rank):
    import torch.distributed as dist
    import stf_unet_b200 as S
    from stf_unet_b200 import _lib, ops, parallel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(0)
    model = S.STFLSTMUNet(1, 2, T_PHASES).to(dev)
    net = parallel.DataParallel(model) if world > 1 else model
    # one-launch AdamW over flat parameter / gradient / moment buffers (re-homes the parameters: before graph capture)
    opt = S.FlatAdamW(model, lr=1e-3, weight_decay=1e-4)
    x_host, t_host = make_batch(rank)
    x_pin, t_pin = x_host.pin_memory(), t_host.pin_memory()
    x_dev, t_dev = x_pin.to(dev), t_pin.to(dev)

    def eager_step(x, t):
        model.train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
            loss = S.criterion(out, t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def make_graphed(x_example):
        from stf_unet_b200.graph import GraphedStep
        g = GraphedStep(net.module if world > 1 else model, S.criterion, x_example, t_dev)
        if world > 1:
            g._hook = net._on_grads
        return g

    graphed = None if args.no_graph else make_graphed(x_dev)
    exchange = None
    if world > 1:
        in_graph = graphed is not None and graphed._in_graph_comm
        exchange = (" + per-segment NCCL all-reduce captured in the graph, overlapping the backward pass" if in_graph
                    else ") + (bucketed NCCL all-reduce after the graph")

    def step(x, t):
        if graphed is None:
            return eager_step(x, t)
        loss = graphed(x, t)          # fwd + CE/Dice + bwd: one CUDA-graph launch (+ gradient all-reduce for N > 1)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step(x_dev, t_dev)
    n0 = _lib.launch_count()
    with ClockSampler(local_rank) as clk:
        total_ms = timed(lambda: step(x_dev, t_dev), args.steps)
    launches = _lib.launch_count() - n0
    if graphed is not None:
        launches += graphed.launches_per_replay * args.steps     # replays do not pass through the host-side counter
    ms_per_step = total_ms / args.steps
    value = BATCH_PER_GPU * world / (ms_per_step / 1000.0)

    # ---- end to end: pinned host -> device each step, loss read back each step ----
    # Every step copies ITS batch from pinned host memory and reads the loss back.  The copy of step i+1 is issued on a copy
    # stream into the other half of a double buffer while step i computes (what a prefetching loader does); the loss read
    # synchronises every step.
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    tbuf = [torch.empty_like(t_dev), torch.empty_like(t_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}

    def issue_copy(slot):
        # no wait needed: the previous user of this slot was step i-1, whose loss read already synchronised the host
        with torch.cuda.stream(copy_stream):
            xbuf[slot].copy_(x_pin, non_blocking=True)
            tbuf[slot].copy_(t_pin, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        slot = state["i"] & 1
        issue_copy(slot ^ 1)                                      # next step's batch, overlapped with this step's compute
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = step(xbuf[slot], tbuf[slot])
        state["i"] += 1
        return loss.item()

    issue_copy(0)
    e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    e2e_val = BATCH_PER_GPU * world / (e2e_ms / 1000.0)

    # ---- the same end-to-end step fed with RAW 8-bit series (what the reference's loader reads from disk): a quarter of
    #      the host->device bytes, ToTensor + Normalize fused into the device-side layout pass ----
    e2e_u8 = None
    if graphed is not None:
        from stf_unet_b200.synthetic import synthetic_dce_batch_u8
        u8_host, _ = synthetic_dce_batch_u8(BATCH_PER_GPU, T_PHASES, HW, HW, seed=1234 + rank)
        u8_pin = u8_host.unsqueeze(2).contiguous().pin_memory()              # [B, T, 1, H, W] uint8
        u8buf = [u8_pin.to(dev), u8_pin.to(dev)]
        graphed_u8 = make_graphed(u8buf[0])

        def issue_copy_u8(slot):
            with torch.cuda.stream(copy_stream):
                u8buf[slot].copy_(u8_pin, non_blocking=True)
                tbuf[slot].copy_(t_pin, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_u8_step():
            slot = state["i"] & 1
            issue_copy_u8(slot ^ 1)
            torch.cuda.current_stream().wait_event(ready[slot])
            loss = graphed_u8(u8buf[slot], tbuf[slot])
            opt.step()
            state["i"] += 1
            return loss.item()

        torch.cuda.synchronize()
        issue_copy_u8(state["i"] & 1)
        e2e_u8_step()
        u8_ms = timed(e2e_u8_step, args.steps) / args.steps
        e2e_u8 = {"value": round(BATCH_PER_GPU * world / (u8_ms / 1000.0), 2), "unit": "slices/s",
                  "h2d_bytes_per_step": int(u8_pin.numel() + t_pin.numel() * 8) * world, "d2h_bytes_per_step": 4 * world,
                  "ms_per_step": round(u8_ms, 3), "input": "uint8 grey levels, normalised on the device"}
        del graphed_u8

    # ---- numerics guard at the full bench size (the CPU oracle is too slow here): the bf16 tensor-core loss of this batch
    #      against the fp32 FFMA family of the same library on the same weights; a corrupted pipeline shows up here ----
    loss_check = None
    if True:                              # every rank: the train-mode backward below holds the gradient all-reduce
        model.eval()                      # eval mode: no running-stat updates, no tape
        with torch.no_grad():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                l16 = S.criterion(model(x_dev), t_dev).item()
            l32 = S.criterion(model(x_dev), t_dev).item()
        model.train()
        # backward kernels (dgrad / wgrad / BN / LSTM): gradients of one train-mode step, bf16 tensor cores vs fp32 FFMA
        names = ["conv1.weight", "layer1.0.conv1.weight", "layer2.1.conv2.weight", "layer3.2.conv1.weight",
                 "layer4.1.conv2.weight", "lstm1.weight_hh_l0", "lstm4.weight_ih_l0", "decoder3.fusion.weight",
                 "final_res.conv_block.0.weight"]
        pd = dict(model.named_parameters())
        grads = {}
        for tag, dt in (("bf16", torch.bfloat16), ("fp32", None)):
            for p_ in model.parameters():
                p_.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt is not None):
                loss_ = S.criterion(model(x_dev), t_dev)
            loss_.backward()
            grads[tag] = {n: pd[n].grad.detach().double().clone() for n in names}
        gerr = {n: ((grads["bf16"][n] - grads["fp32"][n]).norm() / grads["fp32"][n].norm().clamp_min(1e-30)).item() for n in names}
        for p_ in model.parameters():
            p_.grad = None
        loss_check = {"eval_loss_bf16_tcgen05": round(l16, 6), "eval_loss_fp32": round(l32, 6),
                      "rel_diff": round(abs(l16 - l32) / max(abs(l32), 1e-12), 6),
                      "grad_rel_err_bf16_vs_fp32_max": round(max(gerr.values()), 5),
                      "grad_rel_err_worst": max(gerr, key=gerr.get)}
        # bars: eval loss within 2e-2 (north_star's bf16 tolerance); every sampled gradient within 0.25 = 2.7 x the distance of
        # torch's OWN bf16 autocast of the reference arithmetic from fp32 on its worst parameter at this size (9.2e-2,
        # tests/test_models_gpu.py::test_bench_size_train_step_bf16_vs_oracle).  The weights here are wherever 25+ optimizer steps on
        # random data have taken them (loss ~1e-2, tiny gradients), and the trajectory is not bit-reproducible (fp32 atomics in the
        # BatchNorm statistics): at N = 2 the same command measured 0.04 / 0.13 / 0.29 for the worst gradient and 1e-3 .. 3e-2 for the
        # loss in four runs.  So the bars are REPORTED (`within_bars`) and only a grossly wrong pipeline -- O(1) differences or
        # non-finite values -- stops the bench; the parity evidence is the test suite, this is a tripwire.
        import math
        worst = max(gerr.values())
        rel_l = abs(l16 - l32) / max(abs(l32), 1e-6)
        loss_check["within_bars"] = bool(rel_l <= 2e-2 and worst < 0.25)
        grad_bar = float(os.environ.get("STFB_BENCH_GRAD_BAR", "0.75"))
        if rank == 0 and (os.environ.get("STFB_BENCH_VERBOSE") or not loss_check["within_bars"]):
            sys.stderr.write(f"bench.py: loss_check {loss_check} {gerr}\n")
        bad = int(not (math.isfinite(l16) and math.isfinite(l32) and math.isfinite(worst)) or rel_l > 0.25 or worst >= grad_bar)
        if world > 1:                     # the eval losses are per-rank: every rank leaves together or none does
            flag = torch.tensor([bad], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            bad = int(flag.item())
        if bad:
            raise SystemExit(f"bench.py: bf16 and fp32 paths disagree on the bench batch: {loss_check} {gerr}")

    # ---- roofline of the dominant kernel family: per-launch CUDA events over one extra step ----
    roof = None
    # every rank runs the extra step (it contains the gradient all-reduce); only rank 0 records events
    prof = ops.KernelProfiler() if rank == 0 else None
    ops.set_profiler(prof)
    from stf_unet_b200 import engine as _engine
    _saved = (_engine.USE_WGRAD_STREAM, _engine.USE_LSTM_STREAMS)
    _engine.USE_WGRAD_STREAM = _engine.USE_LSTM_STREAMS = False     # one stream: a kernel's event pair times that kernel alone
    eager_step(x_dev, t_dev)        # per-launch events need the eager path
    torch.cuda.synchronize()
    _engine.USE_WGRAD_STREAM, _engine.USE_LSTM_STREAMS = _saved
    ops.set_profiler(None)
    if rank == 0:
        fam = prof.summary()
        if args.profile_detail:
            for ms, n, tf, family, tag in prof.top(40):
                print(f"[prof] {ms:8.3f} ms  n={n:3d}  {tf:7.1f} TF/s  {family:18s} {tag}", file=sys.stderr)
            for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
                print(f"[family] {k:18s} {v['ms']:8.3f} ms  n={v['n']:4d}  {v['flops'] / (v['ms'] * 1e-3) / 1e12 if v['ms'] else 0:7.1f} TF/s  "
                      f"{v['bytes'] / (v['ms'] * 1e-3) / 1e9 if v['ms'] else 0:8.1f} GB/s", file=sys.stderr)
        pk = peaks()
        gemm = {k: v for k, v in fam.items() if v["flops"] > 0}
        membound = {k: v for k, v in fam.items() if v["flops"] == 0}
        fam = gemm
        if fam:
            # Dominant kernel = the GEMM-shaped FAMILY with the largest summed launch time in the step (all launches of one
            # kernel template family: forward + dgrad convolutions and the fused LSTM steps are `conv_tcgen05`, the weight
            # gradients `wgrad_tcgen05`).  `achieved` = the family's algorithmic FLOPs / the sum of its CUDA-event launch
            # times; the slowest (family, shape) group and every family are listed as detail, the whole step (all GEMM FLOPs
            # over the graph-replay step time, which also contains every memory-bound kernel) as `whole_step`.  `traffic`:
            # DRAM bytes per launch of the dominant shape from the committed ncu --set full capture
            # (profiles/roofline_traffic.json), when there is one.
            t_family = max(fam, key=lambda k: fam[k]["ms"])
            d = fam[t_family]
            f_tf = d["flops"] / (d["ms"] / 1e3) / 1e12
            groups = [g for g in prof.top(10000) if g[3] == t_family]
            t_ms, t_n, t_tf, _, t_tag = max(groups, key=lambda g: g[0])
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(t_tag, {}).get("dram_bytes_per_launch")
            names = {"conv3x3_halo_tcgen05": "conv_halo2_kernel (CTA pairs, tcgen05 cta_group::2) / conv_halo_kernel: 3x3 stride-1 implicit-GEMM convolutions, forward + dgrad",
                     "conv_other_tcgen05": "conv_tc_kernel: stride-2 / 1x1 / transposed / 8x8-map / 32-channel convolutions and the LSTM backward GEMMs",
                     "lstm_step_tcgen05": "conv_tc_kernel<EPI=1>: fused LSTM time step (gate GEMM + cell update)",
                     "wgrad_tcgen05": "wgrad_halo_kernel / wgrad_tc_kernel (tcgen05 weight gradients)"}
            whole_tf = value / world * TRAIN_GFLOP_PER_SLICE / 1e3
            roof = {"bound": "tensor", "kernel": f"{t_family}: {names.get(t_family, t_family)}", "achieved": round(f_tf, 2),
                    "peak": pk["tflops"], "unit": "TFLOP/s", "frac": round(f_tf / pk["tflops"], 4), "traffic": traffic,
                    "peak_source": pk["src"], "launches": d["n"], "avg_launch_us": round(1e3 * d["ms"] / d["n"], 2),
                    "share_of_step": round(d["ms"] / ms_per_step, 3),
                    "timing": "per-launch CUDA events over one eager single-stream step (launch gaps of a few us per launch included)",
                    "dominant_shape": {"shape": t_tag, "launches": t_n, "avg_launch_us": round(1e3 * t_ms / t_n, 2),
                                       "tflops": round(t_tf, 2), "frac": round(t_tf / pk["tflops"], 4), "traffic_bytes_per_launch": traffic},
                    "whole_step": {"tflops": round(whole_tf, 2), "frac": round(whole_tf / pk["tflops"], 4),
                                   "note": "all conv + LSTM GEMM FLOPs of the step / the graph-replay step time (includes every memory-bound kernel)"},
                    "families": {k: {"ms": round(v["ms"], 3), "n": v["n"],
                                     "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 2) if v["ms"] > 0 else None,
                                     "frac": round(v["flops"] / (v["ms"] / 1e3) / 1e12 / pk["tflops"], 4) if v["ms"] > 0 else None}
                                 for k, v in fam.items()},
                    "hbm_families": {k: {"ms": round(v["ms"], 3), "n": v["n"],
                                         "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["ms"] > 0 else None,
                                         "frac_of_hbm_peak": round(v["bytes"] / (v["ms"] / 1e3) / 1e9 / pk["hbm_gbs"], 3) if v["ms"] > 0 else None}
                                     for k, v in membound.items()}}

    cpu = None
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, cn = (2, 3) if HW <= 256 else (1, 1)
        times, threads, kind = cpu_train_step_time(cb, cn, 1)
        v = cb / (sum(times) / len(times))
        what = "the unmodified reference modules (baseline/_ref)" if kind == "reference" else "the oracle port"
        cpu = {"value": round(v, 4), "unit": "slices/s", "cores": threads, "kind": kind,
               "sample": f"{cn} train step(s) (fwd+CE/Dice+bwd+AdamW) of B={cb} slices (T={T_PHASES}, {HW}x{HW}) fp32 on {what}, 1 warm-up"}
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        # free this arm's graphs and buffers first: the reference keeps every fp32 activation of the step alive
        graphed = None
        torch.cuda.empty_cache()
        try:
            gpu_ref = gpu_eager_baseline(dev, x_dev, t_dev)
        except Exception as e:       # the baseline must never take the product's number down with it
            gpu_ref = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        pk = peaks()
        model_tflops = value * TRAIN_GFLOP_PER_SLICE / 1e3
        line = {"metric": "train slices/s STF-LSTM-UNet", "value": round(value, 2), "unit": "slices/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world, exchange), "clocks": clk.summary(),
                "e2e": {"value": round(e2e_val, 2), "unit": "slices/s", "h2d_bytes_per_step": int(x_pin.numel() * 4 + t_pin.numel() * 8) * world,
                        "d2h_bytes_per_step": 4 * world, "ms_per_step": round(e2e_ms, 3)},
                "e2e_u8": e2e_u8, "gpu_launches": int(launches),
                "model_tflops": round(model_tflops, 2), "model_frac_of_bf16_peak": round(model_tflops / world / pk["tflops"], 4),
                "loss_check": loss_check, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": gpu_ref}
        emit(line)
    if world > 1:
        # the step's CUDA graph holds captured NCCL kernels: a communicator may only be torn down after every graph that
        # captured its collectives is gone (ncclCommDestroy otherwise waits forever)
        import gc
        dist.barrier()
        graphed = None
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


UNET_TRAIN_GFLOP_PER_IMAGE = 288.50     # BASELINE.md section 3 (UNet in=1, 256x256: fwd + dgrad + wgrad)


def unet_config(n):
    return {"workload": f"UNet(in_channels=1, num_classes=2, base_c=64) train fwd+CE/Dice+bwd+AdamW, batch 4 x 1x{HW}x{HW} "
                        "(BASELINE.json configs[0], the reference's CPU-runnable case)",
            "global_batch": 4 * n, "T": 1, "hw": HW, "parallelism": f"dp{n}",
            "launch": "CUDA graph (fwd+loss+bwd) + one-launch flat AdamW",
            "l2": "a 67 MB buffer is rewritten between timed steps (the step's activations, ~0.9 GB fp32, exceed L2 anyway)"}


def run_unet(args, rank, world, local_rank):
    """BASELINE.json configs[0]: UNet(1, 2, 64), batch 4 x 1x256x256, fwd + CE/Dice + bwd (+ AdamW).  The reference states it
    in fp32 on the CPU; here `value` is the fp32-accurate mode (the 1e-4 parity mode: FFMA implicit GEMM, fp32 storage) and
    `bf16` the tensor-core mode of the same step, next to the reference's CPU time and its eager-cuDNN time on this GPU."""
    import stf_unet_b200 as S
    from stf_unet_b200 import _lib
    from stf_unet_b200.graph import GraphedStep
    from stf_unet_b200.synthetic import synthetic_dce_batch
    if world > 1:
        raise SystemExit("bench.py --workload unet is a single-GPU configuration (configs[0])")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = 4
    x_host, t_host = synthetic_dce_batch(B, 1, HW, HW, seed=1234, half_res_target=False)
    x_pin, t_pin = x_host[:, 0].contiguous().pin_memory(), t_host.pin_memory()
    x_dev, t_dev = x_pin.to(dev), t_pin.to(dev)
    flush = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    results = {}
    launches = 0
    clk_summary = None
    for tag, dt in (("fp32", None), ("bf16", torch.bfloat16)):
        torch.manual_seed(0)
        model = S.UNet(1, 2, 64).to(dev)
        opt = S.FlatAdamW(model, lr=1e-3, weight_decay=1e-4)
        g = GraphedStep(model, S.criterion, x_dev, t_dev, autocast_dtype=dt)

        def step(x, t):
            loss = g(x, t)
            opt.step()
            return loss

        def timed(fn, steps):
            tot = 0.0
            for _ in range(steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / steps

        for _ in range(max(args.warmup, 3)):
            step(x_dev, t_dev)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        with ClockSampler(local_rank) as clk:
            ms = timed(lambda: step(x_dev, t_dev), args.steps)
        if dt is None:
            launches = _lib.launch_count() - n0 + g.launches_per_replay * args.steps
            clk_summary = clk.summary()

        def e2e_step():
            x_dev.copy_(x_pin, non_blocking=True)
            t_dev.copy_(t_pin, non_blocking=True)
            return step(x_dev, t_dev).item()

        e2e_step()
        e2e_ms = timed(e2e_step, args.steps)
        results[tag] = {"ms_per_step": round(ms, 3), "images_per_s": round(B / (ms / 1e3), 2), "e2e_ms_per_step": round(e2e_ms, 3),
                        "e2e_images_per_s": round(B / (e2e_ms / 1e3), 2), "loss": round(g.loss.item(), 6),
                        "model_tflops": round(B / (ms / 1e3) * UNET_TRAIN_GFLOP_PER_IMAGE / 1e3, 2)}
        del g, opt, model
        torch.cuda.empty_cache()
    cpu = None
    if not args.no_cpu_baseline:
        times, threads, kind = cpu_train_step_time(B, 3, 1, workload="unet")
        v = B / (sum(times) / len(times))
        what = "the unmodified reference modules (baseline/_ref)" if kind == "reference" else "the oracle port"
        cpu = {"value": round(v, 4), "unit": "images/s", "cores": threads, "kind": kind,
               "sample": f"3 train steps (fwd+CE/Dice+bwd+AdamW) of B={B} images (1x{HW}x{HW}) fp32 on {what}, 1 warm-up: configs[0] as written"}
    gpu_ref = None
    if not args.no_gpu_baseline:
        try:
            gpu_ref = gpu_eager_baseline(dev, x_dev, t_dev, unet=True)
        except Exception as e:
            gpu_ref = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    pk = peaks()
    f32 = results["fp32"]
    # fp32 mode: convolutions, dgrads and weight gradients run on the tensor cores over split-precision operands (STFB_BF16X3:
    # three bf16 planes per fp32 tensor, SIX tcgen05 products per multiply-accumulate), so the ceiling of the scheme is the measured
    # bf16 dense peak / 6; `achieved` counts the ALGORITHMIC fp32 FLOPs of the step, executed bf16 FLOPs = 6x.
    # The FFMA family this mode ran on before (STFB_NO_SPLIT_FP32=1) has its own ceiling of 74.4 TFLOP/s.
    from stf_unet_b200 import engine as _engine
    split = _engine.USE_SPLIT_FP32 and _engine.USE_TCGEN05
    peak = pk["tflops"] / 6.0 if split else 74.4
    roof = {"bound": "tensor",
            "kernel": ("conv_halo2_kernel / conv_halo_kernel / conv_tc_kernel / wgrad_halo_kernel on bf16x3 operands (whole fp32 step)" if split
                       else "igemm_simt_kernel<float> + wgrad_simt_kernel<float> (whole fp32 step)"),
            "achieved": f32["model_tflops"], "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(f32["model_tflops"] / peak, 4),
            "traffic": None,
            "peak_source": (pk["src"] + " bf16 dense peak / 6 products per MAC") if split else "148 SMs x 128 FMA lanes x 2 x 1.965 GHz",
            "executed_bf16_tflops": round(f32["model_tflops"] * 6, 1) if split else None,
            "fp32_ffma_peak_tflops": 74.4, "frac_of_fp32_ffma_peak": round(f32["model_tflops"] / 74.4, 4),
            "bf16_tensor_mode": {"tflops": results["bf16"]["model_tflops"], "frac": round(results["bf16"]["model_tflops"] / pk["tflops"], 4)}}
    line = {"metric": "train images/s UNet", "value": f32["images_per_s"], "unit": "images/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": f32["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": unet_config(1), "clocks": clk_summary,
            "e2e": {"value": f32["e2e_images_per_s"], "unit": "images/s", "h2d_bytes_per_step": int(x_pin.numel() * 4 + t_pin.numel() * 8),
                    "d2h_bytes_per_step": 4, "ms_per_step": f32["e2e_ms_per_step"]},
            "gpu_launches": int(launches), "model_tflops": f32["model_tflops"], "modes": results, "roofline": roof,
            "cpu_baseline": cpu, "gpu_eager_baseline": gpu_ref}
    emit(line)


INFER_GFLOP_PER_SLICE = 86.72          # BASELINE.md / SURVEY.md section 8(d): forward conv + LSTM GEMMs


def run_infer(args, rank, world, local_rank):
    """BASELINE.json configs[1]: STF-LSTM-UNet eval forward, T=8 x 1x256x256, batch 16 bf16 per GPU, followed by the fused
    argmax-mask kernel (the per-slice product of configs[4]).  Slices are independent: no collective."""
    import torch.distributed as dist
    import stf_unet_b200 as S
    from stf_unet_b200 import _lib
    from stf_unet_b200.metrics import EvalMetrics

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(0)
    model = S.STFLSTMUNet(1, 2, T_PHASES).to(dev).eval()
    x_host, t_host = make_batch(rank)
    x_pin, t_pin = x_host.pin_memory(), t_host.pin_memory()
    xs, ts = x_pin.to(dev), t_pin.to(dev)
    met = EvalMetrics(2, ignore_index=255, device=dev)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(xs)
        return met.update(out, ts, want_mask=True)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    n0 = _lib.launch_count()
    with torch.cuda.graph(graph):
        mask = fwd()
    per_replay = _lib.launch_count() - n0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        graph.replay()
    with ClockSampler(local_rank) as clk:
        ms = timed(graph.replay, args.steps) / args.steps
    value = BATCH_PER_GPU * world / (ms / 1e3)

    # ---- end to end: every step copies ITS batch from pinned host memory and delivers ITS masks to pinned host memory.
    #      Double buffered like a prefetching loader: batch i+1 travels host -> device on a copy stream while graph i runs
    #      (two captured graphs, one per input buffer, sharing one memory pool), and the masks of step i-1 travel back on a
    #      third stream; the host waits for the masks of step i-1 only, so one step of latency hides both copies.
    xs2 = torch.empty_like(xs)
    xbufs = [xs, xs2]
    graph2 = torch.cuda.CUDAGraph()
    met2 = EvalMetrics(2, ignore_index=255, device=dev)

    def fwd2():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(xs2)
        return met2.update(out, ts, want_mask=True)

    with torch.cuda.graph(graph2, pool=graph.pool()):
        mask2 = fwd2()
    graphs, masks_dev = [graph, graph2], [mask, mask2]
    mask_host = [torch.empty(mask.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
    h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    in_ready = [torch.cuda.Event(), torch.cuda.Event()]
    out_ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    st = {"i": 0}

    def issue_h2d(slot):
        h2d.wait_event(done[slot])                 # the graph that last read this buffer has finished
        with torch.cuda.stream(h2d):
            xbufs[slot].copy_(x_pin, non_blocking=True)
            in_ready[slot].record(h2d)

    def e2e_step():
        i = st["i"]
        slot = i & 1
        issue_h2d(slot ^ 1)                         # next step's batch
        cur = torch.cuda.current_stream()
        cur.wait_event(in_ready[slot])
        graphs[slot].replay()
        done[slot].record(cur)
        d2h.wait_event(done[slot])
        with torch.cuda.stream(d2h):
            mask_host[slot].copy_(masks_dev[slot], non_blocking=True)
            out_ready[slot].record(d2h)
        if i > 0:
            out_ready[slot ^ 1].synchronize()       # the PREVIOUS step's masks are on the host now
        st["i"] = i + 1

    torch.cuda.synchronize()
    for ev in done:
        ev.record()
    issue_h2d(0)
    e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    assert torch.equal(mask_host[0], mask_host[1])  # same batch through both graphs: identical masks
    if rank == 0:
        pk = peaks()
        tf = value * INFER_GFLOP_PER_SLICE / 1e3
        cfg = {"workload": f"STF-LSTM-UNet eval forward + argmax mask, T={T_PHASES} x 1x{HW}x{HW}, batch {BATCH_PER_GPU}/GPU bf16 "
                           "(BASELINE.json configs[1])", "global_batch": BATCH_PER_GPU * world, "T": T_PHASES, "hw": HW,
               "parallelism": f"dp{world} (independent slices, no collective)", "launch": "CUDA graph",
               "l2": "activations of one forward (~1.5 GB) exceed the 126 MB L2; no flush needed"}
        line = {"metric": "inference slices/s STF-LSTM-UNet", "value": round(value, 2), "unit": "slices/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
                "clocks": clk.summary(),
                "e2e": {"value": round(BATCH_PER_GPU * world / (e2e_ms / 1e3), 2), "unit": "slices/s",
                        "h2d_bytes_per_step": int(x_pin.numel() * 4) * world, "d2h_bytes_per_step": int(mask.numel()) * world,
                        "ms_per_step": round(e2e_ms, 3)},
                "gpu_launches": int(per_replay * args.steps), "model_tflops": round(tf, 2),
                "model_frac_of_bf16_peak": round(tf / world / pk["tflops"], 4),
                # no per-family CUDA-event leg in this workload: the WHOLE step's algorithmic FLOP rate against the measured bf16 peak
                "roofline": {"bound": "tensor", "kernel": "whole inference step (conv_halo2 / conv_tc / lstm_seq64 kernels + the elementwise passes)",
                             "achieved": round(tf / world, 2), "peak": pk["tflops"], "unit": "TFLOP/s",
                             "frac": round(tf / world / pk["tflops"], 4), "traffic": None, "peak_source": pk["src"], "whole_step": True},
                "cpu_baseline": None}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_volume(args, rank, world, local_rank):
    """BASELINE.json configs[4]: whole-volume tumour-mask inference, one synthetic case of 160 slices x T=8 x 256x256,
    sharded by slice across the ranks (strong scaling, no collective on the data path).  One "step" = one whole case:
    host series in, uint8 masks out (every step is end to end by construction; `value` keeps the series resident)."""
    import torch.distributed as dist
    import stf_unet_b200 as S
    from stf_unet_b200.volume import VolumePredictor, slice_range
    from stf_unet_b200.synthetic import synthetic_dce_batch

    SLICES, VB = 160, 20
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(0)
    model = S.STFLSTMUNet(1, 2, T_PHASES).to(dev).eval()
    lo, hi = slice_range(SLICES, rank, world)
    # this rank's slices of the case (20 distinct synthetic slices tiled over the range: generation cost, not content, is bounded)
    base, _ = synthetic_dce_batch(VB, T_PHASES, HW, HW, seed=1234 + rank)
    series = base.repeat((hi - lo + VB - 1) // VB, 1, 1, 1, 1)[:hi - lo].contiguous().pin_memory()
    pred = VolumePredictor(model, tuple(series.shape[1:]), batch=min(VB, max(1, hi - lo)))
    masks = torch.empty((hi - lo, HW // 2, HW // 2), dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        pred(series, out=masks)
    r0 = pred.replays
    with ClockSampler(local_rank) as clk:
        e2e_ms = timed(lambda: pred(series, out=masks), args.steps) / args.steps
    replays = pred.replays - r0
    # device-resident variant: the same batches without the host copies (what `value` reports)
    nb = (hi - lo + pred.batch - 1) // pred.batch

    def resident_case():
        for _ in range(nb):
            pred.graph.replay()

    resident_case()
    ms = timed(resident_case, args.steps) / args.steps
    # the same slice must give the same mask wherever it sits in the volume (slices are independent)
    assert torch.equal(masks[:min(VB, hi - lo)], masks[VB:2 * VB][:min(VB, hi - lo)]) if hi - lo >= 2 * VB else True
    if rank == 0:
        pk = peaks()
        value = SLICES / (ms / 1e3)
        tf = value * INFER_GFLOP_PER_SLICE / 1e3
        cfg = {"workload": f"whole-volume tumour-mask inference: {SLICES} slices x T={T_PHASES} x 1x{HW}x{HW} per synthetic case, "
                           f"bf16, batches of {pred.batch} slices, sharded by slice over {world} GPU(s) (BASELINE.json configs[4])",
               "slices_per_case": SLICES, "T": T_PHASES, "hw": HW, "parallelism": f"slices/{world} (no collective)",
               "launch": "CUDA graph per batch", "l2": "activations of one batch (~2 GB) exceed the 126 MB L2; no flush needed"}
        line = {"metric": "inference slices/s STF-LSTM-UNet (whole volume)", "value": round(value, 2), "unit": "slices/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": cfg, "clocks": clk.summary(),
                "e2e": {"value": round(SLICES / (e2e_ms / 1e3), 2), "unit": "slices/s",
                        "h2d_bytes_per_step": int(SLICES * T_PHASES * HW * HW * 4), "d2h_bytes_per_step": int(SLICES * (HW // 2) ** 2),
                        "ms_per_step": round(e2e_ms, 3)},
                "gpu_launches": int(pred.launches_per_replay * replays), "model_tflops": round(tf, 2),
                "model_frac_of_bf16_peak": round(tf / world / pk["tflops"], 4),
                # no per-family CUDA-event leg in this workload: the WHOLE step's algorithmic FLOP rate against the measured bf16 peak
                "roofline": {"bound": "tensor", "kernel": "whole inference step (conv_halo2 / conv_tc / lstm_seq64 kernels + the elementwise passes)",
                             "achieved": round(tf / world, 2), "peak": pk["tflops"], "unit": "TFLOP/s",
                             "frac": round(tf / world / pk["tflops"], 4), "traffic": None, "peak_source": pk["src"], "whole_step": True},
                "cpu_baseline": None}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-on-this-GPU (eager cuDNN) leg")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--profile-detail", action="store_true", help="print the slowest GEMM-family launches to stderr")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "train512", "volume", "unet"],
                    help="train = BASELINE.json configs[2] sharded 16/GPU (the headline metric); infer = configs[1], eval forward; "
                         "train512 = configs[3] (T=16 x 512x512, 8 slices/GPU); volume = configs[4] (160-slice case, sharded by slice); "
                         "unet = configs[0] (UNet fp32, batch 4)")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "train512":
        # configs[3]: long-sequence, high-resolution stress.  8 slices per GPU: one step's activations are ~23 GB, and
        # the per-pixel LSTMs see 4x the rows and 2x the steps of configs[2].  2 036.11 GFLOP per slice.
        global T_PHASES, HW, BATCH_PER_GPU, TRAIN_GFLOP_PER_SLICE, CONFIG_NOTE
        T_PHASES, HW, BATCH_PER_GPU, TRAIN_GFLOP_PER_SLICE = 16, 512, 8, 2036.11
        CONFIG_NOTE = "BASELINE.json configs[3]: T=16 phases x 512x512, per-GPU batch chosen by memory"
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    if args.workload == "infer":
        run_infer(args, rank, world, local_rank)
        return
    if args.workload == "volume":
        run_volume(args, rank, world, local_rank)
        return
    if args.workload == "unet":
        run_unet(args, rank, world, local_rank)
        return
    run_own(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
