"""Launch a few representative GEMM-family kernels in isolation (for ncu captures and quick CUDA-event timing).

    python tools/kernel_probe.py [conv|wgrad|bn] [--iters N]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stf_unet_b200 import ops  # noqa: E402

DEV = "cuda"
CONV_SHAPES = [  # N, H, W, Cin, Cout, k  (encoder shapes at B=16, T=8, 256x256)
    (128, 64, 64, 64, 64, 3), (128, 32, 32, 128, 128, 3), (128, 16, 16, 256, 256, 3), (128, 8, 8, 512, 512, 3),
    (16, 64, 64, 64, 256, 1), (128, 64, 64, 64, 256, 1)]


def timeit(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="conv")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nmul", type=int, default=1, help="multiply the batch of every shape (scaling experiments)")
    a = ap.parse_args()
    global CONV_SHAPES
    CONV_SHAPES = [(n * a.nmul, h, w, ci, co, k) for (n, h, w, ci, co, k) in CONV_SHAPES]
    bf = torch.bfloat16
    if a.what in ("conv", "all"):
        for (N, H, W, Cin, Cout, k) in CONV_SHAPES:
            x = torch.randn(N, H, W, Cin, device=DEV).to(bf)
            w = (torch.randn(Cout, Cin, k, k, device=DEV) * 0.05)
            wp = ops.pack_weight(w, True, bf, n_major=True)
            y = torch.empty(N, H, W, Cout, device=DEV, dtype=bf)
            ms = timeit(lambda: ops.conv2d(x, wp, Cout, k, 1, (k - 1) // 2, out=y, impl=ops.IMPL_TCGEN05), a.iters)
            fl = 2.0 * N * H * W * Cin * Cout * k * k
            print(f"conv   N{N} {H}x{W} {Cin}->{Cout} k{k}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s")
    if a.what in ("wgrad", "all"):
        for (N, H, W, Cin, Cout, k) in CONV_SHAPES[:4]:
            x = torch.randn(N, H, W, Cin, device=DEV).to(bf)
            dy = torch.randn(N, H, W, Cout, device=DEV).to(bf)
            dW = torch.zeros(Cout, Cin, k, k, device=DEV)
            acc = torch.zeros(Cout * Cin * k * k, device=DEV)      # deferred mode: accumulate only (what the engine does)
            ms = timeit(lambda: ops.conv2d_wgrad(dy, x, dW, k, 1, (k - 1) // 2, 0, Cin, impl=ops.IMPL_TCGEN05, acc=acc), a.iters)
            fl = 2.0 * N * H * W * Cin * Cout * k * k
            print(f"wgrad  N{N} {H}x{W} {Cin}->{Cout} k{k}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s")
    if a.what in ("lstm", "all"):
        for (B, h, w, C) in [(16, 64, 64, 64), (16, 32, 32, 128), (16, 16, 16, 256), (16, 8, 8, 512)]:
            R = B * h * w
            xt = torch.randn(B, h, w, C, device=DEV).to(bf)
            hp = torch.randn(B, h, w, C, device=DEV).to(bf)
            wih = torch.randn(4 * C, C, device=DEV) / C ** 0.5
            whh = torch.randn(4 * C, C, device=DEV) / C ** 0.5
            bih, bhh = torch.randn(4 * C, device=DEV) * 0.1, torch.randn(4 * C, device=DEV) * 0.1
            cp = torch.randn(R, C, device=DEV)
            wp = ops.pack_lstm_xh(wih, whh, bf)
            c_out = torch.empty(R, C, device=DEV)
            h_out = torch.empty(B, h, w, C, device=DEV, dtype=bf)
            acts = torch.empty(B, h, w, 4 * C, device=DEV, dtype=bf)
            ms = timeit(lambda: ops.lstm_step_fused(xt, hp, wp, bih, bhh, cp, c_out, h_out, acts), a.iters)
            nbytes = R * C * (2 + 2 + 4 + 4 + 2 + 8)
            print(f"lstm   rows{R} C{C}: {ms * 1e3:8.1f} us  {2.0 * R * 2 * C * 4 * C / ms / 1e9:8.1f} TFLOP/s  {nbytes / ms / 1e6:8.1f} GB/s")
    if a.what in ("bn", "all"):
        for (N, H, W, C) in [(128, 128, 128, 64), (128, 64, 64, 64), (128, 32, 32, 128), (128, 16, 16, 256), (128, 8, 8, 512)]:
            G, R = 8, (N // 8) * H * W
            x = torch.randn(N, H, W, C, device=DEV).to(bf)
            dy = torch.randn_like(x)
            gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
            S = x.numel() * 2
            ms = timeit(lambda: ops.bn_stats(x, G, R, C), a.iters)
            st = ops.bn_finalize_train(ops.bn_stats(x, G, R, C), gamma, beta, None, None, None, G, R, C)
            y = torch.empty_like(x)
            ms2 = timeit(lambda: ops.bn_apply(x, st[0], st[1], G, R, C, True, x, out=y), a.iters)
            dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            ms3 = timeit(lambda: ops.bn_bwd(dy, y, x, st[2], st[3], gamma, dg, db, G, R, C, True, True), a.iters)
            print(f"bn     N{N} {H}x{W} C{C}: stats {ms * 1e3:7.1f} us ({S / ms / 1e6:7.1f} GB/s)  apply+res {ms2 * 1e3:7.1f} us "
                  f"({4 * S / ms2 / 1e6:7.1f} GB/s)  bwd(reduce+fin+apply) {ms3 * 1e3:7.1f} us ({8 * S / ms3 / 1e6:7.1f} GB/s)")
    if a.what == "pool":
        def gtime(fn, iters):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        N, H, W, C = 128 * a.nmul, 128, 128, 64
        x = torch.randn(N, H, W, C, device=DEV).to(bf)
        y, idx = ops.maxpool_fwd_idx(x, 3, 2, 1)
        dy = torch.randn_like(y)
        t1 = gtime(lambda: ops.maxpool_fwd_idx(x, 3, 2, 1), a.iters)
        t2 = gtime(lambda: ops.maxpool_bwd_idx(idx, dy, tuple(x.shape), 3, 2, 1), a.iters)
        t3 = gtime(lambda: ops.maxpool_fwd(x, 3, 2, 1), a.iters)
        xe, ye = x.numel() * 2 / 1e3, y.numel() * 2 / 1e3
        print(f"pool   N{N} {H}x{W} C{C}: fwd_idx {t1:6.1f} us ({(xe + 1.5 * ye) / t1:6.0f} GB/s)  bwd_idx {t2:6.1f} us "
              f"({(xe + 1.5 * ye) / t2:6.0f} GB/s)  fwd(eval) {t3:6.1f} us ({(xe + ye) / t3:6.0f} GB/s)")
        img = torch.randn(N, 256, 256, 1, device=DEV).to(bf)
        t4 = gtime(lambda: ops.im2col_small(img, 7, 2, 3, 64), a.iters)
        print(f"im2col N{N} 256x256 k7 s2 -> [{N * 128 * 128}, 64]: {t4:6.1f} us ({(img.numel() * 2 + N * 128 * 128 * 128) / 1e3 / t4:6.0f} GB/s)")
    if a.what == "bn2":
        # every BatchNorm kernel of the training step on the five encoder shapes, each timed as a CUDA graph of `iters`
        # back-to-back launches (no host launch overhead, L2 state as in a chain of kernels over the same tensors)
        def graph_time(fn, iters):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3      # us

        for (N, H, W, C) in [(128, 128, 128, 64), (128, 64, 64, 64), (128, 32, 32, 128), (128, 16, 16, 256), (128, 8, 8, 512)]:
            G, R = 8, (N // 8) * H * W
            x = torch.randn(N, H, W, C, device=DEV).to(bf)
            res = torch.randn_like(x)
            dy = torch.randn_like(x)
            gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
            rm, rv, nbt = torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros((), device=DEV, dtype=torch.long)
            S = x.numel() * 2 / 1e3     # KB of one tensor -> KB / us = GB/s
            part = ops.bn_stats(x, G, R, C)
            st = ops.bn_finalize_train(part, gamma, beta, rm, rv, nbt, G, R, C)
            y = torch.empty_like(x)
            dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            dres = torch.zeros_like(x)
            t_fin = graph_time(lambda: ops.bn_finalize_train(part, gamma, beta, rm, rv, nbt, G, R, C), a.iters)
            t_app = graph_time(lambda: ops.bn_apply(x, st[0], st[1], G, R, C, True, None, out=y), a.iters)
            t_apr = graph_time(lambda: ops.bn_apply(x, st[0], st[1], G, R, C, True, res, out=y), a.iters)
            t_bwd = graph_time(lambda: ops.bn_bwd(dy, y, x, st[2], st[3], gamma, dg, db, G, R, C, True, False, scale=st[0], shift=st[1]), a.iters)
            t_bwr = graph_time(lambda: ops.bn_bwd(dy, y, x, st[2], st[3], gamma, dg, db, G, R, C, True, True, dres_acc=dres), a.iters)
            nsc = ops.bn_bwd_scratch_floats(G, C)
            arena = torch.zeros(nsc * a.iters * 2 + 64, device=DEV)       # one zeroed scratch per launch inside the graph
            cnt = [0]

            def fused(res_mode):
                k = cnt[0] % (2 * a.iters)
                cnt[0] += 1
                sc = arena[k * nsc:(k + 1) * nsc]
                if res_mode:
                    ops.bn_bwd(dy, y, x, st[2], st[3], gamma, dg, db, G, R, C, True, True, dres_acc=dres, scratch=sc)
                else:
                    ops.bn_bwd(dy, y, x, st[2], st[3], gamma, dg, db, G, R, C, True, False, scale=st[0], shift=st[1], scratch=sc)

            def timed_fused(res_mode):
                # the graph consumes 2 * iters scratch slices (warm-up replay + timed replay): zero them, capture, time
                def fn():
                    fused(res_mode)
                cnt[0] = 0
                torch.cuda.synchronize()
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    fn()
                torch.cuda.synchronize()
                cnt[0] = 0
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(a.iters):
                        fn()
                arena.zero_()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / a.iters * 1e3

            t_fu, t_fur = timed_fused(False), timed_fused(True)
            print(f"bn2    N{N} {H}x{W} C{C}: ONE-LAUNCH bwd {t_fu:6.1f} us ({5 * S / t_fu:6.0f} GB/s)  +res {t_fur:6.1f} us ({10 * S / t_fur:6.0f} GB/s)")
            print(f"bn2    N{N} {H}x{W} C{C}: fin {t_fin:5.1f} us | apply {t_app:6.1f} us ({2 * S / t_app:6.0f} GB/s)  +res {t_apr:6.1f} us "
                  f"({3 * S / t_apr:6.0f} GB/s) | bwd chain {t_bwd:6.1f} us ({5 * S / t_bwd:6.0f} GB/s)  +res {t_bwr:6.1f} us ({10 * S / t_bwr:6.0f} GB/s)")


if __name__ == "__main__":
    main()
