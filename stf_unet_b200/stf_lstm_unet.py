"""STFLSTMUNet on libstfb200 -- drop-in for /root/reference/src/stf_lstm_unet.py:89-256.

Same constructor and ``forward(x, pk_maps=None)`` signature, same ``{'out': logits[B, classes, H/2, W/2]}`` return,
same ``state_dict`` keys/shapes, no ``input_format`` attribute (the reference resolves to
"time_sequence" through the getattr default at train_utils/train_and_eval.py:10).

Differences that are deliberate (B200-first, not a port):
  * the T encoder passes run as ONE batch of T*B images in time-major order; BatchNorm statistics are still
    reduced per time step and the running stats get T sequential updates, exactly like the reference's loop;
  * activations are NHWC (bf16 under autocast, fp32 otherwise); the permute/reshape copies around nn.LSTM vanish
    because a pixel's time series is addressable in place;
  * eval mode folds every BatchNorm (+ReLU, +residual) into the conv epilogue.
"""
from __future__ import annotations

import torch.nn as nn

from . import engine, ops
from .modules import B200Module, BNParams, ConvParams, LSTMParams, seq

RESNET34_LAYERS = ((64, 3), (128, 4), (256, 6), (512, 3))


class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.conv1 = ConvParams(cin, cout, 3, bias=False, resnet_init=True)
        self.bn1 = BNParams(cout)
        self.conv2 = ConvParams(cout, cout, 3, bias=False, resnet_init=True)
        self.bn2 = BNParams(cout)
        if down:
            self.downsample = seq(**{"0": ConvParams(cin, cout, 1, bias=False, resnet_init=True), "1": BNParams(cout)})


class _ResidualConvBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_block = seq(**{"0": ConvParams(c, c, 3, bias=False), "1": BNParams(c),
                                 "3": ConvParams(c, c, 3, bias=False), "4": BNParams(c)})


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.up = ConvParams(cin, cout, 3, bias=True, transposed=True)
        self.fusion = ConvParams(cout + cskip, cout, 1, bias=True)
        self.res_conv = _ResidualConvBlock(cout)


class STFLSTMUNet(B200Module):
    #: normalisation the reference's loader applies on the CPU (/root/reference/train.py:147-148); used when the series is
    #: handed over as raw 8-bit grey levels (uint8 [B, T, 1, H, W]): the device then does ToTensor + Normalize itself,
    #: fused into the layout pass.  Float input is taken as already normalised.
    input_mean = 0.709
    input_std = 0.127

    @staticmethod
    def _early_pack(name):
        return name.startswith("conv1.")          # the stem; every other weight is packed beside it

    def __init__(self, in_channels=1, num_classes=2, time_steps=8, use_pk_maps=False, pk_channels=3):
        super().__init__()
        self.time_steps = time_steps
        self.use_pk_maps = use_pk_maps
        self.pk_channels = pk_channels if use_pk_maps else 0
        cin = in_channels + self.pk_channels
        self.conv1 = ConvParams(cin, 64, 7, bias=False)
        self.bn1 = BNParams(64)
        prev = 64
        for li, (c, n) in enumerate(RESNET34_LAYERS, start=1):
            blocks = [_BasicBlock(prev if b == 0 else c, c, down=(b == 0 and li > 1)) for b in range(n)]
            setattr(self, f"layer{li}", nn.ModuleList(blocks))
            prev = c
        if use_pk_maps:
            for k, c in enumerate((64, 128, 256, 512), start=1):
                setattr(self, f"pk_fusion{k}", ConvParams(c + pk_channels, c, 1, bias=True))
        for k, c in enumerate((64, 128, 256, 512), start=1):
            setattr(self, f"lstm{k}", LSTMParams(c))
        self.decoder4 = _DecoderBlock(512, 256, 256)
        self.decoder3 = _DecoderBlock(256, 128, 128)
        self.decoder2 = _DecoderBlock(128, 64, 64)
        self.upconv1 = ConvParams(64, 32, 3, bias=True, transposed=True)
        self.final_res = _ResidualConvBlock(32)
        self.final = ConvParams(32, num_classes, 1, bias=True)

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, x, pk_maps=None):  # pk_maps is ignored, as in the reference (:146-160)
        return self._call(x)

    @staticmethod
    def _residual_block(ex, x, p):
        o = ex.conv_bn(x, p + ".conv_block.0.weight", p + ".conv_block.1", k=3, pad=1, relu=True)
        return ex.conv_bn(o, p + ".conv_block.3.weight", p + ".conv_block.4", k=3, pad=1, relu=True, residual=x)

    def _decoder(self, ex, x, skip, p):
        u = ex.conv(x, p + ".up.weight", k=3, stride=2, pad=1, transposed=True, out_pad=1, bname=p + ".up.bias")
        if u.data.shape[1:3] != skip.data.shape[1:3]:
            u = ex.resize(u, skip.data.shape[1], skip.data.shape[2])
        f = ex.conv(u, p + ".fusion.weight", k=1, bname=p + ".fusion.bias", x2=skip)
        return self._residual_block(ex, f, p + ".res_conv")

    def _forward_impl(self, ex, x):
        if x.dim() != 5:
            raise ValueError(f"STFLSTMUNet expects [B, T, C, H, W], got {tuple(x.shape)}")
        B, total, C, H, W = x.shape
        pk = None
        if self.use_pk_maps:
            T = total - self.pk_channels
            if T < 1 or C != 1:
                raise ValueError("use_pk_maps needs T + pk_channels steps of single-channel images")
            if not x.is_floating_point():
                raise TypeError("use_pk_maps: the PK maps are real-valued; pass a float tensor")
            maps = x[:, T:, 0].contiguous()                                   # [B, pk, H, W] (reference :149-153)
            series = x[:, :T].contiguous()
            pk = ops.nchw_to_nhwc(maps, ex.dtype)                             # [B, H, W, pk], resized per scale below
            xin = engine.Var(ops.pack_series_maps(series, maps, ex.dtype), needs_grad=False)   # cat([x_t, pk]) for all t
        elif x.dtype == __import__("torch").uint8:
            T = total
            xin = engine.Var(ops.pack_series_u8(x, ex.dtype, self.input_mean, self.input_std), needs_grad=False)
        else:
            T = total
            xin = engine.Var(ops.pack_series(x, ex.dtype), needs_grad=False)  # [T*B, H, W, C] time-major
        ex.begin_split(T)      # two half-batch chains through the encoder (training forward, bf16): engine.USE_FWD_SPLIT
        if ex.train:     # stem: conv (statistics in the epilogue) -> BatchNorm + ReLU + max-pool in one pass
            raw = ex.conv(xin, "conv1.weight", k=7, stride=2, pad=3, stats_G=T)
            e = ex.bn_relu_pool(raw, "bn1", T, 3, 2, 1)
        else:
            s = ex.conv_bn(xin, "conv1.weight", "bn1", k=7, stride=2, pad=3, relu=True, G=T)
            e = ex.maxpool(s, 3, 2, 1)
        feats = []
        for li, (c, n) in enumerate(RESNET34_LAYERS, start=1):
            if li >= 3:        # layers 3 and 4 hold 73 % of the parameters: their gradients are exchanged while layers 2, 1 run
                ex.mark_grad_segment(f"layer{li}.0.conv1.weight")
            for b in range(n):
                p = f"layer{li}.{b}"
                down = b == 0 and li > 1
                o = ex.conv_bn(e, p + ".conv1.weight", p + ".bn1", k=3, stride=2 if down else 1, pad=1, relu=True, G=T)
                idt = e
                if down:
                    idt = ex.conv_bn(e, p + ".downsample.0.weight", p + ".downsample.1", k=1, stride=2, pad=0,
                                     relu=False, G=T)
                e = ex.conv_bn(o, p + ".conv2.weight", p + ".bn2", k=3, pad=1, relu=True, residual=idt, G=T)
            feats.append(e)
        ex.end_split()
        # everything registered after the encoder (PK fusion, LSTMs, decoder, head) is final once its backward is done
        ex.mark_grad_segment("pk_fusion1.weight" if pk is not None else "lstm1.weight_ih_l0")
        if pk is not None:
            feats = [self._pk_fuse(ex, f, pk, k + 1, T) for k, f in enumerate(feats)]
        enc = ex.lstm_levels(feats, [f"lstm{k + 1}" for k in range(len(feats))], T)
        d = self._decoder(ex, enc[3], enc[2], "decoder4")
        d = self._decoder(ex, d, enc[1], "decoder3")
        d = self._decoder(ex, d, enc[0], "decoder2")
        d = ex.conv(d, "upconv1.weight", k=3, stride=2, pad=1, transposed=True, out_pad=1, bname="upconv1.bias")
        d = self._residual_block(ex, d, "final_res")
        return ex.conv(d, "final.weight", k=1, bname="final.bias", y_dtype=__import__("torch").float32)

    # -- PK-map branch (use_pk_maps=True; reference :189-200) -----------------------------------------
    def _pk_fuse(self, ex, f, pk, k, T):
        """e_k = pk_fusion_k(cat([e_k, bilinear(pk_maps -> e_k size, align_corners=True)])) for every time step."""
        h, w = f.data.shape[1], f.data.shape[2]
        pk_k = ops.bilinear_fwd(pk, h, w)                                      # [B, h, w, pk]
        pk_t = engine.Var(ops.repeat_batch(pk_k, T), needs_grad=False)         # same maps at every time step
        return ex.conv(f, f"pk_fusion{k}.weight", k=1, bname=f"pk_fusion{k}.bias", x2=pk_t)
