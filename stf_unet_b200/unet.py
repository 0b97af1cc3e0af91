"""UNet on libstfb200 -- drop-in for /root/reference/src/unet.py:5-57.

Same constructor / forward signature, ``{"out": logits[B, classes, H, W]}`` return, ``input_format =
"flat_channels"`` class attribute (train_utils/train_and_eval.py:10-14 keys on it) and ``state_dict`` layout.
"""
from __future__ import annotations

import torch

from . import engine, ops
from .modules import B200Module, BNParams, ConvParams, seq


def _block(cin, cout):
    return seq(**{"0": ConvParams(cin, cout, 3, bias=True), "1": BNParams(cout),
                  "3": ConvParams(cout, cout, 3, bias=True), "4": BNParams(cout)})


class UNet(B200Module):
    input_format = "flat_channels"

    def __init__(self, in_channels=8, num_classes=2, base_c=64):
        super().__init__()
        c = base_c
        self.enc1 = _block(in_channels, c)
        self.enc2 = _block(c, 2 * c)
        self.enc3 = _block(2 * c, 4 * c)
        self.enc4 = _block(4 * c, 8 * c)
        self.bottleneck = _block(8 * c, 16 * c)
        self.up4 = ConvParams(16 * c, 8 * c, 2, bias=True, transposed=True)
        self.dec4 = _block(16 * c, 8 * c)
        self.up3 = ConvParams(8 * c, 4 * c, 2, bias=True, transposed=True)
        self.dec3 = _block(8 * c, 4 * c)
        self.up2 = ConvParams(4 * c, 2 * c, 2, bias=True, transposed=True)
        self.dec2 = _block(4 * c, 2 * c)
        self.up1 = ConvParams(2 * c, c, 2, bias=True, transposed=True)
        self.dec1 = _block(2 * c, c)
        self.out_conv = ConvParams(c, num_classes, 1, bias=True)

    def forward(self, x):
        return self._call(x)

    @staticmethod
    def _double_conv(ex, x, p, x2=None):
        o = ex.conv_bn(x, p + ".0.weight", p + ".1", k=3, pad=1, relu=True, bname=p + ".0.bias", x2=x2)
        return ex.conv_bn(o, p + ".3.weight", p + ".4", k=3, pad=1, relu=True, bname=p + ".3.bias")

    def _forward_impl(self, ex, x):
        if x.dim() != 4:
            raise ValueError(f"UNet expects [B, C, H, W], got {tuple(x.shape)}")
        if not x.is_floating_point():
            raise TypeError("UNet takes normalised float input (the 8-bit input path belongs to STFLSTMUNet)")
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise ValueError("UNet needs H and W divisible by 16 (the reference's torch.cat fails otherwise)")
        xin = engine.Var(ops.nchw_to_nhwc(x, ex.dtype), needs_grad=False)
        e1 = self._double_conv(ex, xin, "enc1")
        e2 = self._double_conv(ex, ex.maxpool(e1, 2, 2, 0), "enc2")
        e3 = self._double_conv(ex, ex.maxpool(e2, 2, 2, 0), "enc3")
        e4 = self._double_conv(ex, ex.maxpool(e3, 2, 2, 0), "enc4")
        d = self._double_conv(ex, ex.maxpool(e4, 2, 2, 0), "bottleneck")
        for k, skip in ((4, e4), (3, e3), (2, e2), (1, e1)):
            u = ex.conv(d, f"up{k}.weight", k=2, stride=2, pad=0, transposed=True, bname=f"up{k}.bias")
            d = self._double_conv(ex, u, f"dec{k}", x2=skip)
        return ex.conv(d, "out_conv.weight", k=1, bname="out_conv.bias", y_dtype=torch.float32)
