"""TEST INFRASTRUCTURE ONLY -- deterministic weights and synthetic DCE-MRI series.

The reference pins no checkpoints and a full state_dict is
~110 MB, so weights are regenerated from a seed on every machine.  Names and
shapes follow the reference's ``state_dict`` contract (SURVEY.md Appendix B;
``/root/reference/src/stf_lstm_unet.py:90-137``, ``/root/reference/src/unet.py:7-37``)
and are cross-checked key-by-key against the live reference module in
``tests/golden/make_golden.py``.

numpy's PCG64 stream is platform independent, which is what lets the golden
fixtures generated in the build container be reproduced on the GPU box.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch


# --------------------------------------------------------------------------
# state_dict specs
# --------------------------------------------------------------------------
def _bn(prefix, c):
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
            (prefix + ".running_mean", (c,), "bn_rm"), (prefix + ".running_var", (c,), "bn_rv"),
            (prefix + ".num_batches_tracked", (), "bn_nbt")]


RESNET34_LAYERS = ((64, 3), (128, 4), (256, 6), (512, 3))  # torchvision resnet34 [3,4,6,3]


def stf_param_spec(in_channels=1, num_classes=2, use_pk_maps=False, pk_channels=3):
    """(name, shape, kind) for every state_dict entry of STFLSTMUNet.

    Follows /root/reference/src/stf_lstm_unet.py:90-137 (registration order)."""
    cin = in_channels + (pk_channels if use_pk_maps else 0)
    spec = [("conv1.weight", (64, cin, 7, 7), "conv")] + _bn("bn1", 64)
    prev = 64
    for li, (c, nblocks) in enumerate(RESNET34_LAYERS, start=1):
        for b in range(nblocks):
            p = f"layer{li}.{b}"
            spec.append((p + ".conv1.weight", (c, prev if b == 0 else c, 3, 3), "conv"))
            spec += _bn(p + ".bn1", c)
            spec.append((p + ".conv2.weight", (c, c, 3, 3), "conv"))
            spec += _bn(p + ".bn2", c)
            if b == 0 and li > 1:
                spec.append((p + ".downsample.0.weight", (c, prev, 1, 1), "conv"))
                spec += _bn(p + ".downsample.1", c)
        prev = c
    if use_pk_maps:
        for k, c in enumerate((64, 128, 256, 512), start=1):
            spec.append((f"pk_fusion{k}.weight", (c, c + pk_channels, 1, 1), "conv"))
            spec.append((f"pk_fusion{k}.bias", (c,), "bias"))
    for k, c in enumerate((64, 128, 256, 512), start=1):
        spec.append((f"lstm{k}.weight_ih_l0", (4 * c, c), "lstm"))
        spec.append((f"lstm{k}.weight_hh_l0", (4 * c, c), "lstm"))
        spec.append((f"lstm{k}.bias_ih_l0", (4 * c,), "lstm"))
        spec.append((f"lstm{k}.bias_hh_l0", (4 * c,), "lstm"))
    for name, cin_d, cskip, cout in (("decoder4", 512, 256, 256), ("decoder3", 256, 128, 128),
                                     ("decoder2", 128, 64, 64)):
        spec.append((name + ".up.weight", (cin_d, cout, 3, 3), "convT"))
        spec.append((name + ".up.bias", (cout,), "bias"))
        spec.append((name + ".fusion.weight", (cout, cout + cskip, 1, 1), "conv"))
        spec.append((name + ".fusion.bias", (cout,), "bias"))
        spec.append((name + ".res_conv.conv_block.0.weight", (cout, cout, 3, 3), "conv"))
        spec += _bn(name + ".res_conv.conv_block.1", cout)
        spec.append((name + ".res_conv.conv_block.3.weight", (cout, cout, 3, 3), "conv"))
        spec += _bn(name + ".res_conv.conv_block.4", cout)
    spec.append(("upconv1.weight", (64, 32, 3, 3), "convT"))
    spec.append(("upconv1.bias", (32,), "bias"))
    spec.append(("final_res.conv_block.0.weight", (32, 32, 3, 3), "conv"))
    spec += _bn("final_res.conv_block.1", 32)
    spec.append(("final_res.conv_block.3.weight", (32, 32, 3, 3), "conv"))
    spec += _bn("final_res.conv_block.4", 32)
    spec.append(("final.weight", (num_classes, 32, 1, 1), "conv"))
    spec.append(("final.bias", (num_classes,), "bias"))
    return spec


def unet_param_spec(in_channels=8, num_classes=2, base_c=64):
    """(name, shape, kind) for UNet; /root/reference/src/unet.py:7-37."""
    c = base_c
    spec = []

    def block(name, ci, co):
        s = [(f"{name}.0.weight", (co, ci, 3, 3), "conv"), (f"{name}.0.bias", (co,), "bias")]
        s += _bn(f"{name}.1", co)
        s += [(f"{name}.3.weight", (co, co, 3, 3), "conv"), (f"{name}.3.bias", (co,), "bias")]
        s += _bn(f"{name}.4", co)
        return s

    spec += block("enc1", in_channels, c)
    spec += block("enc2", c, 2 * c)
    spec += block("enc3", 2 * c, 4 * c)
    spec += block("enc4", 4 * c, 8 * c)
    spec += block("bottleneck", 8 * c, 16 * c)
    for k, (ci, co) in zip((4, 3, 2, 1), ((16 * c, 8 * c), (8 * c, 4 * c), (4 * c, 2 * c), (2 * c, c))):
        spec.append((f"up{k}.weight", (ci, co, 2, 2), "convT"))
        spec.append((f"up{k}.bias", (co,), "bias"))
        spec += block(f"dec{k}", ci, co)
    spec.append(("out_conv.weight", (num_classes, c, 1, 1), "conv"))
    spec.append(("out_conv.bias", (num_classes,), "bias"))
    return spec


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(name.encode())]))


def make_state_dict(spec, seed=0):
    """Deterministic, non-degenerate weights for a spec.

    Not the reference's init (its RNG order cannot be reproduced portably); chosen
    so that eval-mode logits have real spread (SURVEY.md section 7 "hard parts" #1):
    He-normal convs, BN affine / running stats away from (1, 0, 0, 1)."""
    sd = OrderedDict()
    for name, shape, kind in spec:
        g = _rng(seed, name)
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            v = g.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        elif kind == "convT":  # [Cin, Cout, kh, kw]; each output sees ~kh*kw/4 taps (stride 2)
            fan_in = shape[0] * shape[2] * shape[3] / 4.0
            v = g.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        elif kind == "bias":
            v = g.uniform(-0.1, 0.1, shape)
        elif kind == "lstm":
            c = shape[-1] if len(shape) == 2 else shape[0] // 4
            v = g.uniform(-1.0, 1.0, shape) * (1.5 / np.sqrt(c))
        elif kind == "bn_w":
            v = g.uniform(0.6, 1.4, shape)
        elif kind == "bn_b":
            v = g.standard_normal(shape) * 0.15
        elif kind == "bn_rm":
            v = g.standard_normal(shape) * 0.15
        elif kind == "bn_rv":
            v = g.uniform(0.6, 1.4, shape)
        elif kind == "bn_nbt":
            sd[name] = torch.zeros((), dtype=torch.int64)
            continue
        else:
            raise ValueError(kind)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


# --------------------------------------------------------------------------
# synthetic DCE-MRI series
# --------------------------------------------------------------------------
def synthetic_dce_batch(batch, T, H, W, seed=1234, half_res_target=True, channels=1):
    """x [B,T,C,H,W] float32 (normalised like /root/reference/train.py:147-148) and
    target int64 {0,1} ([B,H/2,W/2] for STF, whose logits are half resolution --
    /root/reference/src/stf_lstm_unet.py:245-256; full-res otherwise)."""
    g = np.random.Generator(np.random.PCG64(int(seed)))
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    x = np.empty((batch, T, channels, H, W), dtype=np.float32)
    tgt = np.empty((batch, H, W), dtype=np.int64)
    tt = np.arange(T, dtype=np.float32)
    for b in range(batch):
        phase = g.uniform(0, 2 * np.pi)
        base = 0.5 + 0.2 * np.sin(xx / 17.0 + phase) * np.cos(yy / 23.0)
        cy = g.uniform(H * 0.25, H * 0.75)
        cx = g.uniform(W * 0.25, W * 0.75)
        r = g.uniform(H / 16.0, H / 6.0)
        mask = ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.float32)
        for t in range(T):
            enh = mask * 0.4 * (1 - np.exp(-tt[t] / 2.0)) + 0.05 * (1 - np.exp(-tt[t] / 4.0))
            for c in range(channels):
                img = base + enh + g.standard_normal((H, W)).astype(np.float32) * 0.02
                img = np.clip(img, 0.0, 1.0)
                x[b, t, c] = (img - 0.709) / 0.127
        tgt[b] = mask.astype(np.int64)
    if half_res_target:
        tgt = tgt[:, ::2, ::2]
    return torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(tgt))
