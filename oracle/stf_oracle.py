"""TEST INFRASTRUCTURE ONLY -- functional fp32 restatement of the reference hot path.

The reference's arithmetic lives entirely in third-party libraries that are not
under /root/reference: ``torch`` (2.11.0+cu128 in this image; the reference pins
no version) and ``torchvision.models.resnet34`` (0.26.0).  This module restates
the reference's *composition* of those operators as pure functions over a plain
``state_dict`` (no nn.Module, no torchvision import), each function citing the
reference lines it follows.  The per-pixel LSTM is written out gate by gate
instead of calling ``nn.LSTM``.  Checked against the live reference through the
fixtures of ``tests/golden/`` (see ``oracle/__init__.py``).

Never imported by the product package.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from .weights import RESNET34_LAYERS

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class _State:
    """Parameter lookup + functional BatchNorm with reference buffer semantics."""

    def __init__(self, sd, train):
        self.sd = sd
        self.train = train
        # running buffers are updated on a private copy (returned to the caller)
        self.buffers = OrderedDict()

    def p(self, name):
        return self.sd[name]

    def bn(self, x, prefix):
        """nn.BatchNorm2d(eps=1e-5, momentum=0.1): train -> batch mean / biased var for
        normalisation, unbiased var into running stats, num_batches_tracked += 1 per CALL
        (the STF encoder calls each BN T times per forward:
        /root/reference/src/stf_lstm_unet.py:168-186)."""
        rm_k, rv_k, nbt_k = prefix + ".running_mean", prefix + ".running_var", prefix + ".num_batches_tracked"
        if self.train:
            rm = self.buffers.get(rm_k, self.sd[rm_k]).clone()
            rv = self.buffers.get(rv_k, self.sd[rv_k]).clone()
            nbt = self.buffers.get(nbt_k, self.sd[nbt_k]).clone()
            y = F.batch_norm(x, rm, rv, self.sd[prefix + ".weight"], self.sd[prefix + ".bias"],
                             True, BN_MOMENTUM, BN_EPS)
            self.buffers[rm_k], self.buffers[rv_k], self.buffers[nbt_k] = rm, rv, nbt + 1
            return y
        return F.batch_norm(x, self.sd[rm_k], self.sd[rv_k], self.sd[prefix + ".weight"],
                            self.sd[prefix + ".bias"], False, BN_MOMENTUM, BN_EPS)


def _basic_block(st, x, p, stride, has_down):
    """torchvision BasicBlock.forward (not under /root/reference):
    conv3x3 -> bn -> relu -> conv3x3 -> bn -> (+ downsample(x)) -> relu."""
    out = F.conv2d(x, st.p(p + ".conv1.weight"), None, stride, 1)
    out = F.relu(st.bn(out, p + ".bn1"))
    out = F.conv2d(out, st.p(p + ".conv2.weight"), None, 1, 1)
    out = st.bn(out, p + ".bn2")
    if has_down:
        idt = F.conv2d(x, st.p(p + ".downsample.0.weight"), None, stride, 0)
        idt = st.bn(idt, p + ".downsample.1")
    else:
        idt = x
    return F.relu(out + idt)


def _residual_conv_block(st, x, p):
    """/root/reference/src/stf_lstm_unet.py:7-35 (in==out everywhere, identity shortcut)."""
    out = F.conv2d(x, st.p(p + ".conv_block.0.weight"), None, 1, 1)
    out = F.relu(st.bn(out, p + ".conv_block.1"))
    out = F.conv2d(out, st.p(p + ".conv_block.3.weight"), None, 1, 1)
    out = st.bn(out, p + ".conv_block.4")
    return F.relu(out + x)


def _decoder_block(st, x, skip, p):
    """/root/reference/src/stf_lstm_unet.py:51-68."""
    x = F.conv_transpose2d(x, st.p(p + ".up.weight"), st.p(p + ".up.bias"), stride=2, padding=1,
                           output_padding=1)
    if x.shape[2:] != skip.shape[2:]:
        x = F.interpolate(x, size=skip.shape[2:], mode="bilinear", align_corners=True)
    x = torch.cat([x, skip], dim=1)
    x = F.conv2d(x, st.p(p + ".fusion.weight"), st.p(p + ".fusion.bias"))
    return _residual_conv_block(st, x, p + ".res_conv")


def pixel_lstm_last(seq, w_ih, w_hh, b_ih, b_hh):
    """Per-pixel single-layer LSTM, zero (h0, c0), PyTorch gate order i,f,g,o; returns h_T.

    seq [B,T,C,h,w] -> [B,C,h,w].  Restates /root/reference/src/stf_lstm_unet.py:216-242
    (permute to [B*h*w, T, C], nn.LSTM(C, C, batch_first=True), take the last step)."""
    B, T, C, h, w = seq.shape
    rows = seq.permute(0, 3, 4, 1, 2).reshape(B * h * w, T, C)
    hs = rows.new_zeros(B * h * w, C)
    cs = rows.new_zeros(B * h * w, C)
    for t in range(T):
        gates = rows[:, t] @ w_ih.t() + b_ih + hs @ w_hh.t() + b_hh
        i, f, g, o = gates.split(C, dim=1)
        cs = torch.sigmoid(f) * cs + torch.sigmoid(i) * torch.tanh(g)
        hs = torch.sigmoid(o) * torch.tanh(cs)
    return hs.reshape(B, h, w, C).permute(0, 3, 1, 2)


def stf_forward(sd, x, train=False, use_pk_maps=False, pk_channels=3, return_buffers=False):
    """STFLSTMUNet.forward restated: /root/reference/src/stf_lstm_unet.py:139-256.

    x [B, T(+pk), C, H, W] -> logits [B, num_classes, H/2, W/2]."""
    st = _State(sd, train)
    B, total, C, H, W = x.shape
    if use_pk_maps:                                    # :146-156
        T = total - pk_channels
        pk = x[:, T:].reshape(B, pk_channels, C, H, W).squeeze(2)
        x = x[:, :T]
    else:
        T, pk = total, None
    feats = [[], [], [], []]
    for t in range(T):                                 # :168-206
        xt = x[:, t]
        if pk is not None:
            xt = torch.cat([xt, pk], dim=1)
        xt = F.conv2d(xt, st.p("conv1.weight"), None, 2, 3)
        xt = F.relu(st.bn(xt, "bn1"))
        e = F.max_pool2d(xt, 3, 2, 1)
        es = []
        for li, (c, nblocks) in enumerate(RESNET34_LAYERS, start=1):
            for b in range(nblocks):
                first_down = (b == 0 and li > 1)
                e = _basic_block(st, e, f"layer{li}.{b}", 2 if first_down else 1, first_down)
            es.append(e)
        if pk is not None:                             # :189-200
            for k in range(4):
                pkk = F.interpolate(pk, size=es[k].shape[2:], mode="bilinear", align_corners=True)
                es[k] = F.conv2d(torch.cat([es[k], pkk], dim=1), st.p(f"pk_fusion{k + 1}.weight"),
                                 st.p(f"pk_fusion{k + 1}.bias"))
        for k in range(4):
            feats[k].append(es[k])
    enc = []
    for k in range(4):                                 # :209-242
        seq = torch.stack(feats[k], dim=1)
        enc.append(pixel_lstm_last(seq, st.p(f"lstm{k + 1}.weight_ih_l0"), st.p(f"lstm{k + 1}.weight_hh_l0"),
                                   st.p(f"lstm{k + 1}.bias_ih_l0"), st.p(f"lstm{k + 1}.bias_hh_l0")))
    d = _decoder_block(st, enc[3], enc[2], "decoder4")  # :245-247
    d = _decoder_block(st, d, enc[1], "decoder3")
    d = _decoder_block(st, d, enc[0], "decoder2")
    d = F.conv_transpose2d(d, st.p("upconv1.weight"), st.p("upconv1.bias"), stride=2, padding=1,
                           output_padding=1)           # :250
    d = _residual_conv_block(st, d, "final_res")       # :251
    out = F.conv2d(d, st.p("final.weight"), st.p("final.bias"))  # :254
    return (out, st.buffers) if return_buffers else out


def unet_forward(sd, x, train=False, return_buffers=False):
    """UNet.forward restated: /root/reference/src/unet.py:39-57; conv_block :10-18."""
    st = _State(sd, train)

    def block(x, p):
        x = F.conv2d(x, st.p(p + ".0.weight"), st.p(p + ".0.bias"), 1, 1)
        x = F.relu(st.bn(x, p + ".1"))
        x = F.conv2d(x, st.p(p + ".3.weight"), st.p(p + ".3.bias"), 1, 1)
        return F.relu(st.bn(x, p + ".4"))

    e1 = block(x, "enc1")
    e2 = block(F.max_pool2d(e1, 2), "enc2")
    e3 = block(F.max_pool2d(e2, 2), "enc3")
    e4 = block(F.max_pool2d(e3, 2), "enc4")
    d = block(F.max_pool2d(e4, 2), "bottleneck")
    for k, skip in ((4, e4), (3, e3), (2, e2), (1, e1)):
        d = F.conv_transpose2d(d, st.p(f"up{k}.weight"), st.p(f"up{k}.bias"), stride=2)
        d = block(torch.cat([d, skip], dim=1), f"dec{k}")
    out = F.conv2d(d, st.p("out_conv.weight"), st.p("out_conv.bias"))
    return (out, st.buffers) if return_buffers else out


def dice_loss(logits, target, eps=1e-6, ignore_index=-100):
    """1 - mean_c mean_b (2*sum(p*t)+eps)/(sum(p)+sum(t)+eps), p = softmax(logits), sums over the pixels whose
    label is not ignore_index.

    /root/reference/train_utils/dice_coefficient_loss.py:5-55 with multiclass=True.  build_target (:5-17) writes
    ignore_index into every channel of an ignored pixel and dice_coeff (:27-31) drops those pixels from both x
    and t; a negative ignore_index makes both branches dead.  The ``sets_sum == 0`` branch (:34-35) fires only for
    an image with no valid pixel (softmax sums to one elsewhere): then sets_sum = 2*inter = 0 and d += eps/eps."""
    C = logits.shape[1]
    p = torch.softmax(logits.float(), dim=1)
    if ignore_index >= 0:
        valid = target != ignore_index
        t = F.one_hot(torch.where(valid, target, torch.zeros_like(target)), C).permute(0, 3, 1, 2).to(p.dtype)
        m = valid.unsqueeze(1).to(p.dtype)
        p, t = p * m, t * m
    else:
        t = F.one_hot(target, C).permute(0, 3, 1, 2).to(p.dtype)
    inter = (p * t).sum(dim=(2, 3))
    sets = p.sum(dim=(2, 3)) + t.sum(dim=(2, 3))
    dice = torch.where(sets == 0, torch.ones_like(sets), (2 * inter + eps) / (sets + eps))   # [B, C]
    return 1 - dice.mean(dim=0).mean()


def criterion(logits, target, loss_weight=None, dice=True, ignore_index=-100):
    """CE(weighted mean over the non-ignored pixels) + Dice: /root/reference/train_utils/train_and_eval.py:299-313
    for the single 'out' head."""
    loss = F.cross_entropy(logits.float(), target, ignore_index=ignore_index, weight=loss_weight)
    if dice:
        loss = loss + dice_loss(logits, target, ignore_index=ignore_index)
    return loss


def loss_and_grads(sd, x, target, model="stf", train=True, **kw):
    """fwd + criterion + bwd through the oracle; returns (logits, loss, grads, new_buffers)."""
    params = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
            params[k] = v.detach().clone().requires_grad_(True)
        else:
            params[k] = v
    fwd = stf_forward if model == "stf" else unet_forward
    logits, bufs = fwd(params, x, train=train, return_buffers=True, **kw)
    loss = criterion(logits, target)
    names = [k for k, v in params.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    grads = OrderedDict((k, g) for k, g in zip(names, gs))
    return logits.detach(), loss.detach(), grads, bufs


# ---------------------------------------------------------------------------------------------------------------------
# evaluation metrics (test infrastructure, like everything in oracle/): restatement of ConfusionMatrix.update
# (train_utils/train_and_eval.py:30-39) and DiceCoefficient.update (:80-118) for ONE batch, as integer counts
# ---------------------------------------------------------------------------------------------------------------------
def eval_metrics_batch(logits, target, num_classes=2, ignore_index=255):
    """-> (confmat int64 [C,C] with rows = target, dice_per_class float [C]) of one update() call.

    The confusion matrix counts every pixel whose target is a valid class (:36-38); the Dice update first turns
    ignore_index pixels into class 0 on BOTH sides (`pred * mask`, `target * mask`, :88-91), then per class
    dice = 2*|pred==c & target==c| / (|pred==c| + |target==c|), 1.0 for an empty union (:103-107)."""
    pred = torch.argmax(torch.softmax(logits.float(), dim=1), dim=1)
    a, b = target.flatten(), pred.flatten()
    k = (a >= 0) & (a < num_classes)
    inds = num_classes * a[k].to(torch.int64) + b[k]
    mat = torch.bincount(inds, minlength=num_classes ** 2).reshape(num_classes, num_classes)
    p, t = pred, target
    if ignore_index is not None:
        m = target != ignore_index
        p, t = pred * m, target * m
    p, t = p.reshape(-1), t.reshape(-1)
    dice = []
    for c in range(num_classes):
        pc, tc = (p == c).double(), (t == c).double()
        union = pc.sum() + tc.sum()
        dice.append((2.0 * (pc * tc).sum() / union).item() if union > 0 else 1.0)
    return mat, torch.tensor(dice, dtype=torch.float64)
