"""stf_unet_b200 -- B200-native (sm_100a) implementation of the STF-Unet hot path.

Public surface mirrors the reference's ``src`` package (/root/reference/src/__init__.py:2-3) plus the loss:

    from stf_unet_b200 import STFLSTMUNet, UNet, criterion
"""
from .stf_lstm_unet import STFLSTMUNet
from .unet import UNet
from .loss import criterion, ce_dice, dice_loss
from .optim import FlatAdamW

__all__ = ["STFLSTMUNet", "UNet", "criterion", "ce_dice", "dice_loss", "FlatAdamW"]
