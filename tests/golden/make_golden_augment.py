"""Golden fixture for the paired augmentation: the UNMODIFIED reference transforms (/root/reference/transforms.py) composed as
train.py:58-67 does (get_transform(train=True)) and train.py:70-75 (validation), applied to synthetic 8-bit DCE series.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_augment.py

Shared randomness: the reference calls its Compose once per phase with independent draws (my_dataset.py:173-179, a bug);
here Python's `random` is re-seeded with the SAME seed before every phase of a sample, so all phases and the mask see one
set of draws -- which is what the B200 pipeline implements.  The script also asserts that oracle/augment_oracle.py reproduces
every output bit for bit (including its replay of the draw order) before it writes the fixture."""
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import transforms as T  # noqa: E402  (reference)

from oracle import augment_oracle as AO  # noqa: E402

MEAN, STD = 0.709, 0.127


def train_pipeline():       # train.py:58-67
    base_size, crop_size = 256, 224
    return T.Compose([T.RandomResize(int(0.5 * base_size), int(1.2 * base_size)), T.RandomHorizontalFlip(0.5), T.RandomVerticalFlip(0.5),
                      T.RandomRotation(degrees=30), T.RandomCrop(crop_size), T.ToTensor(), T.Normalize(mean=MEAN, std=STD)])


def val_pipeline():         # train.py:70-75
    return T.Compose([T.RandomResize(224), T.ToTensor(), T.Normalize(mean=MEAN, std=STD)])


def main():
    B, Tn, H = 6, 3, 256
    u8, masks = AO.fixture_inputs(B, Tn, H)       # noisy synthetic series (texture exercises the filters) + disks with stray pixels
    pipe = train_pipeline()
    out = {"series_digest": np.array(AO.digest(u8)), "masks_digest": np.array(AO.digest(masks))}
    seeds = [11, 12, 13, 14, 15, 16]
    n_rot = n_pad = 0
    for b in range(B):
        phases = []
        for t in range(Tn):
            random.seed(seeds[b])
            img_t, m_t = pipe(Image.fromarray(u8[b, t]), Image.fromarray(masks[b]))
            phases.append(img_t.numpy())
        x_ref = np.stack(phases)                                        # [T, 1, 224, 224] float32
        params = AO.draw_params(random.Random(seeds[b]))
        x_or, t_or = AO.apply(u8[b], masks[b], params)
        assert np.array_equal(t_or, m_t.numpy()), f"sample {b}: oracle target != reference"
        assert np.array_equal(x_or, x_ref), f"sample {b}: oracle image != reference (max diff {np.abs(x_or - x_ref).max()})"
        n_rot += int(params["rot"]); n_pad += int(params["rh"] < 224)
        out[f"xdigest{b}"] = np.array(AO.digest(x_ref.astype(np.float32)))
        out[f"tdigest{b}"] = np.array(AO.digest(m_t.numpy().astype(np.int64)))
        if b == 1:                                                     # one sample in full (resized + flipped + rotated + cropped)
            out["x1"] = x_ref.astype(np.float32)
            out["t1"] = m_t.numpy().astype(np.uint8)
        print(b, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in params.items()})
    out["seeds"] = np.array(seeds)
    assert n_rot >= 2 and n_rot < B and n_pad >= 1, (n_rot, n_pad)      # rotated and unrotated, padded and unpadded samples
    # validation pipeline: resize to 224 only (no randomness that matters: randint(224, 224))
    vp = val_pipeline()
    random.seed(0)
    xv, tv = vp(Image.fromarray(u8[0, 0]), Image.fromarray(masks[0]))
    pv = AO.draw_params(random.Random(0), min_size=224, max_size=224, hflip=0.0, vflip=0.0, degrees=0.0, crop=224)
    pv.update(rot=False, hflip=False, vflip=False, h0=0, w0=0)
    xo, to = AO.apply(u8[0, :1], masks[0], pv)
    assert np.array_equal(xo[0], xv.numpy()) and np.array_equal(to, tv.numpy())
    out["xval_digest"], out["tval_digest"] = np.array(AO.digest(xv.numpy().astype(np.float32))), np.array(AO.digest(tv.numpy().astype(np.int64)))
    np.savez_compressed(os.path.join(HERE, "augment_6x3x256.npz"), **out)
    print("saved", os.path.getsize(os.path.join(HERE, "augment_6x3x256.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
