"""Device-side paired augmentation (csrc/augment.cu) == the reference's PIL pipeline, bit for bit: against the fixture
generated from /root/reference/transforms.py (tests/golden/make_golden_augment.py) and against the oracle on fresh draws."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import augment_oracle as AO  # noqa: E402
from stf_unet_b200.augment import PairedAugment  # noqa: E402

DEV = torch.device("cuda")


def test_train_pipeline_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "augment_6x3x256.npz"))
    u8, masks = AO.fixture_inputs()
    assert AO.digest(u8) == str(g["series_digest"])
    aug = PairedAugment(train=True)
    # one random.Random per sample, seeded like the fixture: the kernel must reproduce the reference's draws AND arithmetic
    params = []
    for seed in g["seeds"]:
        aug.rng = random.Random(int(seed))
        params.append(aug.draw((256, 256)))
        assert params[-1] == {k: v for k, v in AO.draw_params(random.Random(int(seed))).items() if k != "crop"}
    x, t = aug(torch.from_numpy(u8).to(DEV), torch.from_numpy(masks).to(DEV), params=params)
    assert x.shape == (6, 3, 1, 224, 224) and x.dtype == torch.float32 and t.shape == (6, 224, 224) and t.dtype == torch.int64
    x, t = x.cpu().numpy(), t.cpu().numpy()
    for b in range(6):
        assert AO.digest(x[b]) == str(g[f"xdigest{b}"]), f"sample {b}: image differs from the PIL pipeline"
        assert AO.digest(t[b]) == str(g[f"tdigest{b}"]), f"sample {b}: mask differs from the PIL pipeline"
    assert np.array_equal(x[1], g["x1"]) and np.array_equal(t[1], g["t1"].astype(np.int64))


def test_validation_pipeline_and_strided_target(golden_dir):
    g = np.load(os.path.join(golden_dir, "augment_6x3x256.npz"))
    u8, masks = AO.fixture_inputs()
    aug = PairedAugment(train=False, target_stride=2)
    x, t = aug(torch.from_numpy(u8[:1, :1]).to(DEV), torch.from_numpy(masks[:1]).to(DEV))
    assert AO.digest(x[0].cpu().numpy()) == str(g["xval_digest"])
    full = PairedAugment(train=False)(torch.from_numpy(u8[:1, :1]).to(DEV), torch.from_numpy(masks[:1]).to(DEV))[1]
    assert AO.digest(full[0].cpu().numpy()) == str(g["tval_digest"])
    assert torch.equal(t, full[:, ::2, ::2])               # STF-LSTM-UNet's half-resolution target = every second pixel


@pytest.mark.parametrize("hw", [(256, 256), (200, 312), (150, 150)])
def test_fresh_draws_match_oracle(hw):
    """Random geometry on inputs the fixture never saw (non-square, upscaling only, ...): kernel == oracle, every sample."""
    H, W = hw
    rng = np.random.default_rng(H * 7 + W)
    B, T = 5, 2
    u8 = rng.integers(0, 256, size=(B, T, H, W), dtype=np.uint8)
    masks = (rng.random((B, H, W)) < 0.3).astype(np.uint8)
    aug = PairedAugment(train=True, rng=random.Random(1000 + H))
    params = [aug.draw((H, W)) for _ in range(B)]
    x, t = aug(torch.from_numpy(u8).to(DEV), torch.from_numpy(masks).to(DEV), params=params)
    for b in range(B):
        xo, to = AO.apply(u8[b], masks[b], dict(params[b], crop=224))
        assert np.array_equal(x[b].cpu().numpy(), xo), (b, params[b])
        assert np.array_equal(t[b].cpu().numpy(), to), (b, params[b])


def test_augment_feeds_the_model_and_has_no_cpu_path():
    import stf_unet_b200 as S
    u8, masks = AO.fixture_inputs(B=2, T=2)
    aug = PairedAugment(train=True, target_stride=2, rng=random.Random(3))
    x, t = aug(torch.from_numpy(u8).to(DEV), torch.from_numpy(masks).to(DEV))
    m = S.STFLSTMUNet(1, 2, 2).to(DEV).train()
    loss = S.criterion(m(x), t)
    loss.backward()
    assert torch.isfinite(loss) and t.shape == (2, 112, 112)
    with pytest.raises(RuntimeError, match="CUDA"):
        aug(torch.from_numpy(u8), torch.from_numpy(masks))
    xe, te = aug(torch.empty((0, 2, 256, 256), dtype=torch.uint8, device=DEV), torch.empty((0, 256, 256), dtype=torch.uint8, device=DEV))
    assert xe.shape == (0, 2, 1, 224, 224) and te.shape == (0, 112, 112)
