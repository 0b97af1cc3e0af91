import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    """The oracle is an fp32 reference: never let cuDNN / cuBLAS run it in TF32 on the GPU box."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _reference_root():
    """The unmodified reference: baseline/_ref (tools/install_ref.sh; travels to the GPU box) or /root/reference (build
    container only)."""
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "src")) and os.path.isdir(os.path.join(p, "train_utils")):
            return p
    return None


@pytest.fixture(scope="session")
def reference():
    """Namespace over the reference's own modules / driver functions; skips when the install is missing."""
    root = _reference_root()
    if root is None:
        pytest.skip("reference not installed: run tools/install_ref.sh in the build container (baseline/_ref)")
    import types
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    from src import STFLSTMUNet, UNet
    from train_utils import train_and_eval as te
    from train_utils import dice_coefficient_loss as dl
    return types.SimpleNamespace(root=root, STFLSTMUNet=STFLSTMUNet, UNet=UNet, te=te, dl=dl)
