"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the STF-Unet hot path.

Nothing in ``stf_unet_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / CPU baseline.

Parity status: the reference ships NO golden vectors, known-answer tests or
checkpoints for this path, so the oracle is pinned
against outputs of the reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the build
container, loads the deterministic weights of ``oracle.weights`` into it,
and commits logits / loss / gradient fixtures under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks the oracle against those fixtures.
"""
from .weights import (stf_param_spec, unet_param_spec, make_state_dict,
                      synthetic_dce_batch)
from .stf_oracle import (stf_forward, unet_forward, criterion, dice_loss,
                         loss_and_grads)

__all__ = [
    "stf_param_spec", "unet_param_spec", "make_state_dict",
    "synthetic_dce_batch", "stf_forward", "unet_forward", "criterion",
    "dice_loss", "loss_and_grads",
]
