"""The accelerated ops (B200 build).  Mirrors /root/reference/torchopt/accelerated_op/__init__.py:30-50."""
from __future__ import annotations

from typing import Iterable, Union

import torch

from torchopt_b200.accelerated_op.adam_op import AdamOp
from torchopt_b200.accelerated_op.adam_step import AdamStep

Device = Union[torch.device, str, int]

__all__ = ["AdamOp", "AdamStep", "is_available"]


def is_available(devices: Device | Iterable[Device] | None = None) -> bool:
    """True iff the accelerated Adam operator runs on every device in ``devices``.

    Like the reference it decides by actually running the op on a one-element tensor and treating any exception
    as "unavailable".  Differences that follow from "no CPU fallback": ``'cpu'`` is False here (the reference's
    OpenMP kernels are not part of this product) and the default device list is ``['cuda']``.
    ``'meta'`` is False, as with the reference's compiled extension (tests/test_accelerated_op.py:37-40).
    """
    if devices is None:
        devices = [torch.device("cuda")]
    elif isinstance(devices, (torch.device, int, str)):
        devices = [devices]
    op = AdamOp()
    try:
        for device in devices:
            device = torch.device(device)
            if device.type == "cuda" and not torch.cuda.is_available():
                return False
            probe = torch.tensor(1.0, device=device)
            op(probe, probe, probe, 1)  # all three arguments alias one 0-dim tensor, as in the reference
    except Exception:  # noqa: BLE001  pylint: disable=broad-except
        return False
    return True
