"""Shared host-side plumbing for the drop-in modules: device buffers owned by PyTorch
(packed weights, workspaces) and the loud no-fallback guard."""
from __future__ import annotations

import os
from typing import Dict, Iterable, Tuple

import torch

from .. import _native as N


def default_precision() -> str:
    return os.environ.get("I2L_PRECISION", "fp32")


def require_cuda(t: torch.Tensor, what: str) -> None:
    """The reference runs on cpu/mps/cuda (utils/mps_utils.py:50-75); this build is
    CUDA sm_100a only by mandate -- fail loudly instead of falling back."""
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: tensor is on '{t.device}'.  hmer-img2latex_b200 runs only on CUDA (sm_100a); "
            "there is no CPU fallback -- move the module and its inputs to a B200.")


def params_key(tensors: Iterable[torch.Tensor], precision: str) -> Tuple:
    return (precision,) + tuple((t.data_ptr(), t._version, t.device.index) for t in tensors)


class Workspace:
    """Grow-only uint8 device buffer reused across calls (allocated by PyTorch's caching
    allocator, i.e. owned by the caller of the C-ABI)."""

    def __init__(self):
        self._buf: Dict[int, torch.Tensor] = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        key = device.index if device.index is not None else torch.cuda.current_device()
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


def f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
