"""Frame ingest — SURVEY.md §8(f) row 4.

``images_to_torch`` mirrors interact/interactive_utils.py:18-23 (the GUI's loader) and the
``ToTensor + im_normalization`` transform of dataset/davis_test_dataset.py:49-52: u8 ``[T,H,W,3]``
frames in, normalised fp32 ``[1,T,3,H,W]`` on the device out.  The reference normalises on the CPU
and uploads 12 bytes per pixel; here the u8 frames cross PCIe (3 bytes per pixel, from pinned
memory) and one HBM-bound kernel produces the same tensor bit for bit
(tests/test_gpu_clients.py::test_ingest_is_bit_identical)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import MivosError


def images_to_torch(frames, device="cuda:0") -> torch.Tensor:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise MivosError("mivos_b200.ingest.images_to_torch needs a CUDA device (no CPU path)")
    if isinstance(frames, np.ndarray):
        frames = torch.from_numpy(np.ascontiguousarray(frames))
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise MivosError("images_to_torch expects uint8 frames [T,H,W,3]")
    if not frames.is_cuda:
        frames = frames.contiguous()
        if not frames.is_pinned():
            frames = frames.pin_memory()
        frames = frames.to(dev, non_blocking=True)
    return ops.frames_u8_normalize(frames.contiguous()).unsqueeze(0)
