"""Mask egress: what the reference's callers do with the ``np.uint8[T,h,w]`` label
maps ``InferenceCore.interact`` returns.

  * ``davis_palette()``      — the DAVIS colour map (util/palette.py:3-22), as the 768-entry list
                               ``Image.putpalette`` takes
  * ``save_masks_png()``     — one indexed-colour PNG per frame, ``{:05d}.png``
                               (eval_interactive_davis.py:88-94, interactive_gui.py:335-339); a
                               self-contained encoder (zlib + CRC, no imaging library on the path)
  * ``overlay_davis()``      — the GUI's per-frame display composite (interact/interactive_utils.py:
                               119-130) on the device: u8 frame + label map -> u8 overlay, one
                               HBM-bound kernel (``mivos_overlay_davis``)
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Optional, Sequence

import numpy as np


def davis_color_map(n: int = 256) -> np.ndarray:
    """uint8 [n,3]: label i -> colour whose channel bits are the bits of i dealt round-robin to r, g, b
    from the most significant position down (util/palette.py:3-20)."""
    idx = np.arange(n, dtype=np.uint32)
    cmap = np.zeros((n, 3), dtype=np.uint32)
    for j in range(8):
        for ch in range(3):
            cmap[:, ch] |= ((idx >> (3 * j + ch)) & 1) << (7 - j)
    return cmap.astype(np.uint8)


def davis_palette() -> list:
    return davis_color_map().reshape(-1).tolist()


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_indexed_png(mask: np.ndarray, palette: Optional[Sequence[int]] = None, level: int = 6) -> bytes:
    """An 8-bit indexed-colour PNG (colour type 3) of a uint8 [h,w] label map."""
    if mask.dtype != np.uint8 or mask.ndim != 2:
        raise ValueError(f"expected a uint8 [h,w] label map, got {mask.dtype} {mask.shape}")
    pal = bytes(davis_palette() if palette is None else [int(v) & 255 for v in palette])
    if len(pal) == 0 or len(pal) % 3 or len(pal) > 768:
        raise ValueError("palette must hold 1..256 RGB triples")
    h, w = mask.shape
    raw = np.empty((h, w + 1), dtype=np.uint8)
    raw[:, 0] = 0  # filter type 0 (None) on every scanline
    raw[:, 1:] = mask
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 3, 0, 0, 0)) + _chunk(b"PLTE", pal)
            + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b""))


def save_masks_png(masks: np.ndarray, out_dir: str, palette: Optional[Sequence[int]] = None, start: int = 0) -> list:
    """masks uint8 [T,h,w] -> out_dir/{start+i:05d}.png (eval_interactive_davis.py:88-94).  Returns the paths."""
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for i in range(len(masks)):
        p = os.path.join(out_dir, "{:05d}.png".format(start + i))
        with open(p, "wb") as f:
            f.write(encode_indexed_png(np.ascontiguousarray(masks[i]), palette))
        paths.append(p)
    return paths


def overlay_davis(image, mask, alpha: float = 0.5):
    """interact/interactive_utils.py:119-130 on the device.  image: uint8 [h,w,3] (or [T,h,w,3]) CUDA
    tensor, mask: uint8 [h,w] ([T,h,w]) CUDA tensor -> uint8 overlay of the image's shape."""
    from . import ops
    return ops.overlay_davis(image, mask, alpha)
