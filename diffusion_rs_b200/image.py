"""PNG encoding of the pipeline's output images (host side).

The reference's Python binding returns PNG-encoded `bytes` per prompt (`diffusion_rs_py/src/lib.rs:140-154`:
`DynamicImage::write_to(.., ImageFormat::Png)`; `diffuse_rs.pyi`: `Pipeline.forward(...) -> list[bytes]`), and its CLI
saves `images[0]` (`diffusion_rs_cli/src/main.rs:142`).  This is the same container written with the standard
library only (zlib + struct): 8-bit RGB, no interlace, filter type 0 on every scanline.
"""
from __future__ import annotations

import struct
import zlib


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(image, compress_level: int = 3) -> bytes:
    """image: HWC uint8 (torch tensor or numpy array) with 3 channels (RGB) or 1 (grey) -> PNG file contents."""
    if hasattr(image, "detach"):
        image = image.detach().cpu().contiguous().numpy()
    if image.ndim != 3 or image.shape[2] not in (1, 3) or str(image.dtype) != "uint8":
        raise ValueError(f"encode_png expects an HWC uint8 image with 1 or 3 channels, got {image.shape} {image.dtype}")
    h, w, c = image.shape
    rows = image.reshape(h, w * c)
    raw = b"".join(b"\x00" + rows[y].tobytes() for y in range(h))  # filter type 0 (None) per scanline
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw, compress_level)) +
            _chunk(b"IEND", b""))


def save_png(image, path) -> None:
    with open(path, "wb") as f:
        f.write(encode_png(image))
