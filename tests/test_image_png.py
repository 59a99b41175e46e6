"""CPU: the PNG container written by diffusion_rs_b200.image decodes (with an independent decoder) to the same pixels."""
import io
import struct
import zlib

import numpy as np
import pytest
import torch

from diffusion_rs_b200.image import encode_png, save_png


def _decode_minimal(png: bytes):
    """Independent of the encoder's helpers: walk the chunks, check CRCs, inflate, undo filter 0."""
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(png):
        n, tag = struct.unpack(">I4s", png[pos:pos + 8])
        data = png[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(tag + data) & 0xFFFFFFFF
        chunks.append((tag, data))
        pos += 12 + n
    assert [t for t, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, flt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, comp, flt, lace) == (8, 0, 0, 0)
    c = 3 if ctype == 2 else 1
    raw = zlib.decompress(chunks[1][1])
    rows = np.frombuffer(raw, np.uint8).reshape(h, 1 + w * c)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, c)


@pytest.mark.parametrize("shape", [(64, 96, 3), (5, 7, 3), (16, 16, 1)])
def test_png_round_trip(shape, tmp_path):
    img = torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    png = encode_png(img)
    assert np.array_equal(_decode_minimal(png), img.numpy())
    try:
        from PIL import Image
    except ImportError:
        Image = None
    if Image is not None:  # a full third-party decoder, when the image has one
        got = np.asarray(Image.open(io.BytesIO(png)))
        assert np.array_equal(got.reshape(shape), img.numpy())
    save_png(img, tmp_path / "x.png")
    assert (tmp_path / "x.png").read_bytes() == png


def test_png_rejects_bad_input():
    with pytest.raises(ValueError):
        encode_png(torch.zeros(4, 4, 4, dtype=torch.uint8))
    with pytest.raises(ValueError):
        encode_png(torch.zeros(4, 4, 3))
