"""`--record`: the per-step PNG dump of GameImage.save_image_obs (misc/game/gameimage.py:9-62) for
images produced by gc_render.  File names and directory layout are the reference's
(`misc/game/record/<filename>/t=%03d.png`, cleared when recording starts); the encoder is a minimal
PNG writer over zlib, so no image library is needed on the GPU box."""
import os
import struct
import zlib

import numpy as np


def encode_png(rgb):
    """uint8[H][W][3] -> PNG bytes (8-bit RGB, filter 0 on every row)"""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, c = rgb.shape
    if c != 3:
        raise ValueError("expected an RGB image")
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def decode_png(data):
    """inverse of encode_png for the files it writes (tests)"""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", 0
    while pos < len(data):
        (n,), tag = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 3).copy()


class GameImage:
    """Recorder half of the reference's GameImage: `save_image_obs(t)` writes the current frame."""

    def __init__(self, filename, render_fn, record=False, root="misc/game/record"):
        self.game_record_dir = os.path.join(root, filename)
        self.record = record
        self._render = render_fn  # () -> uint8[H][W][3] numpy image of the current state
        if record:  # :22-29
            os.makedirs(self.game_record_dir, exist_ok=True)
            for f in os.listdir(self.game_record_dir):
                os.remove(os.path.join(self.game_record_dir, f))

    def get_image_obs(self):
        return self._render()

    def save_image_obs(self, t):  # :54-62
        path = os.path.join(self.game_record_dir, "t=%03d.png" % t)
        with open(path, "wb") as f:
            f.write(encode_png(self._render()))
        return path
