"""Tiny PNG writer + ACES-ish tonemap for eyeballing HDR framebuffers (no external deps)."""
import struct
import zlib

import numpy as np


def write_png(path, rgb_u8):
    h, w = rgb_u8.shape[:2]
    raw = b"".join(b"\x00" + rgb_u8[y].tobytes() for y in range(h))
    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def tonemap(hdr, exposure=1.0):
    x = np.maximum(hdr[..., :3] * exposure, 0.0)
    y = (x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14)
    return (np.clip(y, 0, 1) ** (1 / 2.2) * 255 + 0.5).astype(np.uint8)


def save_hdr_png(path, fb, width, exposure=1.0, flip=True):
    img = tonemap(fb[:, :width], exposure)
    write_png(path, np.ascontiguousarray(img[::-1] if flip else img))
