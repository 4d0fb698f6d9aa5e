"""Output / inspection path of the reference, headless:

  * the post-processing pass the window applies before it shows or captures a frame -- clamp to >= 0, ACES filmic curve, gamma 2.2
    (Src/Shaders/post.frag:18-41) -- as `tonemap_aces` (host, numpy) and `Pathtracer.present()` (device kernel k_present, 8-bit RGBA);
  * PPMExporter (Src/Exporters/PPMExporter.cpp:7-41): binary P6, image flipped vertically, channel = (unsigned char) clamp(255 x);
  * EXRExporter (Src/Exporters/EXRExporter.cpp:10-60): scan-line OpenEXR, channels B G R stored as HALF, no compression (tinyexr's
    zero-initialised header), image flipped vertically;
  * the capture of Src/Main.cpp:195-249: `<name>.ppm` gets the tone-mapped frame, `<name>.exr` the linear one, plus
    albedo.exr / normal.exr / position.exr for every enabled AOV.

Only numpy and struct: no OpenEXR dependency.  `load_exr` reads back what `save_exr` writes (uncompressed HALF / FLOAT scan lines).
"""
from __future__ import annotations

import os
import struct

import numpy as np


def tonemap_aces(rgb):
    """post.frag: max(0, c) -> ACES filmic (Narkowicz) -> clamp -> pow(1 / 2.2).  float32 in, float32 out."""
    c = np.maximum(np.asarray(rgb, dtype=np.float32), np.float32(0.0))
    a, b, cc, d, e = (np.float32(v) for v in (2.51, 0.03, 2.43, 0.59, 0.14))
    c = np.clip((c * (a * c + b)) / (c * (cc * c + d) + e), np.float32(0.0), np.float32(1.0))
    return np.power(c, np.float32(1.0 / 2.2)).astype(np.float32)


def to_bytes(rgb):
    """PPMExporter's conversion: (unsigned char) clamp(255 x, 0, 255) -- truncation, not rounding."""
    return np.clip(np.asarray(rgb, dtype=np.float32) * np.float32(255.0), 0.0, 255.0).astype(np.uint8)


def save_ppm(path, rgb):
    """rgb: [height, width, >=3] floats in [0, 1], row 0 = bottom of the image (the device frame layout); written top row first."""
    img = to_bytes(np.asarray(rgb)[::-1, :, :3])
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n %d\n %d\n %d\n" % (w, h, 255))
        f.write(np.ascontiguousarray(img).tobytes())


def load_ppm(path):
    data = open(path, "rb").read()
    tokens, pos = [], 0
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        tokens.append(data[pos:end]); pos = end
    pos += 1
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(data, dtype=np.uint8, count=w * h * 3, offset=pos).reshape(h, w, 3)[::-1]


def _attr(name, type_name, payload):
    return name.encode() + b"\0" + type_name.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def save_exr(path, rgb, half=True):
    """Scan-line OpenEXR 2.0, channels B, G, R (alphabetical, as the format requires), HALF (or FLOAT) samples, no compression,
    increasing-y line order; the device frame (row 0 = bottom) is flipped so that row 0 of the file is the top of the image."""
    img = np.asarray(rgb, dtype=np.float32)[::-1, :, :3]
    h, w = img.shape[:2]
    ptype, dtype, size = (1, np.float16, 2) if half else (2, np.float32, 4)
    chlist = b"".join(n + b"\0" + struct.pack("<iBBBBii", ptype, 0, 0, 0, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = (struct.pack("<ii", 20000630, 2)
              + _attr("channels", "chlist", chlist) + _attr("compression", "compression", b"\0")
              + _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
              + _attr("lineOrder", "lineOrder", b"\0") + _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
              + _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) + _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
              + b"\0")
    line_bytes = 3 * w * size
    first = len(header) + 8 * h
    with np.errstate(over="ignore"):
        planes = np.stack([img[..., 2], img[..., 1], img[..., 0]], axis=1).astype(dtype)     # [h, (B, G, R), w]
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack("<%dQ" % h, *[first + y * (8 + line_bytes) for y in range(h)]))
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(planes[y].tobytes())


def load_exr(path):
    """Reads the files save_exr writes: [height, width, 3] float32 RGB, row 0 = bottom (device layout)."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", data, 0)
    if magic != 20000630 or (version & 0xFF) != 2:
        raise ValueError("not an OpenEXR 2 file")
    pos, attrs = 8, {}
    while data[pos] != 0:
        end = data.index(b"\0", pos); name = data[pos:end].decode(); pos = end + 1
        end = data.index(b"\0", pos); typ = data[pos:end].decode(); pos = end + 1
        (n,) = struct.unpack_from("<i", data, pos); pos += 4
        attrs[name] = (typ, data[pos:pos + n]); pos += n
    pos += 1
    if attrs["compression"][1] != b"\0":
        raise ValueError("only uncompressed files are supported")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    channels, cp, raw = [], 0, attrs["channels"][1]
    while raw[cp] != 0:
        end = raw.index(b"\0", cp); name = raw[cp:end].decode(); cp = end + 1
        (ptype,) = struct.unpack_from("<i", raw, cp); cp += 16
        channels.append((name, ptype))
    offsets = struct.unpack_from("<%dQ" % h, data, pos)
    out = np.zeros((h, w, 3), dtype=np.float32)
    slot = {"R": 0, "G": 1, "B": 2}
    for off in offsets:
        y, _ = struct.unpack_from("<ii", data, off); p = off + 8
        for name, ptype in channels:
            dt, size = (np.float16, 2) if ptype == 1 else (np.float32, 4)
            row = np.frombuffer(data, dtype=dt, count=w, offset=p).astype(np.float32); p += w * size
            if name in slot:
                out[y - y0, :, slot[name]] = row
    return out[::-1]


def capture(pathtracer, filename, aov_dir=None):
    """What pressing the capture key / `-O` does in the reference (Src/Main.cpp:195-249): the displayed frame to `filename`
    (.ppm: tone-mapped like the window shows it; .exr: linear), and albedo / normal / position .exr for every enabled AOV."""
    from . import pathtracer as pt
    w, h = pathtracer.screen_width, pathtracer.screen_height
    frame = pathtracer.get_display()[:h, :w, :3]
    ext = os.path.splitext(filename)[1].lower()
    if ext == ".ppm":
        save_ppm(filename, tonemap_aces(frame))
    elif ext == ".exr":
        save_exr(filename, frame)
    else:
        raise ValueError(f"unsupported output file extension: {ext}")
    written = [filename]
    aov_dir = aov_dir if aov_dir is not None else os.path.dirname(os.path.abspath(filename))
    for kind, name in ((pt.AOV_ALBEDO, "albedo.exr"), (pt.AOV_NORMAL, "normal.exr"), (pt.AOV_POSITION, "position.exr")):
        if pathtracer.gpu_config.aov_mask & (1 << kind):
            path = os.path.join(aov_dir, name)
            save_exr(path, pathtracer.get_aov(kind, True)[:h, :w, :3])
            written.append(path)
    return written
