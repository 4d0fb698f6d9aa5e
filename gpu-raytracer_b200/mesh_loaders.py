"""The reference's other mesh formats, host side: Stanford PLY, Mitsuba `.serialized`, Mitsuba hair.

Each loader returns what `scene.load_obj` returns -- (positions[F,3,3], normals[F,3,3], tex_coords[F,3,2]) after the reference's
Triangle constructor fix-ups (`scene.finish_triangles`) -- so the meshes go through the same BVH builders and blob.

  * load_ply        Src/Assets/PLYLoader.cpp:152-346.  ascii / binary_little_endian / binary_big_endian; vertex properties x y z nx ny nz
                    (u|s) (v|t) of any scalar type, unknown properties skipped; faces = list property vertex_index / vertex_indices, fan
                    triangulated; v is flipped (1 - v) like the OBJ loader does.
  * load_serialized Src/Assets/Mitsuba/SerializedLoader.cpp:8-221.  Format id 0x041c, versions 3 and 4, end-of-file dictionary with 32-
                    or 64-bit offsets, one zlib stream per sub-mesh (`shapeIndex`), flags: normals 0x1, uvs 0x2, colours 0x8 (skipped),
                    face normals 0x10, single 0x1000 / double 0x2000 precision.
  * load_hair       Src/Assets/Mitsuba/MitshairLoader.cpp:9-123.  "BINARY_HAIR" (+inf separates strands) or ASCII (blank line separates
                    strands); every segment becomes a two-triangle ribbon of half-width `radius`, tapering to 0 at the strand's end, rolled
                    about the strand by a per-strand pseudo-random angle (PCG seeded with the FNV-1a hash of the file name, like the
                    reference; the value of that angle only orients the ribbon).

The writers (`save_ply`, `save_serialized`, `save_hair`) exist for the tests: no file of these formats ships with the reference, so the
loaders are pinned against files written from the format descriptions above, in every variant the loaders accept.
"""
from __future__ import annotations

import math
import struct
import zlib

import numpy as np

f32 = np.float32

_PLY_TYPES = {"int8": "i1", "char": "i1", "int16": "i2", "short": "i2", "int32": "i4", "int": "i4", "uint8": "u1", "uchar": "u1",
              "uint16": "u2", "ushort": "u2", "uint32": "u4", "uint": "u4", "float32": "f4", "float": "f4", "float64": "f8", "double": "f8"}
_PLY_VERTEX_SLOTS = {"x": 0, "y": 1, "z": 2, "nx": 3, "ny": 4, "nz": 5, "u": 6, "s": 6, "v": 7, "t": 7}


def _finish(p, n, t):
    from .scene import finish_triangles
    if len(p) == 0:            # AssetManager.cpp:66-80: an empty mesh becomes one dummy triangle
        p = [[[-1, -1, 0], [0, 1, 0], [1, -1, 0]]]; n = [[[0, 0, 1]] * 3]; t = [[[0, 1], [0.5, 0], [1, 1]]]
    return finish_triangles(np.asarray(p, dtype=f32), np.asarray(n, dtype=f32), np.asarray(t, dtype=f32))


# ----------------------------------------------------------------------------------------------------------------- PLY
def load_ply(path):
    data = open(path, "rb").read()
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        raise ValueError("not a PLY file")
    body = data.index(b"\n", end) + 1
    fmt, elements = None, []
    for line in data[:end].decode("ascii", "replace").splitlines()[1:]:
        tok = line.split()
        if not tok or tok[0] == "comment":
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            if tok[1] not in ("vertex", "face"):
                raise ValueError(f"unsupported PLY element '{tok[1]}'")
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if not elements:
                raise ValueError("PLY property defined without element")
            if tok[1] == "list":
                elements[-1][2].append(("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]], tok[4]))
            else:
                elements[-1][2].append(("scalar", _PLY_TYPES[tok[1]], None, tok[2]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError("invalid PLY format")
    order = ">" if fmt == "binary_big_endian" else "<"
    verts = np.zeros((0, 9), dtype=np.float64)
    tri_idx = []
    if fmt == "ascii":
        tokens = data[body:].split()
        pos = 0
    else:
        pos = body
    for kind, count, props in elements:
        if kind == "vertex":
            if any(p[0] == "list" for p in props):
                raise ValueError("list property on PLY vertices")
            verts = np.zeros((count, 9), dtype=np.float64)           # x y z nx ny nz u v ignored
            slots = [_PLY_VERTEX_SLOTS.get(p[3], 8) for p in props]
            if fmt == "ascii":
                block = np.array(tokens[pos:pos + count * len(props)], dtype=np.float64).reshape(count, len(props)); pos += count * len(props)
                for c, s in enumerate(slots):
                    verts[:, s] = block[:, c]
            else:
                dt = np.dtype([("f%d" % i, order + p[1]) for i, p in enumerate(props)])
                block = np.frombuffer(data, dtype=dt, count=count, offset=pos); pos += count * dt.itemsize
                for i, s in enumerate(slots):
                    verts[:, s] = block["f%d" % i]
        else:
            for _ in range(count):
                for p in props:
                    is_index = p[0] == "list" and p[3] in ("vertex_index", "vertex_indices")
                    if fmt == "ascii":
                        if p[0] == "list":
                            n = int(float(tokens[pos])); vals = [int(float(v)) for v in tokens[pos + 1:pos + 1 + n]]; pos += 1 + n
                        else:
                            vals = None; pos += 1
                    elif p[0] == "list":
                        n = int(np.frombuffer(data, dtype=order + p[1], count=1, offset=pos)[0]); pos += np.dtype(p[1]).itemsize
                        vals = np.frombuffer(data, dtype=order + p[2], count=n, offset=pos).astype(np.int64).tolist(); pos += n * np.dtype(p[2]).itemsize
                    else:
                        vals = None; pos += np.dtype(p[1]).itemsize
                    if is_index:
                        if len(vals) <= 2:
                            raise ValueError("a PLY face needs at least 3 indices")
                        for k in range(1, len(vals) - 1):
                            tri_idx.append((vals[0], vals[k], vals[k + 1]))
    idx = np.asarray(tri_idx, dtype=np.int64).reshape(-1, 3)
    v = verts.astype(f32)
    uv = np.stack([v[:, 6], f32(1.0) - v[:, 7]], axis=1)
    return _finish(v[idx][:, :, 0:3], v[idx][:, :, 3:6], uv[idx])


def save_ply(path, vertices, faces, normals=None, uvs=None, fmt="ascii", vertex_type="float", index_types=("uchar", "int"), extra=False):
    """vertices [N,3], faces = list of index lists (polygons allowed), normals [N,3] / uvs [N,2] optional; `extra` adds a property the
    loader must ignore."""
    vertices = np.asarray(vertices, dtype=np.float64); n = len(vertices)
    cols = [("x", vertices[:, 0]), ("y", vertices[:, 1]), ("z", vertices[:, 2])]
    if extra:
        cols.append(("confidence", np.linspace(0, 1, n)))
    if normals is not None:
        nm = np.asarray(normals, dtype=np.float64); cols += [("nx", nm[:, 0]), ("ny", nm[:, 1]), ("nz", nm[:, 2])]
    if uvs is not None:
        uv = np.asarray(uvs, dtype=np.float64); cols += [("s", uv[:, 0]), ("t", uv[:, 1])]
    head = ["ply", f"format {fmt} 1.0", "comment written by gpu_raytracer_b200.mesh_loaders", f"element vertex {n}"]
    head += [f"property {vertex_type} {name}" for name, _ in cols]
    head += [f"element face {len(faces)}", f"property list {index_types[0]} {index_types[1]} vertex_indices", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        if fmt == "ascii":
            for i in range(n):
                f.write((" ".join(repr(float(c[i])) for _, c in cols) + "\n").encode())
            for face in faces:
                f.write((" ".join([str(len(face))] + [str(int(i)) for i in face]) + "\n").encode())
        else:
            order = ">" if fmt == "binary_big_endian" else "<"
            vt = order + _PLY_TYPES[vertex_type]
            f.write(np.stack([c for _, c in cols], axis=1).astype(vt).tobytes())
            for face in faces:
                f.write(np.array([len(face)], dtype=order + _PLY_TYPES[index_types[0]]).tobytes())
                f.write(np.array(face, dtype=order + _PLY_TYPES[index_types[1]]).tobytes())


# ----------------------------------------------------------------------------------------------------------------- Mitsuba .serialized
def load_serialized(path, shape_index=0):
    data = open(path, "rb").read()
    fmt_id, version = struct.unpack_from("<HH", data, 0)
    if fmt_id != 0x041C:
        raise ValueError("serialized file does not start with format id 0x041c")
    (num_meshes,) = struct.unpack_from("<I", data, len(data) - 4)
    if version <= 3:
        dict_off = len(data) - 4 - 4 * num_meshes
        offsets = list(struct.unpack_from("<%dI" % num_meshes, data, dict_off))
    else:
        dict_off = len(data) - 4 - 8 * num_meshes
        offsets = list(struct.unpack_from("<%dQ" % num_meshes, data, dict_off))
    offsets.append(dict_off)
    if not 0 <= shape_index < num_meshes:
        raise ValueError("shapeIndex out of range")
    raw = zlib.decompress(data[offsets[shape_index] + 4:offsets[shape_index + 1]])
    (flags,) = struct.unpack_from("<I", raw, 0); pos = 4
    single, double = bool(flags & 0x1000), bool(flags & 0x2000)
    if version <= 3:
        single = True
    else:
        pos = raw.index(b"\0", pos) + 1                      # null-terminated name
    nv, nt = struct.unpack_from("<QQ", raw, pos); pos += 16
    if nv == 0 or nt == 0:
        return _finish([], [], [])
    if not (single or double):
        raise ValueError("neither single nor double precision specified")
    et = "<f4" if single else "<f8"
    size = np.dtype(et).itemsize

    def take(count, width):
        nonlocal pos
        a = np.frombuffer(raw, dtype=et, count=count * width, offset=pos).reshape(count, width).astype(f32); pos += count * width * size
        return a
    vp = take(nv, 3)
    vn = take(nv, 3) if flags & 0x0001 else None
    vt = take(nv, 2) if flags & 0x0002 else None
    if flags & 0x0008:
        take(nv, 3)
    idx = np.frombuffer(raw, dtype="<u4" if nv <= 0xFFFFFFFF else "<u8", count=nt * 3, offset=pos).astype(np.int64).reshape(nt, 3)
    p = vp[idx]
    if flags & 0x0010:                                          # face normals
        g = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]); g = g / np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-30)
        n = np.repeat(g[:, None, :], 3, axis=1)
    elif vn is not None:
        n = vn[idx]
    else:
        n = np.zeros_like(p)
    t = vt[idx] if vt is not None else np.zeros((nt, 3, 2), dtype=f32)
    return _finish(p, n, t)


def save_serialized(path, meshes, version=4, double=False):
    """meshes: list of dicts {vertices [N,3], faces [T,3], normals?, uvs?, face_normals?: bool, name?: str}."""
    et = "<f8" if double else "<f4"
    blob, offsets = b"", []
    for m in meshes:
        offsets.append(len(blob))
        v = np.asarray(m["vertices"]); fc = np.asarray(m["faces"], dtype="<u4")
        flags = (0x2000 if double else 0x1000) | (0x1 if m.get("normals") is not None else 0) | (0x2 if m.get("uvs") is not None else 0) | (0x10 if m.get("face_normals") else 0)
        if version <= 3:
            flags &= ~0x3000
        body = struct.pack("<I", flags)
        if version > 3:
            body += m.get("name", "mesh").encode() + b"\0"
        body += struct.pack("<QQ", len(v), len(fc)) + v.astype("<f4" if version <= 3 else et).tobytes()
        if m.get("normals") is not None:
            body += np.asarray(m["normals"]).astype("<f4" if version <= 3 else et).tobytes()
        if m.get("uvs") is not None:
            body += np.asarray(m["uvs"]).astype("<f4" if version <= 3 else et).tobytes()
        body += fc.tobytes()
        blob += struct.pack("<HH", 0x041C, version) + zlib.compress(body)
    tail = b"".join(struct.pack("<I" if version <= 3 else "<Q", o) for o in offsets) + struct.pack("<I", len(meshes))
    with open(path, "wb") as f:
        f.write(blob + tail)


# ----------------------------------------------------------------------------------------------------------------- Mitsuba hair
class _PCG:
    """Src/Core/Random.h:8-52."""

    def __init__(self, seed):
        mul, inc = 747796405, 2891336453
        self.state = (((seed + inc) & 0xFFFFFFFFFFFFFFFF) * mul + inc) & 0xFFFFFFFFFFFFFFFF

    def get_float(self):
        s = self.state
        x = (((s >> 18) ^ s) >> 27) & 0xFFFFFFFF
        r = s >> 59
        self.state = (s * 6364136223846793005 + 1) & 0xFFFFFFFFFFFFFFFF
        u = ((x >> r) | (x << ((-r) & 31))) & 0xFFFFFFFF
        return float(f32(u) * np.frombuffer(struct.pack("<I", 0x2F7FFFFF), dtype=f32)[0])


def _fnv1a(text):
    h = 14695981039346656037
    for b in text.encode():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _orthogonal(v):                                                # Math.h:75-81
    s = math.copysign(1.0, v[2]); a = -1.0 / (s + v[2]); b = v[0] * v[1] * a
    return np.array([1.0 + s * v[0] * v[0] * a, s * b, -s * v[0]])


def _rotate(axis, angle, v):                                       # Quaternion::axis_angle(axis, angle) * v
    c, s = math.cos(angle), math.sin(angle)
    return v * c + np.cross(axis, v) * s + axis * float(np.dot(axis, v)) * (1.0 - c)


def load_hair(path, radius=0.0025):
    data = open(path, "rb").read()
    strands, cur = [], []
    if data.startswith(b"BINARY_HAIR"):
        pos = 11 + 4                                               # magic + vertex count
        while pos + 4 <= len(data):
            (x,) = struct.unpack_from("<f", data, pos)
            if math.isinf(x):
                strands.append(cur); cur = []; pos += 4
            else:
                y, z = struct.unpack_from("<ff", data, pos + 4); cur.append((x, y, z)); pos += 12
    else:
        for line in data.decode("ascii", "replace").split("\n"):
            tok = line.split()
            if not tok:
                strands.append(cur); cur = []
            else:
                cur.append((float(tok[0]), float(tok[1]), float(tok[2])))
    rng = _PCG(_fnv1a(path))
    p, t = [], []
    for strand in strands:
        if len(strand) < 2:
            continue                                               # "a hair strand was defined with less than 2 vertices"
        s = np.asarray(strand, dtype=np.float64)
        angle = math.pi * rng.get_float()
        d = s[1] - s[0]; d = d / np.linalg.norm(d)
        o = _rotate(d, angle, _orthogonal(d))
        prev = (s[0] + radius * o, s[0] - radius * o)
        for v in range(1, len(s)):
            d = s[v] - s[v - 1]; ln = np.linalg.norm(d)
            o = np.array([1.0, 0.0, 0.0]) if not ln > 0 else _rotate(d / ln, angle, _orthogonal(d / ln))
            r = radius + (0.0 - radius) * (float(v) / float(len(s) - 1))
            cur2 = (s[v] + r * o, s[v] - r * o)
            p.append([prev[0], prev[1], cur2[0]]); p.append([prev[1], cur2[1], cur2[0]])
            t.append([(0, 0), (1, 0), (0, 1)]); t.append([(0, 0), (1, 0), (0, 1)])
            prev = cur2
    return _finish(p, np.zeros((len(p), 3, 3)), t)


def save_hair(path, strands, binary=True):
    """strands: list of [k,3] arrays."""
    with open(path, "wb") as f:
        if binary:
            f.write(b"BINARY_HAIR" + struct.pack("<I", sum(len(s) for s in strands)))
            for s in strands:
                f.write(np.asarray(s, dtype="<f4").tobytes()); f.write(struct.pack("<f", float("inf")))
        else:
            for s in strands:
                for v in np.asarray(s, dtype=np.float64):
                    f.write(("%r %r %r\n" % (float(v[0]), float(v[1]), float(v[2]))).encode())
                f.write(b"\n")
