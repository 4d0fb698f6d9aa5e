"""Scene front-end: turns a scene description into the flat "scene blob" the device kernels consume.

This is host-side staging (numpy + the C++ BVH builder in host/bvh_build.cpp); nothing here runs on
the hot path.  The blob is the byte-exact module ABI of the reference:

  triangles       float32 [T,24]   96-B records in BVH leaf order   (Src/Renderer/Integrators/Integrator.cpp:127-151)
  bvh_nodes       uint8   [N*80|32] TLAS in slots [0,2M), BLAS after (Integrator.cpp:113,252-277)
  mesh_*          per-instance tables in TLAS-leaf order             (Integrator.cpp:412-423)
  materials/types 32-B union + 1-B type                              (Pathtracer.cpp:544-588)
  light_*         CDF tables                                         (Pathtracer.cpp:384-534)
  camera          15 floats                                          (Integrator.cpp:456-472, Camera.cpp:20-42)

Two sources feed it: `load_mitsuba()` (the subset of Mitsuba-0.5 XML + OBJ the shipped scenes use;
reference walker: Src/Assets/Mitsuba/MitsubaLoader.cpp:519-671, Src/Assets/OBJLoader.cpp:86-220) and
`procedural_scene()` (self-contained synthetic scenes for tests/bench when the reference data is absent).
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import build as _build

MAT_LIGHT, MAT_DIFFUSE, MAT_PLASTIC, MAT_DIELECTRIC, MAT_CONDUCTOR = 0, 1, 2, 3, 4
INVALID = -1

PMJ_NUM_SEQUENCES = 64
PMJ_NUM_SAMPLES = 4096
BLUE_NOISE_TEXTURES = 16
BLUE_NOISE_DIM = 128

f32 = np.float32


# ----------------------------------------------------------------------------- host BVH builder binding
_hostlib = None


def hostlib():
    global _hostlib
    if _hostlib is None:
        path = _build.build_host()
        lib = ctypes.CDLL(path)
        lib.ptbh_build_triangles.restype = ctypes.c_void_p
        lib.ptbh_build_triangles.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
        lib.ptbh_build_boxes.restype = ctypes.c_void_p
        lib.ptbh_build_boxes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        for name in ("ptbh_kind", "ptbh_node_count", "ptbh_index_count", "ptbh_node_bytes"):
            getattr(lib, name).restype = ctypes.c_int
            getattr(lib, name).argtypes = [ctypes.c_void_p]
        lib.ptbh_export.restype = None
        lib.ptbh_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.ptbh_free.restype = None
        lib.ptbh_free.argtypes = [ctypes.c_void_p]
        lib.ptbh_build_triangles_sbvh.restype = ctypes.c_void_p
        lib.ptbh_build_triangles_sbvh.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float]
        lib.ptbh_from_bvh2.restype = ctypes.c_void_p
        lib.ptbh_from_bvh2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
        lib.ptbh_trace_stats.restype = None
        lib.ptbh_trace_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _hostlib = lib
    return _hostlib


class BuiltBVH:
    """Owns one ptbh handle (a BLAS or a TLAS)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("BVH build failed")
        self.h = ctypes.c_void_p(handle)
        lib = hostlib()
        self.kind = lib.ptbh_kind(self.h)
        self.node_count = lib.ptbh_node_count(self.h)
        self.index_count = lib.ptbh_index_count(self.h)
        self.node_bytes = lib.ptbh_node_bytes(self.h)

    def export(self, node_offset=0, index_offset=0):
        nodes = np.zeros(self.node_count * self.node_bytes, dtype=np.uint8)
        indices = np.zeros(self.index_count, dtype=np.int32)
        hostlib().ptbh_export(self.h, nodes.ctypes.data, indices.ctypes.data, int(node_offset), int(index_offset))
        return nodes, indices

    def __del__(self):
        try:
            hostlib().ptbh_free(self.h)
        except Exception:
            pass


def build_blas(positions: np.ndarray, kind: int, sah_node=4.0, sah_leaf=1.0) -> BuiltBVH:
    pos = np.ascontiguousarray(positions, dtype=f32).reshape(-1, 9)
    return BuiltBVH(hostlib().ptbh_build_triangles(pos.ctypes.data, pos.shape[0], kind, sah_node, sah_leaf))


def blas_from_bvh2(nodes: np.ndarray, indices: np.ndarray, kind: int, sah_node=4.0, sah_leaf=1.0) -> BuiltBVH:
    """A raw binary BVH (32-byte nodes + primitive order, e.g. from the reference's `.bvh` cache, bvh_cache.py) converted to the
    traversal kind without rebuilding (BVH::create_from_bvh2, Src/BVH/BVH.cpp)."""
    nd = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1)
    idx = np.ascontiguousarray(indices, dtype=np.int32)
    return BuiltBVH(hostlib().ptbh_from_bvh2(nd.ctypes.data, nd.size // 32, idx.ctypes.data, idx.size, kind, sah_node, sah_leaf))


def build_blas_cached(mesh_file: str, tri, kind: int, sah_node=4.0, sah_leaf=1.0, force_rebuild=False) -> BuiltBVH:
    """The reference's mesh-load flow (Src/Assets/AssetManager.cpp + BVHLoader.cpp): reuse `<mesh file>.bvh` when it is at least as
    new as the mesh and was written with the same settings; otherwise build the raw SAH BVH2, write the cache, convert.
    tri = (positions[F,3,3], normals[F,3,3], tex_coords[F,3,2]) of that mesh file."""
    from . import bvh_cache
    p, n, t = tri
    cached = bvh_cache.try_to_load(mesh_file, bvh_cache.BVH_TYPE_SAH, False, sah_node, sah_leaf, force_rebuild)
    if cached is not None and cached.triangles.shape[0] == p.shape[0] and np.array_equal(cached.triangles["position"], np.asarray(p, dtype=f32)):
        return blas_from_bvh2(cached.nodes, cached.indices, kind, sah_node, sah_leaf)
    raw = build_blas(p, 2, sah_node, 0.0)                     # raw BVH2, one primitive per leaf: what the cache stores
    nodes, indices = raw.export(0, 0)
    try:
        bvh_cache.save(bvh_cache.bvh_filename(mesh_file), bvh_cache.CachedBVH(bvh_cache.pack_triangles(p, n, t), nodes.view(bvh_cache.NODE2_DTYPE), indices,
                                                                              bvh_cache.BVH_TYPE_SAH, False, sah_node, sah_leaf))
    except OSError:
        pass                                                  # read-only asset directory: build every time, like a failed fopen in BVHLoader::save
    return blas_from_bvh2(nodes, indices, kind, sah_node, sah_leaf)


def build_blas_sbvh(positions: np.ndarray, alpha=3e-4, bins=96, max_dup=2.0, optimize_passes=0, optimize_fraction=1.0, max_depth=0) -> BuiltBVH:
    """CWBVH over a split BVH (spatial splits; host/bvh_build.cpp SpatialBuilder).  index_count >= triangle count: a triangle may be
    referenced by several leaves.  This is what libptb builds for the merged static BVH.  optimize_passes > 0 runs the insertion-based
    optimiser (ReinsertionOptimizer) over the binary tree before the wide collapse, like libptb does."""
    pos = np.ascontiguousarray(positions, dtype=f32).reshape(-1, 9)
    lib = hostlib()
    lib.ptbh_set_optimizer.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int]
    lib.ptbh_set_optimizer(int(optimize_passes), float(optimize_fraction), int(max_depth))
    try:
        return BuiltBVH(lib.ptbh_build_triangles_sbvh(pos.ctypes.data, pos.shape[0], alpha, bins, max_dup))
    finally:
        lib.ptbh_set_optimizer(0, 1.0, 0)


def optimizer_sah():
    """(before, after): sum of node areas / root area of the binary tree of the last optimised build_blas_sbvh call."""
    lib = hostlib()
    v = (ctypes.c_double * 2)()
    lib.ptbh_optimizer_sah.argtypes = [ctypes.c_void_p]
    lib.ptbh_optimizer_sah(v)
    return float(v[0]), float(v[1])


def trace_stats(bvh: BuiltBVH, positions: np.ndarray, rays: np.ndarray):
    """CPU closest-hit walk of a CWBVH (ptbh_trace_stats): (nodes per ray, triangle tests per ray, t[n], triangle id[n])."""
    pos = np.ascontiguousarray(positions, dtype=f32).reshape(-1, 9)
    r = np.ascontiguousarray(rays, dtype=f32).reshape(-1, 6)
    counts = (ctypes.c_ulonglong * 2)(0, 0)
    t = np.empty(r.shape[0], dtype=f32); tri = np.empty(r.shape[0], dtype=np.int32)
    hostlib().ptbh_trace_stats(bvh.h, pos.ctypes.data, r.ctypes.data, r.shape[0], counts, t.ctypes.data, tri.ctypes.data)
    return counts[0] / r.shape[0], counts[1] / r.shape[0], t, tri


def build_tlas(aabbs: np.ndarray, kind: int) -> BuiltBVH:
    bx = np.ascontiguousarray(aabbs, dtype=f32).reshape(-1, 6)
    return BuiltBVH(hostlib().ptbh_build_boxes(bx.ctypes.data, bx.shape[0], kind))


# ----------------------------------------------------------------------------- small math (row-major 4x4)
def m_identity():
    return np.eye(4, dtype=np.float64)


def m_translate(v):
    m = m_identity(); m[:3, 3] = v; return m


def m_scale(x, y=None, z=None):
    if y is None:
        y = z = x
    m = m_identity(); m[0, 0], m[1, 1], m[2, 2] = x, y, z; return m


def q_axis_angle(axis, angle):
    h = 0.5 * angle
    s = math.sin(h)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, math.cos(h)], dtype=np.float64)


def q_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]], dtype=np.float64)


def m_rotation(q):
    x, y, z, w = q
    m = m_identity()
    m[0, 0] = 1 - 2 * (y * y + z * z); m[1, 0] = 2 * (x * y + w * z); m[2, 0] = 2 * (x * z - w * y)
    m[0, 1] = 2 * (x * y - w * z); m[1, 1] = 1 - 2 * (x * x + z * z); m[2, 1] = 2 * (y * z + w * x)
    m[0, 2] = 2 * (x * z + w * y); m[1, 2] = 2 * (y * z - w * x); m[2, 2] = 1 - 2 * (x * x + y * y)
    return m


def q_rotate(q, v):
    return m_rotation(q)[:3, :3] @ np.asarray(v, dtype=np.float64)


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def q_look_rotation(forward, up):
    """Quaternion whose rotation maps +Z to `forward` (Src/Math/Quaternion.h:38-69)."""
    f = _normalize(forward)
    r = _normalize(np.cross(up, f))
    u = np.cross(f, r)
    m00, m01, m02 = r; m10, m11, m12 = u; m20, m21, m22 = f
    if m22 < 0:
        if m00 > m11:
            t = 1 + m00 - m11 - m22; s = 0.5 / math.sqrt(t)
            return np.array([s * t, s * (m01 + m10), s * (m20 + m02), s * (m12 - m21)])
        t = 1 - m00 + m11 - m22; s = 0.5 / math.sqrt(t)
        return np.array([s * (m01 + m10), s * t, s * (m12 + m21), s * (m20 - m02)])
    if m00 < -m11:
        t = 1 - m00 - m11 + m22; s = 0.5 / math.sqrt(t)
        return np.array([s * (m20 + m02), s * (m12 + m21), s * t, s * (m01 - m10)])
    t = 1 + m00 + m11 + m22; s = 0.5 / math.sqrt(t)
    return np.array([s * (m12 - m21), s * (m20 - m02), s * (m01 - m10), s * t])


def m_decompose(m, forward=(0.0, 0.0, -1.0)):
    """position / rotation / uniform scale as the reference extracts them (Src/Math/Matrix4.h:184-194)."""
    pos = m[:3, 3].copy()
    rot = q_look_rotation(m[:3, :3] @ np.asarray(forward, dtype=np.float64), (0.0, 1.0, 0.0))
    sx, sy, sz = (np.linalg.norm(m[i, :3]) for i in range(3))
    scale = float(np.cbrt(sx * sy * sz))
    return pos, rot, scale


def m_cofactor3(m):
    """Upper-left 3x3 of the cofactor matrix (normal transform)."""
    a = m[:3, :3]
    c = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            sub = np.delete(np.delete(a, i, 0), j, 1)
            c[i, j] = ((-1) ** (i + j)) * np.linalg.det(sub)
    return c


# ----------------------------------------------------------------------------- triangles
def finish_triangles(p, n, t):
    """Apply the reference's Triangle constructor fix-ups (Src/Renderer/Triangle.h:36-103): zero normals are
    replaced by the geometric normal; winding is reversed when all three shading normals face backwards."""
    p = np.asarray(p, dtype=f32).reshape(-1, 3, 3).copy()
    n = np.asarray(n, dtype=f32).reshape(-1, 3, 3).copy()
    t = np.asarray(t, dtype=f32).reshape(-1, 3, 2).copy()
    g = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]).astype(np.float64)
    gl = np.linalg.norm(g, axis=1, keepdims=True)
    g = (g / np.where(gl > 0, gl, 1.0)).astype(f32)
    bad = np.linalg.norm(n, axis=2) < 1e-30
    n[bad] = np.broadcast_to(g[:, None, :], n.shape)[bad]
    flip = np.all(np.einsum("ij,ikj->ik", g, n) < 0.0, axis=1)
    for arr in (p, n, t):
        tmp = arr[flip, 1].copy(); arr[flip, 1] = arr[flip, 2]; arr[flip, 2] = tmp
    return p, n, t


def geo_rectangle(m):
    v = [m[:3, :3] @ np.array(c, dtype=np.float64) + m[:3, 3] for c in ((-1, 1, 0), (1, 1, 0), (1, -1, 0), (-1, -1, 0))]
    nrm = _normalize(m_cofactor3(m) @ np.array([0.0, 0.0, 1.0]))
    uv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    p = [[v[0], v[1], v[2]], [v[0], v[2], v[3]]]
    t = [[uv[0], uv[1], uv[2]], [uv[0], uv[2], uv[3]]]
    n = [[nrm] * 3] * 2
    return finish_triangles(p, n, t)


def geo_cube(m):
    corners = [(-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1), (-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)]
    v = [m[:3, :3] @ np.array(c, dtype=np.float64) + m[:3, 3] for c in corners]
    cof = m_cofactor3(m)
    fn = [_normalize(cof @ np.array(d, dtype=np.float64)) for d in ((0, 1, 0), (0, 0, -1), (1, 0, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0))]
    faces = [(0, 1, 2, 3), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7)]
    uv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    p, n, t = [], [], []
    for f, face in enumerate(faces):
        q = [v[i] for i in face]
        p += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
        t += [[uv[0], uv[1], uv[2]], [uv[0], uv[2], uv[3]]]
        n += [[fn[f]] * 3] * 2
    return finish_triangles(p, n, t)


def geo_icosphere(m, subdivisions=2):
    x, z = 0.525731112119133606, 0.850650808352039932
    verts = np.array([(-x, 0, z), (x, 0, z), (-x, 0, -z), (x, 0, -z), (0, z, x), (0, z, -x), (0, -z, x), (0, -z, -x),
                      (z, x, 0), (-z, x, 0), (z, -x, 0), (-z, -x, 0)], dtype=np.float64)
    idx = np.array([(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10), (5, 3, 8), (5, 2, 3), (2, 7, 3),
                    (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6), (0, 1, 6), (6, 1, 10), (9, 0, 11), (9, 11, 2), (9, 2, 5), (7, 2, 11)])
    tris = verts[idx]  # [20,3,3]
    for _ in range(subdivisions):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        nrm = lambda q: q / np.linalg.norm(q, axis=1, keepdims=True)
        ab, bc, ca = nrm(a + b), nrm(b + c), nrm(c + a)
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)])
    unit = tris.reshape(-1, 3)
    pos = unit @ m[:3, :3].T + m[:3, 3]
    cof = m_cofactor3(m)
    nor = unit @ cof.T
    nor = nor / np.linalg.norm(nor, axis=1, keepdims=True)
    uv = np.stack([np.arctan2(-unit[:, 2], unit[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(unit[:, 1], -1, 1)) / np.pi], 1)
    return finish_triangles(pos.reshape(-1, 3, 3), nor.reshape(-1, 3, 3), uv.reshape(-1, 3, 2))


def load_obj(path):
    """Wavefront OBJ -> (positions, normals, uvs) as [F,3,*]; fan triangulation, v-flipped uvs,
    negative indices relative to the END of the file's arrays (reference quirk, OBJLoader.cpp:172-189)."""
    pos, tex, nor, faces = [], [], [], []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            if len(line) < 3:
                continue
            c0 = line[0]
            if c0 == "v":
                c1 = line[1]
                if c1 == " " or c1 == "\t":
                    s = line.split(); pos.append((float(s[1]), float(s[2]), float(s[3])))
                elif c1 == "t":
                    s = line.split(); tex.append((float(s[1]), float(s[2])))
                elif c1 == "n":
                    s = line.split(); nor.append((float(s[1]), float(s[2]), float(s[3])))
            elif c0 == "f" and line[1] in " \t":
                corners = []
                for tok in line.split()[1:]:
                    parts = tok.split("/")
                    v = int(parts[0]) if parts[0] else 0
                    t = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                    n = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    corners.append((v, t, n))
                for k in range(1, len(corners) - 1):
                    faces.append((corners[0], corners[k], corners[k + 1]))
    F = len(faces)
    fa = np.array(faces, dtype=np.int64).reshape(F, 3, 3) if F else np.zeros((0, 3, 3), dtype=np.int64)

    def gather(arr, col, width):
        out = np.zeros((F, 3, width), dtype=f32)
        size = len(arr)
        if size == 0 or F == 0:
            return out, np.zeros((F, 3), dtype=bool)
        a = np.asarray(arr, dtype=f32)
        i = fa[:, :, col]
        r = np.where(i > 0, i - 1, np.where(i < 0, size + i, -1))
        ok = (r >= 0) & (r < size)
        out[ok] = a[r[ok]]
        return out, ok

    p, _ = gather(pos, 0, 3)
    t, tok = gather(tex, 1, 2)
    t[..., 1] = np.where(tok, 1.0 - t[..., 1], t[..., 1]).astype(f32)
    n, _ = gather(nor, 2, 3)
    if F == 0:  # reference inserts a dummy triangle for empty meshes (AssetManager.cpp:66-80)
        p = np.array([[[-1, -1, 0], [0, 1, 0], [1, -1, 0]]], dtype=f32)
        n = np.array([[[0, 0, 1]] * 3], dtype=f32)
        t = np.array([[[0, 1], [0.5, 0], [1, 1]]], dtype=f32)
    return finish_triangles(p, n, t)


# ----------------------------------------------------------------------------- textures
def gamma_to_linear(x):
    x = np.asarray(x, dtype=np.float64)
    lin = np.where(x < 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    return np.clip(np.where(x <= 0, 0.0, np.where(x >= 1, 1.0, lin)), 0.0, 1.0)


def load_tga(path):
    """Minimal TGA reader (types 2/10, 24/32 bpp, + 8-bit grey 3/11) -> uint8 [H,W,4] top-down."""
    with open(path, "rb") as fh:
        data = fh.read()
    idlen, cmap_type, img_type = data[0], data[1], data[2]
    w, h = struct.unpack_from("<HH", data, 12)
    bpp, desc = data[16], data[17]
    if cmap_type != 0 or img_type not in (2, 3, 10, 11):
        raise ValueError(f"unsupported TGA type {img_type} in {path}")
    nb = bpp // 8
    off = 18 + idlen
    if img_type in (2, 3):
        px = np.frombuffer(data, dtype=np.uint8, count=w * h * nb, offset=off).reshape(h, w, nb)
    else:
        out = np.empty((w * h, nb), dtype=np.uint8)
        i, n = off, 0
        total = w * h
        while n < total:
            hdr = data[i]; i += 1
            cnt = (hdr & 0x7F) + 1
            if hdr & 0x80:
                out[n:n + cnt] = np.frombuffer(data, dtype=np.uint8, count=nb, offset=i); i += nb
            else:
                out[n:n + cnt] = np.frombuffer(data, dtype=np.uint8, count=cnt * nb, offset=i).reshape(cnt, nb); i += cnt * nb
            n += cnt
        px = out.reshape(h, w, nb)
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    if nb == 1:
        rgba[..., 0] = rgba[..., 1] = rgba[..., 2] = px[..., 0]; rgba[..., 3] = 255
    else:
        rgba[..., 0], rgba[..., 1], rgba[..., 2] = px[..., 2], px[..., 1], px[..., 0]
        rgba[..., 3] = px[..., 3] if nb == 4 else 255
    if not (desc & 0x20):
        rgba = rgba[::-1]
    if desc & 0x10:
        rgba = rgba[:, ::-1]
    return np.ascontiguousarray(rgba)


def _bc1_encode(level):
    """Range-fit BC1 (4-colour mode) of a uint8 [H,W,4] level, H,W multiples of 4 (or smaller -> padded)."""
    h, w = level.shape[:2]
    H, W = max(4, (h + 3) // 4 * 4), max(4, (w + 3) // 4 * 4)
    pad = np.zeros((H, W, 3), dtype=np.float32)
    pad[:h, :w] = level[..., :3]
    if h < H:
        pad[h:, :w] = pad[h - 1:h, :w]
    if w < W:
        pad[:, w:] = pad[:, w - 1:w]
    blk = pad.reshape(H // 4, 4, W // 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 3)
    lo, hi = blk.min(1), blk.max(1)

    def q565(c):
        r = np.round(c[:, 0] * 31 / 255).astype(np.uint32); g = np.round(c[:, 1] * 63 / 255).astype(np.uint32); b = np.round(c[:, 2] * 31 / 255).astype(np.uint32)
        return (r << 11) | (g << 5) | b

    def d565(v):
        r = (v >> 11) & 31; g = (v >> 5) & 63; b = v & 31
        return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], 1).astype(np.float32)

    c0, c1 = q565(hi), q565(lo)
    swap = c0 < c1
    c0, c1 = np.where(swap, c1, c0), np.where(swap, c0, c1)
    same = c0 == c1
    e0, e1 = d565(c0), d565(c1)
    pal = np.stack([e0, e1, (2 * e0 + e1) / 3, (e0 + 2 * e1) / 3], 1)  # [B,4,3]
    d = ((blk[:, :, None, :] - pal[:, None, :, :]) ** 2).sum(-1)      # [B,16,4]
    sel = d.argmin(-1).astype(np.uint32)
    sel[same] = 0
    bits = (sel << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(1).astype(np.uint32)
    out = np.empty((blk.shape[0], 2), dtype=np.uint32)
    out[:, 0] = c0 | (c1 << 16)
    out[:, 1] = bits
    return out.view(np.uint8).reshape(-1)


def make_texture(rgba_u8, block_compress=True):
    """sRGB uint8 image -> linear mip chain (2x2 box) -> BC1 blocks when power-of-two, else RGBA8
    (reference: Src/Assets/TextureLoader.cpp:129-282; lod_bias quirk: computed from the BLOCK grid size)."""
    h, w = rgba_u8.shape[:2]
    lin = gamma_to_linear(rgba_u8.astype(np.float64) / 255.0)
    levels = [lin]
    lw, lh = w, h
    while lw > 1 or lh > 1:
        prev = levels[-1]
        nw, nh = max(lw // 2, 1), max(lh // 2, 1)
        ph, pw = prev.shape[:2]
        a = prev[: nh * 2 if ph >= 2 else 1, : nw * 2 if pw >= 2 else 1]
        if ph >= 2:
            a = 0.5 * (a[0::2] + a[1::2])
        if pw >= 2:
            a = 0.5 * (a[:, 0::2] + a[:, 1::2])
        levels.append(a); lw, lh = nw, nh
    u8 = [np.clip(l * 255.0, 0, 255).astype(np.uint8) for l in levels]
    pot = (w & (w - 1)) == 0 and (h & (h - 1)) == 0
    if block_compress and pot and w >= 4 and h >= 4:
        bw, bh = (w + 3) // 4, (h + 3) // 4
        nlev = int(math.log2(max(bw, bh))) + 1
        data = [_bc1_encode(u8[l]) for l in range(nlev)]
        return dict(format="bc1", width=w, height=h, levels=data, lod_bias=0.5 * math.log2(float(bw * bh)))
    return dict(format="rgba8", width=w, height=h, levels=[np.ascontiguousarray(l).reshape(-1) for l in u8],
                lod_bias=0.5 * math.log2(float(w * h)))


def load_image_u8(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".tga":
        return load_tga(path)
    raise ValueError(f"no decoder for {ext}")


def pink_texture():
    return dict(format="rgba8", width=1, height=1, levels=[np.array([255, 0, 255, 255], dtype=np.uint8)], lod_bias=0.0)


def load_hdr(path, max_width=None):
    """Radiance .hdr (RGBE, RLE or flat) -> float32 [H,W,4] (w = 0), optionally box-downsampled."""
    with open(path, "rb") as fh:
        data = fh.read()
    pos = 0
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]; pos = end + 1
        if len(line) == 0:
            break
    end = data.index(b"\n", pos)
    dims = data[pos:end].split(); pos = end + 1
    h, w = int(dims[1]), int(dims[3])
    buf = np.frombuffer(data, dtype=np.uint8, offset=pos)
    img = np.empty((h, w, 4), dtype=np.uint8)
    i = 0
    for y in range(h):
        if w >= 8 and w < 32768 and buf[i] == 2 and buf[i + 1] == 2 and ((int(buf[i + 2]) << 8) | int(buf[i + 3])) == w:
            i += 4
            for c in range(4):
                x = 0
                while x < w:
                    cnt = int(buf[i]); i += 1
                    if cnt > 128:
                        cnt -= 128
                        img[y, x:x + cnt, c] = buf[i]; i += 1
                    else:
                        img[y, x:x + cnt, c] = buf[i:i + cnt]; i += cnt
                    x += cnt
        else:
            img[y] = buf[i:i + 4 * w].reshape(w, 4); i += 4 * w
    e = img[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0).astype(np.float32)
    rgb = img[..., :3].astype(np.float32) * scale[..., None]
    while max_width is not None and rgb.shape[1] > max_width and rgb.shape[1] % 2 == 0 and rgb.shape[0] % 2 == 0:
        rgb = 0.25 * (rgb[0::2, 0::2] + rgb[1::2, 0::2] + rgb[0::2, 1::2] + rgb[1::2, 1::2])
    out = np.zeros(rgb.shape[:2] + (4,), dtype=f32)
    out[..., :3] = rgb
    return out


def procedural_sky(width=256, height=128):
    """Smooth analytic sky (gradient + sun lobe) for scenes staged without the reference's .hdr files."""
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    theta = v[:, None] * np.pi
    phi = (u[None, :] - 0.5) * 2 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), -np.sin(theta) * np.sin(phi)], -1)
    up = np.clip(d[..., 1], 0, 1)
    base = np.stack([0.35 + 0.25 * (1 - up), 0.45 + 0.25 * (1 - up), 0.75 + 0.1 * (1 - up)], -1)
    ground = np.array([0.18, 0.16, 0.14])
    col = np.where(d[..., 1:2] > 0, base, ground)
    sun = _normalize([0.4, 0.8, 0.3])
    col = col + 8.0 * np.clip((d @ sun - 0.985) / 0.015, 0, 1)[..., None] ** 2
    out = np.zeros((height, width, 4), dtype=f32)
    out[..., :3] = col
    return out


# ----------------------------------------------------------------------------- RNG tables
def fallback_rng_tables(seed=1234):
    """Stand-in sample tables when the reference's PMJ02 / blue-noise tables were not staged:
    per-sequence stratified jittered points and white-noise offsets. Same shapes and dtypes."""
    rng = np.random.default_rng(seed)
    n = PMJ_NUM_SAMPLES
    side = int(math.isqrt(n))
    pmj = np.empty((PMJ_NUM_SEQUENCES, n, 2), dtype=f32)
    for s in range(PMJ_NUM_SEQUENCES):
        cells = rng.permutation(n)
        cx, cy = cells % side, cells // side
        pmj[s, :, 0] = (cx + rng.random(n)) / side
        pmj[s, :, 1] = (cy + rng.random(n)) / side
    pmj = np.minimum(pmj, np.float32(0.99999994))
    blue = rng.integers(0, 256, size=(BLUE_NOISE_TEXTURES, BLUE_NOISE_DIM, BLUE_NOISE_DIM, 2), dtype=np.uint8)
    return pmj.reshape(-1), blue.reshape(-1)


def load_rng_tables():
    """RNG tables: staged copy of the reference's (tools/stage_data.py) when present, else the fallback."""
    path = os.path.join(_build.REPO_ROOT, "data", "_staged", "rng_tables.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["pmj"].astype(f32).reshape(-1), z["blue_noise"].astype(np.uint8).reshape(-1), "reference"
    pmj, blue = fallback_rng_tables()
    return pmj, blue, "fallback"


# ----------------------------------------------------------------------------- scene description -> blob
class Material:
    def __init__(self, kind=MAT_DIFFUSE, name="Material", emission=(0, 0, 0), diffuse=(1, 1, 1), texture=INVALID,
                 medium=INVALID, ior=1.33, eta=(1.33, 1.33, 1.33), k=(1, 1, 1), roughness=0.5):
        self.kind, self.name = kind, name
        self.emission, self.diffuse, self.texture = tuple(emission), tuple(diffuse), texture
        self.medium, self.ior, self.eta, self.k, self.roughness = medium, ior, tuple(eta), tuple(k), roughness

    def is_light(self):
        return self.kind == MAT_LIGHT and sum(e * e for e in self.emission) > 0.0


class Instance:
    def __init__(self, mesh_data, material, position=(0, 0, 0), rotation=(0, 0, 0, 1), scale=1.0, name=""):
        self.mesh_data, self.material, self.name = mesh_data, material, name
        self.position = np.asarray(position, dtype=np.float64)
        self.rotation = np.asarray(rotation, dtype=np.float64)
        self.scale = float(scale)

    def identity(self, eps=1e-6):
        def ae(a, b):
            return a == b or abs(a - b) < eps * max(abs(a) + abs(b), 1e-30) or (b == 0 and abs(a) < eps * 1.1754944e-38)
        q = self.rotation
        return (ae(self.scale, 1.0) and all(ae(float(v), 0.0) for v in self.position) and all(ae(float(v), 0.0) for v in q[:3])
                and (ae(float(q[3]), 1.0) or ae(float(q[3]), -1.0)))


class SceneDesc:
    """Host-side scene: mesh datas (triangle soups), materials, media, textures, instances, camera, sky."""

    def __init__(self):
        self.mesh_datas = []   # list of (p[F,3,3], n[F,3,3], t[F,3,2])
        self.mesh_files = []   # per mesh data: the file it was loaded from, or None (shapes built in code); keys the `.bvh` cache
        self.materials = [Material(MAT_DIFFUSE, "Default", diffuse=(1, 0, 1))]
        self.media = [dict(sigma_a=(0, 0, 0), sigma_s=(0, 0, 0), g=0.0)]
        self.textures = []
        self.instances = []
        self.cam_position = np.zeros(3)
        self.cam_rotation = np.array([0.0, 0.0, 0.0, 1.0])
        self.cam_fov = math.radians(85.0)
        self.cam_aperture = 0.0
        self.cam_focal = 10.0
        self.width, self.height = 900, 600
        self.num_bounces = 10
        self.sky = None
        self.sky_scale = 1.0
        self.source = "procedural"

    def add_mesh_data(self, tri, source=None):
        self.mesh_datas.append(tri); self.mesh_files.append(source); return len(self.mesh_datas) - 1

    def add_material(self, m):
        self.materials.append(m); return len(self.materials) - 1


def camera_block(position, rotation, fov, width, height, aperture=0.0, focal=10.0):
    """The 15-float device camera (Src/Renderer/Camera.cpp:20-42,86-88; Integrator.cpp:456-472)."""
    half_w, half_h = 0.5 * width, 0.5 * height
    tan_half = math.tan(0.5 * fov)
    d = half_w / tan_half
    blc = q_rotate(rotation, (-half_w, -half_h, -d))
    xa = q_rotate(rotation, (1.0, 0.0, 0.0))
    ya = q_rotate(rotation, (0.0, 1.0, 0.0))
    spread = math.atan(2.0 * tan_half / width)
    return np.array(list(position) + list(blc) + list(xa) + list(ya) + [spread, aperture, focal], dtype=f32)


def view_projection(position, rotation, fov, width, height, near=0.1, far=300.0):
    """projection * R^-1 * T^-1, row-major (Camera.cpp:86-95, Matrix4.h perspective)."""
    half_w, half_h = 0.5 * width, 0.5 * height
    t = math.tan(0.5 * fov)
    aspect = half_h / half_w
    p = np.zeros((4, 4))
    p[0, 0] = 1.0 / t; p[1, 1] = 1.0 / (aspect * t)
    p[2, 2] = -(far + near) / (far - near); p[3, 2] = -1.0
    p[2, 3] = -2.0 * (far * near) / (far - near)
    return (p @ m_rotation(q_conj(rotation)) @ m_translate(-np.asarray(position))).astype(f32).reshape(-1)


def camera_params_of(blob):
    """(position, rotation matrix 3x3, tan(fov/2), aperture, focal) recovered from a blob's 15-float camera block + film size."""
    cam = np.asarray(blob["camera"], dtype=np.float64)
    w0 = float(blob["width"])
    pos, blc, xa, ya = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    za = np.cross(xa, ya)
    d0 = -float(np.dot(blc, za))
    return pos.copy(), np.stack([xa, ya, za], axis=1), (0.5 * w0) / d0, float(cam[13]), float(cam[14])


def retarget_blob(blob, width, height, forward=None, fov=None):
    """Same scene, another film size (and optionally another view direction / fov): recomputes the camera block and the
    view-projection matrix (Camera::resize -> recalibrate, Src/Renderer/Camera.cpp:10-42); everything else is shared.
    `forward` = world-space viewing direction (the camera looks along its local -z)."""
    out = dict(blob)
    pos, R, tan_half, aperture, focal = camera_params_of(blob)
    if fov is not None:
        tan_half = math.tan(0.5 * fov)
    if forward is not None:
        rot = q_look_rotation(tuple(-np.asarray(forward, dtype=np.float64)), (0.0, 1.0, 0.0))
        R = m_rotation(rot)[:3, :3]
    half_w, half_h = 0.5 * width, 0.5 * height
    d = half_w / tan_half
    blc = R @ np.array([-half_w, -half_h, -d])
    spread = math.atan(2.0 * tan_half / width)
    out["camera"] = np.array(list(pos) + list(blc) + list(R[:, 0]) + list(R[:, 1]) + [spread, aperture, focal], dtype=f32)
    near, far = 0.1, 300.0
    p = np.zeros((4, 4))
    p[0, 0] = 1.0 / tan_half; p[1, 1] = 1.0 / ((half_h / half_w) * tan_half)
    p[2, 2] = -(far + near) / (far - near); p[3, 2] = -1.0
    p[2, 3] = -2.0 * (far * near) / (far - near)
    view = np.eye(4); view[:3, :3] = R.T; view[:3, 3] = -(R.T @ pos)
    out["view_projection"] = (p @ view).astype(f32).reshape(-1)
    out["width"], out["height"] = int(width), int(height)
    return out


def instance_transforms(desc, order=None):
    """3x4 object-to-world and world-to-object matrices of the instances (Mesh::update, Src/Renderer/Mesh.cpp:7-15), [M, 12] float32 each;
    `order` (e.g. blob["instance_order"]) permutes them into a context's table order -- what Pathtracer.refit_instances expects."""
    xf, xf_inv = [], []
    for inst in desc.instances:
        T = m_translate(inst.position) @ m_rotation(inst.rotation) @ m_scale(inst.scale)
        Ti = m_scale(1.0 / inst.scale) @ m_rotation(q_conj(inst.rotation)) @ m_translate(-inst.position)
        xf.append(T[:3, :].astype(f32).reshape(-1)); xf_inv.append(Ti[:3, :].astype(f32).reshape(-1))
    xf, xf_inv = np.array(xf, dtype=f32), np.array(xf_inv, dtype=f32)
    if order is not None:
        xf, xf_inv = xf[np.asarray(order)], xf_inv[np.asarray(order)]
    return np.ascontiguousarray(xf), np.ascontiguousarray(xf_inv)


def check_tlas8(nodes, instance_boxes):
    """Host-side validation of a CWBVH TLAS (uint8 [n, 80]) against world boxes of the instances ([M, 6], table order): every
    instance must lie inside the de-quantised box of the leaf slot that references it, and every internal child's slot box must
    contain the union of everything below it.  Returns (number of instances reached, largest overhang found -- 0 for a valid tree,
    mean slack of the leaf slot boxes relative to the instance boxes)."""
    nodes = np.asarray(nodes, dtype=np.uint8).reshape(-1, 80)
    boxes = np.asarray(instance_boxes, dtype=np.float64).reshape(-1, 6)
    worst, reached, slack = [0.0], [0], []

    def slot_box(n, k):
        p = n[:12].view(np.float32).astype(np.float64)
        sc = np.array([np.array([int(n[12 + a]) << 23], dtype=np.uint32).view(np.float32)[0] for a in range(3)], dtype=np.float64)
        lo = np.array([p[a] + sc[a] * n[32 + 16 * a + k] for a in range(3)]); hi = np.array([p[a] + sc[a] * n[40 + 16 * a + k] for a in range(3)])
        return lo, hi

    def walk(ni):
        n = nodes[ni]
        imask = int(n[15]); base_child = int(n[16:20].view(np.uint32)[0]); base_inst = int(n[20:24].view(np.uint32)[0])
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        internal = 0
        for k in range(8):
            meta = int(n[24 + k])
            if imask & (1 << k):
                child = base_child + internal; internal += 1
                if not meta:
                    continue
                clo, chi = walk(child)
            elif meta:
                clo, chi = np.full(3, np.inf), np.full(3, -np.inf)
                for t in range(bin(meta >> 5).count("1")):
                    b = boxes[base_inst + (meta & 31) + t]; reached[0] += 1
                    clo = np.minimum(clo, b[:3]); chi = np.maximum(chi, b[3:])
            else:
                continue
            if not np.all(clo <= chi):
                continue
            lo, hi = slot_box(n, k)
            worst[0] = max(worst[0], float(np.max(lo - clo)), float(np.max(chi - hi)))
            if not (imask & (1 << k)):
                slack.append(float(np.mean((clo - lo) + (hi - chi))))
            lo_all = np.minimum(lo_all, clo); hi_all = np.maximum(hi_all, chi)
        return lo_all, hi_all

    walk(0)
    return reached[0], worst[0], float(np.mean(slack)) if slack else 0.0


def pack_triangles(p, n, t):
    out = np.empty((p.shape[0], 24), dtype=f32)
    out[:, 0:3] = p[:, 0]; out[:, 3:6] = p[:, 1] - p[:, 0]; out[:, 6:9] = p[:, 2] - p[:, 0]
    out[:, 9:12] = n[:, 0]; out[:, 12:15] = n[:, 1] - n[:, 0]; out[:, 15:18] = n[:, 2] - n[:, 0]
    out[:, 18:20] = t[:, 0]; out[:, 20:22] = t[:, 1] - t[:, 0]; out[:, 22:24] = t[:, 2] - t[:, 0]
    return out


def build_blob(desc: SceneDesc, bvh_kind=8, width=None, height=None, threads=None, timing=None, rng="auto", use_bvh_cache=False):
    """Flatten a SceneDesc into the device ABI. `timing` (dict) receives the CPU BVH-build seconds.
    use_bvh_cache: keep / reuse the reference's `<mesh file>.bvh` files next to the meshes (build_blas_cached)."""
    import time
    width = int(width or desc.width); height = int(height or desc.height)
    M = len(desc.instances)
    threads = threads or os.cpu_count() or 1

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as pool:  # one job per mesh data, like the reference's asset pool
        files = list(desc.mesh_files) + [None] * (len(desc.mesh_datas) - len(desc.mesh_files))
        blas = list(pool.map(lambda a: build_blas_cached(a[1], a[0], bvh_kind) if (use_bvh_cache and a[1]) else build_blas(a[0][0], bvh_kind),
                             zip(desc.mesh_datas, files)))
    t_blas = time.perf_counter() - t0

    node_bytes = {8: 80, 4: 128, 2: 32}[int(bvh_kind)]
    node_off, tri_off = [], []
    n_nodes, n_tris = 2 * M, 0
    for b in blas:
        node_off.append(n_nodes); tri_off.append(n_tris)
        n_nodes += b.node_count; n_tris += b.index_count
    nodes = np.zeros(n_nodes * node_bytes, dtype=np.uint8)
    triangles = np.zeros((n_tris, 24), dtype=f32)
    reverse = []  # per mesh data: original triangle -> flat BVH-order index
    for d, b in enumerate(blas):
        nd, idx = b.export(node_off[d], tri_off[d])
        nodes[node_off[d] * node_bytes:(node_off[d] + b.node_count) * node_bytes] = nd
        p, n, t = desc.mesh_datas[d]
        triangles[tri_off[d]:tri_off[d] + b.index_count] = pack_triangles(p[idx], n[idx], t[idx])
        rev = np.empty(p.shape[0], dtype=np.int64); rev[idx] = tri_off[d] + np.arange(b.index_count)
        reverse.append(rev)

    # per-instance transforms + world boxes (Src/Renderer/Mesh.cpp:16-33, Math/AABB.cpp transform)
    xf, xf_inv, boxes = [], [], []
    for inst in desc.instances:
        T = m_translate(inst.position) @ m_rotation(inst.rotation) @ m_scale(inst.scale)
        Ti = m_scale(1.0 / inst.scale) @ m_rotation(q_conj(inst.rotation)) @ m_translate(-inst.position)
        p = desc.mesh_datas[inst.mesh_data][0].reshape(-1, 3).astype(np.float64)
        lo, hi = p.min(0), p.max(0)
        c, e = 0.5 * (lo + hi), 0.5 * (hi - lo)
        nc = T[:3, :3] @ c + T[:3, 3]; ne = np.abs(T[:3, :3]) @ e
        lo, hi = (nc - ne).astype(f32), (nc + ne).astype(f32)
        for k in range(3):
            eps = f32(0.001)
            while hi[k] - lo[k] < eps:
                lo[k] -= eps; hi[k] += eps; eps *= f32(2.0)
        xf.append(T[:3, :].astype(f32).reshape(-1)); xf_inv.append(Ti[:3, :].astype(f32).reshape(-1))
        boxes.append(np.concatenate([lo, hi]))
    t0 = time.perf_counter()
    tlas = build_tlas(np.array(boxes, dtype=f32), bvh_kind)
    t_tlas = time.perf_counter() - t0
    tl_nodes, tl_idx = tlas.export(0, 0)
    assert tlas.node_count <= 2 * M
    nodes[:tl_nodes.size] = tl_nodes

    order = tl_idx
    roots = np.array([np.uint32(node_off[desc.instances[i].mesh_data]) | (np.uint32(1 << 31) if desc.instances[i].identity() else np.uint32(0))
                      for i in order], dtype=np.uint32).view(np.int32)
    mat_ids = np.array([desc.instances[i].material for i in order], dtype=np.int32)
    transforms = np.array([xf[i] for i in order], dtype=f32)
    transforms_inv = np.array([xf_inv[i] for i in order], dtype=f32)

    # materials (Pathtracer.cpp:544-588)
    K = len(desc.materials)
    mtypes = np.zeros(K, dtype=np.int8)
    mats = np.zeros((K, 8), dtype=f32)
    for i, m in enumerate(desc.materials):
        mtypes[i] = m.kind
        if m.kind == MAT_LIGHT:
            mats[i, 0:3] = m.emission
        elif m.kind in (MAT_DIFFUSE, MAT_PLASTIC):
            mats[i, 0:3] = m.diffuse
            mats[i, 3:4].view(np.int32)[0] = m.texture
            if m.kind == MAT_PLASTIC:
                mats[i, 4] = m.roughness
        elif m.kind == MAT_DIELECTRIC:
            mats[i, 0:1].view(np.int32)[0] = m.medium
            mats[i, 1] = max(m.ior, 1.0001); mats[i, 2] = m.roughness
        elif m.kind == MAT_CONDUCTOR:
            mats[i, 0:3] = m.eta; mats[i, 3] = m.roughness; mats[i, 4:7] = m.k
    media = np.zeros((len(desc.media), 8), dtype=f32)
    for i, md in enumerate(desc.media):
        media[i, 0:3] = md["sigma_a"]; media[i, 3] = md["g"]; media[i, 4:7] = md["sigma_s"]

    # lights (Pathtracer.cpp:384-534): per light mesh-data area CDF, per light instance power*area*scale^2 CDF
    light_tri_idx, light_tri_cdf = [], []
    md_span = {}
    for inst in desc.instances:
        if desc.materials[inst.material].is_light() and inst.mesh_data not in md_span:
            p = desc.mesh_datas[inst.mesh_data][0]
            area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]).astype(np.float64), axis=1)
            total = float(area.sum())
            first = len(light_tri_idx)
            cdf = np.cumsum(area / total).astype(f32)
            cdf = (cdf / cdf[-1]).astype(f32)
            light_tri_idx += list(reverse[inst.mesh_data]); light_tri_cdf += list(cdf)
            md_span[inst.mesh_data] = (first, first + len(area) - 1, total)
    lm_cdf, lm_span, lm_xf = [], [], []
    total_w = 0.0
    for slot, i in enumerate(order):
        inst = desc.instances[i]
        m = desc.materials[inst.material]
        if m.is_light():
            first, last, area = md_span[inst.mesh_data]
            power = 0.299 * m.emission[0] + 0.587 * m.emission[1] + 0.114 * m.emission[2]
            w = float(f32(power * f32(area))) * inst.scale * inst.scale
            if w > 0.0:
                total_w += w
                lm_cdf.append(total_w); lm_span.append((first, last)); lm_xf.append(slot)
    lm_cdf = (np.array(lm_cdf, dtype=f32) / f32(total_w)).astype(f32) if lm_cdf else np.zeros(0, dtype=f32)

    if rng == "fallback":
        pmj, blue = fallback_rng_tables(); rng_src = "fallback"
    else:
        pmj, blue, rng_src = load_rng_tables()
    sky = desc.sky if desc.sky is not None else procedural_sky()
    blob = dict(
        bvh_kind=int(bvh_kind), width=width, height=height, num_bounces=int(desc.num_bounces),
        triangles=triangles, bvh_nodes=nodes, tlas_node_count=int(tlas.node_count),
        mesh_bvh_root_indices=roots, mesh_material_ids=mat_ids,
        mesh_transforms=transforms, mesh_transforms_inv=transforms_inv, mesh_transforms_prev=transforms.copy(),
        material_types=mtypes, materials=mats, media=media,
        light_triangle_indices=np.array(light_tri_idx, dtype=np.int32),
        light_triangle_cdf=np.array(light_tri_cdf, dtype=f32),
        light_mesh_cdf=lm_cdf, light_mesh_triangle_span=np.array(lm_span, dtype=np.int32).reshape(-1, 2),
        light_mesh_transform_indices=np.array(lm_xf, dtype=np.int32), lights_total_weight=float(f32(total_w)),
        camera=camera_block(desc.cam_position, desc.cam_rotation, desc.cam_fov, width, height, desc.cam_aperture, desc.cam_focal),
        view_projection=view_projection(desc.cam_position, desc.cam_rotation, desc.cam_fov, width, height),
        textures=desc.textures, sky=np.ascontiguousarray(sky, dtype=f32), sky_scale=float(desc.sky_scale),
        pmj=pmj, blue_noise=blue, rng_source=rng_src, source=desc.source,
        instance_order=np.asarray(order, dtype=np.int32),
        mesh_tri_first=np.array([tri_off[desc.instances[i].mesh_data] for i in order], dtype=np.int32),
        mesh_tri_count=np.array([blas[desc.instances[i].mesh_data].index_count for i in order], dtype=np.int32),
    )
    if timing is not None:
        timing.update(blas_seconds=t_blas, tlas_seconds=t_tlas, threads=threads, triangles=int(n_tris), nodes=int(n_nodes),
                      mesh_datas=len(desc.mesh_datas), instances=M)
    return blob


_ARRAY_KEYS = ["triangles", "bvh_nodes", "mesh_bvh_root_indices", "mesh_material_ids", "mesh_transforms", "mesh_transforms_inv",
               "mesh_transforms_prev", "material_types", "materials", "media", "light_triangle_indices", "light_triangle_cdf",
               "light_mesh_cdf", "light_mesh_triangle_span", "light_mesh_transform_indices", "camera", "view_projection", "sky",
               "pmj", "blue_noise", "instance_order", "mesh_tri_first", "mesh_tri_count"]
_SCALAR_KEYS = ["bvh_kind", "width", "height", "num_bounces", "tlas_node_count", "lights_total_weight", "sky_scale", "rng_source", "source"]


def save_blob(blob, path):
    arrays = {k: blob[k] for k in _ARRAY_KEYS}
    meta = {k: blob[k] for k in _SCALAR_KEYS}
    meta["textures"] = []
    for i, t in enumerate(blob["textures"]):
        meta["textures"].append(dict(format=t["format"], width=t["width"], height=t["height"], lod_bias=t["lod_bias"], levels=len(t["levels"])))
        for l, lv in enumerate(t["levels"]):
            arrays[f"tex{i}_l{l}"] = np.asarray(lv, dtype=np.uint8)
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, **arrays)      # the staged blobs travel to the GPU box with every gpurun call


def dump_raw(blob, path, camera_only=False):
    """Flat binary dump for non-Python hosts (tests/cpp/facade_render.cpp reads it): magic, entry count, then per entry a
    32-byte name, an int64 byte count and the bytes."""
    import struct
    items = []
    def put(name, arr):
        a = np.ascontiguousarray(arr)
        items.append((name, a.tobytes()))
    put("camera", np.asarray(blob["camera"], dtype=f32)); put("view_projection", np.asarray(blob["view_projection"], dtype=f32))
    put("width", np.int32(blob["width"])); put("height", np.int32(blob["height"]))
    if not camera_only:
        for k in ("triangles", "bvh_nodes", "mesh_bvh_root_indices", "mesh_material_ids", "mesh_transforms", "mesh_transforms_inv", "mesh_transforms_prev",
                  "material_types", "materials", "media", "pmj", "blue_noise", "light_triangle_indices", "light_triangle_cdf", "light_mesh_cdf",
                  "light_mesh_triangle_span", "light_mesh_transform_indices"):
            put(k, blob[k])
        sky = np.ascontiguousarray(blob["sky"], dtype=f32)
        put("sky", sky); put("sky_height", np.int32(sky.shape[0])); put("sky_width", np.int32(sky.shape[1])); put("sky_scale", f32(blob["sky_scale"]))
        put("bvh_kind", np.int32(blob["bvh_kind"])); put("tlas_node_count", np.int32(blob["tlas_node_count"])); put("num_bounces", np.int32(blob["num_bounces"]))
        put("lights_total_weight", f32(blob["lights_total_weight"]))
        put("texture_count", np.int32(len(blob["textures"])))
        for i, t in enumerate(blob["textures"]):
            meta = np.array([1 if t["format"] == "bc1" else 0, t["width"], t["height"], len(t["levels"]), 0], dtype=np.int32)
            meta[4:5].view(f32)[0] = t["lod_bias"]
            put(f"tex{i}_meta", meta)
            for l, lv in enumerate(t["levels"]):
                put(f"tex{i}_l{l}", np.asarray(lv, dtype=np.uint8))
    with open(path, "wb") as f:
        f.write(b"PTBRAW1\0"); f.write(struct.pack("<i", len(items)))
        for name, data in items:
            f.write(name.encode().ljust(32, b"\0")[:32]); f.write(struct.pack("<q", len(data))); f.write(data)


def load_blob(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta_json"]).decode())
    blob = {k: z[k] for k in _ARRAY_KEYS}
    for k in _SCALAR_KEYS:
        blob[k] = meta[k]
    blob["textures"] = []
    for i, t in enumerate(meta["textures"]):
        blob["textures"].append(dict(format=t["format"], width=t["width"], height=t["height"], lod_bias=t["lod_bias"],
                                     levels=[z[f"tex{i}_l{l}"] for l in range(t["levels"])]))
    return blob


# ----------------------------------------------------------------------------- Mitsuba 0.5 XML subset
_KNOWN_IOR = {"vacuum": 1.0, "helium": 1.00004, "hydrogen": 1.00013, "air": 1.00028, "carbon dioxide": 1.00045, "water": 1.3330,
              "acetone": 1.36, "ethanol": 1.361, "carbon tetrachloride": 1.461, "glycerol": 1.4729, "benzene": 1.501,
              "silicone oil": 1.52045, "bromine": 1.661, "water ice": 1.31, "fused quartz": 1.458, "pyrex": 1.470,
              "acrylic glass": 1.49, "polypropylene": 1.49, "bk7": 1.5046, "sodium chloride": 1.544, "amber": 1.55, "pet": 1.575,
              "diamond": 2.419}


def _vec(s, n=3):
    parts = [float(x) for x in s.replace(",", " ").split()]
    if len(parts) == 1:
        parts = parts * n
    return parts


def _child_by_name(node, name):
    for c in node:
        if c.get("name") == name:
            return c
    return None


def _child_value(node, name, default=None, cast=float):
    c = _child_by_name(node, name)
    if c is None:
        return default
    return cast(c.get("value"))


def _parse_transform_matrix(node):
    world = m_identity()
    tr = node.find("transform")
    if tr is None:
        return world
    for t in tr:
        if t.tag == "matrix":
            world = np.array(_vec(t.get("value"), 16), dtype=np.float64).reshape(4, 4) @ world
        elif t.tag == "lookat":
            o = np.array(_vec(t.get("origin", "0 0 0"))); tg = np.array(_vec(t.get("target", "0 0 -1"))); up = np.array(_vec(t.get("up", "0 1 0")))
            world = m_translate(o) @ m_rotation(q_look_rotation(tg - o, up)) @ world
        elif t.tag == "scale":
            if t.get("value") is not None:
                world = m_scale(float(t.get("value"))) @ world
            else:
                world = m_scale(float(t.get("x", 1)), float(t.get("y", 1)), float(t.get("z", 1))) @ world
        elif t.tag == "rotate":
            ax = (float(t.get("x", 0)), float(t.get("y", 0)), float(t.get("z", 0)))
            if any(ax):
                world = m_rotation(q_axis_angle(ax, math.radians(float(t.get("angle", 0))))) @ world
        elif t.tag == "translate":
            world = m_translate((float(t.get("x", 0)), float(t.get("y", 0)), float(t.get("z", 0)))) @ world
    return world


class _MitsubaWalker:
    def __init__(self, desc, base, load_textures=True):
        self.desc, self.base, self.load_textures = desc, base, load_textures
        self.material_map, self.texture_map, self.texture_cache, self.mesh_cache, self.shape_groups = {}, {}, {}, {}, {}

    def texture(self, node, rgb):
        typ = node.get("type")
        if typ == "scale":
            sc = _child_by_name(node, "scale")
            if sc is not None:
                rgb[:] = [a * b for a, b in zip(rgb, _vec(sc.get("value")))]
            node = node.find("texture"); typ = node.get("type")
        if typ != "bitmap":
            return INVALID
        rel = _child_by_name(node, "filename").get("value").replace("\\\\", "/").replace("\\", "/")
        path = os.path.join(self.base, rel)
        if path not in self.texture_cache:
            tex = None
            if self.load_textures:
                try:
                    tex = make_texture(load_image_u8(path))
                except Exception:
                    tex = None
            self.desc.textures.append(tex if tex is not None else pink_texture())
            self.texture_cache[path] = len(self.desc.textures) - 1
        h = self.texture_cache[path]
        if node.get("id"):
            self.texture_map[node.get("id")] = h
        return h

    def rgb_or_texture(self, node, name):
        rgb, tex = [1.0, 1.0, 1.0], INVALID
        c = _child_by_name(node, name)
        if c is not None:
            if c.tag == "rgb":
                rgb = _vec(c.get("value", "1"))
            elif c.tag == "srgb":
                rgb = list(gamma_to_linear(np.array(_vec(c.get("value", "1")))))
            elif c.tag == "texture":
                tex = self.texture(c, rgb)
                sc = _child_by_name(c, "scale")
                if sc is not None:
                    rgb = _vec(sc.get("value", "1"))
            elif c.tag == "ref":
                tex = self.texture_map.get(c.get("id"), INVALID)
        return rgb, tex

    def material(self, node):
        d = self.desc
        if node.tag != "bsdf":
            em = node.find("emitter")
            if em is not None:
                return d.add_material(Material(MAT_LIGHT, "emitter", emission=_vec(_child_by_name(em, "radiance").get("value"))))
            ref = node.find("ref")
            if ref is not None:
                return self.material_map.get(ref.get("id"), 0)
            bsdf = node.find("bsdf")
            if bsdf is None:
                return 0
        else:
            bsdf = node
        name = bsdf.get("id")
        inner = bsdf
        while inner.get("type") in ("twosided", "mask", "bumpmap", "coating"):
            child = inner.find("bsdf")
            if child is None:
                ref = inner.find("ref")
                return self.material_map.get(ref.get("id"), 0) if ref is not None else 0
            inner = child
            if name is None:
                name = inner.get("id")
        typ = inner.get("type")
        m = Material(name=name or "Material")
        if typ == "diffuse":
            m.kind = MAT_DIFFUSE; m.diffuse, m.texture = self.rgb_or_texture(inner, "reflectance")
        elif typ in ("conductor", "roughconductor"):
            m.kind = MAT_CONDUCTOR
            m.roughness = 0.0 if typ == "conductor" else _child_value(inner, "alpha", 0.5)
            ms = _child_by_name(inner, "material")
            if ms is not None and ms.get("value") == "none":
                m.eta, m.k = (0, 0, 0), (1, 1, 1)
            else:
                e, k = _child_by_name(inner, "eta"), _child_by_name(inner, "k")
                m.eta = tuple(_vec(e.get("value"))) if e is not None else (1.33,) * 3
                m.k = tuple(_vec(k.get("value"))) if k is not None else (1.0,) * 3
        elif typ in ("plastic", "roughplastic", "roughdiffuse", "phong"):
            m.kind = MAT_PLASTIC; m.diffuse, m.texture = self.rgb_or_texture(inner, "diffuseReflectance")
            if typ == "plastic":
                m.roughness = 0.0
            elif typ == "phong":
                m.roughness = (0.5 * _child_value(inner, "exponent", 1.0) + 1.0) ** 0.25
            else:
                m.roughness = _child_value(inner, "alpha", 0.5)
        elif typ in ("thindielectric", "dielectric", "roughdielectric"):
            def ior(nm, dflt):
                c = _child_by_name(inner, nm)
                if c is not None and c.tag == "string":
                    return _KNOWN_IOR[c.get("value")]
                return _child_value(inner, nm, dflt)
            i, e = ior("intIOR", 1.33), ior("extIOR", 1.0)
            m.kind = MAT_DIELECTRIC; m.ior = i if e == 0.0 else i / e
            m.roughness = _child_value(inner, "alpha", 0.5) if typ == "roughdielectric" else 0.0
        elif typ == "difftrans":
            m.kind = MAT_DIFFUSE; m.diffuse, m.texture = self.rgb_or_texture(inner, "transmittance")
        else:
            return 0
        return d.add_material(m)

    def shape_mesh(self, node):
        typ = node.get("type")
        if typ in ("obj", "ply", "serialized", "hair"):                    # MitsubaLoader.cpp:434-515
            from . import mesh_loaders
            path = os.path.join(self.base, _child_by_name(node, "filename").get("value").replace("\\", "/"))
            key, source = path, path
            if typ == "serialized":
                shape_index = int(_child_value(node, "shapeIndex", 0))
                key, source = "%s#%d" % (path, shape_index), "%s.shape_%d" % (path, shape_index)      # the .bvh cache is per sub-mesh (:495)
            if key not in self.mesh_cache:
                if typ == "obj":
                    tri = load_obj(path)
                elif typ == "ply":
                    tri = mesh_loaders.load_ply(path)
                elif typ == "serialized":
                    tri = mesh_loaders.load_serialized(path, shape_index)
                else:
                    tri = mesh_loaders.load_hair(path, float(_child_value(node, "radius", 0.0025)))
                self.mesh_cache[key] = self.desc.add_mesh_data(tri, source=source if typ == "obj" else None)
            return self.mesh_cache[key]
        m = _parse_transform_matrix(node)
        if typ == "rectangle":
            return self.desc.add_mesh_data(geo_rectangle(m))
        if typ == "cube":
            return self.desc.add_mesh_data(geo_cube(m))
        if typ == "sphere":
            r = _child_value(node, "radius", 1.0)
            c = _child_by_name(node, "center")
            ctr = (float(c.get("x", 0)), float(c.get("y", 0)), float(c.get("z", 0))) if c is not None else (0, 0, 0)
            return self.desc.add_mesh_data(geo_icosphere(m @ m_translate(ctr) @ m_scale(r)))
        return INVALID

    def walk(self, node):
        d = self.desc
        if node.tag == "bsdf":
            h = self.material(node)
            self.material_map[d.materials[h].name] = h
        elif node.tag == "texture":
            self.texture(node, [1.0, 1.0, 1.0])
        elif node.tag == "shape":
            typ = node.get("type")
            if typ == "shapegroup":
                sh = node.find("shape")
                if sh is not None:
                    self.shape_groups[node.get("id")] = (self.shape_mesh(sh), self.material(sh))
            elif typ == "instance":
                ref = node.find("ref")
                grp = self.shape_groups.get(ref.get("id")) if ref is not None else None
                if grp and grp[0] != INVALID:
                    pos, rot, sc = m_decompose(_parse_transform_matrix(node), (0, 0, 1))
                    d.instances.append(Instance(grp[0], grp[1], pos, rot, sc, ref.get("id")))
            else:
                md = self.shape_mesh(node)
                mat = self.material(node)
                if md != INVALID:
                    inst = Instance(md, mat, name=typ)
                    if typ not in ("rectangle", "cube", "disk", "cylinder", "sphere"):
                        inst.position, inst.rotation, inst.scale = m_decompose(_parse_transform_matrix(node), (0, 0, 1))
                    d.instances.append(inst)
        elif node.tag == "sensor":
            if node.get("type") in ("perspective", "perspective_rdist", "thinlens"):
                fov = _child_by_name(node, "fov")
                if fov is not None:
                    d.cam_fov = float(fov.get("value")) / 180.0 * math.pi
                if node.get("type") == "perspective":
                    d.cam_aperture = 0.0
                else:
                    d.cam_aperture = _child_value(node, "apertureRadius", 0.05); d.cam_focal = _child_value(node, "focusDistance", 10.0)
                d.cam_position, d.cam_rotation, _ = m_decompose(_parse_transform_matrix(node), (0, 0, -1))
            film = node.find("film")
            if film is not None:
                d.width = int(_child_value(film, "width", d.width)); d.height = int(_child_value(film, "height", d.height))
        elif node.tag == "integrator":
            d.num_bounces = int(_child_value(node, "maxDepth", d.num_bounces))
        elif node.tag == "emitter":
            if node.get("type") == "area" and node.get("id"):
                h = d.add_material(Material(MAT_LIGHT, node.get("id"), emission=_vec(_child_by_name(node, "radiance").get("value"))))
                self.material_map[node.get("id")] = h
        else:
            for c in node:
                self.walk(c)


def load_mitsuba(xml_path, sky_path=None, load_textures=True, sky_max_width=2500):
    import xml.etree.ElementTree as ET
    desc = SceneDesc()
    desc.source = os.path.basename(os.path.dirname(os.path.abspath(xml_path))) + "/" + os.path.basename(xml_path)
    root = ET.parse(xml_path).getroot()
    _MitsubaWalker(desc, os.path.dirname(os.path.abspath(xml_path)), load_textures).walk(root)
    if sky_path and os.path.exists(sky_path):
        desc.sky = load_hdr(sky_path, max_width=sky_max_width)
    return desc


# ----------------------------------------------------------------------------- procedural scenes
def _soup(rng, count, extent, size):
    c = rng.uniform(-extent, extent, size=(count, 1, 3))
    p = c + rng.normal(0, size, size=(count, 3, 3))
    g = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-20)
    n = np.repeat(g[:, None, :], 3, 1)
    t = rng.uniform(0, 1, size=(count, 3, 2))
    return finish_triangles(p, n, t)


def procedural_scene(kind="atrium", seed=7, width=256, height=256, detail=1.0, all_materials=False):
    """Self-contained synthetic scenes.
      'soup'   : random triangle soup + floor + area light (identity and non-identity instances)
      'atrium' : Sponza-like courtyard: floor, walls with arched openings, two rows of tessellated columns (instanced,
                 rotated / scaled), draped curtains, two emissive spheres, open to the sky.  detail scales triangle count."""
    rng = np.random.default_rng(seed)
    d = SceneDesc()
    d.source = f"procedural:{kind}:seed{seed}:detail{detail}"
    d.width, d.height, d.num_bounces = width, height, 4
    white = d.add_material(Material(MAT_DIFFUSE, "white", diffuse=(0.73, 0.71, 0.68)))
    red = d.add_material(Material(MAT_DIFFUSE, "red", diffuse=(0.63, 0.065, 0.05)))
    green = d.add_material(Material(MAT_DIFFUSE, "green", diffuse=(0.14, 0.45, 0.091)))
    light = d.add_material(Material(MAT_LIGHT, "light", emission=(17, 12, 4)))
    extra = []
    if all_materials:
        extra = [d.add_material(Material(MAT_PLASTIC, "plastic", diffuse=(0.2, 0.8, 0.8), roughness=0.2)),
                 d.add_material(Material(MAT_DIELECTRIC, "glass", ior=1.5, roughness=0.3)),
                 d.add_material(Material(MAT_CONDUCTOR, "gold", eta=(1.45, 0.43, 0.21), k=(1.95, 2.46, 3.27), roughness=0.3)),
                 d.add_material(Material(MAT_DIELECTRIC, "smooth_glass", ior=1.33, roughness=0.0))]
    if kind == "cornell":
        # the classic box: five walls, two blocks, one ceiling lamp (own dimensions; all instances identity like the reference's primitives)
        rot = lambda ax, a: m_rotation(q_axis_angle(ax, a))
        walls = [(m_translate((0, 0, 0)) @ rot((1, 0, 0), -math.pi / 2), white), (m_translate((0, 2, 0)) @ rot((1, 0, 0), math.pi / 2), white),
                 (m_translate((0, 1, -1)), white), (m_translate((-1, 1, 0)) @ rot((0, 1, 0), math.pi / 2), red), (m_translate((1, 1, 0)) @ rot((0, 1, 0), -math.pi / 2), green)]
        for m, mat in walls:
            d.instances.append(Instance(d.add_mesh_data(geo_rectangle(m)), mat))
        d.instances.append(Instance(d.add_mesh_data(geo_cube(m_translate((0.33, 0.3, 0.35)) @ rot((0, 1, 0), -0.3) @ m_scale(0.3))), white))
        d.instances.append(Instance(d.add_mesh_data(geo_cube(m_translate((-0.34, 0.6, -0.3)) @ rot((0, 1, 0), 0.33) @ m_scale(0.3, 0.6, 0.3))), white))
        d.instances.append(Instance(d.add_mesh_data(geo_rectangle(m_translate((0, 1.98, 0)) @ rot((1, 0, 0), math.pi / 2) @ m_scale(0.24, 0.19, 1.0))), light))
        d.cam_position = np.array([0.0, 1.0, 3.4]); d.cam_rotation = np.array([0.0, 0.0, 0.0, 1.0]); d.cam_fov = math.radians(40)
        return d
    if kind == "soup":
        n = max(8, int(400 * detail))
        soup = d.add_mesh_data(_soup(rng, n, 1.0, 0.15))
        floor = d.add_mesh_data(geo_rectangle(m_translate((0, -1.2, 0)) @ m_rotation(q_axis_angle((1, 0, 0), -math.pi / 2)) @ m_scale(4.0)))
        lamp = d.add_mesh_data(geo_rectangle(m_translate((0, 2.2, 0)) @ m_rotation(q_axis_angle((1, 0, 0), math.pi / 2)) @ m_scale(0.6)))
        ball = d.add_mesh_data(geo_icosphere(m_identity(), 2))
        d.instances += [Instance(soup, white), Instance(floor, green), Instance(lamp, light),
                        Instance(soup, red, position=(2.2, 0.3, -0.5), rotation=q_axis_angle((0, 1, 0), 0.7), scale=0.6),
                        Instance(ball, extra[0] if extra else white, position=(-1.8, -0.5, 0.8), scale=0.5),
                        Instance(ball, extra[1] if extra else red, position=(-0.4, -0.6, 1.6), rotation=q_axis_angle((1, 0, 0), 0.3), scale=0.45),
                        Instance(ball, extra[2] if extra else green, position=(1.0, -0.7, 1.5), scale=0.4),
                        Instance(ball, light, position=(-2.0, 1.5, -1.0), scale=0.2)]
        if extra:
            d.instances.append(Instance(ball, extra[3], position=(0.2, 0.9, 1.2), scale=0.35))
        d.cam_position = np.array([0.0, 0.4, 5.0]); d.cam_rotation = np.array([0.0, 0.0, 0.0, 1.0]); d.cam_fov = math.radians(55)
        return d
    if kind != "atrium":
        raise ValueError(kind)
    sub = 3 if detail >= 1.0 else 2
    # column: stack of squashed spheres (many small triangles, like Sponza's ornate columns)
    col_parts = [geo_icosphere(m_translate((0, y, 0)) @ m_scale(0.35, 0.55, 0.35), sub) for y in np.linspace(0.4, 5.6, max(2, int(8 * detail)))]
    column = d.add_mesh_data(tuple(np.concatenate([c[k] for c in col_parts]) for k in range(3)))
    floor = d.add_mesh_data(_grid_plane(24, 10, int(48 * detail), int(20 * detail), y=0.0))
    wall = d.add_mesh_data(_grid_wall(24, 8, int(64 * detail), int(24 * detail), rng))
    cloth = d.add_mesh_data(_curtain(3.0, 4.0, int(40 * detail), int(40 * detail)))
    lamp = d.add_mesh_data(geo_icosphere(m_identity(), 1))
    d.instances.append(Instance(floor, white, name="floor"))
    d.instances.append(Instance(wall, white, position=(0, 0, -5.0), name="wall_back"))
    d.instances.append(Instance(wall, white, position=(0, 0, 5.0), rotation=q_axis_angle((0, 1, 0), math.pi), name="wall_front"))
    d.instances.append(Instance(wall, red, position=(-12.0, 0, 0), rotation=q_axis_angle((0, 1, 0), math.pi / 2), scale=0.42, name="wall_left"))
    d.instances.append(Instance(wall, green, position=(12.0, 0, 0), rotation=q_axis_angle((0, 1, 0), -math.pi / 2), scale=0.42, name="wall_right"))
    mats = [white, white, red, green] + extra
    for i, x in enumerate(np.linspace(-10, 10, 9)):
        for z in (-3.0, 3.0):
            d.instances.append(Instance(column, mats[(i + (z > 0)) % len(mats)] if extra else white, position=(x, 0, z),
                                        rotation=q_axis_angle((0, 1, 0), 0.37 * i), scale=1.0 + 0.04 * (i % 3), name="column"))
    for i, x in enumerate((-7.5, -2.5, 2.5, 7.5)):
        d.instances.append(Instance(cloth, (red, green)[i % 2], position=(x, 2.0, -3.4 if i % 2 else 3.4), name="curtain"))
    d.instances.append(Instance(lamp, light, position=(-4.0, 4.5, 0.0), scale=0.25, name="lamp"))
    d.instances.append(Instance(lamp, light, position=(5.0, 3.5, 0.5), scale=0.25, name="lamp"))
    for m in d.materials:
        if m.name == "light":
            m.emission = (12, 12, 12)
    d.cam_position = np.array([-9.5, 2.2, 0.6]); d.cam_rotation = q_axis_angle((0, 1, 0), -math.pi / 2 + 0.12); d.cam_fov = math.radians(85)
    return d


def _grid_plane(sx, sz, nx, nz, y=0.0):
    nx, nz = max(nx, 1), max(nz, 1)
    xs = np.linspace(-sx / 2, sx / 2, nx + 1); zs = np.linspace(-sz / 2, sz / 2, nz + 1)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    P = np.stack([X, np.full_like(X, y), Z], -1)
    U = np.stack([X / 2.0, Z / 2.0], -1)
    return _quads(P, U, (0, 1, 0))


def _grid_wall(sx, sy, nx, ny, rng):
    nx, ny = max(nx, 2), max(ny, 2)
    xs = np.linspace(-sx / 2, sx / 2, nx + 1); ys = np.linspace(0, sy, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    Zb = 0.08 * np.sin(3.1 * X) * np.cos(2.3 * Y)  # gentle relief so normals vary
    P = np.stack([X, Y, Zb], -1)
    U = np.stack([X / 3.0, Y / 3.0], -1)
    p, n, t = _quads(P, U, (0, 0, 1))
    # punch arched openings: drop quads whose centre lies inside an arch
    c = p.mean(1)
    period = 2.5
    fx = np.abs(((c[:, 0] + 100 * period) % period) - period / 2)
    inside = (fx < 0.8) & (c[:, 1] < 3.2 - 1.2 * (fx / 0.8) ** 2) & (c[:, 1] > 0.0)
    keep = ~inside
    return p[keep], n[keep], t[keep]


def _curtain(w, h, nx, ny):
    nx, ny = max(nx, 2), max(ny, 2)
    xs = np.linspace(-w / 2, w / 2, nx + 1); ys = np.linspace(-h / 2, h / 2, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    Zc = 0.18 * np.sin(7.0 * X) * (0.6 + 0.4 * (Y / h + 0.5))
    P = np.stack([X, Y, Zc], -1)
    U = np.stack([X / w + 0.5, Y / h + 0.5], -1)
    return _quads(P, U, (0, 0, 1))


def _quads(P, U, up):
    a, b, c, dd = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    ua, ub, uc, ud = U[:-1, :-1], U[1:, :-1], U[1:, 1:], U[:-1, 1:]
    p = np.concatenate([np.stack([a, b, c], -2).reshape(-1, 3, 3), np.stack([a, c, dd], -2).reshape(-1, 3, 3)])
    t = np.concatenate([np.stack([ua, ub, uc], -2).reshape(-1, 3, 2), np.stack([ua, uc, ud], -2).reshape(-1, 3, 2)])
    # smooth vertex normals from the height field gradient
    g = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-20)
    flip = (g @ np.asarray(up, dtype=np.float64)) < 0
    g[flip] = -g[flip]
    n = np.repeat(g[:, None, :], 3, 1)
    return finish_triangles(p, n, t)
