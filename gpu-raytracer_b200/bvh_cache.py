"""The reference's on-disk BVH cache (`<mesh file>.bvh`, file type version 7), reader and writer
(Src/Assets/BVHLoader.cpp:15-260, Src/Assets/BVHLoader.h:13-14).

Layout (little endian, the MSVC x64 struct of BVHLoader.cpp:19-32, 28 bytes):

    char  filetype_identifier[4]   "BVH\\0"
    char  filetype_version         7
    char  underlying_bvh_type      0 = SAH BVH, 1 = SBVH   (Config.h:25-30, BVH.h:95-102)
    bool  bvh_is_optimized         + 1 byte of padding
    float sah_cost_node, sah_cost_leaf
    int   num_triangles, num_nodes, num_indices

followed by ONE raw deflate stream (miniz tdefl without the zlib header: `tdefl_init(.., flags = 256 probes)`, BVHLoader.cpp:232;
read back by tinfl without TINFL_FLAG_PARSE_ZLIB_HEADER, :129) of three arrays back to back:

    Triangle[num_triangles]   96 bytes: position_0..2, normal_0..2 (float3 each), tex_coord_0..2 (float2 each)  (Renderer/Triangle.h:11-22)
    BVHNode2[num_nodes]       32 bytes: aabb min, aabb max, left|first, count:30 | axis:2                      (BVH/BVH.h:11-23)
    int[num_indices]          primitive order of the leaves

A file is reused only if it is at least as new as the mesh file and was written with the same settings (BVHLoader.cpp:35,163-170);
otherwise the caller rebuilds and overwrites it.  Only zlib (raw deflate, wbits = -15) and numpy are needed.
"""
from __future__ import annotations

import os
import struct
import zlib
from dataclasses import dataclass

import numpy as np

FILE_EXTENSION = ".bvh"
FILETYPE_VERSION = 7
BVH_TYPE_SAH, BVH_TYPE_SBVH = 0, 1
_HEADER = struct.Struct("<4sbb?xffiii")          # 28 bytes

TRIANGLE_DTYPE = np.dtype([("position", "<f4", (3, 3)), ("normal", "<f4", (3, 3)), ("tex_coord", "<f4", (3, 2))])
NODE2_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("left_or_first", "<i4"), ("count_axis", "<u4")])
assert _HEADER.size == 28 and TRIANGLE_DTYPE.itemsize == 96 and NODE2_DTYPE.itemsize == 32


@dataclass
class CachedBVH:
    triangles: np.ndarray        # TRIANGLE_DTYPE[num_triangles]
    nodes: np.ndarray            # NODE2_DTYPE[num_nodes]  (node 1 is the reference's unused alignment dummy)
    indices: np.ndarray          # int32[num_indices]
    bvh_type: int = BVH_TYPE_SAH
    optimized: bool = False
    sah_cost_node: float = 4.0
    sah_cost_leaf: float = 1.0


def bvh_filename(mesh_filename: str) -> str:
    """BVHLoader::get_bvh_filename: the extension is appended, not substituted (`sponza.obj.bvh`)."""
    return mesh_filename + FILE_EXTENSION


def save(path: str, bvh: CachedBVH, level: int = 9) -> None:
    tri = np.ascontiguousarray(bvh.triangles, dtype=TRIANGLE_DTYPE)
    nodes = np.ascontiguousarray(bvh.nodes, dtype=NODE2_DTYPE)
    idx = np.ascontiguousarray(bvh.indices, dtype="<i4")
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)            # raw deflate, 32 KB window (TINFL_LZ_DICT_SIZE)
    with open(path, "wb") as f:
        f.write(_HEADER.pack(b"BVH\0", FILETYPE_VERSION, int(bvh.bvh_type), bool(bvh.optimized), float(bvh.sah_cost_node), float(bvh.sah_cost_leaf),
                             tri.shape[0], nodes.shape[0], idx.shape[0]))
        for block in (tri, nodes, idx):
            f.write(comp.compress(block.tobytes()))
        f.write(comp.flush())


def load(path: str) -> CachedBVH:
    """Reads a version-7 file; raises ValueError on a foreign / truncated / older file (the reference returns false and rebuilds)."""
    with open(path, "rb") as f:
        head = f.read(_HEADER.size)
        if len(head) != _HEADER.size:
            raise ValueError("truncated BVH file header")
        magic, version, bvh_type, optimized, cost_node, cost_leaf, n_tri, n_nodes, n_idx = _HEADER.unpack(head)
        if magic != b"BVH\0":
            raise ValueError("not a BVH cache file")
        if version != FILETYPE_VERSION:
            raise ValueError(f"BVH file version {version}, expected {FILETYPE_VERSION}")
        if min(n_tri, n_nodes, n_idx) < 0:
            raise ValueError("corrupt BVH file header")
        want = n_tri * TRIANGLE_DTYPE.itemsize + n_nodes * NODE2_DTYPE.itemsize + n_idx * 4
        d = zlib.decompressobj(-15)
        raw = d.decompress(f.read(), want) if want else b""
    if len(raw) != want:
        raise ValueError("BVH file payload is shorter than its header says")
    a, b = n_tri * TRIANGLE_DTYPE.itemsize, n_tri * TRIANGLE_DTYPE.itemsize + n_nodes * NODE2_DTYPE.itemsize
    return CachedBVH(np.frombuffer(raw, dtype=TRIANGLE_DTYPE, count=n_tri).copy(), np.frombuffer(raw, dtype=NODE2_DTYPE, count=n_nodes, offset=a).copy(),
                     np.frombuffer(raw, dtype="<i4", count=n_idx, offset=b).copy(), bvh_type, optimized, cost_node, cost_leaf)


def try_to_load(mesh_filename: str, bvh_type: int = BVH_TYPE_SAH, optimized: bool = False, sah_cost_node: float = 4.0, sah_cost_leaf: float = 1.0,
                force_rebuild: bool = False):
    """BVHLoader::try_to_load: the cached BVH, or None when there is none / it is older than the mesh / it was made with other settings."""
    path = bvh_filename(mesh_filename)
    if force_rebuild or not os.path.exists(mesh_filename) or not os.path.exists(path) or os.path.getmtime(mesh_filename) > os.path.getmtime(path):
        return None
    try:
        c = load(path)
    except (ValueError, zlib.error, OSError):
        return None
    if c.bvh_type != bvh_type or bool(c.optimized) != bool(optimized) or c.sah_cost_node != np.float32(sah_cost_node) or c.sah_cost_leaf != np.float32(sah_cost_leaf):
        return None
    return c


def pack_triangles(positions, normals, tex_coords) -> np.ndarray:
    """[n,3,3] positions, [n,3,3] normals, [n,3,2] texture coordinates -> Triangle records."""
    n = np.asarray(positions).shape[0]
    t = np.zeros(n, dtype=TRIANGLE_DTYPE)
    t["position"] = np.asarray(positions, dtype=np.float32).reshape(n, 3, 3)
    t["normal"] = np.asarray(normals, dtype=np.float32).reshape(n, 3, 3)
    t["tex_coord"] = np.asarray(tex_coords, dtype=np.float32).reshape(n, 3, 2)
    return t
