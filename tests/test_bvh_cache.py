"""The `.bvh` cache (file type 7) of Src/Assets/BVHLoader.cpp: byte layout, round trip, reuse rules, conversion without rebuild."""
import os
import struct
import time
import zlib

import numpy as np
import pytest

from gpu_raytracer_b200 import bvh_cache, scene


def _soup(n, seed=3):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-4, 4, (n, 1, 3)).astype(np.float32)
    p = c + rng.normal(0, 0.2, (n, 3, 3)).astype(np.float32)
    nrm = np.tile(np.array([0, 1, 0], dtype=np.float32), (n, 3, 1))
    uv = rng.uniform(0, 1, (n, 3, 2)).astype(np.float32)
    return p, nrm, uv


def test_header_and_stream_layout(tmp_path):
    p, n, t = _soup(300)
    raw = scene.build_blas(p, 2, 4.0, 0.0)
    nodes, idx = raw.export(0, 0)
    path = str(tmp_path / "m.obj.bvh")
    bvh_cache.save(path, bvh_cache.CachedBVH(bvh_cache.pack_triangles(p, n, t), nodes.view(bvh_cache.NODE2_DTYPE), idx, bvh_cache.BVH_TYPE_SAH, False, 4.0, 1.0))
    data = open(path, "rb").read()
    # the MSVC struct of BVHLoader.cpp:19-32: 4 + 1 + 1 + 1 (+1 pad) + 4 + 4 + 3 x 4 = 28 bytes
    assert data[:4] == b"BVH\0" and data[4] == 7 and data[5] == 0 and data[6] == 0
    assert struct.unpack_from("<ff", data, 8) == (4.0, 1.0)
    assert struct.unpack_from("<iii", data, 16) == (300, raw.node_count, 300)
    # one raw deflate stream (no zlib header) of Triangle[96 B], BVHNode2[32 B], int[]
    payload = zlib.decompressobj(-15).decompress(data[28:])
    assert len(payload) == 300 * 96 + raw.node_count * 32 + 300 * 4
    first = np.frombuffer(payload, dtype=np.float32, count=24)
    assert np.array_equal(first[:9], p[0].reshape(-1)) and np.array_equal(first[9:18], n[0].reshape(-1)) and np.array_equal(first[18:], t[0].reshape(-1))
    assert payload[300 * 96:300 * 96 + raw.node_count * 32] == nodes.tobytes()


def test_round_trip_and_conversion_equals_direct_build(tmp_path):
    p, n, t = _soup(2000, seed=5)
    raw = scene.build_blas(p, 2, 4.0, 0.0)
    nodes, idx = raw.export(0, 0)
    path = str(tmp_path / "mesh.obj.bvh")
    bvh_cache.save(path, bvh_cache.CachedBVH(bvh_cache.pack_triangles(p, n, t), nodes.view(bvh_cache.NODE2_DTYPE), idx))
    c = bvh_cache.load(path)
    assert np.array_equal(c.triangles["position"], p) and np.array_equal(c.triangles["tex_coord"], t)
    assert c.nodes.tobytes() == nodes.tobytes() and np.array_equal(c.indices, idx)
    for kind, leaf in ((8, 1.0), (2, 1.0), (2, 0.0)):
        direct = scene.build_blas(p, kind, 4.0, leaf)
        conv = scene.blas_from_bvh2(c.nodes, c.indices, kind, 4.0, leaf)
        dn, di = direct.export(0, 0); cn, ci = conv.export(0, 0)
        assert dn.tobytes() == cn.tobytes() and np.array_equal(di, ci), (kind, leaf)


def test_reuse_rules(tmp_path):
    p, n, t = _soup(500, seed=9)
    mesh = str(tmp_path / "thing.obj")
    open(mesh, "w").write("# stand-in for the mesh file: only its mtime matters here\n")
    assert bvh_cache.try_to_load(mesh) is None                         # no cache yet
    first = scene.build_blas_cached(mesh, (p, n, t), 8)
    assert os.path.exists(mesh + ".bvh")
    again = scene.build_blas_cached(mesh, (p, n, t), 8)                # served from the file
    assert first.export()[0].tobytes() == again.export()[0].tobytes()
    assert bvh_cache.try_to_load(mesh) is not None
    assert bvh_cache.try_to_load(mesh, sah_cost_node=2.0) is None      # other settings -> rebuild (BVHLoader.cpp:163-170)
    assert bvh_cache.try_to_load(mesh, bvh_type=bvh_cache.BVH_TYPE_SBVH) is None
    assert bvh_cache.try_to_load(mesh, force_rebuild=True) is None
    future = time.time() + 60
    os.utime(mesh, (future, future))                                   # mesh newer than its cache -> stale (BVHLoader.cpp:35)
    assert bvh_cache.try_to_load(mesh) is None


def test_foreign_and_damaged_files(tmp_path):
    bad = str(tmp_path / "x.bvh")
    open(bad, "wb").write(b"BVH\0" + bytes([6]) + bytes(23))
    with pytest.raises(ValueError):
        bvh_cache.load(bad)                                            # older file version
    open(bad, "wb").write(b"nope" + bytes(24))
    with pytest.raises(ValueError):
        bvh_cache.load(bad)
    p, n, t = _soup(64)
    raw = scene.build_blas(p, 2, 4.0, 0.0)
    nodes, idx = raw.export(0, 0)
    broken = nodes.copy().view(bvh_cache.NODE2_DTYPE)
    broken["left_or_first"][0] = 0                                     # the root points at itself
    with pytest.raises(RuntimeError):
        scene.blas_from_bvh2(broken, idx, 8)
