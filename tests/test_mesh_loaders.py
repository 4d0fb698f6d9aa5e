"""PLY / Mitsuba serialized / Mitsuba hair loaders (Src/Assets/PLYLoader.cpp, Src/Assets/Mitsuba/{Serialized,Mitshair}Loader.cpp).
No file of these formats ships with the reference: the loaders are pinned against files written from the format descriptions, in every
variant they accept."""
import struct
import zlib

import numpy as np
import pytest

from gpu_raytracer_b200 import mesh_loaders as ml, scene


def _quad_mesh():
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1.0]], dtype=np.float64)
    faces = [[0, 1, 2, 3], [0, 1, 4], [1, 2, 4]]                       # one quad (fan -> 2 triangles) + 2 triangles
    n = np.tile([0.0, 0.0, 1.0], (5, 1))
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.25]], dtype=np.float64)
    return v, faces, n, uv


def _expected(v, faces, n, uv):
    tri = []
    for f in faces:
        for k in range(1, len(f) - 1):
            tri.append((f[0], f[k], f[k + 1]))
    idx = np.array(tri)
    uvf = np.stack([uv[:, 0], 1.0 - uv[:, 1]], axis=1)
    return scene.finish_triangles(v[idx].astype(np.float32), n[idx].astype(np.float32), uvf[idx].astype(np.float32))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("vtype,itypes", [("float", ("uchar", "int")), ("double", ("ushort", "uint"))])
def test_ply_variants(tmp_path, fmt, vtype, itypes):
    v, faces, n, uv = _quad_mesh()
    path = str(tmp_path / "m.ply")
    ml.save_ply(path, v, faces, normals=n, uvs=uv, fmt=fmt, vertex_type=vtype, index_types=itypes, extra=True)
    p, nn, t = ml.load_ply(path)
    ep, en, et = _expected(v, faces, n, uv)
    assert p.shape == (4, 3, 3)
    assert np.array_equal(p, ep) and np.array_equal(nn, en) and np.array_equal(t, et)


def test_ply_without_normals_gets_geometric_normals_and_matches_obj(tmp_path):
    v, faces, _, _ = _quad_mesh()
    path = str(tmp_path / "plain.ply")
    ml.save_ply(path, v, faces)
    p, n, t = ml.load_ply(path)
    g = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]); g /= np.linalg.norm(g, axis=1, keepdims=True)
    assert np.allclose(n, np.repeat(g[:, None, :], 3, axis=1), atol=1e-6)
    assert np.array_equal(t[..., 1], np.ones_like(t[..., 1]))              # v = 1 - 0, as the loader flips it (PLYLoader.cpp:276)
    # the same geometry through the OBJ loader
    obj = str(tmp_path / "plain.obj")
    with open(obj, "w") as f:
        for a in v:
            f.write("v %r %r %r\n" % tuple(float(x) for x in a))
        for face in faces:
            f.write("f " + " ".join(str(i + 1) for i in face) + "\n")
    po, no, _ = scene.load_obj(obj)
    assert np.array_equal(p, po) and np.allclose(n, no, atol=1e-6)
    with pytest.raises(ValueError):
        open(str(tmp_path / "bad.ply"), "wb").write(b"ply\nformat ascii 1.0\nelement edge 3\nend_header\n"); ml.load_ply(str(tmp_path / "bad.ply"))


@pytest.mark.parametrize("version,double", [(3, False), (4, False), (4, True)])
def test_serialized_variants(tmp_path, version, double):
    v, faces, n, uv = _quad_mesh()
    tris = np.array([[0, 1, 2], [0, 2, 3], [0, 1, 4]])
    meshes = [dict(vertices=v, faces=tris, normals=n, uvs=uv, name="first"),
              dict(vertices=v * 2.0, faces=tris[:2], name="second"),
              dict(vertices=v + 1.0, faces=tris, face_normals=True)]
    path = str(tmp_path / "scene.serialized")
    ml.save_serialized(path, meshes, version=version, double=double)
    data = open(path, "rb").read()
    assert struct.unpack_from("<HH", data, 0) == (0x041C, version) and struct.unpack_from("<I", data, len(data) - 4)[0] == 3
    p0, n0, t0 = ml.load_serialized(path, 0)
    ep, en, et = scene.finish_triangles(v[tris].astype(np.float32), n[tris].astype(np.float32), uv[tris].astype(np.float32))
    assert np.array_equal(p0, ep) and np.array_equal(n0, en) and np.array_equal(t0, et)     # no v flip in this format (SerializedLoader.cpp:196-200)
    p1, n1, t1 = ml.load_serialized(path, 1)
    assert np.array_equal(p1, (v * 2.0)[tris[:2]].astype(np.float32)) and not t1.any()
    assert np.allclose(n1[:, :, 2], 1.0)                                  # missing normals -> geometric normal
    p2, n2, _ = ml.load_serialized(path, 2)
    assert p2.shape == (3, 3, 3) and np.allclose(np.linalg.norm(n2, axis=2), 1.0, atol=1e-6)
    with pytest.raises(ValueError):
        ml.load_serialized(path, 3)
    open(str(tmp_path / "x.serialized"), "wb").write(b"\x00\x00\x04\x00" + bytes(16))
    with pytest.raises(ValueError):
        ml.load_serialized(str(tmp_path / "x.serialized"))


@pytest.mark.parametrize("binary", [True, False])
def test_hair_ribbons(tmp_path, binary):
    rng = np.random.default_rng(4)
    strands = [np.cumsum(rng.normal(0, 0.1, (k, 3)) + np.array([0, 0.2, 0]), axis=0) + rng.uniform(-1, 1, 3) for k in (5, 2, 9)]
    strands.insert(1, np.zeros((1, 3)))                                   # a one-vertex strand is dropped with a warning
    path = str(tmp_path / ("hair.bin" if binary else "hair.txt"))
    ml.save_hair(path, strands, binary=binary)
    radius = 0.01
    p, n, t = ml.load_hair(path, radius)
    assert p.shape[0] == 2 * ((5 - 1) + (2 - 1) + (9 - 1))                # two triangles per segment
    # every ribbon vertex lies within `radius` of its strand vertex, the width tapers to 0 at the tip, normals are geometric
    k = 0
    for s in (strands[0], strands[2], strands[3]):
        s = s.astype(np.float32 if binary else np.float64)
        for v in range(1, len(s)):
            quad = np.concatenate([p[k], p[k + 1]])
            d = np.min(np.linalg.norm(quad[:, None, :] - s[None, v - 1:v + 1, :], axis=2), axis=1)
            assert np.all(d <= radius * 1.0001 + 1e-6)
            k += 2
        assert np.allclose(p[k - 1][1], s[-1], atol=1e-5) and np.allclose(p[k - 1][2], s[-1], atol=1e-5)      # tip: zero width
    ln = np.linalg.norm(n, axis=2)                                        # the last triangle of a strand has zero width at the tip: no normal
    assert np.all((np.abs(ln - 1.0) < 1e-4) | (ln == 0.0)) and (ln == 0.0).all(axis=1).sum() <= 3
    assert np.array_equal(t[0], np.array([[0, 0], [1, 0], [0, 1]], dtype=np.float32)) or np.array_equal(t[0], np.array([[0, 0], [0, 1], [1, 0]], dtype=np.float32))
    # deterministic: the roll angle comes from a PCG seeded with the file name
    p2, _, _ = ml.load_hair(path, radius)
    assert np.array_equal(p, p2)


def test_mitsuba_scene_dispatch(tmp_path):
    """<shape type="ply" | "serialized" | "hair"> in a scene file reach the loaders (MitsubaLoader.cpp:434-515)."""
    v, faces, n, uv = _quad_mesh()
    ml.save_ply(str(tmp_path / "a.ply"), v, faces, normals=n, uvs=uv, fmt="binary_little_endian")
    ml.save_serialized(str(tmp_path / "b.serialized"), [dict(vertices=v, faces=np.array([[0, 1, 2]])), dict(vertices=v, faces=np.array([[0, 1, 2], [0, 2, 3]]))])
    ml.save_hair(str(tmp_path / "c.hair"), [np.array([[0, 0, 0], [0, 1, 0], [0, 2, 0.5]])])
    xml = """<scene version="0.6.0">
      <shape type="ply"><string name="filename" value="a.ply"/></shape>
      <shape type="serialized"><string name="filename" value="b.serialized"/><integer name="shapeIndex" value="1"/></shape>
      <shape type="hair"><string name="filename" value="c.hair"/><float name="radius" value="0.02"/></shape>
    </scene>"""
    path = str(tmp_path / "scene.xml")
    open(path, "w").write(xml)
    d = scene.load_mitsuba(path)
    counts = sorted(md[0].shape[0] for md in d.mesh_datas)
    assert counts == [2, 4, 4]                                            # serialized sub-mesh 1, ply (quad + 2 tris), hair (2 segments)
    blob = scene.build_blob(d, 8, rng="fallback", width=64, height=48)
    assert int(blob["triangles"].shape[0]) == 10
