"""Scene front-end: Mitsuba/OBJ subset, triangle fix-ups, camera block, light tables, textures, blob round trip."""
import math
import os

import numpy as np

from gpu_raytracer_b200 import scene

XML = """<?xml version="1.0"?>
<scene version="0.5.0">
  <integrator type="path"><integer name="maxDepth" value="5"/></integrator>
  <sensor type="perspective">
    <float name="fov" value="45"/>
    <transform name="toWorld"><translate x="0" y="1" z="4"/></transform>
    <film type="hdrfilm"><integer name="width" value="48"/><integer name="height" value="32"/></film>
  </sensor>
  <bsdf type="twosided" id="grey"><bsdf type="diffuse"><rgb name="reflectance" value="0.5, 0.4, 0.3"/></bsdf></bsdf>
  <bsdf type="roughconductor" id="metal"><rgb name="eta" value="1.45, 0.43, 0.21"/><rgb name="k" value="1.95, 2.46, 3.27"/><float name="alpha" value="0.2"/></bsdf>
  <bsdf type="roughdielectric" id="glass"><string name="intIOR" value="water"/><float name="alpha" value="0.3"/></bsdf>
  <shape type="obj"><string name="filename" value="quad.obj"/><ref id="grey"/></shape>
  <shape type="obj"><string name="filename" value="quad.obj"/><ref id="metal"/><transform><scale value="2"/><translate x="3" y="0" z="0"/></transform></shape>
  <shape type="cube"><transform name="toWorld"><scale value="0.5"/><translate x="-2" y="0.5" z="0"/></transform><ref id="glass"/></shape>
  <shape type="rectangle"><transform name="toWorld"><translate x="0" y="3" z="0"/></transform>
    <emitter type="area"><rgb name="radiance" value="10, 8, 6"/></emitter></shape>
</scene>
"""
OBJ = """# quad
v -1 0 -1
v 1 0 -1
v 1 0 1
v -1 0 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 1 0
f 1/1/1 2/2/1 3/3/1 4/4/1
"""


def load(tmp_path):
    (tmp_path / "scene.xml").write_text(XML); (tmp_path / "quad.obj").write_text(OBJ)
    return scene.load_mitsuba(str(tmp_path / "scene.xml"))


def test_mitsuba_subset(tmp_path):
    d = load(tmp_path)
    assert (d.width, d.height, d.num_bounces) == (48, 32, 5)
    assert abs(d.cam_fov - math.radians(45)) < 1e-6 and np.allclose(d.cam_position, [0, 1, 4])
    kinds = [m.kind for m in d.materials]
    assert kinds == [scene.MAT_DIFFUSE, scene.MAT_DIFFUSE, scene.MAT_CONDUCTOR, scene.MAT_DIELECTRIC, scene.MAT_LIGHT]
    assert abs(d.materials[3].ior - 1.333) < 1e-6 and d.materials[2].roughness == 0.2
    assert len(d.instances) == 4 and len(d.mesh_datas) == 3          # the OBJ is shared by two instances
    assert d.instances[0].identity() and not d.instances[1].identity()
    assert abs(d.instances[1].scale - 2.0) < 1e-5 and np.allclose(d.instances[1].position, [3, 0, 0])
    p, n, t = d.mesh_datas[0]
    assert p.shape == (2, 3, 3)                                       # fan triangulation of the quad
    assert np.allclose(t[0, :, 1], [1, 0, 1])                          # v flipped, then winding reversed (normals faced backwards)


def test_winding_fix_and_zero_normals():
    p = [[[0, 0, 0], [1, 0, 0], [0, 1, 0]]]
    pn, nn, tn = scene.finish_triangles(p, [[[0, 0, -1]] * 3], [[[0, 0], [1, 0], [0, 1]]])
    assert np.allclose(pn[0, 1], [0, 1, 0]) and np.allclose(tn[0, 1], [0, 1])     # reversed: shading normals faced backwards
    pn, nn, tn = scene.finish_triangles(p, [[[0, 0, 0]] * 3], [[[0, 0], [1, 0], [0, 1]]])
    assert np.allclose(nn[0], [[0, 0, 1]] * 3)                                       # replaced by the geometric normal


def test_blob_tables(tmp_path):
    d = load(tmp_path)
    blob = scene.build_blob(d, 8, rng="fallback")
    cam = blob["camera"]
    # Mitsuba sensors look along +Z of their toWorld frame; the renderer's camera looks along -Z, so an un-rotated
    # toWorld decomposes to a half turn about Y (MitsubaLoader.cpp:605 passes forward = (0,0,-1) to Matrix4::decompose)
    assert np.allclose(cam[0:3], [0, 1, 4]) and np.allclose(cam[6:9], [-1, 0, 0], atol=1e-6) and np.allclose(cam[9:12], [0, 1, 0], atol=1e-6)
    dist = 24.0 / math.tan(math.radians(22.5))
    assert np.allclose(cam[3:6], [24, -16, dist], rtol=1e-5, atol=1e-4)
    assert abs(cam[12] - math.atan(2 * math.tan(math.radians(22.5)) / 48)) < 1e-7
    assert blob["light_mesh_cdf"].tolist() == [1.0] and blob["light_triangle_cdf"][-1] == 1.0
    assert abs(blob["lights_total_weight"] - (0.299 * 10 + 0.587 * 8 + 0.114 * 6) * 4.0) < 1e-3      # luminance x area (2x2 rectangle)
    li = blob["light_mesh_transform_indices"][0]
    assert blob["material_types"][blob["mesh_material_ids"][li]] == scene.MAT_LIGHT
    assert blob["triangles"].shape[1] == 24 and blob["bvh_nodes"].size % 80 == 0
    mats = blob["materials"]
    assert np.allclose(mats[1, 0:3], [0.5, 0.4, 0.3]) and mats[1, 3:4].view(np.int32)[0] == -1
    assert np.allclose(mats[2, 4:7], [1.95, 2.46, 3.27]) and abs(mats[3, 1] - 1.333) < 1e-6


def test_blob_roundtrip(tmp_path):
    d = scene.procedural_scene("soup", seed=1, width=32, height=32, detail=0.1)
    d.textures.append(scene.make_texture((np.random.default_rng(0).integers(0, 255, (16, 16, 4))).astype(np.uint8)))
    blob = scene.build_blob(d, 8, rng="fallback")
    path = str(tmp_path / "b.npz")
    scene.save_blob(blob, path)
    back = scene.load_blob(path)
    for k in ("triangles", "bvh_nodes", "camera", "pmj", "blue_noise", "mesh_transforms"):
        assert np.array_equal(blob[k], back[k])
    assert back["textures"][0]["format"] == "bc1" and len(back["textures"][0]["levels"]) == 3
    assert back["width"] == 32 and back["rng_source"] == "fallback"


def test_texture_pipeline():
    img = np.zeros((8, 8, 4), dtype=np.uint8); img[..., 0] = 255; img[..., 3] = 255
    t = scene.make_texture(img)
    assert t["format"] == "bc1" and len(t["levels"]) == 2 and t["levels"][0].size == 4 * 8     # 2x2 blocks of 8 bytes, then 1 block
    assert abs(t["lod_bias"] - 0.5 * math.log2(4.0)) < 1e-9                                  # reference quirk: log2 of the BLOCK grid
    blk = t["levels"][0][:8].view(np.uint16)
    assert blk[0] == 0xF800                                                                   # pure red endpoint in RGB565
    t2 = scene.make_texture(np.zeros((6, 10, 4), dtype=np.uint8))
    assert t2["format"] == "rgba8" and t2["levels"][0].size == 6 * 10 * 4                     # non power of two stays RGBA8
    assert abs(float(scene.gamma_to_linear(0.5)) - 0.21404) < 1e-4


def test_fallback_rng_tables_shape_and_range():
    pmj, blue = scene.fallback_rng_tables()
    assert pmj.size == 64 * 4096 * 2 and blue.size == 16 * 128 * 128 * 2
    assert pmj.min() >= 0.0 and pmj.max() < 1.0
    s = pmj.reshape(64, 4096, 2)[3]
    cells = (np.floor(s[:, 0] * 64).astype(int) * 64 + np.floor(s[:, 1] * 64).astype(int))
    assert len(set(cells.tolist())) == 4096                                                   # one point per stratum
