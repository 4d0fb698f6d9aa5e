"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI (libptb.so via the ctypes façade).

Checkers, in order of strength:
  1. the reference's own kernels (oracle/_ref cubin + oracle/ref_harness.cpp) on the same blob: BIT-EXACT hits, AOVs, counters;
  2. committed golden fixtures produced by (1) (tests/golden/*.npz): bit-exact again;
  3. the CPU restatement (oracle/pt_oracle.c): tolerance based (IEEE libm vs GPU fast-math).
"""
import os

import numpy as np
import pytest

from conftest import ROOT
from gpu_raytracer_b200 import pathtracer as pt, scene

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_available():
    from oracle import ref
    return ref.available()


def rel_l2(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30)))


def small_blobs():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    return make_golden, {n: make_golden.case_blob(c) for n, c in make_golden.CASES.items()}


@pytest.fixture(scope="module")
def cases():
    return small_blobs()


def valid_hits_equal(a, b):
    """triangle id and t always; mesh id and packed uv only where something was hit (the reference leaves them uninitialised on a miss)"""
    valid = b[..., 1] != 0xFFFFFFFF
    same = (a[..., 1] == b[..., 1]) & (a[..., 2] == b[..., 2])
    same &= ~valid | ((a[..., 0] == b[..., 0]) & (a[..., 3] == b[..., 3]))
    return same


@pytest.mark.parametrize("name", ["soup_bvh8", "cornell_bvh8", "cornell_bvh2", "atrium_bvh8"])
def test_golden_fixtures_bit_exact(cases, name):
    mg, blobs = cases
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated yet")
    g = np.load(path)
    blob = blobs[name]; c = mg.CASES[name]; w = c["size"][0]
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=1))
    p.render_pass(1); p.sync()
    hits = p.primary_hits()[:, :w]
    assert valid_hits_equal(hits, g["hits"]).all()
    p.close()
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=c["bounces"], aov_mask=0x3F))
    p.render_frames(int(g["passes"]))
    for k, key in ((0, "radiance"), (3, "albedo"), (4, "normal"), (5, "position")):
        img = p.get_aov(k)[:, :w, :3]
        assert np.array_equal(img.view(np.uint32), g[key].view(np.uint32)), key
    st = p.ray_stats()
    assert np.array_equal(st["trace"][:8].astype(np.int64), g["trace"]) and np.array_equal(st["shadow"][:8].astype(np.int64), g["shadow"])
    p.close()


@pytest.mark.parametrize("name", ["soup_bvh8", "cornell_bvh8", "cornell_bvh2", "atrium_bvh8"])
def test_reference_kernels_bit_exact(cases, name):
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    from oracle import ref
    mg, blobs = cases
    blob = blobs[name]; c = mg.CASES[name]; w = c["size"][0]
    for nb in range(1, c["bounces"] + 1):
        cfg = pt.default_config(num_bounces=nb, aov_mask=0x3F)
        p = pt.Pathtracer(blob, config=cfg); r = ref.Reference(blob, config=cfg)
        p.render_frames(2); r.render_frames(2)
        if nb == 1:
            assert valid_hits_equal(p.primary_hits()[:, :w], r.primary_hits()[:, :w]).all()
        for k in range(6):
            a, b = p.get_aov(k)[:, :w], r.get_aov(k)[:, :w]
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nb, pt.AOV_NAMES[k])
        assert np.array_equal(p.get_display()[:, :w].view(np.uint32), r.get_display()[:, :w].view(np.uint32))
        sp, sr = p.ray_stats(), r.ray_stats()
        assert np.array_equal(sp["trace"], sr["trace"]) and np.array_equal(sp["shadow"], sr["shadow"]) and np.array_equal(sp["shaded"], sr["shaded"])
        p.close(); r.close()


@pytest.mark.parametrize("flags", [dict(enable_next_event_estimation=0), dict(enable_multiple_importance_sampling=0), dict(enable_russian_roulette=0),
                                   dict(reconstruction_filter=0), dict(reconstruction_filter=1), dict(enable_mipmapping=0)])
def test_config_switches_bit_exact(cases, flags):
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    from oracle import ref
    mg, blobs = cases
    blob = blobs["soup_bvh8"]; w = mg.CASES["soup_bvh8"]["size"][0]
    cfg = pt.default_config(num_bounces=3, **flags)
    p = pt.Pathtracer(blob, config=cfg); r = ref.Reference(blob, config=cfg)
    p.render_frames(2); r.render_frames(2)
    assert np.array_equal(p.get_aov(0)[:, :w].view(np.uint32), r.get_aov(0)[:, :w].view(np.uint32))
    p.close(); r.close()


def test_all_material_types_against_reference():
    """plastic / dielectric / conductor + Kulla-Conty LUTs baked by our kernels vs the reference's."""
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    from oracle import ref
    blob = scene.build_blob(scene.procedural_scene("soup", seed=9, width=160, height=96, detail=0.25, all_materials=True), 8, rng="fallback")
    cfg = pt.default_config(num_bounces=4, aov_mask=0x3F)
    p = pt.Pathtracer(blob, config=cfg); r = ref.Reference(blob, config=cfg)
    p.render_frames(4); r.render_frames(4)
    a, b = p.get_aov(0)[:, :160, :3], r.get_aov(0)[:, :160, :3]
    sp, sr = p.ray_stats(), r.ray_stats()
    assert sp["shaded"][1] > 0 and sp["shaded"][2] > 0 and sp["shaded"][3] > 0
    assert rel_l2(a, b) <= 1e-4, rel_l2(a, b)           # north-star tolerance on the HDR framebuffer
    assert abs(int(sp["trace"].sum()) - int(sr["trace"].sum())) <= 1e-3 * int(sr["trace"].sum())
    p.close(); r.close()


def test_svgf_taa_against_reference():
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    from oracle import ref
    blob = scene.build_blob(scene.procedural_scene("atrium", seed=4, width=320, height=180, detail=0.5), 8, rng="fallback")
    cfg = pt.default_config(num_bounces=3, enable_svgf=1, enable_spatial_variance=1, enable_taa=1, num_atrous_iterations=6)
    p = pt.Pathtracer(blob, config=cfg); r = ref.Reference(blob, config=cfg)
    for si in range(6):
        p.render_pass(si); r.render_pass(si)
    p.sync(); r.sync()
    a, b = p.get_display()[:, :320, :3], r.get_display()[:, :320, :3]
    assert np.isfinite(a).all()
    assert rel_l2(a, b) <= 1e-4, rel_l2(a, b)
    p.close(); r.close()


@pytest.mark.parametrize("name", ["soup_bvh8", "cornell_bvh8"])
def test_cpu_oracle_agrees_within_tolerance(cases, name):
    from oracle.oracle import Oracle
    mg, blobs = cases
    blob = blobs[name]; c = mg.CASES[name]; w = c["size"][0]
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=c["bounces"]))
    p.render_frames(4)
    img = p.get_aov(0)[:, :w, :3]
    o = Oracle(blob, num_bounces=c["bounces"])
    acc = o.render(4)["radiance"][:, :w, :3]
    assert rel_l2(img, acc) < 2e-2          # IEEE libm vs fast-math: a few paths flip at triangle edges
    p1 = pt.Pathtracer(blob, config=pt.default_config(num_bounces=1))
    p1.render_pass(1); p1.sync()
    gh = p1.primary_hits()[:, :w]; oh = o.primary_hits(1)[:, :w]
    assert (gh[..., 1] != oh[..., 1]).mean() < 2e-3
    p.close(); p1.close()
