"""The compiled C++ facade (gpu-raytracer_b200/host/ptb_pathtracer.{h,cpp}: the reference's Integrator / Pathtracer entry points --
cuda_init / update / render / resize_free / resize_init -- over the C ABI) renders the reference's Data/cornellbox (a procedural Cornell
box when the scene is not staged) and must match the ctypes path bit for bit, before and after a resize."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from gpu_raytracer_b200 import build, pathtracer as pt, scene

pytestmark = pytest.mark.gpu


def test_cpp_facade_matches_ctypes_path_bit_for_bit(tmp_path):
    build.build_facade()
    drv = os.path.join(ROOT, "tests", "cpp", "facade_render")
    staged = os.path.join(ROOT, "data", "_staged", "cornellbox_bvh8.npz")
    if os.path.exists(staged):
        blob = scene.load_blob(staged)
    else:
        blob = scene.build_blob(scene.procedural_scene("cornell", seed=1, width=256, height=256), 8, rng="fallback")
    blob["num_bounces"] = 3
    w, h = int(blob["width"]), int(blob["height"])
    small = scene.retarget_blob(blob, 320, 192)
    raw, cam2 = str(tmp_path / "scene.raw"), str(tmp_path / "cam2.raw")
    scene.dump_raw(blob, raw); scene.dump_raw(small, cam2, camera_only=True)
    out1, out2 = str(tmp_path / "a.f32"), str(tmp_path / "b.f32")
    passes = 3
    r = subprocess.run([drv, raw, str(passes), out1, "320", "192", cam2, out2], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FACADE-OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
    cfg = pt.default_config(num_bounces=3)
    p = pt.Pathtracer(blob, config=cfg); p.render_frames(passes)
    want = p.get_aov(pt.AOV_RADIANCE)
    got = np.fromfile(out1, dtype=np.float32).reshape(want.shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # resize through the ctypes path too (ptb_resize), and against a context created at that size
    p.resize(small); p.render_frames(passes)
    resized = p.get_aov(pt.AOV_RADIANCE); p.close()
    q = pt.Pathtracer(small, config=cfg); q.render_frames(passes); fresh = q.get_aov(pt.AOV_RADIANCE); q.close()
    got2 = np.fromfile(out2, dtype=np.float32).reshape(fresh.shape)
    assert np.array_equal(resized.view(np.uint32), fresh.view(np.uint32))
    assert np.array_equal(got2.view(np.uint32), fresh.view(np.uint32))
    assert float(np.abs(fresh[..., :3]).sum()) > 0.0
