"""BASELINE.json's other configs on one B200 (the bench line is configs[1]): device ms and Mrays/s, CUDA events on the ctx stream.
  c0 cornellbox 512x512, 1 spp, binary SAH BVH, 1 bounce          c2 Sponza 1080p, SVGF + TAA, 8 displayed frames of 1 spp
  c3 instancing 1080p, 8 spp (camera turned towards the grid)      c4 Sponza 3840x2160, SVGF + TAA, 16 displayed frames of 1 spp"""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpu_raytracer_b200 import pathtracer as pt, scene

def staged(name):
    return scene.load_blob(os.path.join(ROOT, "data", "_staged", name))

def timed(p, fn, reps):
    s = torch.cuda.ExternalStream(p.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p.sync(); p.ray_stats(reset=True)
    with torch.cuda.stream(s): e0.record()
    for _ in range(reps): fn()
    with torch.cuda.stream(s): e1.record()
    p.sync()
    st = p.ray_stats(reset=True)
    rays = float(st["trace"].sum() + st["shadow"].sum())
    ms = e0.elapsed_time(e1)
    return ms / reps, rays / (ms * 1e-3) / 1e6

def accumulate_case(tag, blob, bounces, passes, wave):
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=bounces)); p.reserve_wave(wave)
    for _ in range(3): p.render_frame(passes)
    ms, mr = timed(p, lambda: p.render_frame(passes), 10)
    print(f"{tag}: {ms:.3f} ms per {passes}-spp frame, {mr:.0f} Mrays/s", flush=True); p.close()

def svgf_case(tag, blob, frames):
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4, enable_svgf=1, enable_taa=1))
    def one():
        p.update(); p.render()
    for _ in range(4): one()
    ms, mr = timed(p, one, frames)
    print(f"{tag}: {ms:.3f} ms per displayed frame ({1000.0 / ms:.0f} fps), {ms * frames:.1f} ms for {frames} frames, {mr:.0f} Mrays/s", flush=True); p.close()

accumulate_case("c0 cornellbox 512x512 BVH2 1 bounce 1 spp", staged("cornellbox.npz"), 1, 1, 2)
sp = staged("sponza.npz")
svgf_case("c2 Sponza 1920x1080 SVGF+TAA", sp, 8)
inst = dict(staged("instancing.npz"))
pos = np.array(inst["camera"][:3], dtype=np.float64)
look = np.array([0.70710678, -0.15, -0.70710678]); look /= np.linalg.norm(look)
rot = scene.q_look_rotation(tuple(-look), (0.0, 1.0, 0.0))
inst["camera"] = scene.camera_block(tuple(pos), rot, math.radians(80.0), 1920, 1080)
inst["view_projection"] = scene.view_projection(tuple(pos), rot, math.radians(80.0), 1920, 1080)
accumulate_case("c3 instancing 1920x1080 4 bounces 8 spp (444 instances, all BSDFs)", inst, 4, 8, 9)
b4 = dict(sp); cam = np.array(sp["camera"], dtype=np.float32).copy(); cam[3:6] *= 2; cam[12] = math.atan(math.tan(float(cam[12])) / 2)
b4["camera"] = cam; b4["width"] = 3840; b4["height"] = 2160
svgf_case("c4 Sponza 3840x2160 SVGF+TAA", b4, 16)
