"""Profiling driver for `ncu --profile-from-start off`: Sponza 1080p, 4 bounces.
   range 1: ONE pass of the wavefront pipeline, pass by pass (wave 1) -- generate, trace x4, sort x4, shade x4, shadow x4, accumulate
   range 2: ONE SVGF + TAA frame (push-free, 1 GPU) -- the same plus reproject, variance, 6 a-trous, finalize, taa x2, clear
Everything before the ranges (upload, LUT bake, merged-BVH build, warm-up passes) is not profiled."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402
import torch  # noqa: E402  (locates libcudart)

rt = ctypes.CDLL("libcudart.so.12")        # already in the process (torch)
blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", "sponza.npz"))
merge = int(os.environ.get("PTB_PROF_MERGE", "1"))

wave = int(os.environ.get("PTB_PROF_WAVE", "1"))
p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4)); p.set_static_merge(merge)
if wave > 1:                       # the bench configuration: all 9 passes of a frame in one wave (PTB_TRACE_OVERLAP=0 keeps one stream)
    p.reserve_wave(wave); p.set_timing(True)     # timing mode = plain launches on one stream, no graph
    p.render_frame(8); p.sync()
    rt.cudaProfilerStart()
    p.render_frame(8); p.sync()
    rt.cudaProfilerStop()
else:
    for si in range(3):
        p.render_pass(si)
    p.sync()
    rt.cudaProfilerStart()
    p.render_pass(3); p.sync()
    rt.cudaProfilerStop()
p.close()

if os.environ.get("PTB_PROF_SVGF", "1") == "1":
    q = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4, enable_svgf=1, enable_taa=1)); q.set_static_merge(merge)
    for _ in range(3):
        q.update(); q.render()
    q.sync()
    rt.cudaProfilerStart()
    q.update(); q.render(); q.sync()
    rt.cudaProfilerStop()
    q.close()
print("done")
