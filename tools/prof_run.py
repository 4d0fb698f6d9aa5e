"""Short profiling driver (for ncu): Sponza 1080p, N passes of the 4-bounce wavefront pipeline, nothing else."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
staged = os.path.join(ROOT, "data", "_staged", "sponza.npz")
blob = scene.load_blob(staged) if os.path.exists(staged) else scene.build_blob(scene.procedural_scene("atrium", seed=7, width=1920, height=1080, detail=4.0), 8, 1920, 1080)
blob["num_bounces"] = 4
p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4))
for si in range(passes):
    p.render_pass(si)
p.sync()
print("rays", int(p.ray_stats()["trace"].sum() + p.ray_stats()["shadow"].sum()))
p.close()
