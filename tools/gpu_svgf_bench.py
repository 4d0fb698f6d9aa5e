"""BASELINE.json configs[4]: Sponza 3840x2160 (and 1920x1080), 1 spp per displayed frame, SVGF (variance + 6 a-trous) + TAA.
Per-stage device ms per frame, averaged."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene

src = scene.load_blob(os.path.join(ROOT, "data", "_staged", "sponza.npz"))
for scale in (1, 2):
    blob = dict(src)
    cam = np.array(src["camera"], dtype=np.float32).copy()
    cam[3:6] *= scale
    cam[12] = math.atan(math.tan(float(cam[12])) / scale)
    blob["camera"] = cam; blob["width"] = 1920 * scale; blob["height"] = 1080 * scale
    cfg = pt.default_config(num_bounces=4, enable_svgf=1, enable_taa=1, aov_mask=0x3F)
    p = pt.Pathtracer(blob, config=cfg)
    for _ in range(4):
        p.update(); p.render()
    p.sync(); p.set_timing(True)
    tot = {}; N = 8
    for _ in range(N):
        p.update(); p.render(); p.sync()
        for k, v in p.stage_ms().items(): tot[k] = tot.get(k, 0.0) + v / N
    st = p.ray_stats(reset=True)
    print(f"{1920 * scale}x{1080 * scale} svgf+taa: " + " ".join(f"{k} {v:.3f}" for k, v in tot.items()) + f" | total {sum(tot.values()):.3f} ms/frame", flush=True)
    p.close()
