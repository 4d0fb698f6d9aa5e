"""A/B of the static merge (and ray ordering) on a staged scene at its own size: per-stage device ms per 9-pass frame, CRC."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene
name = sys.argv[1] if len(sys.argv) > 1 else "sponza.npz"
blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", name))
for merge, bins in ((False, 0), (True, 0), (True, 8)):
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4))
    p.reserve_wave(9); p.set_static_merge(merge); p.set_ray_ordering(bins)
    for _ in range(2): p.render_frame(8)
    p.sync(); p.set_timing(True)
    tot = {}
    for _ in range(4):
        p.render_frame(8); p.sync()
        for k, v in p.stage_ms().items(): tot[k] = tot.get(k, 0.0) + v / 4
    crc = zlib.crc32(p.get_aov(0).tobytes())
    tr = p.measure_traversal(1)
    print(f"{name} merge {int(merge)} bins {bins}: " + " ".join(f"{k} {v:.2f}" for k, v in tot.items()) + f" | total {sum(tot.values()):.2f} ms/frame  nodes/ray {tr['nodes'][0] / tr['rays'][0]:.2f}/{tr['nodes'][1] / max(tr['rays'][1],1):.2f} tris/ray {tr['triangles'][0] / tr['rays'][0]:.2f} crc {crc:08x}", flush=True)
    p.close()

# how different is the merged image from the un-merged one?
import numpy as np
imgs = []
for merge in (False, True):
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4))
    p.reserve_wave(9); p.set_static_merge(merge)
    p.render_frame(8); p.sync()
    imgs.append(p.get_aov(0)[..., :3].astype(np.float64)); p.close()
d = imgs[0] - imgs[1]
npx = int((np.abs(d).sum(-1) > 0).sum())
print(f"merge on vs off after 9 passes: {npx} of {d.shape[0] * d.shape[1]} pixels differ, rel-L2 {np.sqrt((d ** 2).sum() / (imgs[0] ** 2).sum()):.3e}, max abs {np.abs(d).max():.3e}", flush=True)
