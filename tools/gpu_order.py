"""A/B of ray ordering (ptb_set_ray_ordering 0 / 8 / 64) on Sponza 1080p: per-stage device time per 9-pass frame (wave 9) and
pass by pass, plus a CRC of the accumulated image (must not change)."""
import os, sys, json, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene

blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", sys.argv[1] if len(sys.argv) > 1 else "sponza.npz"))
for wave in (9, 1):
    for bins in (0, 8, 64):
        p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4))
        p.reserve_wave(wave); p.set_ray_ordering(bins)
        for _ in range(2): p.render_frame(8)
        p.sync(); p.set_timing(True)
        tot = {}
        N = 4
        for _ in range(N):
            p.render_frame(8); p.sync()
            for k, v in p.stage_ms().items(): tot[k] = tot.get(k, 0.0) + v / N
        crc = zlib.crc32(p.get_aov(0).tobytes())
        total = sum(tot.values())
        print(f"wave {wave} bins {bins:2d}: " + " ".join(f"{k} {v:.2f}" for k, v in tot.items()) + f" | total {total:.2f} ms/frame crc {crc:08x}", flush=True)
        p.close()
