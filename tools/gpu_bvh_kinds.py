"""Sponza 1080p bench configuration through the three traversal kernels in the reference's two-level mode (no merged tree): CWBVH
(k_trace8), BVH4 (k_trace4), SAH-collapsed BVH2 (k_trace2) -- device ms per 9-pass frame, per-stage ms, node visits / triangle tests per
ray.  Question behind it: on a machine where the scene is cache resident and the trace kernel is issue bound, does the wide compressed
tree still win?  Blobs: data/_staged/sponza{,_bvh4,_bvh2}.npz (tools/stage_data.py stage_scene with bvh_kind 8 / 4 / 2)."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpu_raytracer_b200 import pathtracer as pt, scene
for name in ("sponza.npz", "sponza_bvh4.npz", "sponza_bvh2.npz"):
    path = os.path.join(ROOT, "data", "_staged", name)
    if not os.path.exists(path): print(name, "missing"); continue
    blob = scene.load_blob(path)
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4)); p.set_static_merge(0); p.reserve_wave(9)
    for _ in range(3): p.render_frame(8)
    p.sync()
    s = torch.cuda.ExternalStream(p.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s): e0.record()
    for _ in range(6): p.render_frame(8)
    with torch.cuda.stream(s): e1.record()
    p.sync()
    ms = e0.elapsed_time(e1) / 6
    crc = zlib.crc32(p.get_aov(0).tobytes())
    p.set_timing(True); p.render_frame(8); p.sync(); st = p.stage_ms(); p.set_timing(False)
    try:
        tr = p.measure_traversal(1)
        trs = "closest nodes/ray %.2f tris/ray %.2f | shadow nodes/ray %.2f tris/ray %.2f" % (tr["nodes"][0] / tr["rays"][0], tr["triangles"][0] / tr["rays"][0], tr["nodes"][1] / max(tr["rays"][1], 1), tr["triangles"][1] / max(tr["rays"][1], 1))
    except Exception as e:
        trs = "no traversal stats: %s" % e
    print("%-16s kind %d two-level: %.3f ms/frame crc %08x trace %.2f shadow %.2f sort %.2f shade %.2f | %s" % (name, int(blob["bvh_kind"]), ms, crc, st["trace"], st["shadow_trace"], st["sort"], st["shade"], trs), flush=True)
    del s, e0, e1
    p.close()
