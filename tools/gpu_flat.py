"""Experiment: how much of the Sponza traversal cost is the two-level (384-instance) structure?  Same geometry, one material,
NEE off: (A) the reference's TLAS + 384 BLASes, (B) all triangles in ONE BLAS.  Prints trace time, nodes/triangles per ray, CRC."""
import os, sys, zlib, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene

src = scene.load_blob(os.path.join(ROOT, "data", "_staged", "sponza.npz"))
f32 = np.float32
mt = np.asarray(src["material_types"]); mats = np.asarray(src["materials"]).copy()
m = int(np.where(mt == 1)[0][0])
mats[m, 0:3] = 0.7; mats[m, 3:4].view(np.int32)[0] = -1          # untextured grey diffuse

A = dict(src); A["materials"] = mats
A["mesh_material_ids"] = np.full_like(np.asarray(src["mesh_material_ids"]), m)

tri = np.asarray(src["triangles"])
P = np.stack([tri[:, 0:3], tri[:, 0:3] + tri[:, 3:6], tri[:, 0:3] + tri[:, 6:9]], axis=1).astype(f32)
t0 = time.perf_counter()
blas = scene.build_blas(P, 8)
print("flat build s", time.perf_counter() - t0, "nodes", blas.node_count, flush=True)
nd, idx = blas.export(2, 0)
lo, hi = P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0)
tlas = scene.build_tlas(np.array([np.concatenate([lo, hi])], dtype=f32), 8)
tn, _ = tlas.export(0, 0)
nodes = np.zeros((2 + blas.node_count) * 80, dtype=np.uint8)
nodes[:tn.size] = tn; nodes[160:160 + nd.size] = nd
B = dict(A)
B["triangles"] = np.ascontiguousarray(tri[idx]); B["bvh_nodes"] = nodes; B["tlas_node_count"] = int(tlas.node_count)
B["mesh_bvh_root_indices"] = np.array([np.uint32(2) | np.uint32(1 << 31)], dtype=np.uint32).view(np.int32)
B["mesh_material_ids"] = np.array([m], dtype=np.int32)
ident = np.array([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]], dtype=f32)
B["mesh_transforms"] = ident; B["mesh_transforms_inv"] = ident.copy(); B["mesh_transforms_prev"] = ident.copy()
B["mesh_tri_first"] = np.array([0], dtype=np.int32); B["mesh_tri_count"] = np.array([tri.shape[0]], dtype=np.int32)
B["instance_order"] = np.array([0], dtype=np.int32)
for k in ("light_triangle_indices", "light_triangle_cdf", "light_mesh_cdf", "light_mesh_transform_indices"):
    A[k] = np.asarray(src[k])[:0]; B[k] = A[k]
A["light_mesh_triangle_span"] = np.zeros((0, 2), dtype=np.int32); B["light_mesh_triangle_span"] = A["light_mesh_triangle_span"]
A["lights_total_weight"] = 0.0; B["lights_total_weight"] = 0.0

for name, blob in (("two-level", A), ("flat", B)):
    cfg = pt.default_config(num_bounces=4, enable_next_event_estimation=0)
    p = pt.Pathtracer(blob, config=cfg)
    p.reserve_wave(9)
    for _ in range(2): p.render_frame(8)
    p.sync(); p.set_timing(True)
    tot = {}
    for _ in range(4):
        p.render_frame(8); p.sync()
        for k, v in p.stage_ms().items(): tot[k] = tot.get(k, 0.0) + v / 4
    img = p.get_aov(0)
    tr = p.measure_traversal(1)
    rays = p.ray_stats()["trace"][:4]
    print(f"{name:10s} trace {tot['trace']:.2f} ms/frame  nodes/ray {tr['nodes'][0] / tr['rays'][0]:.2f} tris/ray {tr['triangles'][0] / tr['rays'][0]:.2f} crc {zlib.crc32(img.tobytes()):08x}", flush=True)
    p.close()
