import copy, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402

d = scene.procedural_scene("atrium", seed=5, width=320, height=192, detail=0.5)
blob = scene.build_blob(d, 8, rng="fallback")
order = np.asarray(blob["instance_order"])
cfg = pt.default_config(num_bounces=3)


def frame(p):
    p.invalidated_gpu_config = True
    p.render_frames(3)
    return p.get_aov(0)


def apply(moved, what):
    for i, inst in enumerate(moved.instances):
        if inst.name == "column" and i % 2 and "col_move" in what:
            inst.position = inst.position + np.array([0.6, 0.0, -0.4])
        if inst.name == "column" and i % 2 and "col_rot" in what:
            inst.rotation = scene.q_axis_angle((0, 1, 0), 0.2 * i)
        if inst.name == "curtain" and "curtain" in what:
            inst.position = inst.position + np.array([0.0, 0.5, 0.0])
        if inst.name == "floor" and "floor" in what:
            inst.position = inst.position + np.array([0.0, -0.15, 0.0])
        if inst.name == "lamp" and i % 2 and "lamp" in what:
            inst.position = inst.position + np.array([1.0, -0.5, 0.5])


for what in (["col_move"], ["col_rot"], ["curtain"], ["floor"], ["lamp"], ["col_move", "col_rot"], ["col_move", "col_rot", "curtain"], ["col_move", "col_rot", "curtain", "floor"],
             ["col_move", "col_rot", "curtain", "floor", "lamp"]):
    moved = copy.deepcopy(d); apply(moved, what)
    rb = scene.build_blob(moved, 8, rng="fallback")
    q = pt.Pathtracer(rb, config=cfg); q.set_static_merge(0); want = frame(q); q.close()
    p = pt.Pathtracer(blob, config=cfg); p.set_static_merge(0); p.render_frames(2)
    xf, xi = scene.instance_transforms(moved, order)
    p.refit_instances(xf, xi)
    got = frame(p); p.close()
    differ = np.any(got.view(np.uint32) != want.view(np.uint32), axis=-1).mean()
    print(what, "pixels differing", round(float(differ), 5), "rel-L2", float(np.linalg.norm(got - want) / np.linalg.norm(want)), "tlas nodes", int(rb["tlas_node_count"]), int(blob["tlas_node_count"]), flush=True)
