"""TLAS refit diagnosis: after ptb_refit_instances, read the TLAS back and validate it on the host against the moved instance boxes
(scene.check_tlas8), then compare the frame with a context that was given a host-rebuilt TLAS of the moved scene."""
import copy, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402


def world_boxes(desc, order):
    xf, _ = scene.instance_transforms(desc, order)
    out = []
    for j, i in enumerate(order):
        inst = desc.instances[i]; p = desc.mesh_datas[inst.mesh_data][0].reshape(-1, 3).astype(np.float64)
        T = xf[j].reshape(3, 4).astype(np.float64); lo, hi = p.min(0), p.max(0); c, e = 0.5 * (lo + hi), 0.5 * (hi - lo)
        nc = T[:, :3] @ c + T[:, 3]; ne = np.abs(T[:, :3]) @ e
        out.append(np.concatenate([nc - ne, nc + ne]))
    return np.array(out)


d = scene.procedural_scene("atrium", seed=5, width=320, height=192, detail=0.5)
blob = scene.build_blob(d, 8, rng="fallback")
moved = copy.deepcopy(d)
for i, inst in enumerate(moved.instances):
    if inst.name == "column" and i % 2:
        inst.position = inst.position + np.array([0.6, 0.0, -0.4]); inst.rotation = scene.q_axis_angle((0, 1, 0), 0.2 * i)
    elif inst.name == "curtain":
        inst.position = inst.position + np.array([0.0, 0.5, 0.0])
    elif inst.name == "floor":
        inst.position = inst.position + np.array([0.0, -0.15, 0.0])
    elif inst.name == "lamp" and i % 2:
        inst.position = inst.position + np.array([1.0, -0.5, 0.5])
order = np.asarray(blob["instance_order"])
n_tlas = int(blob["tlas_node_count"])
cfg = pt.default_config(num_bounces=3)
for merge in (0, 1):
    p = pt.Pathtracer(blob, config=cfg); p.set_static_merge(merge)
    p.render_frames(2)
    before = p.tlas_nodes(n_tlas)
    print("merge", merge, "before refit:", scene.check_tlas8(before, world_boxes(d, order)))
    xf, xi = scene.instance_transforms(moved, order)
    p.refit_instances(xf, xi)
    p.sync()
    after = p.tlas_nodes(n_tlas)
    print("  nodes changed:", int(np.any(before != after, axis=1).sum()), "of", n_tlas)
    print("  after refit vs moved boxes:", scene.check_tlas8(after, world_boxes(moved, order)))
    print("  after refit vs OLD boxes  :", scene.check_tlas8(after, world_boxes(d, order)))
    p.invalidated_gpu_config = True
    p.render_frames(3)
    got = p.get_aov(0); hits = p.primary_hits()
    p.close()
    rebuilt = scene.build_blob(moved, 8, rng="fallback")
    q = pt.Pathtracer(rebuilt, config=cfg); q.set_static_merge(merge); q.render_frames(3)
    want = q.get_aov(0); whits = q.primary_hits(); q.close()
    differ = np.any(got.view(np.uint32) != want.view(np.uint32), axis=-1)
    print("  pixels differing:", differ.mean(), "rel-L2", np.linalg.norm(got - want) / np.linalg.norm(want))
    ys, xs = np.nonzero(differ)
    if len(ys):
        print("  rows with differences:", ys.min(), ys.max(), "cols", xs.min(), xs.max())
    t_got = hits[..., 2].view(np.float32); t_want = whits[..., 2].view(np.float32)
    print("  primary t differs on", float((t_got != t_want).mean()), "of pixels; triangle ids differ on", float((hits[..., 1] != whits[..., 1]).mean()))
    bad = t_got != t_want
    worder = np.asarray(rebuilt["instance_order"])
    names_w = {}; names_g = {}
    for m in whits[..., 0][bad].astype(np.int32):
        n = moved.instances[worder[m]].name if 0 <= m < len(worder) else "sky"; names_w[n] = names_w.get(n, 0) + 1
    for m in hits[..., 0][bad].astype(np.int32):
        n = moved.instances[order[m]].name if 0 <= m < len(order) else "sky"; names_g[n] = names_g.get(n, 0) + 1
    print("  differing pixels by instance (rebuilt ctx):", names_w)
    print("  differing pixels by instance (refit ctx)  :", names_g)
    sel = np.argwhere(bad)[:: max(1, bad.sum() // 6)][:6]
    for y, x in sel:
        print("   pixel", y, x, "refit: mesh", int(hits[y, x, 0]), "tri", int(hits[y, x, 1]), "t", float(t_got[y, x]), "| rebuilt: mesh", int(whits[y, x, 0]), "tri", int(whits[y, x, 1]), "t", float(t_want[y, x]))
