"""SVGF + TAA parity hunt: strict-mode context vs the reference kernels, frame by frame, buffer by buffer (count, bounding box and a
few samples of the mismatching pixels).  usage: gpu_svgf_diag.py [sponza|atrium] [frames] [width height]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene
from oracle import ref

which = sys.argv[1] if len(sys.argv) > 1 else "sponza"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if which == "sponza":
    blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", "sponza.npz"))
    if len(sys.argv) > 4:
        blob = scene.retarget_blob(blob, int(sys.argv[3]), int(sys.argv[4]))
else:
    blob = scene.build_blob(scene.procedural_scene("atrium", seed=4, width=320, height=180, detail=0.5), 8, rng="fallback")
w, h = int(blob["width"]), int(blob["height"])
cfg = pt.default_config(num_bounces=4, enable_svgf=1, enable_spatial_variance=1, enable_taa=1, num_atrous_iterations=6)
r = ref.Reference(blob, config=cfg)
p = pt.Pathtracer(blob, config=cfg); p.set_static_merge(False)
names = ["history_direct", "history_indirect", "history_moment", "history_normal_and_depth", "history_length", "frame_buffer_moment", "taa_frame_curr", "taa_frame_prev"]

def report(tag, a, b):
    a = a[:h, :w]; b = b[:h, :w]
    ne = (a.view(np.uint32) != b.view(np.uint32))
    if ne.ndim == 3:
        ne = ne.any(-1)
    n = int(ne.sum())
    if n == 0:
        print(f"   {tag:26s} ok"); return
    ys, xs = np.nonzero(ne)
    print(f"   {tag:26s} {n:8d} differ  bbox x[{xs.min()},{xs.max()}] y[{ys.min()},{ys.max()}]")
    for k in range(min(3, n)):
        print(f"        ({xs[k]},{ys[k]}) ours {a[ys[k], xs[k]]} ref {b[ys[k], xs[k]]}")

mode = os.environ.get("DIAG_MODE", "sync")         # "sync": compare after every frame; "nosync": queue all frames, compare the last;
extra = None                                       # "third": nosync + a second ptb context rendering alongside (what the pytest case does)
if mode == "third":
    extra = pt.Pathtracer(blob, config=cfg)
for si in range(frames):
    if mode == "sync":
        p.render_pass(si); r.render_pass(si); p.sync(); r.sync()
    else:
        r.render_pass(si); p.render_pass(si)
        if extra is not None:
            extra.render_pass(si)
        if si != frames - 1:
            continue
        p.sync(); r.sync()
    print(f"frame {si}")
    report("display", p.get_display(), r.get_display())
    for k, nm in ((1, "acc_direct"), (2, "acc_indirect")):
        report(nm, p.get_aov(k, True), r.get_aov(k, True))
    for nm in names:
        report(nm, p.svgf_buffer(nm), r.svgf_buffer(nm))
st, sr = p.ray_stats(), r.ray_stats()
print("counters equal:", np.array_equal(st["trace"], sr["trace"]), np.array_equal(st["shadow"], sr["shadow"]))
