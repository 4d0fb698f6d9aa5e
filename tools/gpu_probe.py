"""First-contact GPU script: renders a few scenes with the product path and the reference kernels, prints parity
numbers and timings, and drops PNGs + npz into gpurun_out/ for inspection."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402
import imgio  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def rel_l2(a, b, width):
    a = a[:, :width, :3].astype(np.float64); b = b[:, :width, :3].astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30)))


def compare(name, blob, passes=2, with_ref=True, with_oracle=False, cfg=None):
    print(f"=== {name}: {blob['width']}x{blob['height']} tris={blob['triangles'].shape[0]} bvh{blob['bvh_kind']} bounces={blob['num_bounces']}", flush=True)
    cfg = cfg or pt.default_config(num_bounces=int(blob["num_bounces"]))
    t0 = time.time()
    p = pt.Pathtracer(blob, config=cfg)
    print(f"  ptb create+upload {time.time() - t0:.2f}s", flush=True)
    p.render_pass(0); p.sync()
    hits = p.primary_hits()
    cnt = p.last_pass_counters()
    print("  ptb pass0 trace", cnt["trace"][:5], "diffuse", cnt["diffuse"][:5], "shadow", cnt["shadow"][:5], flush=True)
    p.set_timing(True)
    t0 = time.time()
    for si in range(1, passes + 1):
        p.render_pass(si)
    p.sync()
    dt = time.time() - t0
    print(f"  ptb {passes} passes {dt * 1e3:.1f} ms wall; last-pass stages {p.stage_ms()}", flush=True)
    img = p.get_aov(0, True)
    imgio.save_hdr_png(os.path.join(OUT, f"{name}_ptb.png"), img, blob["width"])
    st = p.ray_stats()
    print("  ptb rays", int(st["trace"].sum() + st["shadow"].sum()), "nan", int(np.isnan(img).sum()), "mean", float(img[:, :blob['width'], :3].mean()), flush=True)
    if with_ref:
        from oracle import ref
        t0 = time.time()
        r = ref.Reference(blob, config=cfg)
        print(f"  ref create+upload {time.time() - t0:.2f}s geometry {r.launch_geometry()}", flush=True)
        r.render_pass(0); r.sync()
        rh = r.primary_hits()
        covered = rh[..., 1] != 0xFFFFFFFF
        covered[:, blob["width"]:] = False
        same = (rh == hits).all(-1)
        print(f"  primary hits: covered {int(covered.sum())} mismatching {int((covered & ~same).sum())} "
              f"(tri id differs {int((covered & (rh[..., 1] != hits[..., 1])).sum())}, t bits differ {int((covered & (rh[..., 2] != hits[..., 2])).sum())})", flush=True)
        r.set_timing(True)
        t0 = time.time()
        for si in range(1, passes + 1):
            r.render_pass(si)
        r.sync()
        dt = time.time() - t0
        print(f"  ref {passes} passes {dt * 1e3:.1f} ms wall; last-pass stages {r.stage_ms()}", flush=True)
        rimg = r.get_aov(0, True)
        imgio.save_hdr_png(os.path.join(OUT, f"{name}_ref.png"), rimg, blob["width"])
        rs = r.ray_stats()
        print("  ref rays", int(rs["trace"].sum() + rs["shadow"].sum()), "trace/bounce", rs["trace"][:5], "shadow", rs["shadow"][:5], flush=True)
        print("  ptb      trace/bounce", st["trace"][:5], "shadow", st["shadow"][:5], flush=True)
        print(f"  rel-L2 ptb vs ref (accumulated radiance, {passes} passes): {rel_l2(img, rimg, blob['width']):.3e}", flush=True)
        d = np.abs(img[:, :blob['width'], :3] - rimg[:, :blob['width'], :3]).max(-1)
        print(f"  pixels differing >1e-6: {int((d > 1e-6).sum())} of {d.size}; max abs {float(d.max()):.3e}", flush=True)
        r.close()
    if with_oracle:
        from oracle.oracle import Oracle
        o = Oracle(blob, num_bounces=cfg.num_bounces)
        acc = o.render(passes)
        print(f"  rel-L2 ptb vs CPU oracle: {rel_l2(img, acc['radiance'], blob['width']):.3e}", flush=True)
        oh = o.primary_hits(0)
        print(f"  primary tri ids ptb vs CPU oracle differ: {int((oh[:, :blob['width'], 1] != hits[:, :blob['width'], 1]).sum())}", flush=True)
    p.close()


def main():
    which = sys.argv[1:] or ["soup", "cornell8", "cornell2", "atrium", "sponza"]
    staged = os.path.join(ROOT, "data", "_staged")
    if "soup" in which:
        compare("soup", scene.build_blob(scene.procedural_scene("soup", seed=3, width=256, height=256), 8), passes=4, with_oracle=True)
    if "cornell8" in which and os.path.exists(os.path.join(staged, "cornellbox_bvh8.npz")):
        compare("cornell8", scene.load_blob(os.path.join(staged, "cornellbox_bvh8.npz")), passes=4, with_oracle=True)
    if "cornell2" in which and os.path.exists(os.path.join(staged, "cornellbox.npz")):
        b = scene.load_blob(os.path.join(staged, "cornellbox.npz")); b["num_bounces"] = 4
        compare("cornell2", b, passes=4, with_oracle=True)
    if "atrium" in which:
        compare("atrium", scene.build_blob(scene.procedural_scene("atrium", seed=3, width=640, height=360), 8), passes=4)
    if "soupmat" in which:
        compare("soupmat", scene.build_blob(scene.procedural_scene("soup", seed=3, width=256, height=256, all_materials=True), 8), passes=4)
    if "sponza" in which and os.path.exists(os.path.join(staged, "sponza.npz")):
        compare("sponza", scene.load_blob(os.path.join(staged, "sponza.npz")), passes=8)


if __name__ == "__main__":
    main()
