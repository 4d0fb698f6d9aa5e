"""Parity diagnostics on the GPU box: per-bounce-depth, single-pass, bitwise comparison of product vs reference kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402
from oracle import ref  # noqa: E402

NAMES = ("radiance", "direct", "indirect", "albedo", "normal", "position")


def diag(name, blob, sample_index=1, max_nb=4, **cfg_over):
    w = int(blob["width"])
    print(f"=== {name} {w}x{blob['height']}", flush=True)
    for nb in range(1, max_nb + 1):
        cfg = pt.default_config(num_bounces=nb, aov_mask=0x3F, **cfg_over)
        p = pt.Pathtracer(blob, config=cfg)
        r = ref.Reference(blob, config=cfg)
        p.render_pass(sample_index); p.sync()
        r.render_pass(sample_index); r.sync()
        line = [f"nb={nb}"]
        if nb == 1:
            ph, rh = p.primary_hits()[:, :w], r.primary_hits()[:, :w]
            covered = rh[..., 1] != 0xFFFFFFFF if blob["width"] * blob["height"] > 1080 * 720 else np.ones(rh.shape[:2], bool)
            valid = rh[..., 1] != 0xFFFFFFFF
            neq = ((ph[..., 1:3] != rh[..., 1:3]).any(-1) | (valid & (ph != rh).any(-1))) & covered   # mesh id / uv of a miss are uninitialised in the reference
            line.append(f"hits differ {int(neq.sum())}/{int(covered.sum())} (tri {int(((ph[..., 1] != rh[..., 1]) & covered).sum())}, t {int(((ph[..., 2] != rh[..., 2]) & covered).sum())}, uv {int(((ph[..., 3] != rh[..., 3]) & covered).sum())})")
            if neq.any():
                ys, xs = np.nonzero(neq)
                for k in range(min(3, len(ys))):
                    print("    hit diff at", xs[k], ys[k], "ptb", ph[ys[k], xs[k]], "ref", rh[ys[k], xs[k]])
        for k in range(6):
            a = p.get_aov(k, False)[:, :w]; b = r.get_aov(k, False)[:, :w]
        # framebuffers were cleared by the pass; compare accumulators instead (sample_index=1 -> acc == fb of this pass)
        for k in range(6):
            a = p.get_aov(k, True)[:, :w, :3]; b = r.get_aov(k, True)[:, :w, :3]
            bit = (a.view(np.uint32) != b.view(np.uint32)).any(-1)
            d = np.abs(a.astype(np.float64) - b.astype(np.float64)).max(-1)
            big = d > 1e-3 * (1e-3 + np.abs(b).max(-1))
            line.append(f"{NAMES[k]}: bitdiff {int(bit.sum())} big {int(big.sum())} max {float(d.max()):.2e}")
        print("  " + " | ".join(line), flush=True)
        sp, sr = p.ray_stats(), r.ray_stats()
        print(f"    rays ptb {sp['trace'][:nb]} {sp['shadow'][:nb]} ref {sr['trace'][:nb]} {sr['shadow'][:nb]}", flush=True)
        p.close(); r.close()


def main():
    which = sys.argv[1:] or ["soup", "cornell8", "atrium", "sponza"]
    staged = os.path.join(ROOT, "data", "_staged")
    if "soup" in which:
        diag("soup", scene.build_blob(scene.procedural_scene("soup", seed=3, width=256, height=256), 8))
    if "cornell8" in which:
        diag("cornell8", scene.load_blob(os.path.join(staged, "cornellbox_bvh8.npz")))
    if "atrium" in which:
        diag("atrium", scene.build_blob(scene.procedural_scene("atrium", seed=3, width=640, height=360), 8))
    if "sponza" in which:
        diag("sponza", scene.load_blob(os.path.join(staged, "sponza.npz")), max_nb=3)
    for kind, mat in (("plastic", scene.Material(scene.MAT_PLASTIC, "p", diffuse=(0.2, 0.8, 0.8), roughness=0.2)),
                      ("dielectric", scene.Material(scene.MAT_DIELECTRIC, "d", ior=1.5, roughness=0.3)),
                      ("smoothglass", scene.Material(scene.MAT_DIELECTRIC, "s", ior=1.33, roughness=0.0)),
                      ("conductor", scene.Material(scene.MAT_CONDUCTOR, "c", eta=(1.45, 0.43, 0.21), k=(1.95, 2.46, 3.27), roughness=0.3))):
        if kind in which:
            d = scene.procedural_scene("soup", seed=3, width=256, height=256)
            m = d.add_material(mat)
            for inst in d.instances[3:7]:
                inst.material = m
            diag(kind, scene.build_blob(d, 8, rng="fallback"))
    if "sponza_nomip" in which:
        diag("sponza_nomip", scene.load_blob(os.path.join(staged, "sponza.npz")), max_nb=2, enable_mipmapping=0)


if __name__ == "__main__":
    main()
