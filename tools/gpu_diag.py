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


def bits(a, b, w):
    a = a[:, :w]; b = b[:, :w]
    if a.dtype != np.int32:
        a3, b3 = a.view(np.uint32), b.view(np.uint32)
    else:
        a3, b3 = a, b
    nd = (a3 != b3)
    nd = nd.any(-1) if nd.ndim == 3 else nd
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return f"bitdiff {int(nd.sum())} max {float(d.max()):.2e}"


def svgf_diag():
    blob = scene.build_blob(scene.procedural_scene("atrium", seed=4, width=320, height=180, detail=0.5), 8, rng="fallback")
    w = 320
    for label, over in (("svgf+var+taa", dict(enable_svgf=1, enable_spatial_variance=1, enable_taa=1)),
                        ("svgf only, 0 atrous", dict(enable_svgf=1, enable_spatial_variance=0, enable_taa=0, num_atrous_iterations=0))):
        cfg = pt.default_config(num_bounces=3, **over)
        p = pt.Pathtracer(blob, config=cfg); r = ref.Reference(blob, config=cfg)
        print("=== svgf", label, flush=True)
        for si in range(4):
            p.render_pass(si); r.render_pass(si); p.sync(); r.sync()
            line = [f"frame {si}: display {bits(p.get_display(), r.get_display(), w)}"]
            for name in ("history_normal_and_depth", "frame_buffer_moment", "history_moment", "history_direct", "history_indirect", "history_length", "taa_frame_curr", "taa_frame_prev"):
                line.append(f"{name} {bits(p.svgf_buffer(name), r.svgf_buffer(name), w)}")
            for k, nm in ((1, "direct.acc"), (2, "indirect.acc"), (3, "albedo.acc")):
                line.append(f"{nm} {bits(p.get_aov(k, True), r.get_aov(k, True), w)}")
            print("  " + " | ".join(line), flush=True)
        p.close(); r.close()


def material_diag(kind, mat):
    d = scene.procedural_scene("soup", seed=3, width=256, height=256)
    m = d.add_material(mat)
    for inst in d.instances[3:7]:
        inst.material = m
    blob = scene.build_blob(d, 8, rng="fallback")
    w = 256
    print("=== material", kind, flush=True)
    for label, over in (("default", {}), ("nee off", dict(enable_next_event_estimation=0)), ("mis off", dict(enable_multiple_importance_sampling=0))):
        cfg = pt.default_config(num_bounces=2, aov_mask=0x3F, **over)
        p = pt.Pathtracer(blob, config=cfg); r = ref.Reference(blob, config=cfg)
        p.render_pass(1); r.render_pass(1); p.sync(); r.sync()
        a, b = p.get_aov(0)[:, :w, :3], r.get_aov(0)[:, :w, :3]
        dd = np.abs(a.astype(np.float64) - b).max(-1)
        print(f"  {label}: {bits(p.get_aov(0), r.get_aov(0), w)} rays {p.ray_stats()['trace'][:2]} {r.ray_stats()['trace'][:2]} shadow {p.ray_stats()['shadow'][:2]} {r.ray_stats()['shadow'][:2]}", flush=True)
        ys, xs = np.unravel_index(np.argsort(dd, axis=None)[::-1][:4], dd.shape)
        ph = None
        for y, x in zip(ys, xs):
            print(f"     px ({x},{y}) ptb {a[y, x]} ref {b[y, x]} normal {p.get_aov(4)[y, x, :3]} pos {p.get_aov(5)[y, x, :3]}", flush=True)
        if label == "nee off":
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"matdump_{kind}.npz"), ptb=p.get_aov(0)[:, :w, :3], ref=r.get_aov(0)[:, :w, :3],
                                normal=p.get_aov(4)[:, :w, :3], position=p.get_aov(5)[:, :w, :3], luts=p.lut_contents(), luts_ref=r.lut_contents(),
                                camera=blob["camera"], pmj=blob["pmj"], blue=blob["blue_noise"])
        p.close(); r.close()


def lut_diag():
    d = scene.procedural_scene("soup", seed=3, width=64, height=64, all_materials=True)
    blob = scene.build_blob(d, 8, rng="fallback")
    p = pt.Pathtracer(blob); r = ref.Reference(blob)
    a, b = p.lut_contents(), r.lut_contents()
    names = [("diel_dir_enter", 4096), ("diel_dir_leave", 4096), ("diel_enter", 256), ("diel_leave", 256), ("cond_dir", 1024), ("cond", 32)]
    off = 0
    for n, sz in names:
        x, y = a[off:off + sz], b[off:off + sz]
        print(f"  LUT {n}: ptb [{x.min():.4f},{x.max():.4f}] ref [{y.min():.4f},{y.max():.4f}] max abs diff {np.abs(x - y).max():.3e} bitdiff {(x.view(np.uint32) != y.view(np.uint32)).sum()}/{sz}", flush=True)
        if np.abs(x - y).max() > 1e-3:
            k = int(np.abs(x - y).argmax()); print("     worst at", k, x[k], y[k], "first values", x[:4], y[:4])
        off += sz
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
    if "svgf" in which:
        svgf_diag()
    if "dielectric2" in which:
        material_diag("dielectric", scene.Material(scene.MAT_DIELECTRIC, "d", ior=1.5, roughness=0.3))
    if "luts" in which:
        lut_diag()
    if "sponza_nomip" in which:
        diag("sponza_nomip", scene.load_blob(os.path.join(staged, "sponza.npz")), max_nb=2, enable_mipmapping=0)


if __name__ == "__main__":
    main()
