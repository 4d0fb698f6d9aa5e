"""Generates tests/golden/*.npz on the GPU box by running THE REFERENCE'S OWN KERNELS (oracle/_ref/pathtracer_ref.cubin,
Src/CUDA/Pathtracer.cu compiled unmodified) through oracle/ref_harness.cpp on small deterministic scenes.

    gpurun -- python tests/golden/make_golden.py        # writes into gpurun_out/golden/, copy to tests/golden/

The scenes are procedural (gpu-raytracer_b200/scene.py, fixed seeds) and use the FALLBACK RNG tables so the fixtures can
be regenerated and checked from a clean clone without /root/reference.  Each fixture holds, for one scene/config:
  hits      [H,W,4] uint32   bounce-0 hit table of pass sample_index=1 (mesh, triangle, t bits, u16|v16<<16), pixel keyed
  radiance  [H,W,3] float32  accumulator after passes sample_index 0..PASSES
  albedo/normal/position     bounce-0 AOV accumulators
  trace/shadow  per-bounce ray counts summed over those passes
These pin the CPU restatement (oracle/pt_oracle.c, `pytest -m "not gpu"`) and the CUDA path (`pytest -m gpu`).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402

PASSES = 4
CASES = {
    "soup_bvh8":    dict(kind="soup", seed=5, size=(96, 64), bvh=8, bounces=3),
    "cornell_bvh8": dict(kind="cornell", seed=1, size=(64, 64), bvh=8, bounces=4),
    "cornell_bvh2": dict(kind="cornell", seed=1, size=(64, 64), bvh=2, bounces=4),
    "atrium_bvh8":  dict(kind="atrium", seed=2, size=(128, 72), bvh=8, bounces=3, detail=0.5),
}


def case_blob(c):
    desc = scene.procedural_scene(c["kind"], seed=c["seed"], width=c["size"][0], height=c["size"][1], detail=c.get("detail", 0.25 if c["kind"] == "soup" else 1.0))
    blob = scene.build_blob(desc, c["bvh"], rng="fallback")
    blob["num_bounces"] = c["bounces"]
    return blob


def main():
    from oracle import ref
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    for name, c in CASES.items():
        blob = case_blob(c)
        w = c["size"][0]
        cfg1 = pt.default_config(num_bounces=1)
        r = ref.Reference(blob, config=cfg1)
        r.render_pass(1); r.sync()
        hits = r.primary_hits()[:, :w].copy()
        miss = hits[..., 1] == 0xFFFFFFFF
        hits[miss, 0] = 0; hits[miss, 3] = 0       # uninitialised in the reference for misses
        r.close()
        cfg = pt.default_config(num_bounces=c["bounces"], aov_mask=0x3F)
        r = ref.Reference(blob, config=cfg)
        r.render_frames(PASSES)
        st = r.ray_stats()
        np.savez_compressed(os.path.join(out, name + ".npz"), hits=hits,
                            radiance=r.get_aov(0)[:, :w, :3], albedo=r.get_aov(3)[:, :w, :3], normal=r.get_aov(4)[:, :w, :3], position=r.get_aov(5)[:, :w, :3],
                            trace=st["trace"][:8].astype(np.int64), shadow=st["shadow"][:8].astype(np.int64), passes=np.int64(PASSES),
                            triangles=np.int64(blob["triangles"].shape[0]), nodes=np.int64(blob["bvh_nodes"].size))
        print("golden", name, "rays", int(st["trace"].sum()), int(st["shadow"].sum()), flush=True)
        r.close()


if __name__ == "__main__":
    main()
