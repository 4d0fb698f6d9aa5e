"""Stage inputs that live in /root/reference into data/_staged/ (git-ignored, travels with gpurun).

  rng_tables.npz   PMJ02 sample table (64 x 4096 points, sequences 1..63 block-shuffled exactly like the reference
                   does at start-up: Src/Util/PMJ.cpp:8-27 driven from Src/Main.cpp:88-92) and the 16 blue-noise
                   tiles (Src/Util/BlueNoise.cpp).  These are *inputs* of the hot path.
  cornellbox.npz / sponza.npz / instancing.npz   scene blobs (gpu-raytracer_b200/scene.py) for BASELINE.json's configs.

Runs only where /root/reference exists (the authoring container); everything downstream reads data/_staged/.
"""
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PTB_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "data", "_staged")


class PCG:  # Src/Core/Random.h:9-52
    def __init__(self, seed):
        self.state = ((seed + 2891336453) * 747796405 + 2891336453) & 0xFFFFFFFFFFFFFFFF

    def u32(self):
        s = self.state
        x = (((s >> 18) ^ s) >> 27) & 0xFFFFFFFF
        r = s >> 59
        self.state = (s * 6364136223846793005 + 1) & 0xFFFFFFFFFFFFFFFF
        return ((x >> r) | (x << ((-r) & 31))) & 0xFFFFFFFF

    def bounded(self, mx):
        x = self.u32(); m = x * mx; l = m & 0xFFFFFFFF
        if l < mx:
            t = (-mx) & 0xFFFFFFFF
            if t >= mx:
                t -= mx
                if t >= mx:
                    t %= mx
            while l < t:
                x = self.u32(); m = x * mx; l = m & 0xFFFFFFFF
        return m >> 32

    def ranged(self, lo, hi):
        return lo + self.bounded(hi - lo)


def stage_rng():
    out = os.path.join(OUT, "rng_tables.npz")
    if os.path.exists(out):
        return out
    txt = open(os.path.join(REF, "Src/Util/PMJ.cpp")).read()
    body = txt[txt.index("PMJ::samples["):]
    vals = np.array([int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]{8}", body)], dtype=np.uint32)
    assert vals.size == 64 * 4096 * 2, vals.size
    pts = vals.reshape(64, 4096, 2).copy()
    odd, even = [0, 1, 4, 5, 10, 11, 14, 15], [2, 3, 6, 7, 8, 9, 12, 13]
    for seq in range(1, 64):
        rng = PCG(seq)
        s = pts[seq]
        for j in range(0, 4096, 16):
            for grp in (odd, even):
                for i in range(8):
                    k = rng.ranged(i, 8)
                    a, b = j + grp[i], j + grp[k]
                    tmp = s[a].copy(); s[a] = s[b]; s[b] = tmp
    pmj = pts.view(np.float32).reshape(-1)
    txt = open(os.path.join(REF, "Src/Util/BlueNoise.cpp")).read()
    body = txt[txt.index("BlueNoise::textures["):]
    body = body[body.index("{"):]
    bn = np.array([int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]{4}\b", body)], dtype=np.uint16)
    assert bn.size == 16 * 128 * 128, bn.size
    os.makedirs(OUT, exist_ok=True)
    np.savez(out, pmj=pmj, blue_noise=bn.view(np.uint8))
    return out


def stage_scene(name, xml, width, height, bounces, bvh_kind, sky="Data/Skies/sky_15.hdr", sky_max_width=2500, textures=True):
    from gpu_raytracer_b200 import scene
    out = os.path.join(OUT, f"{name}.npz")
    if os.path.exists(out):
        return out
    t0 = time.time()
    desc = scene.load_mitsuba(os.path.join(REF, xml), os.path.join(REF, sky), load_textures=textures, sky_max_width=sky_max_width)
    desc.num_bounces = bounces
    t1 = time.time()
    timing = {}
    blob = scene.build_blob(desc, bvh_kind, width, height, timing=timing)
    scene.save_blob(blob, out)
    print(f"staged {name}: load {t1 - t0:.1f}s build {timing} -> {os.path.getsize(out) / 1e6:.1f} MB")
    return out


def main():
    if not os.path.isdir(REF):
        print("no reference tree; nothing staged"); return
    stage_rng()
    which = sys.argv[1:] or ["cornellbox", "cornellbox_bvh8", "sponza", "instancing"]
    if "cornellbox" in which:
        stage_scene("cornellbox", "Data/cornellbox/scene.xml", 512, 512, 1, 2, sky_max_width=625)
    if "cornellbox_bvh8" in which:
        stage_scene("cornellbox_bvh8", "Data/cornellbox/scene.xml", 512, 512, 4, 8, sky_max_width=625)
    if "sponza" in which:
        stage_scene("sponza", "Data/Sponza/scene.xml", 1920, 1080, 4, 8, sky_max_width=1250)
    if "instancing" in which:
        stage_scene("instancing", "Data/instancing/scene.xml", 1920, 1080, 4, 8, sky_max_width=625)


if __name__ == "__main__":
    main()
