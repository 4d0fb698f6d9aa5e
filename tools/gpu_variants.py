"""A/B timing of tuning variants of libptb (built with -D flags by build.build_cuda(defines=..., suffix=...)).
Each variant runs in its own process: Sponza 1080p, 4 bounces, per-stage device time averaged over PASSES passes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
from gpu_raytracer_b200 import pathtracer as pt, scene
blob = scene.load_blob(os.path.join(%r, "data", "_staged", "sponza.npz")); blob["num_bounces"] = 4
p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4))
for si in range(4): p.render_pass(si)
p.sync(); p.set_timing(True)
tot = {}
N = 18
for si in range(N):
    p.render_pass(si %% 9); p.sync()
    for k, v in p.stage_ms().items(): tot[k] = tot.get(k, 0.0) + v / N
img = p.get_aov(0)
import zlib
tot["crc"] = zlib.crc32(img.tobytes())
print("RESULT " + json.dumps(tot))
'''


def main():
    variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {}
    for name, path in variants.items():
        env = dict(os.environ, PTB_LIB_PATH=path)
        out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT)], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(name, "FAILED", out.stderr[-400:]); continue
        r = json.loads(line[0][7:])
        total = sum(v for k, v in r.items() if k != "crc")
        print(f"{name:28s} trace {r['trace']:.3f} shadow {r['shadow_trace']:.3f} sort {r['sort']:.3f} shade {r['shade']:.3f} total {total:.3f} ms/pass crc {r['crc']:08x}", flush=True)


if __name__ == "__main__":
    main()
