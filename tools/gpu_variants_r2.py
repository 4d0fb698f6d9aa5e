"""Round-2 A/B of libptb variants (PTB_LIB_PATH) in the bench configuration (Sponza 1080p, 9-pass wave, graph replay), per merge mode:
device ms per frame (CUDA events, 10 frames), per-stage ms, CRC of the frame, traversal visit counts.  Each variant runs in its own
process under a timeout (a bad traversal variant must not take the box down).
usage: gpu_variants_r2.py '{"name": "lib path", ...}' [merge modes, default "1,2,0"]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, zlib
sys.path.insert(0, %r)
import torch
from gpu_raytracer_b200 import pathtracer as pt, scene
blob = scene.load_blob(os.path.join(%r, "data", "_staged", "sponza.npz"))
mode = int(sys.argv[1])
p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4)); p.set_static_merge(mode); p.reserve_wave(9)
if os.environ.get("PTB_WOOP"): p.set_intersector("woop")
for _ in range(3): p.render_frame(8)
p.sync()
s = torch.cuda.ExternalStream(p.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s): e0.record()
for _ in range(10): p.render_frame(8)
with torch.cuda.stream(s): e1.record()
p.sync()
ms = e0.elapsed_time(e1) / 10
crc = zlib.crc32(p.get_aov(0).tobytes())
p.set_timing(True); p.render_frame(8); p.sync(); st = p.stage_ms(); p.set_timing(False)
tr = p.measure_traversal(1)
print("RESULT mode=%%d %%.3f ms crc=%%08x trace=%%.2f shadow=%%.2f sort=%%.3f shade=%%.3f acc=%%.3f | closest nodes/ray %%.2f tris/ray %%.2f | shadow nodes/ray %%.2f tris/ray %%.2f" %% (
    mode, ms, crc, st["trace"], st["shadow_trace"], st["sort"], st["shade"], st["accumulate_or_svgf"], tr["nodes"][0] / tr["rays"][0], tr["triangles"][0] / tr["rays"][0],
    tr["nodes"][1] / max(tr["rays"][1], 1), tr["triangles"][1] / max(tr["rays"][1], 1)))
del s, e0, e1
p.close()
'''
modes = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "1,2,0").split(",")]
for name, path in json.loads(sys.argv[1]).items():
    for mode in modes:
        try:
            env = dict(os.environ, PTB_LIB_PATH=os.path.join(ROOT, path.split("+")[0]))
            for extra in path.split("+")[1:]:                 # "lib.so+woop", "lib.so+PTB_NODE_TEST_EXACT=1"
                if extra == "woop": env["PTB_WOOP"] = "1"
                elif "=" in extra: env[extra.split("=")[0]] = extra.split("=", 1)[1]
            out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT), str(mode)], env=env, capture_output=True, text=True, timeout=240)
            line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
            print(name, line[0] if line else "FAILED " + out.stderr[-400:], flush=True)
        except subprocess.TimeoutExpired:
            print(name, f"mode={mode} TIMEOUT", flush=True)
