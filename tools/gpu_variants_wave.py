"""A/B of libptb tuning variants (PTB_LIB_PATH) in the bench configuration: Sponza 1080p, 9-pass wave, graph replay.
Device ms per frame (CUDA events over 10 frames) + CRC of the frame."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, zlib
sys.path.insert(0, %r)
import torch
from gpu_raytracer_b200 import pathtracer as pt, scene
blob = scene.load_blob(os.path.join(%r, "data", "_staged", "sponza.npz"))
p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=4)); p.reserve_wave(9)
for _ in range(3): p.render_frame(8)
p.sync()
s = torch.cuda.ExternalStream(p.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s): e0.record()
for _ in range(10): p.render_frame(8)
with torch.cuda.stream(s): e1.record()
p.sync()
print("RESULT %%.3f %%08x" %% (e0.elapsed_time(e1) / 10, zlib.crc32(p.get_aov(0).tobytes())))
del s, e0, e1
p.close()
'''
for name, path in json.loads(sys.argv[1]).items():
    out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT)], env=dict(os.environ, PTB_LIB_PATH=path), capture_output=True, text=True, timeout=300)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print(name, line[0] if line else "FAILED " + out.stderr[-300:], flush=True)
