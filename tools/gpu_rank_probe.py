"""What one rank of an N-way tile shard costs, measured on ONE GPU (no exchange): device ms per 9-pass frame for rank 0 of
world = 1, 2, 4, 8, against the ideal 1/N of the whole frame; per-stage breakdown from the event timers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpu_raytracer_b200 import pathtracer as pt, scene
blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", "sponza.npz"))
base = None
for world in (1, 2, 4, 8):
    for rank in ((0,) if world < 8 else (0, 3)):
        p = pt.Pathtracer(blob, rank=rank, world=world, band_rows=8, config=pt.default_config(num_bounces=4))
        p.reserve_wave(9)
        for _ in range(3): p.render_frame(8)
        p.sync()
        s = torch.cuda.ExternalStream(p.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N = 10
        with torch.cuda.stream(s): e0.record()
        for _ in range(N): p.render_frame(8)
        with torch.cuda.stream(s): e1.record()
        p.sync()
        ms = e0.elapsed_time(e1) / N
        p.set_timing(True); p.render_frame(8); p.sync(); st = p.stage_ms(); p.set_timing(False)
        if base is None: base = ms
        print(f"world {world} rank {rank}: {ms:.3f} ms/frame (ideal {base / world:.3f}, efficiency {base / world / ms:.3f}) stages " +
              " ".join(f"{k} {v:.2f}" for k, v in st.items() if v > 0), flush=True)
        del s, e0, e1
        p.close()
