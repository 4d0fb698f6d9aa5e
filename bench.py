#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

Metric : Mrays/s (and ms per step) of the wavefront path tracer.  `--config K` picks BASELINE.json's configs[K] (default 1, the
         configuration the headline metric is quoted on):
           0  Data/cornellbox 512x512, 1 spp, binary SAH BVH, 1 bounce, SVGF off
           1  Data/Sponza 1920x1080, 8 spp, CWBVH, 4 bounces, NEE + MIS                         (headline)
           2  Data/Sponza 1920x1080, CWBVH, SVGF + TAA on, 8 displayed frames of 1 spp
           3  Data/instancing 1920x1080, TLAS/BLAS CWBVH, 8 spp (tile-sharded with --gpus N)
           4  Data/Sponza 3840x2160, CWBVH + SVGF + TAA, 16 displayed frames of 1 spp (tile-sharded with --gpus N)
Step   : accumulation configs (0, 1, 3): ONE n-spp frame = passes sample_index 0..n through the wavefront pipeline (n + 1 traced
         passes: exactly what the reference's `-N n` capture does, Src/Main.cpp:142 -- its accumulator overwrites pass 0,
         AOV.h:35-46).  SVGF configs (2, 4): F consecutive displayed frames, one traced pass + reproject / variance / 6 a-trous /
         finalize / TAA each (Pathtracer.cpp:798-837).  rays = closest-hit + shadow rays, read from the device counters.
value  : device-timed (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks),
         scene and queues resident in HBM.
e2e    : the same steps through the host facade (`Pathtracer.update()/render()`) with HOST buffers: every displayed frame uploads
         the per-frame inputs the reference uploads (TLAS nodes + per-instance tables + camera, Integrator.cpp:399-481) from pinned
         host memory and reads the displayed frame back to pinned host memory (double-buffered, the same in both arms).
N > 1  : one process per GPU (torchrun); the frame is sharded by interleaved row bands, every rank traces its rows; the frame
         gather is fused into the accumulate kernel over NVLink peer memory (`--gather nccl`: export -> all_gather -> assemble).
         Total work is fixed as N grows -> "scaling": "strong".  Rank 0 also re-renders the last frame on one GPU and compares it
         bit for bit with the gathered frame (`gather_check`).

`--impl reference` times the reference's OWN CUDA kernels (oracle/_ref/pathtracer_ref.cubin, compiled unmodified from
/root/reference) through oracle/ref_harness.cpp with the reference's launch recipe, on 1 GPU, same scene/config/metric.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOUNCES = 4
PASSES_PER_STEP = 9          # configs[1]: sample_index 0..8

CONFIGS = {
    0: dict(blob="cornellbox.npz", width=512, height=512, bounces=1, mode="accumulate", passes=2,
            name="Data/cornellbox 512x512 1spp binary SAH BVH 1 bounce SVGF off"),
    1: dict(blob="sponza.npz", width=1920, height=1080, bounces=4, mode="accumulate", passes=9,
            name="Data/Sponza 1920x1080 8spp CWBVH 4 bounces NEE+MIS"),
    2: dict(blob="sponza.npz", width=1920, height=1080, bounces=4, mode="svgf", frames=8,
            name="Data/Sponza 1920x1080 CWBVH 4 bounces SVGF+TAA, 8 displayed frames of 1 spp"),
    3: dict(blob="instancing.npz", width=1920, height=1080, bounces=4, mode="accumulate", passes=9, forward=(0.70710678, -0.15, -0.70710678),
            name="Data/instancing 1920x1080 8spp TLAS/BLAS CWBVH 4 bounces (444 instances, all BSDFs; camera turned towards the instance grid -- the file's own sensor looks away from it)"),
    4: dict(blob="sponza.npz", width=3840, height=2160, bounces=4, mode="svgf", frames=16,
            name="Data/Sponza 3840x2160 CWBVH 4 bounces SVGF+TAA, 16 displayed frames of 1 spp"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def load_workload(args):
    """The staged blob of BASELINE.json's configs[args.config] (tools/stage_data.py, from the reference's scene files); when the
    reference data was never staged, a procedural stand-in of the same film size (said so in the workload name)."""
    from gpu_raytracer_b200 import scene
    c = CONFIGS[args.config]
    staged = os.path.join(ROOT, "data", "_staged", c["blob"])
    if args.scene != "procedural" and os.path.exists(staged):
        blob = scene.load_blob(staged)
        name = c["name"] + " (blob staged from the reference's scene.xml)"
    else:
        if args.scene == "sponza":
            raise SystemExit(f"{staged} is not staged")
        kind = "cornell" if args.config == 0 else "atrium"
        desc = scene.procedural_scene(kind, seed=7, width=c["width"], height=c["height"], detail=args.detail if kind == "atrium" else 1.0)
        blob = scene.build_blob(desc, 2 if args.config == 0 else 8, c["width"], c["height"])
        name = f"procedural {kind} ({blob['triangles'].shape[0]} tris) standing in for: " + c["name"] + " (reference data not staged)"
    if (int(blob["width"]), int(blob["height"])) != (c["width"], c["height"]) or "forward" in c:
        blob = scene.retarget_blob(blob, c["width"], c["height"], forward=c.get("forward"))
    blob["num_bounces"] = c["bounces"]
    return blob, name


def workload_config(args, name):
    """The `config` object of the JSON line: identical in both arms (the driver compares them)."""
    c = CONFIGS[args.config]
    d = {"workload": name, "config_index": args.config, "mode": c["mode"], "width": c["width"], "height": c["height"], "bounces": c["bounces"],
         "l2": "ray queues + framebuffers of a pass exceed the 126 MB L2 (the scene itself is cache resident; see roofline.dram_frac)"}
    if c["mode"] == "accumulate":
        d["passes_per_step"] = c["passes"]
    else:
        d["frames_per_step"] = c["frames"]
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


def algorithmic_bytes(trav, node_bytes=80):
    """closest-hit = 24 R + 16 W + 80/node + 48/triangle + 48/instance transform per ray;
    shadow = 28 R + 80/node + 48/triangle + 48/instance (+ 16 R + 64 RMW on a miss).  (binary BVH: 32-byte nodes)"""
    closest = trav["rays"][0] * (24 + 16) + node_bytes * trav["nodes"][0] + 48 * trav["triangles"][0] + 48 * trav["instance_transforms"][0]
    shadow = trav["rays"][1] * 28 + node_bytes * trav["nodes"][1] + 48 * trav["triangles"][1] + 48 * trav["instance_transforms"][1] + trav["shadow_misses"] * (16 + 64)
    return closest, shadow


def pipeline_algorithmic_bytes(stats, closest_bytes, shadow_bytes, pixels, passes, svgf, aovs):
    """Whole-step algorithmic bytes with SURVEY.md 8d's per-unit figures: generate 48 B per primary ray; sort 69 B per extension
    ray at bounce 0, 109 B after; shade 264 B per shaded ray (+20 after bounce 0) + 44 B per shadow ray written + 52 B per
    extension ray written; accumulate 64 B per enabled AOV per pixel per pass; SVGF + TAA ~1020 B per pixel per displayed frame
    (reproject 104, variance 96, 6 a-trous x 80 + 32 feedback, finalize 196, TAA + finalize 112)."""
    trace = np.asarray(stats["trace"], dtype=np.float64); shadow = np.asarray(stats["shadow"], dtype=np.float64)
    gen = 48.0 * trace[0]
    sort = 69.0 * trace[0] + 109.0 * trace[1:].sum()
    shaded = float(np.asarray(stats["shaded"], dtype=np.float64).sum())
    shade = 264.0 * shaded + 20.0 * max(shaded - trace[0], 0.0) + 44.0 * shadow.sum() + 52.0 * trace[1:].sum()
    post = (1020.0 if svgf else 64.0 * aovs) * pixels * passes
    return gen + closest_bytes + shadow_bytes + sort + shade + post


def cpu_baseline(blob, bounces, budget_s=12.0):
    """The oracle port (CPU restatement, OpenMP over all host cores) on a bounded sample of the same workload: 72-row bands of
    the frame, pass after pass (sample_index 1, 2, ...), until the time budget is used up."""
    from oracle.oracle import Oracle
    cores = os.cpu_count() or 1
    o = Oracle(blob, num_bounces=bounces, threads=cores)
    step = min(72, o.height)                  # tall bands keep all host cores busy (OpenMP over rows)
    o.render_pass(1, rows=(0, step))            # thread pool spin-up and first touch stay outside the timed sample
    bands, t_total, rays, sample = 0, 0.0, 0, 1
    while t_total < budget_s and sample <= 8:
        for y in range(0, o.height - step + 1, step):
            before = int(o.counters.sum())
            t0 = time.perf_counter()
            o.render_pass(sample, rows=(y, y + step))
            t_total += time.perf_counter() - t0
            rays += int(o.counters.sum()) - before
            bands += 1
            if t_total >= budget_s:
                break
        sample += 1
    return dict(value=rays / t_total / 1e6, unit="Mrays/s", cores=cores, kind="port",
                sample=f"{bands} bands of {step} rows x {o.width} px of the same frame, passes 1..{sample - 1} ({rays} rays in {t_total:.1f} s), oracle/pt_oracle.c with OpenMP (diffuse + emitters + sky + NEE/MIS/RR; SVGF not included)")


def cpu_bvh_build(blob):
    """Reference CPU path the north star keeps: SAH sweep + CWBVH conversion, one job per mesh on all host cores."""
    from concurrent.futures import ThreadPoolExecutor
    from gpu_raytracer_b200 import scene
    first, count = np.asarray(blob["mesh_tri_first"]), np.asarray(blob["mesh_tri_count"])
    ranges = sorted(set(zip(first.tolist(), count.tolist())))
    tri = np.asarray(blob["triangles"])
    soups = []
    for f, c in ranges:
        t = tri[f:f + c]
        p0 = t[:, 0:3]; soups.append(np.ascontiguousarray(np.stack([p0, p0 + t[:, 3:6], p0 + t[:, 6:9]], 1), dtype=np.float32))
    cores = os.cpu_count() or 1
    kind = int(blob["bvh_kind"])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        built = list(pool.map(lambda s: scene.build_blas(s, kind), soups))
    dt = time.perf_counter() - t0
    ntri = int(sum(s.shape[0] for s in soups))
    return dict(seconds=dt, triangles=ntri, triangles_per_s=ntri / dt, meshes=len(soups), nodes=int(sum(b.node_count for b in built)), cores=cores)


def ncu_counters(kernel_prefix="k_trace8<0"):
    """Hardware counters of the dominant kernel from the committed ncu capture (bench.py itself never runs under a profiler):
    DRAM bytes per launch, issue-slot utilisation, average active lanes per executed instruction."""
    import csv
    for name in ("r2_trace8_ncu.csv", "r1_trace8_wave9_ncu.csv"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        rows = [r for r in csv.DictReader(open(path)) if r.get("kernel", "").startswith(kernel_prefix)]
        if not rows:
            continue
        def mean(col, scale=1.0):
            v = [float(r[col]) * scale for r in rows if r.get(col) not in (None, "")]
            return sum(v) / len(v) if v else None
        dram = None
        if "dram__bytes_read.sum" in rows[0]:
            dram = mean("dram__bytes_read.sum", 1e6) + mean("dram__bytes_write.sum", 1e6)
        return dict(traffic=dram, issue_active=mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    lanes_per_inst=mean("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    source=f"profiles/{name} (ncu --set full, mean over the closest-hit launches of one frame)")
    return dict(traffic=None, issue_active=None, lanes_per_inst=None, source=None)


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, blob, workload):
    """Reference arm: the reference's own kernels on GPU 0 (rank 0 only)."""
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/pathtracer_ref.cubin not built (needs /root/reference at build time)"}))
        return
    import torch
    from gpu_raytracer_b200 import pathtracer as pt
    c = CONFIGS[args.config]
    svgf = c["mode"] == "svgf"
    torch.cuda.set_device(0)
    cfg = pt.default_config(num_bounces=c["bounces"], enable_svgf=1 if svgf else 0)
    r = ref.Reference(blob, config=cfg)
    sample_no, dl_no = [0], [0]

    def displayed():
        """queue the pipelined read-back of the frame just rendered; await the previous one"""
        dl_no[0] += 1
        k = dl_no[0] & 1
        r.begin_display_download(k)
        if dl_no[0] > 1:
            r.wait_display_download(k ^ 1)

    def step(download=False):
        if svgf:
            for _ in range(c["frames"]):
                r.render_pass(sample_no[0]); sample_no[0] += 1
                if download:
                    displayed()
        else:
            for si in range(c["passes"]):
                r.render_pass(si)
            if download:
                displayed()

    for _ in range(args.warmup):
        step()
    r.sync(); r.ray_stats(reset=True)
    sampler = ClockSampler(0); sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    r.sync()
    dt = time.perf_counter() - t0          # the harness launches on the NULL stream of the primary context; wall clock brackets a full sync
    clocks = sampler.stop()
    st = r.ray_stats(reset=True)
    rays = int(st["trace"].sum() + st["shadow"].sum())
    # e2e: + read-back of every displayed frame to pinned host memory, double-buffered behind the next frame (as in the product arm)
    step(download=True); step(download=True); r.wait_display_download(dl_no[0] & 1)
    r.sync(); r.ray_stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(download=True)
    r.wait_display_download(dl_no[0] & 1)
    r.sync()
    dt_e2e = time.perf_counter() - t0
    st2 = r.ray_stats(reset=True)
    rays_e2e = int(st2["trace"].sum() + st2["shadow"].sum())
    frame_bytes = r.height * r.pitch * 16
    reads_per_step = c["frames"] if svgf else 1
    value = rays / dt / 1e6
    launched = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": launched, "reference_uses_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, workload),
            "details": {"reference_arm": "reference CUDA kernels (Pathtracer.cu compiled unmodified) driven by oracle/ref_harness.cpp, reference launch recipe; the reference is single-GPU: whatever N it is launched with, rank 0 runs it on GPU 0 and the other ranks exit",
                        "launch_geometry": r.launch_geometry(), "timing": "wall clock around a full device synchronise (the harness launches on the NULL stream)"},
            "rays_per_step": rays // args.steps, "clocks": clocks, "gpu_launches": 0,
            "e2e": {"value": rays_e2e / dt_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(frame_bytes * reads_per_step),
                    "ms_per_step": dt_e2e / args.steps * 1e3, "readback": "every displayed frame, double-buffered pinned host memory (same scheme as the product arm)"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 0, "kind": "reference-cuda", "sample": "full workload on 1 GPU (the reference has no CPU implementation of this path)"}}
    print(json.dumps(line))
    r.close()


# ------------------------------------------------------------------------------------------------- product arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ptb", choices=["ptb", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="index into BASELINE.json's configs (default 1 = the headline configuration)")
    ap.add_argument("--scene", default="auto", choices=["auto", "sponza", "procedural"])
    ap.add_argument("--detail", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--band-rows", type=int, default=8)
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="multi-GPU frame gather: fused peer-memory stores or NCCL all-gather")
    ap.add_argument("--wave", type=int, default=0, help="passes traced together per wave (0 = all passes of a frame; 1 = pass by pass like the reference)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    c = CONFIGS[args.config]
    svgf = c["mode"] == "svgf"
    W, H = c["width"], c["height"]
    passes = c["passes"] if not svgf else 1
    frames = c["frames"] if svgf else 1
    wave = 1 if svgf else (args.wave or passes)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        blob, workload = load_workload(args)
        run_reference(args, blob, workload)
        return

    import ctypes
    import torch
    import torch.distributed as dist
    from gpu_raytracer_b200 import pathtracer as pt, tiles
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    blob, workload = load_workload(args)
    cfg = pt.default_config(num_bounces=c["bounces"], enable_svgf=1 if svgf else 0)
    p = pt.Pathtracer(blob, device=local_rank, rank=rank, world=world, band_rows=args.band_rows, config=cfg)
    p.reserve_wave(wave)                        # all passes of a frame travel through the pipeline together (bit-identical result)
    stream = torch.cuda.ExternalStream(p.stream())
    hbm_peak, peak_src = measured_peaks()
    lib = pt.lib()
    rt_path = os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12")
    lib_rt = ctypes.CDLL(rt_path) if os.path.exists(rt_path) else ctypes.CDLL("libcudart.so")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mx_rows = tiles.max_owned_rows(H, world, args.band_rows)
    use_nccl_buffers = world > 1 and not svgf
    packed = torch.zeros((mx_rows, p.screen_pitch, 4), dtype=torch.float32, device="cuda") if use_nccl_buffers else None
    gathered = torch.empty((world * mx_rows, p.screen_pitch, 4), dtype=torch.float32, device="cuda") if use_nccl_buffers else None   # rank-major
    frame = torch.empty((H, p.screen_pitch, 4), dtype=torch.float32, device="cuda") if use_nccl_buffers else None

    # ---- the per-frame gather.  "p2p": fused into the last accumulate kernel over NVLink peer memory (ptb_exchange_*, every rank
    # ends the frame holding the whole image, no extra launches); "nccl": export -> all_gather -> assemble (3 launches + NCCL).
    gather_mode = "none"
    if world > 1:
        gather_mode = args.gather if not svgf else "p2p"
        if gather_mode == "p2p":
            ok, handle, why = 1, b"", ""
            try:
                _, handle = p.exchange_create()
            except Exception as e:                      # e.g. IPC not permitted in this container
                ok, why = 0, str(e)
            handles = [None] * world
            dist.all_gather_object(handles, handle)     # every rank takes part in every collective, whatever happened above
            if ok and all(len(h) == 64 for h in handles):
                try:
                    p.exchange_connect_ipc(handles)
                except Exception as e:
                    ok, why = 0, str(e)
            else:
                ok = 0
            if not ok:
                print(f"[bench] rank {rank}: peer-memory exchange unavailable ({why or 'a peer could not export its block'}); using the NCCL gather", file=sys.stderr, flush=True)
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if svgf:
                    raise SystemExit("SVGF across ranks needs the peer-memory exchange (CUDA IPC); it is unavailable here")
                if ok:
                    p.exchange_disconnect()
                gather_mode = "nccl"

    def one_step():
        """one step of the workload on the ctx stream, asynchronous"""
        if svgf:
            for _ in range(frames):
                p.update(); p.render()            # Integrator::update + Pathtracer::render of one displayed frame (Main.cpp:137-138)
        else:
            p.render_frame(passes - 1)            # sample_index 0..passes-1 (Integrator.cpp:518-526), replayed as one CUDA graph
            if gather_mode == "nccl":
                with torch.cuda.stream(stream):
                    p.export_rows(packed.data_ptr(), pt.AOV_RADIANCE)
                    dist.all_gather_into_tensor(gathered, packed)
                    p.assemble_rows(gathered.data_ptr(), mx_rows, frame.data_ptr())

    def timed_steps(n):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record()
        for _ in range(n):
            one_step()
        with torch.cuda.stream(stream):
            ev1.record()
        barrier()
        p.sync()
        st = p.ray_stats(reset=True)
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        r = torch.tensor([float(st["trace"].sum() + st["shadow"].sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(r, op=dist.ReduceOp.SUM)
        return float(t.item()), float(r.item()), st

    # ---- warm-up, then the roofline accounting pass (instrumented traversal; outside every timed region)
    for _ in range(args.warmup):
        one_step()
    p.sync()
    trav = p.measure_traversal(1) if not svgf else None
    p.ray_stats(reset=True)
    if trav is None:                              # SVGF keeps temporal state: measure the traversal on a scratch context instead
        q = pt.Pathtracer(blob, device=local_rank, rank=rank, world=world, band_rows=args.band_rows, config=pt.default_config(num_bounces=c["bounces"]))
        trav = q.measure_traversal(1); q.close()

    # ---- timed region: device time, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = p.launch_count()
    ms_max, rays_total, st_main = timed_steps(args.steps)
    launches = p.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- the same steps with the static merge off: the reference's two-level TLAS -> BLAS walk only (bit-exact mode)
    two_level = None
    if int(blob["bvh_kind"]) == 8:
        p.set_static_merge(False)
        for _ in range(2):
            one_step()
        t_tl, r_tl, _ = timed_steps(args.steps)
        r_tl *= args.steps / (args.steps + 2)      # the counters also hold the two warm-up steps
        two_level = {"value": r_tl / (t_tl * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": t_tl / args.steps,
                     "note": "ptb_set_static_merge(0): the reference's TLAS->BLAS traversal only; every pixel bit-identical to the reference kernels"}
        p.set_static_merge(True)
        one_step(); p.sync(); p.ray_stats(reset=True)

    # ---- per-kernel timing of the dominant kernel: every launch of one step bracketed by CUDA events on its stream
    p.set_timing(True)
    one_step(); p.sync()
    stage_step = p.stage_ms()
    if svgf:                                      # stage_ms holds the last displayed frame only: scale to the step
        stage_step = {k: v * frames for k, v in stage_step.items()}
    p.set_timing(False)
    trace_ms, shadow_ms = stage_step["trace"], stage_step["shadow_trace"]
    waves = -(-passes // wave) * frames
    n_trace_launch = c["bounces"] * waves
    st_one = p.ray_stats(reset=True)
    node_bytes = 80 if int(blob["bvh_kind"]) == 8 else 32
    closest_bytes, shadow_bytes = algorithmic_bytes(trav, node_bytes)
    hw = ncu_counters() if (args.config == 1 and world == 1) else dict(traffic=None, issue_active=None, lanes_per_inst=None, source=None)

    # trav is one pass; scale node/triangle work to the passes of a step by the measured ray ratio of that step
    scale_c = float(st_one["trace"].sum()) / max(trav["rays"][0], 1)
    scale_s = float(st_one["shadow"].sum()) / max(trav["rays"][1], 1)
    ach_gbs = closest_bytes * scale_c / (max(trace_ms, 1e-9) * 1e-3) / 1e9
    avg_launch_ms = trace_ms / n_trace_launch
    pipeline_bytes = pipeline_algorithmic_bytes(st_one, closest_bytes * scale_c, shadow_bytes * scale_s, p.owned_rows() * W, passes * frames, svgf,
                                                bin(int(cfg.aov_mask) | 1).count("1"))
    step_ms_local = ms_max / args.steps
    roofline = {"kernel": "k_trace8<closest-hit>" if node_bytes == 80 else "k_trace2<closest-hit>", "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs, burst copy)", "traffic": hw["traffic"], "traffic_source": hw["source"],
                "dram_frac": (hw["traffic"] / (avg_launch_ms * 1e-3) / 1e9 / hbm_peak) if hw["traffic"] else None,
                "issue_active": hw["issue_active"], "lanes_per_inst": hw["lanes_per_inst"],
                "limiter": "instruction issue (the scene is L1/L2 resident: DRAM traffic is a few percent of the algorithmic bytes; `frac` is the contract's HBM-roofline fraction of the ALGORITHMIC bytes, not the kernel's limiter)",
                "algorithmic_bytes_per_launch": closest_bytes * scale_c / n_trace_launch,
                "avg_launch_ms": avg_launch_ms, "launches_timed": n_trace_launch,
                "algorithmic_bytes_per_ray": closest_bytes / max(trav["rays"][0], 1),
                "nodes_per_ray": trav["nodes"][0] / max(trav["rays"][0], 1), "triangles_per_ray": trav["triangles"][0] / max(trav["rays"][0], 1),
                "shadow_kernel": {"achieved": shadow_bytes * scale_s / (max(shadow_ms, 1e-9) * 1e-3) / 1e9,
                                  "algorithmic_bytes_per_ray": shadow_bytes / max(trav["rays"][1], 1)},
                "whole_pipeline_algorithmic_gbs": pipeline_bytes / (step_ms_local * 1e-3) / 1e9}

    # ---- multi-GPU self-check: rank 0 re-renders the last step on ONE GPU and compares with the frame the ranks gathered
    gather_check = None
    if world > 1:
        one_step(); p.sync()
        if gather_mode == "p2p":
            src = p.exchange_frame()
        else:
            src = frame.data_ptr()
        got = torch.empty((H, p.screen_pitch, 4), dtype=torch.float32, device="cuda")
        lib_rt.cudaMemcpyAsync(ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(src), ctypes.c_size_t(got.numel() * 4), 3, ctypes.c_void_p(p.stream()))
        p.sync()
        if rank == 0 and not svgf:                # (SVGF keeps temporal state across the run: tests/test_gpu_properties.py covers that path)
            solo = pt.Pathtracer(blob, device=local_rank, config=cfg); solo.reserve_wave(wave)
            solo.render_frame(passes - 1); solo.sync()
            want = torch.from_numpy(solo.get_aov(pt.AOV_RADIANCE)).cuda()
            solo.close()
            diff = (got[:, :W].view(torch.int32) != want[:, :W].view(torch.int32)).any(-1)
            gather_check = {"pixels_differing": int(diff.sum().item()), "pixels": W * H, "equal": bool(not diff.any().item()),
                            "what": "frame gathered from all ranks over " + gather_mode + " vs the same frame rendered by one rank, compared bit for bit"}
            del want
        p.ray_stats(reset=True)
        barrier()

    # ---- e2e: host buffers in, host frame out, every displayed frame
    tl_nodes = np.ascontiguousarray(np.asarray(blob["bvh_nodes"])[: int(blob["tlas_node_count"]) * node_bytes])
    inst = [np.ascontiguousarray(blob[k]) for k in ("mesh_bvh_root_indices", "mesh_material_ids", "mesh_transforms", "mesh_transforms_inv", "mesh_transforms_prev")]
    pinned = [torch.from_numpy(a.copy()).pin_memory() for a in [tl_nodes] + inst]
    h2d = int(sum(a.numel() * a.element_size() for a in pinned)) + 60 + 44

    # The host<->device traffic of a frame is pipelined behind the rendering of the next one (separate copy stream, two pinned host
    # frames): frame k uploads its inputs, renders, parks its displayed image in a staging buffer (device to device, on the render
    # stream) and queues the device->host read of THAT buffer; the read of frame k-1 is awaited at the end of frame k.  Every frame
    # still pays its upload and its read-back inside the timed region.
    copy_stream = torch.cuda.Stream()
    host_frames = [torch.empty((H, p.screen_pitch, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_stage = [torch.empty((H, p.screen_pitch, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
    frame_bytes = host_frames[0].numel() * 4
    reads = [None, None]
    frame_no = [0]
    host_trace = [] if os.environ.get("PTB_BENCH_TRACE") else None

    def e2e_displayed_frame():
        k = frame_no[0] & 1
        t_a = time.perf_counter()
        lib.ptb_update_instances(p._ctx, ctypes.c_void_p(pinned[0].data_ptr()), int(blob["tlas_node_count"]), int(inst[0].size),
                                 *[ctypes.c_void_p(x.data_ptr()) for x in pinned[1:]])
        t_b = time.perf_counter()
        if svgf:
            p.update(); p.render()
        else:
            p.invalidated_camera = True
            p.render_frame(passes - 1)
            if gather_mode == "nccl":
                with torch.cuda.stream(stream):
                    p.export_rows(packed.data_ptr(), pt.AOV_RADIANCE)
                    dist.all_gather_into_tensor(gathered, packed)
                    p.assemble_rows(gathered.data_ptr(), mx_rows, frame.data_ptr())
        t_c = time.perf_counter()
        if rank == 0:
            if gather_mode == "p2p":
                src0 = p.exchange_frame()
            elif gather_mode == "nccl":
                src0 = frame.data_ptr()
            else:
                src0 = p.display_device_ptr()[0]
            lib_rt.cudaMemcpyAsync(ctypes.c_void_p(dev_stage[k].data_ptr()), ctypes.c_void_p(src0), ctypes.c_size_t(frame_bytes), 3, ctypes.c_void_p(p.stream()))
            rendered = torch.cuda.Event()
            rendered.record(stream)
            copy_stream.wait_event(rendered)
            lib_rt.cudaMemcpyAsync(ctypes.c_void_p(host_frames[k].data_ptr()), ctypes.c_void_p(dev_stage[k].data_ptr()), ctypes.c_size_t(frame_bytes), 2, ctypes.c_void_p(copy_stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(copy_stream)
            reads[k] = done
            t_d = time.perf_counter()
            if reads[k ^ 1] is not None:
                reads[k ^ 1].synchronize()              # the previous frame is now in host memory
                stream.wait_event(reads[k ^ 1])         # ... and its staging buffer may be overwritten two frames later
            if host_trace is not None:
                host_trace.append((t_b - t_a, t_c - t_b, t_d - t_c, time.perf_counter() - t_d))
        frame_no[0] += 1

    def e2e_step():
        for _ in range(frames):
            e2e_displayed_frame()

    def e2e_drain():
        for ev in reads:
            if ev is not None:
                ev.synchronize()
        p.sync()

    e2e_step(); e2e_drain()
    p.ray_stats(reset=True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    if host_trace:
        print("[bench] host ms per e2e frame (update_instances, render, queue read-back, wait previous read-back):", file=sys.stderr)
        for h in host_trace[-args.steps:]:
            print("   " + " ".join(f"{x * 1e3:7.3f}" for x in h), file=sys.stderr)
    ms_e2e = max(ev0.elapsed_time(ev1), 0.0)
    st2 = p.ray_stats(reset=True)
    t = torch.tensor([max(ms_e2e, wall_e2e)], dtype=torch.float64, device="cuda"); r2 = torch.tensor([float(st2["trace"].sum() + st2["shadow"].sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(r2, op=dist.ReduceOp.SUM)
    e2e_val = float(r2.item()) / (float(t.item()) * 1e-3) / 1e6

    if rank == 0:
        value = rays_total / (ms_max * 1e-3) / 1e6
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, workload),
                "details": {"parallelism": f"tiles{world}x(bands of {args.band_rows} rows)" if world > 1 else "1 GPU", "passes_per_wave": wave, "gather": gather_mode,
                            "static_merge": "identity-transform instances traced through one merged CWBVH (<= 1e-4 rel-L2 vs the reference, tests/test_gpu_parity.py)",
                            "rng_tables": blob.get("rng_source", "?"), "ms_per_displayed_frame": ms_max / args.steps / frames},
                "rays_per_step": int(rays_total / args.steps), "clocks": clocks, "gpu_launches": int(launches),
                "stage_ms_per_step": stage_step,
                "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * frames, "d2h_bytes_per_step": int(frame_bytes) * frames, "ms_per_step": float(t.item()) / args.steps,
                        "readback": "every displayed frame, double-buffered pinned host memory"},
                "roofline": roofline}
        if two_level:
            line["two_level_only"] = two_level
        if gather_check:
            line["gather_check"] = gather_check
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(blob, c["bounces"])
            line["cpu_bvh_build"] = cpu_bvh_build(blob)
        print(json.dumps(line), flush=True)
    # teardown order matters: tensors that were used on the ctx stream must die before the stream does
    del packed, gathered, frame, host_frames, dev_stage, pinned, reads
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    torch.cuda.empty_cache()
    p.close()


if __name__ == "__main__":
    main()
