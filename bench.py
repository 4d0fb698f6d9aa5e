#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Metric : Mrays/s (and ms/frame) at 1920x1080, 8 spp, Sponza CWBVH, 4 bounces, NEE + MIS    (configs[1])
Step   : ONE 8-spp frame = passes sample_index 0..8 through the wavefront pipeline (9 traced passes: exactly what the
         reference's `-N 8` capture does, Src/Main.cpp:142 -- its accumulator overwrites pass 0, AOV.h:35-46).
         rays = closest-hit + shadow rays of all 9 passes, read from the device counters (never estimated).
value  : device-timed (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks),
         scene and queues resident in HBM.
e2e    : the same frames through the host façade (`Pathtracer.update()/render()`) with HOST buffers: every step
         uploads the per-frame inputs the reference uploads (TLAS nodes + per-instance tables + camera,
         Integrator.cpp:399-481) from pinned host memory and reads the finished frame back to pinned host memory.
N > 1  : one process per GPU (torchrun); the frame is sharded by interleaved row bands, every rank traces its rows, one
         NCCL all-gather of the packed tile framebuffers per frame (inside the timed region).  Total work is fixed as N
         grows -> "scaling": "strong".

`--impl reference` times the reference's OWN CUDA kernels (oracle/_ref/pathtracer_ref.cubin, compiled unmodified from
/root/reference) through oracle/ref_harness.cpp with the reference's launch recipe, on 1 GPU, same scene/config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PASSES_PER_STEP = 9          # sample_index 0..8
WIDTH, HEIGHT, BOUNCES = 1920, 1080, 4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def load_workload(args):
    """Sponza blob staged from the reference data when present; otherwise a procedural Sponza-like atrium."""
    from gpu_raytracer_b200 import scene
    staged = os.path.join(ROOT, "data", "_staged", "sponza.npz")
    if args.scene == "sponza" or (args.scene == "auto" and os.path.exists(staged)):
        blob = scene.load_blob(staged)
        name = "Data/Sponza 1920x1080 8spp CWBVH 4 bounces NEE+MIS (blob staged from the reference's scene.xml)"
    else:
        desc = scene.procedural_scene("atrium", seed=7, width=WIDTH, height=HEIGHT, detail=args.detail)
        blob = scene.build_blob(desc, 8, WIDTH, HEIGHT)
        name = f"procedural atrium ({blob['triangles'].shape[0]} tris) 1920x1080 8spp CWBVH 4 bounces NEE+MIS (reference data not staged)"
    blob["num_bounces"] = BOUNCES
    assert (int(blob["width"]), int(blob["height"])) == (WIDTH, HEIGHT)
    return blob, name


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


def algorithmic_bytes(trav):
    """closest-hit = 24 R + 16 W + 80/node + 48/triangle + 48/instance transform per ray;
    shadow = 28 R + 80/node + 48/triangle + 48/instance (+ 16 R + 64 RMW on a miss)."""
    closest = trav["rays"][0] * (24 + 16) + 80 * trav["nodes"][0] + 48 * trav["triangles"][0] + 48 * trav["instance_transforms"][0]
    shadow = trav["rays"][1] * 28 + 80 * trav["nodes"][1] + 48 * trav["triangles"][1] + 48 * trav["instance_transforms"][1] + trav["shadow_misses"] * (16 + 64)
    return closest, shadow


def cpu_baseline(blob, budget_s=12.0):
    """The oracle port (CPU restatement, OpenMP over all host cores) on a bounded sample of the same workload: 72-row bands of
    the 1080p frame, pass after pass (sample_index 1, 2, ...), until the time budget is used up."""
    from oracle.oracle import Oracle
    cores = os.cpu_count() or 1
    o = Oracle(blob, num_bounces=BOUNCES, threads=cores)
    step = 72                  # tall bands keep all host cores busy (OpenMP over rows)
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
                sample=f"{bands} bands of {step} rows x {o.width} px of the same frame, passes 1..{sample - 1} ({rays} rays in {t_total:.1f} s), oracle/pt_oracle.c with OpenMP")


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
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        built = list(pool.map(lambda s: scene.build_blas(s, 8), soups))
    dt = time.perf_counter() - t0
    ntri = int(sum(s.shape[0] for s in soups))
    return dict(seconds=dt, triangles=ntri, triangles_per_s=ntri / dt, meshes=len(soups), nodes=int(sum(b.node_count for b in built)), cores=cores)


def run_reference(args, blob, workload):
    """Reference arm: the reference's own kernels on GPU 0."""
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/pathtracer_ref.cubin not built (needs /root/reference at build time)"}))
        return
    import torch
    from gpu_raytracer_b200 import pathtracer as pt
    torch.cuda.set_device(0)
    cfg = pt.default_config(num_bounces=BOUNCES)
    r = ref.Reference(blob, config=cfg)
    for _ in range(args.warmup):
        for si in range(PASSES_PER_STEP):
            r.render_pass(si)
    r.sync(); r.ray_stats(reset=True)
    host = np.empty((r.height, r.pitch, 4), dtype=np.float32)
    sampler = ClockSampler(0); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for si in range(PASSES_PER_STEP):
            r.render_pass(si)
    r.sync()
    dt = time.perf_counter() - t0          # the harness launches on the NULL stream of its own context; wall clock brackets a full sync
    clocks = sampler.stop()
    st = r.ray_stats()
    rays = int(st["trace"].sum() + st["shadow"].sum())
    # e2e: + frame readback to host each step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for si in range(PASSES_PER_STEP):
            r.render_pass(si)
        host[...] = r.get_display()
    r.sync()
    dt_e2e = time.perf_counter() - t0
    value = rays / dt / 1e6
    launched = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": launched, "reference_uses_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "passes_per_step": PASSES_PER_STEP, "l2": "working set (ray queues + AOVs > 400 MB per pass) exceeds the 126 MB L2",
                       "reference_arm": "reference CUDA kernels (Pathtracer.cu compiled unmodified) driven by oracle/ref_harness.cpp, reference launch recipe; the reference is single-GPU: whatever N it is launched with, rank 0 runs it on GPU 0 and the other ranks exit",
                       "launch_geometry": r.launch_geometry()},
            "rays_per_step": rays // args.steps, "clocks": clocks, "gpu_launches": 0,
            "e2e": {"value": rays / dt_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(host.nbytes)},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 0, "kind": "reference-cuda", "sample": "full workload on 1 GPU (the reference has no CPU implementation of this path)"}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ptb", choices=["ptb", "reference"])
    ap.add_argument("--scene", default="auto", choices=["auto", "sponza", "procedural"])
    ap.add_argument("--detail", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--band-rows", type=int, default=8)
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="multi-GPU frame gather: fused peer-memory stores or NCCL all-gather")
    ap.add_argument("--wave", type=int, default=PASSES_PER_STEP, help="passes traced together per wave (1 = pass by pass like the reference)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        blob, workload = load_workload(args)
        run_reference(args, blob, workload)
        return

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
    cfg = pt.default_config(num_bounces=BOUNCES)
    p = pt.Pathtracer(blob, device=local_rank, rank=rank, world=world, band_rows=args.band_rows, config=cfg)
    p.reserve_wave(args.wave)                   # all passes of a frame travel through the pipeline together (bit-identical result)
    stream = torch.cuda.ExternalStream(p.stream())
    hbm_peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mx_rows = tiles.max_owned_rows(HEIGHT, world, args.band_rows)
    packed = torch.zeros((mx_rows, p.screen_pitch, 4), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world * mx_rows, p.screen_pitch, 4), dtype=torch.float32, device="cuda") if world > 1 else None   # rank-major
    frame = torch.empty((HEIGHT, p.screen_pitch, 4), dtype=torch.float32, device="cuda") if world > 1 else None
    host_frame = torch.empty((HEIGHT, p.screen_pitch, 4), dtype=torch.float32).pin_memory()

    # ---- the per-frame gather.  "p2p": fused into the last accumulate kernel over NVLink peer memory (ptb_exchange_*, every rank
    # ends the frame holding the whole image, no extra launches); "nccl": export -> all_gather -> assemble (3 launches + NCCL).
    gather_mode = "none"
    if world > 1:
        gather_mode = args.gather
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
                if ok:
                    p.exchange_disconnect()
                gather_mode = "nccl"

    def one_frame(gather=True):
        p.render_frame(PASSES_PER_STEP - 1)       # sample_index 0..8 (Integrator.cpp:518-526), replayed as one CUDA graph
        if world > 1 and gather and gather_mode == "nccl":
            with torch.cuda.stream(stream):
                p.export_rows(packed.data_ptr(), pt.AOV_RADIANCE)
                dist.all_gather_into_tensor(gathered, packed)
                p.assemble_rows(gathered.data_ptr(), mx_rows, frame.data_ptr())

    # ---- warm-up, then the roofline accounting pass (instrumented traversal; outside every timed region)
    for _ in range(args.warmup):
        one_frame()
    p.sync()
    trav = p.measure_traversal(1)
    p.ray_stats(reset=True)

    # ---- timed region: device time, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = p.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        one_frame()
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    p.sync()
    launches = p.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    st = p.ray_stats(reset=True)
    rays_local = int(st["trace"].sum() + st["shadow"].sum())
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); r = torch.tensor([rays_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(r, op=dist.ReduceOp.SUM)
    ms_max, rays_total = float(t.item()), float(r.item())

    # ---- the same frames with the static merge off: the reference's two-level TLAS -> BLAS walk only (bit-exact mode)
    p.set_static_merge(False)
    for _ in range(2):
        one_frame()
    barrier()
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        one_frame()
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    st_tl = p.ray_stats(reset=True)
    t_tl = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    r_tl = torch.tensor([float(st_tl["trace"].sum() + st_tl["shadow"].sum()) * args.steps / (args.steps + 2)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_tl, op=dist.ReduceOp.MAX); dist.all_reduce(r_tl, op=dist.ReduceOp.SUM)
    two_level = {"value": float(r_tl.item()) / (float(t_tl.item()) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": float(t_tl.item()) / args.steps,
                 "note": "ptb_set_static_merge(0): the reference's TLAS->BLAS traversal only; every pixel bit-identical to the reference kernels"}
    p.set_static_merge(True)
    one_frame(); p.sync(); p.ray_stats(reset=True)

    # ---- per-kernel timing of the dominant kernel: every launch of one frame bracketed by CUDA events on its stream
    p.set_timing(True)
    p.render_frame(PASSES_PER_STEP - 1); p.sync()
    stage_frame = p.stage_ms()
    p.set_timing(False)
    trace_ms, shadow_ms = stage_frame["trace"], stage_frame["shadow_trace"]
    waves = -(-PASSES_PER_STEP // args.wave)
    n_trace_launch = n_shadow_launch = BOUNCES * waves
    st_one = p.ray_stats(reset=True)
    closest_bytes, shadow_bytes = algorithmic_bytes(trav)
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
    # capture of this configuration (profiles/r1_trace8_wave9_ncu.csv; bench.py cannot run under ncu itself)
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_trace8_wave9_ncu.csv")
    if os.path.exists(tpath) and args.wave == PASSES_PER_STEP and world == 1:
        import csv
        rows = [r for r in csv.DictReader(open(tpath)) if r["kernel"].startswith("k_trace8<0")]
        if rows:
            traffic = sum((float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e6 for r in rows) / len(rows)
            traffic_src = "profiles/r1_trace8_wave9_ncu.csv (ncu --set full, mean over the closest-hit launches of one frame)"

    # trav is one pass; scale node/triangle work to the 9 passes of a frame by the measured ray ratio of that frame
    scale_c = float(st_one["trace"].sum()) / max(trav["rays"][0], 1)
    ach_gbs = closest_bytes * scale_c / (trace_ms * 1e-3) / 1e9
    roofline = {"kernel": "k_trace8<closest-hit>", "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs, burst copy)", "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": closest_bytes * scale_c / n_trace_launch,
                "avg_launch_ms": trace_ms / n_trace_launch, "launches_timed": n_trace_launch,
                "algorithmic_bytes_per_ray": closest_bytes / max(trav["rays"][0], 1),
                "nodes_per_ray": trav["nodes"][0] / max(trav["rays"][0], 1), "triangles_per_ray": trav["triangles"][0] / max(trav["rays"][0], 1),
                "shadow_kernel": {"achieved": shadow_bytes * (float(st_one["shadow"].sum()) / max(trav["rays"][1], 1)) / (max(shadow_ms, 1e-9) * 1e-3) / 1e9,
                                  "algorithmic_bytes_per_ray": shadow_bytes / max(trav["rays"][1], 1)},
                "whole_pipeline_algorithmic_gbs": None}

    # ---- e2e: host buffers in, host frame out, every step
    tl_nodes = np.ascontiguousarray(np.asarray(blob["bvh_nodes"])[: int(blob["tlas_node_count"]) * 80])
    inst = [np.ascontiguousarray(blob[k]) for k in ("mesh_bvh_root_indices", "mesh_material_ids", "mesh_transforms", "mesh_transforms_inv", "mesh_transforms_prev")]
    pinned = [torch.from_numpy(a.copy()).pin_memory() for a in [tl_nodes] + inst]
    h2d = int(sum(a.numel() * a.element_size() for a in pinned)) + 60 + 44
    import ctypes
    lib = pt.lib()

    # The host<->device traffic of a step is pipelined behind the rendering of the next one (separate copy stream, two pinned host
    # frames): step k uploads its inputs, renders, and queues the device->host read of ITS frame; the read of step k-1 is awaited
    # at the end of step k.  Every step still pays its upload and its read-back inside the timed region.
    copy_stream = torch.cuda.Stream()
    host_frames = [host_frame, torch.empty_like(host_frame).pin_memory()]
    dev_stage = [torch.empty((HEIGHT, p.screen_pitch, 4), dtype=torch.float32, device="cuda") for _ in range(2)] if gather_mode != "p2p" else None
    frame_bytes = host_frame.numel() * 4
    reads = [None, None]
    step_no = [0]

    host_trace = [] if os.environ.get("PTB_BENCH_TRACE") else None

    def e2e_frame():
        k = step_no[0] & 1
        t_a = time.perf_counter()
        lib.ptb_update_instances(p._ctx, ctypes.c_void_p(pinned[0].data_ptr()), int(blob["tlas_node_count"]), int(inst[0].size),
                                 *[ctypes.c_void_p(x.data_ptr()) for x in pinned[1:]])
        p.invalidated_camera = True
        t_b = time.perf_counter()
        one_frame()
        t_c = time.perf_counter()
        if rank == 0:
            if gather_mode == "p2p":
                src = p.exchange_frame()                # the two exchange buffers alternate: frame k stays intact while k+1 renders
            else:                                       # single display / assembled buffer: park the frame in a staging copy first (device to device)
                src0 = frame.data_ptr() if gather_mode == "nccl" else p.display_device_ptr()[0]
                lib_rt.cudaMemcpyAsync(ctypes.c_void_p(dev_stage[k].data_ptr()), ctypes.c_void_p(src0), ctypes.c_size_t(frame_bytes), 3, ctypes.c_void_p(p.stream()))
                src = dev_stage[k].data_ptr()
            rendered = torch.cuda.Event()
            rendered.record(stream)
            copy_stream.wait_event(rendered)
            lib_rt.cudaMemcpyAsync(ctypes.c_void_p(host_frames[k].data_ptr()), ctypes.c_void_p(src), ctypes.c_size_t(frame_bytes), 2, ctypes.c_void_p(copy_stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(copy_stream)
            reads[k] = done
            t_d = time.perf_counter()
            if reads[k ^ 1] is not None:
                reads[k ^ 1].synchronize()              # the previous step's frame is now in host memory
            if host_trace is not None:
                host_trace.append((t_b - t_a, t_c - t_b, t_d - t_c, time.perf_counter() - t_d))
        step_no[0] += 1

    def e2e_drain():
        for ev in reads:
            if ev is not None:
                ev.synchronize()
        p.sync()

    lib_rt = ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12")) if os.path.exists(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12")) else ctypes.CDLL("libcudart.so")
    e2e_frame(); e2e_drain()
    p.ray_stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        e2e_frame()
    e2e_drain()
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    if host_trace:
        print("[bench] host ms per e2e step (update_instances, render_frame, queue read-back, wait previous read-back):", file=sys.stderr)
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
                "config": {"workload": workload, "passes_per_step": PASSES_PER_STEP, "parallelism": f"tiles{world}x(bands of {args.band_rows} rows)" if world > 1 else "1 GPU", "passes_per_wave": args.wave, "gather": gather_mode,
                           "static_merge": "identity-transform instances traced through one merged CWBVH (<= 1e-4 rel-L2 vs the reference, tests/test_gpu_parity.py)",
                           "l2": "working set (ray queues + AOVs > 400 MB per pass) exceeds the 126 MB L2", "rng_tables": blob.get("rng_source", "?"),
                           "ms_per_frame": ms_max / args.steps},
                "rays_per_step": int(rays_total / args.steps), "clocks": clocks, "gpu_launches": int(launches),
                "stage_ms_per_frame": stage_frame,
                "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(host_frame.numel() * 4), "ms_per_step": float(t.item()) / args.steps},
                "two_level_only": two_level, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(blob)
            line["cpu_bvh_build"] = cpu_bvh_build(blob)
        print(json.dumps(line), flush=True)
    # teardown order matters: tensors that were used on the ctx stream must die before the stream does
    del packed, gathered, frame, host_frame, host_frames, dev_stage, pinned, reads
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    torch.cuda.empty_cache()
    p.close()


if __name__ == "__main__":
    main()
