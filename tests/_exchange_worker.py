"""Worker of test_peer_memory_exchange_between_processes: one rank per GPU, frame exchange over CUDA IPC (not collected by pytest)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gpu_raytracer_b200 import pathtracer as pt, scene  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = rank % torch.cuda.device_count()          # a 1-GPU box: both processes share the GPU (CUDA IPC works within a device too)
    torch.cuda.set_device(device)
    dist.init_process_group("gloo")
    d = scene.procedural_scene("atrium", seed=9, width=416, height=250, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=3)
    whole = pt.Pathtracer(blob, device=device, config=cfg)
    p = pt.Pathtracer(blob, device=device, rank=rank, world=world, band_rows=8, config=cfg)
    p.reserve_wave(5)
    _, handle = p.exchange_create()
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    p.exchange_connect_ipc(handles)
    rt = ctypes.CDLL("libcudart.so.12")
    for _ in range(3):
        whole.render_frame(4); whole.sync()
        want = whole.get_aov(0)
        p.render_frame(4); p.sync()
        got = np.empty_like(want)
        assert rt.cudaMemcpy(ctypes.c_void_p(got.ctypes.data), ctypes.c_void_p(p.exchange_frame()), ctypes.c_size_t(got.nbytes), 2) == 0
        assert np.array_equal(got[:, :416].view(np.uint32), want[:, :416].view(np.uint32))
    dist.barrier()
    p.close(); whole.close()
    print("EXCHANGE-OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
