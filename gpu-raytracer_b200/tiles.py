"""Screen-tile sharding across the GPUs of one box.

Rows are dealt to ranks in interleaved bands of `band_rows` rows (load balance: sky vs interior); every rank traces
only its rows but uses the GLOBAL pixel index, so a pixel's value does not depend on the partition.  After the last
pass each rank packs its rows (ptb_export_rows), one all-gather moves the packed tiles over NVLink, and
ptb_assemble_rows scatters them back into the full frame.  No reduction is needed: no pixel is shared.

The functions here are the host-side index logic (numpy mirrors of k_export_rows / k_assemble_rows, used by the CPU
tests with the gloo backend) plus the torch.distributed plumbing used by bench.py on NCCL.
"""
from __future__ import annotations

import numpy as np


def owner_of_row(y, world, band_rows):
    return (y // band_rows) % world


def owned_rows(height, rank, world, band_rows):
    return [y for y in range(height) if owner_of_row(y, world, band_rows) == rank]


def max_owned_rows(height, world, band_rows):
    return max(len(owned_rows(height, r, world, band_rows)) for r in range(world))


def local_row_to_y(local_row, rank, world, band_rows):
    band = local_row // band_rows
    return (band * world + rank) * band_rows + (local_row - band * band_rows)


def pack_rows(image, rank, world, band_rows, max_rows=None):
    """[height, pitch, C] full frame -> [max_rows, pitch, C] packed rows of `rank` (zero padded)."""
    rows = owned_rows(image.shape[0], rank, world, band_rows)
    out = np.zeros((max_rows or len(rows),) + image.shape[1:], dtype=image.dtype)
    out[:len(rows)] = image[rows]
    return out


def assemble_rows(packed, height, world, band_rows):
    """[world, max_rows, pitch, C] gathered tiles -> [height, pitch, C] full frame."""
    out = np.zeros((height,) + packed.shape[2:], dtype=packed.dtype)
    for r in range(world):
        rows = owned_rows(height, r, world, band_rows)
        out[rows] = packed[r, :len(rows)]
    return out


def gather_frame_torch(pathtracer, aov_type=0):
    """All ranks: pack owned rows -> all_gather_into_tensor (NCCL over NVLink) -> assemble. Returns a CUDA tensor
    [height, pitch, 4] on every rank.  Timed by the caller with CUDA events on the ctx stream."""
    import torch
    import torch.distributed as dist
    p = pathtracer
    world = dist.get_world_size()
    mx = max_owned_rows(p.screen_height, world, p.band_rows)
    st = torch.cuda.ExternalStream(p.stream())
    with torch.cuda.stream(st):
        packed = torch.zeros((mx, p.screen_pitch, 4), dtype=torch.float32, device="cuda")
        p.export_rows(packed.data_ptr(), aov_type)
        gathered = torch.empty((world * mx, p.screen_pitch, 4), dtype=torch.float32, device="cuda")   # rank-major concatenation
        dist.all_gather_into_tensor(gathered, packed)
        frame = torch.empty((p.screen_height, p.screen_pitch, 4), dtype=torch.float32, device="cuda")
        p.assemble_rows(gathered.data_ptr(), mx, frame.data_ptr())
    return frame
