"""Multi-GPU host logic on CPU: tile ownership math and the gather path with torch.distributed (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from gpu_raytracer_b200 import tiles


@pytest.mark.parametrize("height,world,band", [(1080, 8, 8), (1080, 3, 8), (33, 2, 8), (7, 4, 2), (2160, 8, 32)])
def test_row_ownership_is_a_partition(height, world, band):
    rows = [tiles.owned_rows(height, r, world, band) for r in range(world)]
    assert sorted(sum(rows, [])) == list(range(height))
    for r in range(world):
        assert [tiles.local_row_to_y(i, r, world, band) for i in range(len(rows[r]))] == rows[r]
    assert tiles.max_owned_rows(height, world, band) == max(len(x) for x in rows)
    assert max(len(x) for x in rows) - min(len(x) for x in rows) <= band     # balanced to within one band


def test_pack_assemble_roundtrip():
    rng = np.random.default_rng(0)
    img = rng.normal(size=(45, 64, 4)).astype(np.float32)
    world, band = 3, 8
    mx = tiles.max_owned_rows(45, world, band)
    packed = np.stack([tiles.pack_rows(img, r, world, band, mx) for r in range(world)])
    assert np.array_equal(tiles.assemble_rows(packed, 45, world, band), img)


def _worker(rank, world, port, height, pitch, band, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.arange(height * pitch * 4, dtype=np.float32).reshape(height, pitch, 4)
    mine = np.zeros_like(full)
    rows = tiles.owned_rows(height, rank, world, band)
    mine[rows] = full[rows]                                    # each rank only "rendered" its own rows
    mx = tiles.max_owned_rows(height, world, band)
    packed = torch.from_numpy(tiles.pack_rows(mine, rank, world, band, mx))
    gathered = torch.empty((world * mx,) + tuple(packed.shape[1:]), dtype=torch.float32)      # rank-major concatenation
    dist.all_gather_into_tensor(gathered, packed)
    frame = tiles.assemble_rows(gathered.numpy().reshape((world, mx) + tuple(packed.shape[1:])), height, world, band)
    ok = np.array_equal(frame, full)
    t = torch.tensor([1.0 if ok else 0.0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(t.item() == 1.0))
    dist.destroy_process_group()


def test_gather_with_gloo_world_2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 32, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
