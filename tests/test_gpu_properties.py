"""Size-independent properties at BASELINE.json's full sizes + edge cases (GPU, through the C ABI)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from gpu_raytracer_b200 import pathtracer as pt, scene, tiles

pytestmark = pytest.mark.gpu


def full_size_blob():
    staged = os.path.join(ROOT, "data", "_staged", "sponza.npz")
    if os.path.exists(staged):
        return scene.load_blob(staged)
    return scene.build_blob(scene.procedural_scene("atrium", seed=7, width=1920, height=1080, detail=2.0), 8, 1920, 1080)


@pytest.fixture(scope="module")
def big():
    return full_size_blob()


def test_full_size_determinism_and_accumulation(big):
    """1080p, 4 bounces: two independent contexts give bit-identical frames; pass 0 does not survive in the accumulator;
    acc after passes 0..2 == mean of the framebuffers of passes 1 and 2 (online mean, AOV.h:35-46)."""
    cfg = pt.default_config(num_bounces=4)
    a = pt.Pathtracer(big, config=cfg); b = pt.Pathtracer(big, config=cfg)
    a.render_frames(2)
    b.render_pass(1); b.sync(); f1 = b.get_aov(0).astype(np.float32)       # sample_index 1 on a fresh accumulator: acc = fb
    c = pt.Pathtracer(big, config=cfg); c.render_pass(0); c.render_pass(1); c.sync()
    assert np.allclose(c.get_aov(0), f1, rtol=1e-5, atol=1e-7)                # pass 0 overwritten: acc + (fb - acc) / 1 == fb up to rounding
    b2 = pt.Pathtracer(big, config=cfg); b2.render_frames(2)
    assert np.array_equal(a.get_aov(0).view(np.uint32), b2.get_aov(0).view(np.uint32))
    st = a.ray_stats()
    assert st["trace"][0] == 3 * 1920 * 1080 and st["trace"][4] == 0
    assert (st["trace"][1:4] <= st["trace"][0:3]).all() and (st["shadow"][:4] <= st["trace"][:4]).all()
    for p in (a, b, c, b2):
        p.close()


def test_tile_sharding_is_partition_invariant(big):
    """world=3 ranks (all on this GPU) each trace their row bands; packed+assembled frame == the 1-GPU frame bit for bit."""
    import torch
    cfg = pt.default_config(num_bounces=3)
    whole = pt.Pathtracer(big, config=cfg); whole.render_frames(1)
    ref_img = whole.get_aov(0)
    world, band = 3, 8
    mx = tiles.max_owned_rows(1080, world, band)
    gathered = torch.zeros((world, mx, whole.screen_pitch, 4), dtype=torch.float32, device="cuda")
    frame = torch.zeros((1080, whole.screen_pitch, 4), dtype=torch.float32, device="cuda")
    rays = 0
    for r in range(world):
        p = pt.Pathtracer(big, rank=r, world=world, band_rows=band, config=cfg)
        p.render_frames(1)
        rows = p.export_rows(gathered[r].data_ptr()); p.sync()
        assert rows == len(tiles.owned_rows(1080, r, world, band))
        rays += int(p.ray_stats()["trace"].sum())
        if r == world - 1:
            p.assemble_rows(gathered.data_ptr(), mx, frame.data_ptr()); p.sync()
        p.close()
    out = frame.cpu().numpy()
    assert np.array_equal(out[:, :1920].view(np.uint32), ref_img[:, :1920].view(np.uint32))
    assert rays == int(whole.ray_stats()["trace"].sum())
    # numpy mirror of the assemble kernel agrees with the kernel
    assert np.array_equal(tiles.assemble_rows(gathered.cpu().numpy(), 1080, world, band)[:, :1920], out[:, :1920])
    whole.close()


def test_cuda_graph_frame_replay_is_identical():
    """ptb_render_frame (one CUDA graph per frame) == the same passes launched one by one; replaying the cached graph again
    gives the same frame; changing the camera drops the cached graph (new image), same camera keeps it."""
    d = scene.procedural_scene("atrium", seed=5, width=320, height=200, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=4)
    a = pt.Pathtracer(blob, config=cfg); a.render_frames(4)
    b = pt.Pathtracer(blob, config=cfg); b.render_frame(4); b.sync()
    assert np.array_equal(a.get_aov(0).view(np.uint32), b.get_aov(0).view(np.uint32))
    assert np.array_equal(a.ray_stats()["trace"], b.ray_stats()["trace"]) and a.launch_count() == b.launch_count()
    first = b.get_aov(0).copy()
    b.render_frame(4); b.sync()
    assert np.array_equal(first.view(np.uint32), b.get_aov(0).view(np.uint32))
    cam = np.array(blob["camera"]); cam[0] += 0.5
    b.set_camera(cam); b.render_frame(4); b.sync()
    assert not np.array_equal(first, b.get_aov(0))
    b.set_camera(blob["camera"]); b.render_frame(4); b.sync()
    assert np.array_equal(first.view(np.uint32), b.get_aov(0).view(np.uint32))
    a.close(); b.close()


@pytest.mark.parametrize("wave,world", [(5, 1), (3, 1), (5, 2)])
def test_multi_pass_waves_are_bit_identical_to_pass_by_pass(wave, world):
    """ptb_reserve_wave: several passes traced together (per-ray pass slot, per-slot framebuffer planes, ordered fold) give the
    same accumulators, display image and ray counts as tracing the passes one after another -- also under tile sharding."""
    d = scene.procedural_scene("atrium", seed=6, width=256, height=160, detail=0.5, all_materials=False)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=4, aov_mask=0x3F)
    base = pt.Pathtracer(blob, config=cfg); base.render_frames(4)
    out = np.zeros_like(base.get_aov(0)); rays = 0
    outs = {k: np.zeros_like(out) for k in (0, 3, 4, 5)}
    for rank in range(world):
        q = pt.Pathtracer(blob, rank=rank, world=world, band_rows=8, config=cfg)
        q.reserve_wave(wave)
        q.render_frame(4); q.sync()
        rows = [y for y in range(160) if (y // 8) % world == rank]
        for k in outs:
            outs[k][rows] = q.get_aov(k)[rows]
        rays += int(q.ray_stats()["trace"].sum() + q.ray_stats()["shadow"].sum())
        if world == 1:
            assert np.array_equal(q.get_display().view(np.uint32), base.get_display().view(np.uint32))
            assert not q.get_aov(0, accumulated=False).any()            # every framebuffer plane was cleared by the fold
        q.close()
    for k in outs:
        assert np.array_equal(outs[k].view(np.uint32), base.get_aov(k).view(np.uint32)), pt.AOV_NAMES[k]
    assert rays == int(base.ray_stats()["trace"].sum() + base.ray_stats()["shadow"].sum())
    base.close()


def test_present_and_capture(tmp_path):
    """ptb_present: the device-side tone mapping (post.frag: ACES + gamma, 8-bit RGBA) agrees with the numpy restatement to one
    code value; exporters.capture writes what the reference's capture writes (Main.cpp:195-249): tone-mapped .ppm, linear .exr and
    the AOV .exr files, which read back as the HALF-rounded frame."""
    from gpu_raytracer_b200 import exporters as ex
    d = scene.procedural_scene("cornell", seed=2, width=160, height=120)
    blob = scene.build_blob(d, 8, rng="fallback")
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=3, aov_mask=0x3F)); p.render_frames(4)
    w, h = 160, 120
    ldr = p.present()[:h, :w]
    frame = p.get_display()[:h, :w, :3]
    want = np.rint(255.0 * ex.tonemap_aces(frame).astype(np.float64)).astype(np.int32)
    assert (ldr[..., 3] == 255).all() and np.abs(ldr[..., :3].astype(np.int32) - want).max() <= 1
    assert ldr[..., :3].max() > 40                                  # something was rendered
    files = ex.capture(p, str(tmp_path / "shot.exr"))
    assert [os.path.basename(f) for f in files] == ["shot.exr", "albedo.exr", "normal.exr", "position.exr"]
    assert np.array_equal(ex.load_exr(files[0]), frame.astype(np.float16).astype(np.float32))
    assert np.array_equal(ex.load_exr(files[2]), p.get_aov(pt.AOV_NORMAL)[:h, :w, :3].astype(np.float16).astype(np.float32))
    ex.capture(p, str(tmp_path / "shot.ppm"))
    assert np.abs(ex.load_ppm(str(tmp_path / "shot.ppm")).astype(np.int32) - ldr[..., :3].astype(np.int32)).max() <= 1
    p.close()


def test_edge_cases():
    # width not a multiple of 32 (pitch padding), one bounce, no lights, a single triangle
    d = scene.procedural_scene("soup", seed=2, width=70, height=33, detail=0.1)
    blob = scene.build_blob(d, 8, rng="fallback")
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=1)); p.render_frames(1)
    img = p.get_aov(0)
    assert img.shape == (33, 96, 4) and np.isfinite(img).all() and not img[:, 70:].any()
    assert p.ray_stats()["shadow"].sum() == 0       # the last bounce terminates in the sort pass
    p.close()
    d = scene.SceneDesc(); d.width, d.height, d.num_bounces = 64, 48, 3
    m = d.add_material(scene.Material(scene.MAT_DIFFUSE, diffuse=(0.5, 0.5, 0.5)))
    tri = scene.finish_triangles([[[-1, -1, -3], [1, -1, -3], [0, 1, -3]]], [[[0, 0, 1]] * 3], [[[0, 0], [1, 0], [0, 1]]])
    d.instances.append(scene.Instance(d.add_mesh_data(tri), m))
    blob = scene.build_blob(d, 8, rng="fallback")
    assert blob["light_mesh_cdf"].size == 0
    p = pt.Pathtracer(blob); p.render_frames(2)
    st = p.ray_stats()
    assert st["shadow"].sum() == 0 and st["trace"][0] == 3 * 64 * 48 and 0 < st["shaded"][0] < st["trace"][0]
    assert np.isfinite(p.get_aov(0)).all()
    p.close()


def test_errors_are_reported_not_swallowed():
    import ctypes
    lib = pt.lib()
    ctx = ctypes.c_void_p()
    assert lib.ptb_create(ctypes.byref(ctx), 0, 64, 64, 0, 1, 8) == 0
    assert lib.ptb_render(ctx, 0) == -2                      # no scene uploaded
    bad = pt.default_config(num_bounces=0)
    assert lib.ptb_set_config(ctx, ctypes.byref(bad)) == -1
    lib.ptb_destroy(ctx)


def test_peer_memory_frame_exchange_is_the_whole_frame():
    """ptb_exchange_*: world=3 ranks (three ctxs on this GPU, blocks connected by raw pointers) render a frame each; the last
    accumulate kernel stores every rank's rows into every rank's frame, so each rank's exchange frame == the 1-GPU frame bit for
    bit -- with no export / all_gather / assemble.  Two frames (both buffer parities), second one with a moved camera."""
    import ctypes
    d = scene.procedural_scene("atrium", seed=9, width=416, height=250, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=3)
    world = 3
    whole = pt.Pathtracer(blob, config=cfg)
    ranks = [pt.Pathtracer(blob, rank=r, world=world, band_rows=8, config=cfg) for r in range(world)]
    for p in ranks:
        p.reserve_wave(5)
    bases = [p.exchange_create()[0] for p in ranks]
    for p in ranks:
        p.exchange_connect(bases)
    cam2 = np.array(blob["camera"]); cam2[0] += 0.25
    rt = ctypes.CDLL("libcudart.so.12")
    for frame_no, cam in enumerate([blob["camera"], cam2, blob["camera"]]):
        whole.set_camera(cam); whole.render_frame(4); whole.sync()
        want = whole.get_aov(0)
        for p in ranks:                       # enqueue all three before waiting on any: each frame ends with a device-side wait for its peers
            p.set_camera(cam); p.render_frame(4)
        for p in ranks:
            p.sync()
            got = np.empty_like(want)
            assert rt.cudaMemcpy(ctypes.c_void_p(got.ctypes.data), ctypes.c_void_p(p.exchange_frame()), ctypes.c_size_t(got.nbytes), 2) == 0
            assert np.array_equal(got[:, :416].view(np.uint32), want[:, :416].view(np.uint32)), (frame_no, p.rank)
    # after a disconnect a rank renders on its own again
    ranks[0].exchange_disconnect(); ranks[0].render_frame(4); ranks[0].sync()
    for p in ranks:
        p.close()
    whole.close()


def test_peer_memory_exchange_times_out_instead_of_hanging():
    """A peer that never delivers must not hang the GPU: the device-side wait gives up after 4 s and ptb_sync reports it."""
    d = scene.procedural_scene("soup", seed=2, width=96, height=64, detail=0.3)
    blob = scene.build_blob(d, 8, rng="fallback")
    a = pt.Pathtracer(blob, rank=0, world=2, band_rows=8)
    b = pt.Pathtracer(blob, rank=1, world=2, band_rows=8)
    bases = [a.exchange_create()[0], b.exchange_create()[0]]
    a.exchange_connect(bases); b.exchange_connect(bases)
    a.render_frame(1)                         # b never renders
    with pytest.raises(RuntimeError, match="exchange"):
        a.sync()
    a.close(); b.close()


def test_peer_memory_exchange_between_processes():
    """Two processes (one GPU each when the box has two, otherwise sharing GPU 0), blocks connected through CUDA IPC handles:
    both end up with the 1-GPU frame.  This is the data plane bench.py --gpus N rides on."""
    import subprocess, sys
    worker = os.path.join(os.path.dirname(__file__), "_exchange_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", worker], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("EXCHANGE-OK") == 2, r.stdout[-2000:]


def _hits_and_image(blob, cfg, merge):
    one = pt.Pathtracer(blob, config=pt.default_config(num_bounces=1)); one.set_static_merge(merge)
    one.render_frames(1); hits = one.primary_hits(); one.close()
    p = pt.Pathtracer(blob, config=cfg); p.set_static_merge(merge)
    p.render_frames(2)
    out = hits, p.get_aov(0), p.ray_stats()
    p.close()
    return out


def test_static_merge_is_invisible(big):
    """ptb_set_static_merge: tracing the identity-transform instances through ONE merged CWBVH (default) gives the same
    primary-hit table (mesh_id, triangle_id, t, uv) and the same image as the reference's TLAS -> BLAS walk, bit for bit --
    on the full-size scene and on a mixed scene with rotated / scaled instances, all four BSDFs and NEE."""
    cfg = pt.default_config(num_bounces=4)
    mixed = scene.build_blob(scene.procedural_scene("soup", seed=11, width=320, height=200, detail=0.6, all_materials=True), 8, rng="fallback")
    for blob, w in ((big, 1920), (mixed, 320)):
        ha, ia, sa = _hits_and_image(blob, cfg, True)
        hb, ib, sb = _hits_and_image(blob, cfg, False)
        hit = ha[..., 1] != 0xFFFFFFFF
        assert hit.any() and np.array_equal(ha[..., 1], hb[..., 1])
        assert np.array_equal(ha[hit], hb[hit])
        assert np.array_equal(ia[:, :w].view(np.uint32), ib[:, :w].view(np.uint32))
        assert np.array_equal(sa["trace"], sb["trace"]) and np.array_equal(sa["shadow"], sb["shadow"])


def test_static_merge_follows_instance_updates():
    """ptb_update_instances re-sending the same tables (per-frame fast path) and with an identity instance turning into a
    moving one (its slot in the merged BVH is retired -- no rebuild, no stall -- and the instance is traced through the TLAS
    again): the image stays equal to the un-merged one."""
    import ctypes
    d = scene.procedural_scene("soup", seed=4, width=256, height=160, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=3)
    M = len(blob["mesh_material_ids"])
    ident = [i for i in range(M) if int(np.asarray(blob["mesh_bvh_root_indices"]).view(np.uint32)[i]) >> 31]
    assert len(ident) >= 2

    def permuted(blob, move=None):
        xf = np.array(blob["mesh_transforms"]).copy(); xi = np.array(blob["mesh_transforms_inv"]).copy()
        roots = np.array(blob["mesh_bvh_root_indices"]).copy()
        if move is not None:
            xf[move, 3] += 0.37; xi[move, 3] -= 0.37
            roots[move] = roots[move] & 0x7FFFFFFF
        return roots, xf, xi

    for move in (None, ident[0]):
        roots, xf, xi = permuted(blob, move)
        outs = []
        for merge in (True, False):
            p = pt.Pathtracer(blob, config=cfg); p.set_static_merge(merge)
            p.render_frames(1)                                      # something is cached before the update
            tl = np.ascontiguousarray(np.asarray(blob["bvh_nodes"])[: int(blob["tlas_node_count"]) * 80])
            mats = np.ascontiguousarray(blob["mesh_material_ids"])
            pt._check(pt.lib().ptb_update_instances(p._ctx, ctypes.c_void_p(tl.ctypes.data), int(blob["tlas_node_count"]), M,
                                                    ctypes.c_void_p(roots.ctypes.data), ctypes.c_void_p(mats.ctypes.data),
                                                    ctypes.c_void_p(xf.ctypes.data), ctypes.c_void_p(xi.ctypes.data), ctypes.c_void_p(xf.ctypes.data)), "update")
            p.sample_index = 0; p.invalidated_gpu_config = True
            p.render_frames(2)
            outs.append(p.get_aov(0))
            p.close()
        assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


@pytest.mark.parametrize("bvh_kind,merge", [(8, 1), (8, 0), (2, 0)])
def test_tlas_refit_on_device_equals_host_rebuilt_tlas(bvh_kind, merge):
    """ptb_refit_instances: instances move (translations, rotations, an identity instance starting to move, a lamp), the
    TLAS keeps its topology and is refitted by k_refit_tlas on the device.  The frame must equal the one of a context that was given
    the moved scene with a TLAS rebuilt on the host -- any valid BVH yields the same closest hits; only the order in which
    equal-distance hits are met may differ, hence the tie allowance."""
    import copy
    d = scene.procedural_scene("atrium", seed=5, width=320, height=192, detail=0.5)
    blob = scene.build_blob(d, bvh_kind, rng="fallback")
    moved = copy.deepcopy(d)
    for i, inst in enumerate(moved.instances):
        if inst.name == "column" and i % 2:
            inst.position = inst.position + np.array([0.6, 0.0, -0.4]); inst.rotation = scene.q_axis_angle((0, 1, 0), 0.2 * i)
        elif inst.name == "curtain":
            inst.position = inst.position + np.array([0.0, 0.5, 0.0])
        elif inst.name == "floor":
            inst.position = inst.position + np.array([0.0, -0.15, 0.0])          # identity -> moving: leaves the merged static BVH
        elif inst.name == "lamp" and i % 2:
            inst.position = inst.position + np.array([1.0, -0.5, 0.5])
    rebuilt = scene.build_blob(moved, bvh_kind, rng="fallback")
    # NEE off: with next-event estimation the light is picked by its position in the light table, which follows the TLAS leaf order --
    # a rebuilt TLAS may list the two lamps the other way round and then the same random number picks the other lamp (equally valid,
    # different noise).  Without it the frame depends on the geometry alone.
    cfg = pt.default_config(num_bounces=3, enable_next_event_estimation=0, aov_mask=0x39)
    want = pt.Pathtracer(rebuilt, config=cfg); want.set_static_merge(merge)
    want.render_frames(3)
    ref_img = want.get_aov(0); ref_aovs = [want.get_aov(k) for k in (pt.AOV_ALBEDO, pt.AOV_NORMAL, pt.AOV_POSITION)]
    want.close()

    p = pt.Pathtracer(blob, config=cfg); p.set_static_merge(merge)
    p.render_frames(2)                                                           # graphs, merged tree and pruned TLAS exist before the move
    xf, xi = scene.instance_transforms(moved, blob["instance_order"])
    p.refit_instances(xf, xi)
    if bvh_kind == 8:
        # the refitted TLAS itself, read back: every instance inside the slot box that references it, every internal slot box around
        # everything below it (overhang 0), and about as tight as a TLAS built on the host for the moved scene
        p.sync()
        order = np.asarray(blob["instance_order"])
        boxes = []
        for j, i in enumerate(order):
            inst = moved.instances[i]; pts = moved.mesh_datas[inst.mesh_data][0].reshape(-1, 3).astype(np.float64)
            T = xf[j].reshape(3, 4).astype(np.float64); lo, hi = pts.min(0), pts.max(0); c, e = 0.5 * (lo + hi), 0.5 * (hi - lo)
            nc = T[:, :3] @ c + T[:, 3]; ne = np.abs(T[:, :3]) @ e
            boxes.append(np.concatenate([nc - ne, nc + ne]))
        reached, overhang, slack = scene.check_tlas8(p.tlas_nodes(int(blob["tlas_node_count"])), np.array(boxes))
        assert reached == len(order) and overhang <= 1e-5 and slack < 0.2, (reached, overhang, slack)
    p.invalidated_gpu_config = True
    p.render_frames(3)
    got = p.get_aov(0); got_aovs = [p.get_aov(k) for k in (pt.AOV_ALBEDO, pt.AOV_NORMAL, pt.AOV_POSITION)]
    # a second refit back to the start must restore the original picture (the blanked / retired bookkeeping survives round trips)
    xf0, xi0 = scene.instance_transforms(d, blob["instance_order"])
    p.refit_instances(xf0, xi0); p.invalidated_gpu_config = True
    p.render_frames(3)
    back = p.get_aov(0)
    p.close()
    fresh = pt.Pathtracer(blob, config=cfg); fresh.set_static_merge(merge); fresh.render_frames(3); orig = fresh.get_aov(0); fresh.close()

    def close(a, b):
        w = a.shape[1]
        differ = np.any(a.view(np.uint32) != b.view(np.uint32), axis=-1).mean()
        err = np.linalg.norm((a - b).astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30)
        return differ, err
    differ, err = close(got, ref_img)
    assert differ <= 0.01 and err <= 1e-5, (differ, err)
    for a, b in zip(got_aovs, ref_aovs):
        differ, err = close(a, b)
        assert differ <= 0.01 and err <= 1e-5, (differ, err)
    differ, err = close(back, orig)
    assert differ <= 0.01 and err <= 1e-5, (differ, err)
    assert np.abs(got - orig).max() > 0.05                                       # the move is visible: the refit did something


def _exchange_frame_host(p, like):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    got = np.empty_like(like)
    assert rt.cudaMemcpy(ctypes.c_void_p(got.ctypes.data), ctypes.c_void_p(p.exchange_frame()), ctypes.c_size_t(got.nbytes), 2) == 0
    return got


@pytest.mark.parametrize("width,height,world", [(320, 184, 3), (256, 400, 2)])
def test_svgf_taa_sharded_equals_single_gpu(width, height, world):
    """BASELINE configs[4] shape (SVGF + TAA, tile-sharded).  Tracing is sharded by interleaved 8-row bands, FILTERING by contiguous
    blocks of rows: every rank stores the rows it traced into the input planes of the ranks that filter them (block + 72-row halo),
    filters only its block, reads last frame's history across block borders from the owning rank over peer memory, and ships its
    displayed rows to everybody.  The gathered frame on every rank and the temporal buffers of every rank's block stay bit-identical
    to the 1-GPU run over 6 frames with a moving camera (reprojection crosses block borders).  Second case: blocks taller than the
    halo (the production shape)."""
    d = scene.procedural_scene("atrium", seed=3, width=width, height=height, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=3, enable_svgf=1, enable_taa=1)
    whole = pt.Pathtracer(blob, config=cfg)
    ranks = [pt.Pathtracer(blob, rank=r, world=world, band_rows=8, config=cfg) for r in range(world)]
    bases = [p.exchange_create()[0] for p in ranks]
    for p in ranks:
        p.exchange_connect(bases)
    cam = np.array(blob["camera"])
    for p in [whole] + ranks:
        p.update()                     # uploads the config (allocations synchronise the device: keep them out of the frame loop,
                                       # where a rank's frame ends in a device-side wait for peers that share this GPU)
    for frame in range(6):
        if frame in (2, 3, 4):
            cam = cam.copy(); cam[0] += 0.05; cam[1] += 0.03 * (frame - 2)          # sideways and up / down: history taps move across rows
            for p in [whole] + ranks:
                p.set_camera(cam)
        if frame > 0:
            for p in [whole] + ranks:
                p.update()
        whole.render(); whole.sync()
        for p in ranks:
            p.render()
        for p in ranks:
            p.sync()
        want = whole.get_display()
        for p in ranks:
            got = _exchange_frame_host(p, want)
            assert np.array_equal(got[:, :width].view(np.uint32), want[:, :width].view(np.uint32)), (frame, p.rank)
    rows_per_block = -(-height // world)
    for name in ("history_direct", "history_indirect", "history_moment", "history_normal_and_depth", "history_length", "taa_frame_prev"):
        want = whole.svgf_buffer(name)
        for p in ranks:
            y0 = p.rank * rows_per_block; y1 = height if p.rank == world - 1 else min(height, (p.rank + 1) * rows_per_block)
            assert np.array_equal(p.svgf_buffer(name)[y0:y1, :width], want[y0:y1, :width]), (name, p.rank)
    for p in ranks:
        p.close()
    whole.close()


@pytest.mark.parametrize("bins", [8, 64])
def test_ray_ordering_is_invisible(bins):
    """ptb_set_ray_ordering: tracing secondary / shadow rays in direction-bin order is a scheduling choice -- same image, same
    ray counts, in wave mode and pass by pass."""
    d = scene.procedural_scene("atrium", seed=12, width=352, height=208, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    cfg = pt.default_config(num_bounces=4)
    a = pt.Pathtracer(blob, config=cfg); a.reserve_wave(5); a.render_frame(4); a.sync()
    for wave in (5, 1):
        b = pt.Pathtracer(blob, config=cfg); b.reserve_wave(wave); b.set_ray_ordering(bins); b.render_frame(4); b.sync()
        assert np.array_equal(a.get_aov(0).view(np.uint32), b.get_aov(0).view(np.uint32))
        assert np.array_equal(a.ray_stats()["trace"], b.ray_stats()["trace"]) and np.array_equal(a.ray_stats()["shadow"], b.ray_stats()["shadow"])
        b.close()
    a.close()


def test_pixel_query_reports_the_primary_hit():
    """set_pixel_query (Integrator.h:266-277 / Pathtracer.cu:345-348): after the next pass the query holds the (mesh_id,
    triangle_id) of that pixel's primary hit -- the same ids as the primary-hit table -- and (-1, -1) for a sky pixel; reading
    it clears the query."""
    d = scene.procedural_scene("soup", seed=6, width=160, height=96, detail=0.5)
    blob = scene.build_blob(d, 8, rng="fallback")
    p = pt.Pathtracer(blob, config=pt.default_config(num_bounces=1))
    p.render_pass(1); p.sync()
    hits = p.primary_hits()[:96, :160]
    hit_px = np.argwhere(hits[..., 1] != 0xFFFFFFFF); sky_px = np.argwhere(hits[..., 1] == 0xFFFFFFFF)
    assert len(hit_px) and len(sky_px)
    y, x = (int(v) for v in hit_px[len(hit_px) // 2])
    p.set_pixel_query(x, y); p.render_pass(1)
    assert p.get_pixel_query() == (int(hits[y, x, 0]), int(hits[y, x, 1]))
    p.render_pass(1)
    assert p.get_pixel_query() == (-1, -1)                     # consumed: nothing pending any more
    y, x = (int(v) for v in sky_px[0])
    p.set_pixel_query(x, y); p.render_pass(1)
    assert p.get_pixel_query() == (-1, -1)
    p.close()
