"""Host BVH builders (SAH sweep, CWBVH conversion, BVH2 leaf collapse): structural invariants and traversal-vs-brute-force."""
import numpy as np
import pytest

from gpu_raytracer_b200 import scene
from oracle.oracle import Oracle


def decode_nodes8(raw):
    n = raw.reshape(-1, 80)
    p = n[:, 0:12].copy().view(np.float32)
    e = n[:, 12:15]; imask = n[:, 15]
    base_child = n[:, 16:20].copy().view(np.uint32)[:, 0]; base_tri = n[:, 20:24].copy().view(np.uint32)[:, 0]
    meta = n[:, 24:32]
    q = n[:, 32:80].reshape(-1, 6, 8)   # qlo_x, qhi_x, qlo_y, qhi_y, qlo_z, qhi_z
    return p, e, imask, base_child, base_tri, meta, q


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000])
def test_cwbvh_references_every_triangle_once_and_boxes_contain_children(n):
    rng = np.random.default_rng(n)
    tri = rng.uniform(-1, 1, size=(n, 3, 3)).astype(np.float32)
    b = scene.build_blas(tri, 8)
    nodes, idx = b.export()
    assert sorted(idx.tolist()) == list(range(n))
    p, e, imask, base_child, base_tri, meta, q = decode_nodes8(nodes)
    seen = np.zeros(n, dtype=int)
    tri_sorted = tri[idx]
    for ni in range(b.node_count):
        scale = (e[ni].astype(np.uint32) << 23).view(np.float32)
        rel = 0
        for s in range(8):
            m = int(meta[ni, s])
            if m == 0:
                continue
            lo = p[ni] + scale * q[ni, 0::2, s]; hi = p[ni] + scale * q[ni, 1::2, s]
            if (m & 0x1F) >= 24:      # internal
                assert imask[ni] & (1 << s) and (m & 0x1F) == 24 + s and (m >> 5) == 1
                child = base_child[ni] + rel; rel += 1
                assert child < b.node_count
            else:
                cnt = bin(m >> 5).count("1"); off = m & 0x1F
                assert 1 <= cnt <= 3 and off + cnt <= 24
                ts = tri_sorted[base_tri[ni] + off: base_tri[ni] + off + cnt]
                seen[base_tri[ni] + off: base_tri[ni] + off + cnt] += 1
                eps = 1e-4 * (1 + np.abs(ts).max())
                assert (ts.min((0, 1)) >= lo - eps).all() and (ts.max((0, 1)) <= hi + eps).all()
        assert rel == bin(int(imask[ni])).count("1")
    assert (seen == 1).all()


def test_bvh2_collapse_keeps_all_triangles_and_sane_nodes():
    rng = np.random.default_rng(3)
    tri = rng.uniform(-1, 1, size=(500, 3, 3)).astype(np.float32)
    b = scene.build_blas(tri, 2)
    nodes, idx = b.export()
    assert sorted(idx.tolist()) == list(range(500))
    n = nodes.reshape(-1, 32)
    first = n[:, 24:28].copy().view(np.int32)[:, 0]; ca = n[:, 28:32].copy().view(np.uint32)[:, 0]
    count = ca & 0x3FFFFFFF
    leaves = [i for i in range(b.node_count) if i != 1 and count[i] > 0]
    assert sum(int(count[i]) for i in leaves) == 500
    raw = scene.BuiltBVH(scene.hostlib().ptbh_build_triangles(np.ascontiguousarray(tri.reshape(-1, 9)).ctypes.data, 500, 2, 4.0, 0.0))
    assert raw.node_count >= b.node_count      # collapsing never adds nodes


@pytest.mark.parametrize("kind,bvh", [("soup", 8), ("soup", 2), ("cornell", 8), ("cornell", 2), ("atrium", 8)])
def test_traversal_matches_brute_force(kind, bvh):
    """TLAS/BLAS traversal over the built BVH finds exactly the closest hit an exhaustive instance x triangle loop finds."""
    d = scene.procedural_scene(kind, seed=4, width=64, height=40, detail=0.2)
    blob = scene.build_blob(d, bvh, rng="fallback")
    o = Oracle(blob)
    h = o.primary_hits(1)[:, :64]
    bf = o.brute_force_primary(1, blob["mesh_tri_first"], blob["mesh_tri_count"])[:, :64]
    assert np.array_equal(h[..., 2], bf[..., 2])                       # identical t bits
    differ = h[..., 1] != bf[..., 1]
    assert differ.mean() < 1e-3                                        # ids may differ only on exact ties (shared edges)


def test_tlas_slots_and_instance_tables():
    d = scene.procedural_scene("atrium", seed=1, width=32, height=32, detail=0.2)
    blob = scene.build_blob(d, 8, rng="fallback")
    m = len(d.instances)
    assert blob["tlas_node_count"] <= 2 * m
    roots = blob["mesh_bvh_root_indices"].view(np.uint32)
    assert ((roots & 0x7FFFFFFF) >= 2 * m).all()                       # BLAS live after the TLAS slots
    ident = (roots >> 31).astype(bool)
    order = blob["instance_order"]
    assert [d.instances[i].identity() for i in order] == ident.tolist()
    t = blob["mesh_transforms"].reshape(-1, 3, 4); ti = blob["mesh_transforms_inv"].reshape(-1, 3, 4)
    for a, b in zip(t, ti):
        A = np.vstack([a, [0, 0, 0, 1]]); B = np.vstack([b, [0, 0, 0, 1]])
        assert np.allclose(A @ B, np.eye(4), atol=1e-4)


# ---------------------------------------------------------------------------------------------- static-merge host helpers
def _merge_lib():
    import ctypes
    lib = scene.hostlib()
    lib.ptbh_collect_leaf_primitives.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int]
    lib.ptbh_prune_tlas.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    lib.ptbh_bfs_relayout.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


def _collect(lib, nodes, root, cap):
    out = np.empty(cap, dtype=np.int32)
    n = lib.ptbh_collect_leaf_primitives(nodes.ctypes.data, int(root), out.ctypes.data, cap)
    assert n <= cap
    return out[:n]


def test_static_merge_collects_each_instance_triangles():
    """host/static_merge.h::collect_leaf_primitives walks a BLAS of the uploaded node array and must return exactly the
    triangle range of that mesh (what rebuild_static_merge feeds to the merged SAH build)."""
    lib = _merge_lib()
    blob = scene.build_blob(scene.procedural_scene("soup", seed=5, width=64, height=64, detail=0.5), 8, rng="fallback")
    nodes = np.ascontiguousarray(blob["bvh_nodes"])
    roots = np.asarray(blob["mesh_bvh_root_indices"]).view(np.uint32)
    T = len(blob["triangles"])
    for i, root in enumerate(roots):
        got = _collect(lib, nodes, root & 0x3FFFFFFF, T)
        first, count = int(blob["mesh_tri_first"][i]), int(blob["mesh_tri_count"][i])
        assert sorted(got.tolist()) == list(range(first, first + count))
    # the TLAS lists every instance exactly once
    inst = _collect(lib, nodes, 0, 4 * len(roots))
    assert sorted(inst.tolist()) == list(range(len(roots)))


def test_static_merge_bfs_relayout_preserves_the_tree():
    """bfs_relayout: same leaves, root first, children contiguous, and the first nodes are the top levels (depth never decreases)."""
    lib = _merge_lib()
    rng = np.random.default_rng(3)
    tris = (rng.random((4000, 3, 3)) * 10.0).astype(np.float32)
    tris[:, 1:] = tris[:, :1] + (rng.random((4000, 2, 3)).astype(np.float32) - 0.5) * 0.3
    b = scene.build_blas(tris, 8)
    dfs, idx = b.export(0, 0)
    bfs = np.zeros_like(dfs)
    base = 1000
    lib.ptbh_bfs_relayout(dfs.ctypes.data, b.node_count, base, bfs.ctypes.data)
    # walk the relaid tree (indices are global: shift the array so that node `base` is at offset base)
    padded = np.concatenate([np.zeros(base * 80, dtype=np.uint8), bfs])
    want = sorted(_collect(lib, dfs, 0, 3 * len(tris)).tolist())
    got = sorted(_collect(lib, padded, base, 3 * len(tris)).tolist())
    assert got == want == list(range(len(tris)))
    # breadth-first: depth of node i is non-decreasing in i
    n = bfs.reshape(-1, 80)
    depth = np.zeros(b.node_count, dtype=np.int32)
    for i in range(b.node_count):
        kids = bin(int(n[i, 15])).count("1")
        first = int(n[i, 16:20].view(np.uint32)[0]) - base
        assert kids == 0 or (i < first and first + kids <= b.node_count)
        depth[first:first + kids] = depth[i] + 1
    assert (np.diff(depth) >= 0).all() and depth[0] == 0


def test_static_merge_prunes_the_tlas():
    """prune_tlas blanks every slot that only leads to merged instances: walking the pruned TLAS reaches all un-merged instances,
    and merged ones only where they share a leaf with an un-merged one; merging everything empties the root."""
    lib = _merge_lib()
    rng = np.random.default_rng(1)
    lo = (rng.random((97, 3)) * 50).astype(np.float32)
    boxes = np.concatenate([lo, lo + 1.0 + rng.random((97, 3)).astype(np.float32)], axis=1)
    tl = scene.build_tlas(boxes, 8)
    nodes, order = tl.export(0, 0)
    M = len(boxes)
    merged = (rng.random(M) < 0.8)
    pruned = nodes.copy()
    all_gone = lib.ptbh_prune_tlas(pruned.ctypes.data, merged.astype(np.int8).tobytes(), M)
    assert all_gone == 0
    reach = set(_collect(lib, pruned, 0, 4 * M).tolist())
    assert set(np.where(~merged)[0].tolist()) <= reach
    assert len(reach) < M                                            # something was actually pruned
    # a merged instance is only still reachable through a leaf it shares with an un-merged one: its leaf group has an un-merged member
    full = _collect(lib, nodes, 0, 4 * M).tolist()
    assert sorted(full) == list(range(M))
    everything = nodes.copy()
    assert lib.ptbh_prune_tlas(everything.ctypes.data, np.ones(M, dtype=np.int8).tobytes(), M) == 1
    assert len(_collect(lib, everything, 0, 4 * M)) == 0


def test_static_merge_depth_bound():
    """max_depth (the guard that keeps the merged tree inside the 32-entry traversal stack): a single node is depth 1, the depth
    of a built tree is what a recursive walk finds, and the real scenes stay far below the bound."""
    import ctypes
    lib = _merge_lib()
    lib.ptbh_max_depth.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    rng = np.random.default_rng(8)
    one = (rng.random((2, 3, 3))).astype(np.float32)
    b1 = scene.build_blas(one, 8); n1, _ = b1.export(0, 0)
    assert lib.ptbh_max_depth(n1.ctypes.data, 0) == 1
    tris = (rng.random((20000, 3, 3)) * 20.0).astype(np.float32)
    tris[:, 1:] = tris[:, :1] + (rng.random((20000, 2, 3)).astype(np.float32) - 0.5) * 0.2
    b = scene.build_blas(tris, 8); nd, _ = b.export(0, 0)
    n = nd.reshape(-1, 80)

    def walk(i):
        kids = bin(int(n[i, 15])).count("1"); base = int(n[i, 16:20].view(np.uint32)[0])
        return 1 + max([walk(base + k) for k in range(kids)], default=0)
    d = lib.ptbh_max_depth(nd.ctypes.data, 0)
    assert d == walk(0) and 3 <= d and 2 * d + 3 <= 32


# ---------------------------------------------------------------------------------------------- split BVH (merged static tree)
def _long_triangles(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-2, 2, size=(n, 1, 3))
    tri = c + (rng.random((n, 3, 3)) - 0.5) * np.array([1.6, 0.08, 0.5])      # long, thin, heavily overlapping boxes: the SBVH case
    flat = rng.integers(0, n, size=n // 8)
    tri[flat, :, 1] = tri[flat, :1, 1]                                            # axis-aligned flat ones (zero-thickness boxes)
    return np.ascontiguousarray(tri, dtype=np.float32)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (40, 2), (3000, 3)])
def test_sbvh_finds_the_same_closest_hits_as_the_sah_tree(n, seed):
    """Spatial splits duplicate references and clip their boxes; the closest hit of every ray must stay what the plain SAH tree
    (verified against brute force above) finds -- same t bits, same triangle -- and every triangle must be referenced."""
    tri = _long_triangles(n, seed)
    rng = np.random.default_rng(seed + 10)
    o = rng.uniform(-3, 3, size=(4000, 3)); d = rng.normal(size=(4000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1).astype(np.float32)
    sah = scene.build_blas(tri, 8); sb = scene.build_blas_sbvh(tri, alpha=1e-5, bins=32)
    _, idx = sb.export()
    assert set(idx.tolist()) == set(range(n)) and sb.index_count >= n
    a = scene.trace_stats(sah, tri, rays); b = scene.trace_stats(sb, tri, rays)
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    hit = np.isfinite(a[2])
    assert not hit.any() or (a[3][hit] != b[3][hit]).mean() < 2e-3
    if n >= 3000:
        assert hit.mean() > 0.2
        assert b[0] < a[0] and b[1] < a[1]                                        # and it is the better tree for this geometry
        assert sb.index_count > n                                                 # ... because it actually split something


def test_sbvh_children_stay_inside_their_parents():
    """Flat leaves are widened (Box::fatten); the refit keeps every child box inside the box its quantisation grid spans --
    a child sticking out would wrap around in the 8-bit grid."""
    tri = _long_triangles(500, 5)
    sb = scene.build_blas_sbvh(tri, alpha=1e-6, bins=32)
    nodes, idx = sb.export()
    p, e, imask, base_child, base_tri, meta, q = decode_nodes8(nodes)
    for ni in range(sb.node_count):
        for slot in range(8):
            if meta[ni, slot] == 0:
                continue
            assert (q[ni, 0::2, slot] <= q[ni, 1::2, slot]).all()                # lo <= hi on every axis
    # every referenced triangle lies inside the dequantised box of the leaf slot that references it (clipped to that box's slab)
    rays = np.concatenate([tri.mean(1) + np.array([0, 0, 5.0]), np.tile([0, 0, -1.0], (len(tri), 1))], 1).astype(np.float32)
    _, _, t, hit_tri = scene.trace_stats(sb, tri, rays)
    _, _, t_ref, _ = scene.trace_stats(scene.build_blas(tri, 8), tri, rays)
    assert np.array_equal(t.view(np.uint32), t_ref.view(np.uint32)) and np.isfinite(t).mean() > 0.9     # a ray through each centroid


@pytest.mark.parametrize("n,seed,passes,fraction", [(2, 1, 1, 1.0), (7, 2, 2, 1.0), (3000, 3, 1, 1.0), (3000, 4, 3, 0.3)])
def test_reinsertion_optimizer_keeps_hits_and_never_raises_the_sah_cost(n, seed, passes, fraction):
    """The insertion-based optimiser (host/bvh_build.cpp ReinsertionOptimizer) re-links subtrees of the binary split BVH before the wide
    collapse: every ray must keep its closest hit bit for bit, every triangle must stay referenced, child boxes must stay inside their
    parents' grids, the tree's SAH cost may only fall."""
    tri = _long_triangles(n, seed)
    rng = np.random.default_rng(seed + 20)
    o = rng.uniform(-3, 3, size=(4000, 3)); d = rng.normal(size=(4000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1).astype(np.float32)
    base = scene.build_blas_sbvh(tri, alpha=1e-5, bins=32)
    opt = scene.build_blas_sbvh(tri, alpha=1e-5, bins=32, optimize_passes=passes, optimize_fraction=fraction)
    before, after = scene.optimizer_sah()
    assert after <= before * (1 + 1e-6)
    _, idx = opt.export()
    assert set(idx.tolist()) == set(range(n)) and opt.index_count == base.index_count
    a = scene.trace_stats(base, tri, rays); b = scene.trace_stats(opt, tri, rays)
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    nodes, _ = opt.export()
    p, e, imask, base_child, base_tri, meta, q = decode_nodes8(nodes)
    for ni in range(opt.node_count):
        for slot in range(8):
            if meta[ni, slot]:
                assert (q[ni, 0::2, slot] <= q[ni, 1::2, slot]).all()
    if n >= 3000:
        assert after < before and b[0] < 1.03 * a[0]                  # random soups gain little (2-3 %); Sponza's merged tree 8 % SAH, 9-20 % node visits
    # the depth guard: an impossible bound keeps the tree as built
    same = scene.build_blas_sbvh(tri, alpha=1e-5, bins=32, optimize_passes=passes, optimize_fraction=fraction, max_depth=1 if n >= 3000 else 0)
    if n >= 3000:
        n0, i0 = base.export(); n1, i1 = same.export()
        assert np.array_equal(n0, n1) and np.array_equal(i0, i1)


def test_bvh4_conversion_is_a_valid_tree():
    """QuadConverter (Src/BVH/Converters/BVH4Converter.cpp): starting at (node 1, slot 0) every primitive is reached exactly once, every
    slot box contains what hangs below it, unused slots trail, and children were adopted (more than two used slots per node on average)."""
    rng = np.random.default_rng(11)
    n = 3000
    c = rng.uniform(-5, 5, (n, 1, 3)).astype(np.float32)
    tri = (c + rng.normal(0, 0.15, (n, 3, 3))).astype(np.float32)
    for leaf_cost in (1.0, 0.0):
        b = scene.build_blas(tri, 4, 4.0, leaf_cost)
        assert b.kind == 4 and b.node_bytes == 128
        raw, idx = b.export(0, 0)
        nodes = raw.view(np.float32).reshape(-1, 32)
        ic = raw.view(np.int32).reshape(-1, 32)[:, 24:].reshape(-1, 4, 2)
        assert np.array_equal(np.sort(idx), np.arange(n))
        assert ic[1, 0, 0] == 0 and ic[1, 0, 1] == 0                  # the entry slot points at node 0
        seen = np.zeros(n, dtype=np.int32); used = []

        def walk(node, slot):
            index, count = int(ic[node, slot, 0]), int(ic[node, slot, 1])
            lo = np.array([np.inf] * 3); hi = -lo
            if count > 0:
                for t in range(index, index + count):
                    seen[idx[t]] += 1
                    lo = np.minimum(lo, tri[idx[t]].min(0)); hi = np.maximum(hi, tri[idx[t]].max(0))
            else:
                k = 0
                while k < 4 and ic[index, k, 1] != -1:
                    k += 1
                assert k >= 1 and all(ic[index, j, 1] == -1 for j in range(k, 4))
                used.append(k)
                for j in range(k):
                    clo, chi = walk(index, j)
                    box_lo = nodes[index, [j, 4 + j, 8 + j]]; box_hi = nodes[index, [12 + j, 16 + j, 20 + j]]
                    assert np.all(box_lo <= clo + 1e-6) and np.all(box_hi >= chi - 1e-6)
                    lo = np.minimum(lo, clo); hi = np.maximum(hi, chi)
            return lo, hi

        import sys
        sys.setrecursionlimit(10000)
        walk(1, 0)
        assert np.all(seen == 1)
        assert np.mean(used) > 2.5
    # offsets as the aggregated arrays need them (Integrator.cpp:216-246)
    raw2, _ = b.export(100, 5000)
    ic2 = raw2.view(np.int32).reshape(-1, 32)[:, 24:].reshape(-1, 4, 2)
    inner = ic[:, :, 1] == 0; leaf = ic[:, :, 1] > 0
    assert np.array_equal(ic2[:, :, 0][inner], ic[:, :, 0][inner] + 100) and np.array_equal(ic2[:, :, 0][leaf], ic[:, :, 0][leaf] + 5000)
