"""CPU-side tuning aid for the merged static BVH: builds it with the plain SAH builder and with the split-BVH builder (host/bvh_build.cpp) and
counts node visits / triangle tests per ray with ptbh_trace_stats on primary rays of the scene's camera and on cosine-distributed bounce
rays from their hit points (what the device's bounce >= 1 launches trace).  No GPU involved."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_raytracer_b200 import scene

def load_lib(path=None):
    lib = ctypes.CDLL(path or scene.hostlib()._name)
    lib.ptbh_build_triangles.restype = ctypes.c_void_p
    lib.ptbh_build_triangles.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
    lib.ptbh_build_triangles_sbvh.restype = ctypes.c_void_p
    lib.ptbh_build_triangles_sbvh.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float]
    lib.ptbh_trace_stats.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 3
    lib.ptbh_node_count.argtypes = [ctypes.c_void_p]; lib.ptbh_index_count.argtypes = [ctypes.c_void_p]
    lib.ptbh_max_depth.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    lib.ptbh_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.ptbh_free.argtypes = [ctypes.c_void_p]
    return lib

def identity_triangles(blob):
    roots = np.asarray(blob["mesh_bvh_root_indices"]).view(np.uint32)
    first, count = np.asarray(blob["mesh_tri_first"]), np.asarray(blob["mesh_tri_count"])
    tri = np.asarray(blob["triangles"])
    out = []
    for i in np.nonzero(roots >> 31)[0]:
        t = tri[first[i]:first[i] + count[i]]
        p0 = t[:, 0:3]; out.append(np.stack([p0, p0 + t[:, 3:6], p0 + t[:, 6:9]], 1))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)

def primary_rays(blob, n, rng):
    cam = np.asarray(blob["camera"], dtype=np.float64); w, h = int(blob["width"]), int(blob["height"])
    x = rng.random(n) * w; y = rng.random(n) * h
    d = cam[3:6][None] + x[:, None] * cam[6:9][None] + y[:, None] * cam[9:12][None]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.repeat(cam[0:3][None], n, 0)
    return np.ascontiguousarray(np.concatenate([o, d], 1), dtype=np.float32)

def trace(lib, h, pos, rays):
    n = rays.shape[0]
    counts = (ctypes.c_ulonglong * 2)(0, 0); t = np.empty(n, np.float32); tri = np.empty(n, np.int32)
    lib.ptbh_trace_stats(h, pos.ctypes.data, rays.ctypes.data, n, counts, t.ctypes.data, tri.ctypes.data)
    return counts[0] / n, counts[1] / n, t, tri

def bounce_rays(pos, rays, t, tri, rng):
    ok = tri >= 0
    r = rays[ok]; tt = t[ok]; tr = pos[tri[ok]]
    p = r[:, :3] + r[:, 3:] * tt[:, None]
    nrm = np.cross(tr[:, 1] - tr[:, 0], tr[:, 2] - tr[:, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-30
    nrm *= -np.sign((nrm * r[:, 3:]).sum(1, keepdims=True))
    u1, u2 = rng.random(len(p)), rng.random(len(p))
    rr, phi = np.sqrt(u1), 2 * np.pi * u2
    a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0, 1, 0]]), np.array([[1, 0, 0]]))
    tx = np.cross(nrm, a); tx /= np.linalg.norm(tx, axis=1, keepdims=True); ty = np.cross(nrm, tx)
    d = tx * (rr * np.cos(phi))[:, None] + ty * (rr * np.sin(phi))[:, None] + nrm * np.sqrt(1 - u1)[:, None]
    return np.ascontiguousarray(np.concatenate([p + 1e-3 * nrm, d], 1), dtype=np.float32)

def main():
    blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", sys.argv[1] if len(sys.argv) > 1 else "sponza.npz"))
    lib = load_lib(os.environ.get("PTB_HOST_LIB"))
    pos = identity_triangles(blob); n = pos.shape[0]
    rng = np.random.default_rng(1)
    prim = primary_rays(blob, 60000, rng)
    variants = [("sah", None)] + [(f"sbvh a={a:g} bins={b}", (a, b, 2.0)) for a, b in ((1e-3, 64), (3e-4, 96))]
    bounce = None
    for name, par in variants:
        t0 = time.time()
        h = lib.ptbh_build_triangles(pos.ctypes.data, n, 8, 4.0, 1.0) if par is None else lib.ptbh_build_triangles_sbvh(pos.ctypes.data, n, par[0], par[1], par[2])
        dt = time.time() - t0
        nn, ni = lib.ptbh_node_count(h), lib.ptbh_index_count(h)
        nodes = np.empty(nn * 80, np.uint8); idx = np.empty(ni, np.int32)
        lib.ptbh_export(h, nodes.ctypes.data, idx.ctypes.data, 0, 0)
        depth = lib.ptbh_max_depth(nodes.ctypes.data, 0)
        a = trace(lib, h, pos, prim)
        if bounce is None:
            bounce = bounce_rays(pos, prim, a[2], a[3], rng); ref_t = a[2]
        b = trace(lib, h, pos, bounce)
        if par is None:
            ref_bt = b[2]
        same = float(np.mean(a[2] == ref_t)), float(np.mean(b[2] == ref_bt))
        print(f"{name:24s} build {dt:6.2f}s nodes {nn:7d} refs {ni:7d} ({ni / n:4.2f}x) depth {depth:2d} | primary nodes/ray {a[0]:6.2f} tris/ray {a[1]:5.2f} | bounce nodes/ray {b[0]:6.2f} tris/ray {b[1]:5.2f} | same t {same[0]:.5f} {same[1]:.5f}", flush=True)
        lib.ptbh_free(h)

if __name__ == "__main__":
    main()
