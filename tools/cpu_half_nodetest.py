"""CPU study of a conservative half-precision (packed half2) CWBVH node test: emulates the fp16 arithmetic exactly (HFMA2 = one
rounding of an exactly representable product-sum, numpy float16 round-to-nearest-even) on real (ray, node) visits of the merged Sponza
tree and checks (a) it never rejects a child the float test accepts, (b) how many extra children it accepts."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import cpu_bvh_quality as q
from gpu_raytracer_b200 import scene

def h16(x):            # round to half, return as float64 (inf on overflow like the hardware)
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)

def main():
    lib = q.load_lib(os.environ.get("PTB_HOST_LIB"))
    lib.ptbh_set_visit_sink.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_longlong]
    lib.ptbh_visit_count.restype = ctypes.c_longlong
    blob = scene.load_blob(os.path.join(ROOT, "data", "_staged", "sponza.npz"))
    pos = q.identity_triangles(blob); n = len(pos)
    rng = np.random.default_rng(2)
    h = lib.ptbh_build_triangles_sbvh(pos.ctypes.data, n, 3e-4, 96, 2.0)
    nn, ni = lib.ptbh_node_count(h), lib.ptbh_index_count(h)
    nodes = np.empty(nn * 80, np.uint8); idx = np.empty(ni, np.int32)
    lib.ptbh_export(h, nodes.ctypes.data, idx.ctypes.data, 0, 0)
    nd = nodes.reshape(nn, 80)
    prim = q.primary_rays(blob, 3000, rng)
    a = q.trace(lib, h, pos, prim)
    rays = np.concatenate([prim, q.bounce_rays(pos, prim, a[2], a[3], rng)])
    cap = 400000
    vr = np.empty(cap, np.int32); vn = np.empty(cap, np.int32); vt = np.empty(cap, np.float32)
    lib.ptbh_set_visit_sink(vr.ctypes.data, vn.ctypes.data, vt.ctypes.data, cap)
    q.trace(lib, h, pos, rays)
    m = int(lib.ptbh_visit_count()); lib.ptbh_set_visit_sink(None, None, None, 0)
    vr, vn, vt = vr[:m], vn[:m], vt[:m]
    print("visits", m)
    N = nd[vn]
    p = N[:, 0:12].copy().view(np.float32).astype(np.float32)
    e = N[:, 12:15].astype(np.uint32)
    meta = N[:, 24:32]
    qq = N[:, 32:80].reshape(m, 6, 8).astype(np.float32)     # qlo_x, qhi_x, qlo_y, qhi_y, qlo_z, qhi_z
    o = rays[vr, 0:3]; d = rays[vr, 3:6]
    scale = (e << 23).view(np.float32)
    inv = (scale / d).astype(np.float32)
    org = ((p - o) / d).astype(np.float32)
    neg = d < 0
    qn = np.where(neg[:, :, None], qq[:, 1::2], qq[:, 0::2]); qf = np.where(neg[:, :, None], qq[:, 0::2], qq[:, 1::2])   # [m, 3, 8]
    # ---- float test (the device's: fma in f32, min/max, strict <)
    tn = (qn.astype(np.float64) * inv[:, :, None] + org[:, :, None]).astype(np.float32)
    tf = (qf.astype(np.float64) * inv[:, :, None] + org[:, :, None]).astype(np.float32)
    tmin = np.maximum(np.maximum(tn[:, 0], tn[:, 1]), np.maximum(tn[:, 2], 0.0))
    tmax = np.minimum(np.minimum(tf[:, 0], tf[:, 1]), np.minimum(tf[:, 2], vt[:, None]))
    valid = meta != 0
    acc32 = (tmin < tmax) & valid
    # ---- half test, time origin shifted to the node's own entry
    span = 255.0 * inv.astype(np.float64)
    entry = np.where(neg, org + span, org)                       # slab entry of the quantisation frame per axis
    t0 = np.maximum(entry.max(1), 0.0)
    t0 = np.where(np.isfinite(t0), t0, 0.0).astype(np.float32).astype(np.float64)
    orgs = org.astype(np.float64) - t0[:, None]
    M = 2.0 ** -10 * (np.abs(orgs) + 255.0 * np.abs(inv))
    inv_h = h16(inv); lo_h = h16(orgs - M); hi_h = h16(orgs + M)
    big = (np.abs(orgs) + 255.0 * np.abs(inv)) > 30000.0        # axis out of half range: ignore its slab
    tn_h = h16(qn * inv_h[:, :, None] + lo_h[:, :, None]); tf_h = h16(qf * inv_h[:, :, None] + hi_h[:, :, None])
    tn_h = np.where(big[:, :, None], -60000.0, tn_h); tf_h = np.where(big[:, :, None], 60000.0, tf_h)
    zero_h = h16(np.minimum(-t0 * (1 - 2.0 ** -10) - 0.0, 0.0)); zero_h = np.minimum(zero_h, h16(-t0))          # <= (0 - t0)
    with np.errstate(invalid="ignore"):
        thit = (vt.astype(np.float64) - t0)
        thit_h = np.where(np.isfinite(thit), h16(np.minimum(thit * (1 + 2.0 ** -9) + 2.0 ** -9 * np.abs(thit) + 1e-6, 60000.0)), 65504.0)
    tmin_h = np.maximum(np.maximum(tn_h[:, 0], tn_h[:, 1]), np.maximum(tn_h[:, 2], zero_h[:, None]))
    tmax_h = np.minimum(np.minimum(tf_h[:, 0], tf_h[:, 1]), np.minimum(tf_h[:, 2], thit_h[:, None]))
    acc16 = (tmin_h <= tmax_h) & valid
    missed = acc32 & ~acc16
    print("float accepts", int(acc32.sum()), "half accepts", int(acc16.sum()), "extra", int((acc16 & ~acc32).sum()), f"(+{100.0 * (acc16 & ~acc32).sum() / acc32.sum():.2f} %)", "MISSED", int(missed.sum()))
    print("axes out of half range:", float(big.mean()))
    if missed.any():
        i, j = np.argwhere(missed)[0]
        print("first miss: visit", i, "child", j, "tmin", tmin[i, j], "tmax", tmax[i, j], "tmin_h", tmin_h[i, j], "tmax_h", tmax_h[i, j], "t0", t0[i], "inv", inv[i], "org", org[i], "best", vt[i])

if __name__ == "__main__":
    main()
