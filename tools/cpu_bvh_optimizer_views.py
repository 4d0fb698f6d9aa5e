import sys, os, time, ctypes
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo")
import numpy as np
import cpu_bvh_quality as q
from gpu_raytracer_b200 import scene
blob = scene.load_blob("/root/repo/data/_staged/" + (sys.argv[1] if len(sys.argv) > 1 else "sponza.npz"))
lib = q.load_lib()
lib.ptbh_set_optimizer.argtypes=[ctypes.c_int, ctypes.c_float, ctypes.c_int]
lib.ptbh_optimizer_sah.argtypes=[ctypes.c_void_p]
pos = q.identity_triangles(blob); n = pos.shape[0]
def rot_y(v, ang):
    c,s=np.cos(ang),np.sin(ang); R=np.array([[c,0,s],[0,1,0],[-s,0,c]]); return v@R.T
def view_rays(ang, shift, nrays, rng):
    cam = np.asarray(blob["camera"], dtype=np.float64).copy(); w, h = int(blob["width"]), int(blob["height"])
    x = rng.random(nrays)*w; y = rng.random(nrays)*h
    d = cam[3:6][None] + x[:,None]*cam[6:9][None] + y[:,None]*cam[9:12][None]
    d = rot_y(d, ang); d /= np.linalg.norm(d,axis=1,keepdims=True)
    o = np.repeat((cam[0:3]+np.array(shift))[None], nrays, 0)
    return np.ascontiguousarray(np.concatenate([o,d],1), dtype=np.float32)
rng = np.random.default_rng(1)
views = [("bench view",0.0,(0,0,0)),("rot 90",np.pi/2,(0,0,0)),("rot 180",np.pi,(0,0,0)),("rot 270 up",3*np.pi/2,(10,6,5))]
sets=[]
ref=None
for passes, frac in ((0,1.0),(1,1.0),(2,1.0),(3,0.5)):
    lib.ptbh_set_optimizer(passes, frac, 14)
    t0=time.time(); h = lib.ptbh_build_triangles_sbvh(pos.ctypes.data, n, 3e-4, 96, 2.0); dt=time.time()-t0
    sah=(ctypes.c_double*2)(); lib.ptbh_optimizer_sah(sah)
    nn, ni = lib.ptbh_node_count(h), lib.ptbh_index_count(h)
    nodes = np.empty(nn*80, np.uint8); idx = np.empty(ni, np.int32); lib.ptbh_export(h, nodes.ctypes.data, idx.ctypes.data, 0, 0)
    depth = lib.ptbh_max_depth(nodes.ctypes.data, 0)
    line=f"passes {passes} frac {frac}: build {dt:.1f}s sah {sah[0]:.1f}->{sah[1]:.1f} nodes8 {nn} depth {depth} |"
    if not sets:
        for name,ang,shift in views:
            pr = view_rays(ang, shift, 40000, rng); a = q.trace(lib, h, pos, pr); bn = q.bounce_rays(pos, pr, a[2], a[3], rng); sets.append((name,pr,bn))
    res=[]
    tot=0
    for name,pr,bn in sets:
        a = q.trace(lib, h, pos, pr); b = q.trace(lib, h, pos, bn)
        res.append((a[2],b[2]))
        cost = 0.66*(a[0]+b[0]) + 0.315*(a[1]+b[1]); tot+=cost
        line += f" {name}: prim n {a[0]:.2f} t {a[1]:.2f} bounce n {b[0]:.2f} t {b[1]:.2f} |"
    if ref is None: ref=res
    same = min(min(float(np.mean(r[0]==rr[0])), float(np.mean(r[1]==rr[1]))) for r,rr in zip(res,ref))
    print(line, f"cost {tot:.2f} same_t {same:.5f}", flush=True)
    lib.ptbh_free(h)
