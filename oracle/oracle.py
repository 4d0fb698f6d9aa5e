"""ctypes binding of oracle/pt_oracle.c -- TEST INFRASTRUCTURE ONLY (see the header of pt_oracle.c).
Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the product package."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class PoScene(ctypes.Structure):
    _fields_ = [
        ("triangles", ctypes.c_void_p), ("bvh_nodes", ctypes.c_void_p), ("bvh_kind", ctypes.c_int),
        ("mesh_roots", ctypes.c_void_p), ("mesh_material_ids", ctypes.c_void_p),
        ("mesh_transforms", ctypes.c_void_p), ("mesh_transforms_inv", ctypes.c_void_p),
        ("material_types", ctypes.c_void_p), ("materials", ctypes.c_void_p),
        ("lights_total_weight", ctypes.c_float),
        ("light_triangle_indices", ctypes.c_void_p), ("light_triangle_cdf", ctypes.c_void_p),
        ("light_mesh_count", ctypes.c_int), ("light_mesh_cdf", ctypes.c_void_p),
        ("light_mesh_triangle_span", ctypes.c_void_p), ("light_mesh_transform_indices", ctypes.c_void_p),
        ("sky", ctypes.c_void_p), ("sky_width", ctypes.c_int), ("sky_height", ctypes.c_int), ("sky_scale", ctypes.c_float),
        ("pmj", ctypes.c_void_p), ("blue_noise", ctypes.c_void_p),
        ("camera", ctypes.c_float * 15),
        ("width", ctypes.c_int), ("height", ctypes.c_int), ("pitch", ctypes.c_int),
    ]


class PoConfig(ctypes.Structure):
    _fields_ = [("reconstruction_filter", ctypes.c_int), ("num_bounces", ctypes.c_int),
                ("enable_nee", ctypes.c_int), ("enable_mis", ctypes.c_int), ("enable_rr", ctypes.c_int)]


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libpt_oracle.so")
        src = os.path.join(_HERE, "pt_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            import subprocess
            subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", so, src, "-lm"])
        _lib = ctypes.CDLL(so)
        assert _lib.po_sizeof_scene() == ctypes.sizeof(PoScene)
    return _lib


def round_up(x, m):
    return (x + m - 1) // m * m


class Oracle:
    """CPU renderer over a scene blob. Framebuffers are [height, pitch, 4] float32."""

    def __init__(self, blob, num_bounces=None, nee=True, mis=True, rr=True, reconstruction_filter=2, threads=None):
        self.blob = blob
        self._keep = {k: np.ascontiguousarray(blob[k]) for k in (
            "triangles", "bvh_nodes", "mesh_bvh_root_indices", "mesh_material_ids", "mesh_transforms", "mesh_transforms_inv",
            "material_types", "materials", "light_triangle_indices", "light_triangle_cdf", "light_mesh_cdf",
            "light_mesh_triangle_span", "light_mesh_transform_indices", "sky", "pmj", "blue_noise")}
        k = self._keep
        s = PoScene()
        s.triangles = k["triangles"].ctypes.data; s.bvh_nodes = k["bvh_nodes"].ctypes.data; s.bvh_kind = int(blob["bvh_kind"])
        s.mesh_roots = k["mesh_bvh_root_indices"].ctypes.data; s.mesh_material_ids = k["mesh_material_ids"].ctypes.data
        s.mesh_transforms = k["mesh_transforms"].ctypes.data; s.mesh_transforms_inv = k["mesh_transforms_inv"].ctypes.data
        s.material_types = k["material_types"].ctypes.data; s.materials = k["materials"].ctypes.data
        s.lights_total_weight = float(blob["lights_total_weight"])
        s.light_triangle_indices = k["light_triangle_indices"].ctypes.data; s.light_triangle_cdf = k["light_triangle_cdf"].ctypes.data
        s.light_mesh_count = int(k["light_mesh_cdf"].size); s.light_mesh_cdf = k["light_mesh_cdf"].ctypes.data
        s.light_mesh_triangle_span = k["light_mesh_triangle_span"].ctypes.data
        s.light_mesh_transform_indices = k["light_mesh_transform_indices"].ctypes.data
        s.sky = k["sky"].ctypes.data; s.sky_height, s.sky_width = k["sky"].shape[:2]; s.sky_scale = float(blob["sky_scale"])
        s.pmj = k["pmj"].ctypes.data; s.blue_noise = k["blue_noise"].ctypes.data
        for i in range(15):
            s.camera[i] = float(blob["camera"][i])
        s.width, s.height = int(blob["width"]), int(blob["height"]); s.pitch = round_up(s.width, 32)
        self.scene = s
        self.cfg = PoConfig(int(reconstruction_filter), int(num_bounces or blob["num_bounces"]), int(nee), int(mis), int(rr))
        self.width, self.height, self.pitch = s.width, s.height, s.pitch
        if threads:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        self.counters = np.zeros(256, dtype=np.int64)

    def render_pass(self, sample_index, rows=None, aovs=("radiance",), primary_hits=False):
        y0, y1 = rows or (0, self.height)
        names = ("radiance", "direct", "indirect", "albedo", "normal", "position")
        fbs = {n: np.zeros((self.height, self.pitch, 4), dtype=np.float32) for n in names if n in aovs}
        hits = np.zeros((self.height, self.pitch, 4), dtype=np.uint32) if primary_hits else None
        ptr = lambda n: fbs[n].ctypes.data if n in fbs else None
        lib().po_render_pass(ctypes.byref(self.scene), ctypes.byref(self.cfg), int(sample_index), int(y0), int(y1),
                             *[ctypes.c_void_p(ptr(n)) for n in names],
                             ctypes.c_void_p(hits.ctypes.data if hits is not None else None), ctypes.c_void_p(self.counters.ctypes.data))
        return (fbs, hits) if primary_hits else fbs

    def render(self, passes, aovs=("radiance",)):
        """sample_index 0..passes: returns accumulators after the reference's online mean (pass 0 is overwritten)."""
        acc = None
        for si in range(passes + 1):
            fbs = self.render_pass(si, aovs=aovs)
            if acc is None:
                acc = {k: v.copy() for k, v in fbs.items()}
            else:
                for k in acc:
                    lib().po_accumulate(ctypes.c_void_p(acc[k].ctypes.data), ctypes.c_void_p(fbs[k].ctypes.data),
                                        ctypes.c_longlong(acc[k].size), ctypes.c_float(float(si)))
        return acc

    def primary_hits(self, sample_index=0):
        hits = np.zeros((self.height, self.pitch, 4), dtype=np.uint32)
        lib().po_primary_hits(ctypes.byref(self.scene), ctypes.byref(self.cfg), int(sample_index), ctypes.c_void_p(hits.ctypes.data))
        return hits

    def brute_force_primary(self, sample_index, mesh_tri_first, mesh_tri_count):
        hits = np.zeros((self.height, self.pitch, 4), dtype=np.uint32)
        a = np.ascontiguousarray(mesh_tri_first, dtype=np.int32); b = np.ascontiguousarray(mesh_tri_count, dtype=np.int32)
        lib().po_brute_force_primary(ctypes.byref(self.scene), ctypes.byref(self.cfg), int(sample_index), int(a.size),
                                     ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data), ctypes.c_void_p(hits.ctypes.data))
        return hits
