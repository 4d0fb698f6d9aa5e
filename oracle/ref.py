"""ctypes binding of oracle/ref_harness.cpp (the reference's own kernels, compiled unmodified into
oracle/_ref/pathtracer_ref.cubin).  TEST / BASELINE INFRASTRUCTURE ONLY -- see the header of ref_harness.cpp."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(_HERE, "_ref", "pathtracer_ref.cubin")
# the same source with the documented one-function patch of oracle/patch_uniform_lut.py (warp-uniform LUT handles): the checker
# for the rough-dielectric BSDF, where the unmodified build is miscompiled by ptxas 12.9
CUBIN_UNIFORM = os.path.join(_HERE, "_ref", "pathtracer_ref_uniform.cubin")
_lib = None


def available():
    return os.path.exists(CUBIN)


CUBIN_AO = os.path.join(_HERE, "_ref", "ao_ref.cubin")      # Src/CUDA/AO.cu, unmodified: the reference's ambient-occlusion integrator


def ao_available():
    return os.path.exists(CUBIN_AO)


def uniform_available():
    return os.path.exists(CUBIN_UNIFORM)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libref_harness.so")
        src = os.path.join(_HERE, "ref_harness.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "harness"])
        _lib = ctypes.CDLL(so)
    return _lib


class Reference:
    """Same surface as gpu_raytracer_b200.pathtracer.Pathtracer (render_pass / get_aov / ray_stats ...)."""

    def __init__(self, blob, config=None, device=0, cubin=None):
        from gpu_raytracer_b200 import pathtracer as pt   # only for the plain ctypes struct definitions of include/ptb.h
        self._pt = pt
        l = lib()
        self.width, self.height = int(blob["width"]), int(blob["height"])
        self.pitch = (self.width + 31) // 32 * 32
        self._ctx = ctypes.c_void_p()
        self._ck(l.ref_create(ctypes.byref(self._ctx), (cubin or CUBIN).encode(), device, self.width, self.height), "ref_create")
        keep = []
        scene = pt.fill_scene_struct(blob, keep)
        self._ck(l.ref_upload_scene(self._ctx, ctypes.byref(scene)), "ref_upload_scene")
        self.config = config or pt.default_config(num_bounces=int(blob["num_bounces"]))
        self._ck(l.ref_set_config(self._ctx, ctypes.byref(self.config)), "ref_set_config")
        cam = pt.camera_struct(blob["camera"])
        if self.config.enable_svgf:
            cam.aperture_radius = 0.0
        vp = np.ascontiguousarray(blob["view_projection"], dtype=np.float32)
        self._ck(l.ref_set_camera(self._ctx, ctypes.byref(cam), ctypes.c_void_p(vp.ctypes.data), ctypes.c_void_p(vp.ctypes.data)), "ref_set_camera")

    @staticmethod
    def _ck(code, what):
        if code != 0:
            raise RuntimeError(f"{what} failed: {code}")

    def render_pass(self, sample_index):
        self._ck(lib().ref_render(self._ctx, int(sample_index)), "ref_render")

    def render_frames(self, passes):
        for si in range(passes + 1):
            self.render_pass(si)
        self.sync()

    def sync(self):
        self._ck(lib().ref_sync(self._ctx), "ref_sync")

    def get_aov(self, aov_type, accumulated=True):
        out = np.empty((self.height, self.pitch, 4), dtype=np.float32)
        self._ck(lib().ref_download(self._ctx, int(aov_type), int(accumulated), ctypes.c_void_p(out.ctypes.data)), "ref_download")
        return out

    def get_display(self):
        out = np.empty((self.height, self.pitch, 4), dtype=np.float32)
        self._ck(lib().ref_download(self._ctx, -1, 1, ctypes.c_void_p(out.ctypes.data)), "ref_download")
        return out

    def primary_hits(self):
        """Pixel-keyed [height, pitch, 4] uint32 table of the LAST batch's bounce-0 hits (0xffffffff where not covered)."""
        n = min(self.width * self.height, 1080 * 720)
        pix = np.empty(n, dtype=np.uint32); hits = np.empty((n, 4), dtype=np.uint32)
        self._ck(lib().ref_read_primary(self._ctx, ctypes.c_void_p(pix.ctypes.data), ctypes.c_void_p(hits.ctypes.data), n), "ref_read_primary")
        out = np.full((self.height * self.pitch, 4), 0xFFFFFFFF, dtype=np.uint32)
        out[pix & 0x3FFFFFFF] = hits
        return out.reshape(self.height, self.pitch, 4)

    def ray_stats(self, reset=False):
        st = self._pt.PtbRayStats()
        self._ck(lib().ref_get_ray_stats(self._ctx, ctypes.byref(st), int(reset)), "ref_get_ray_stats")
        return dict(trace=np.array(st.trace[:], dtype=np.uint64), shadow=np.array(st.shadow[:], dtype=np.uint64),
                    shaded=np.array(st.shaded[:], dtype=np.uint64), frames=int(st.frames))

    def svgf_buffer(self, name):
        if name == "history_length":
            out = np.empty((self.height, self.pitch), dtype=np.int32)
        else:
            out = np.empty((self.height, self.pitch, 4), dtype=np.float32)
        l = lib(); l.ref_read_global_buffer.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        self._ck(l.ref_read_global_buffer(self._ctx, name.encode(), ctypes.c_void_p(out.ctypes.data), out.nbytes), "ref_read_global_buffer")
        return out

    def begin_display_download(self, slot):
        """queue the pipelined device->host read of the displayed frame into pinned slot 0/1 (bench.py e2e leg)"""
        self._ck(lib().ref_e2e_begin_download(self._ctx, int(slot)), "ref_e2e_begin_download")

    def wait_display_download(self, slot):
        """block until slot's read-back has landed; returns it as a [height, pitch, 4] float32 view of the pinned buffer"""
        p = ctypes.c_void_p()
        self._ck(lib().ref_e2e_wait(self._ctx, int(slot), ctypes.byref(p)), "ref_e2e_wait")
        n = self.height * self.pitch * 4
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n,)).reshape(self.height, self.pitch, 4)

    def lut_contents(self):
        n = 2 * 16 ** 3 + 2 * 16 ** 2 + 32 ** 2 + 32
        out = np.empty(n, dtype=np.float32)
        self._ck(lib().ref_read_luts(self._ctx, ctypes.c_void_p(out.ctypes.data)), "ref_read_luts")
        return out

    def set_timing(self, on=True):
        lib().ref_set_timing(self._ctx, int(on))

    def stage_ms(self):
        ms = (ctypes.c_float * 6)()
        lib().ref_get_stage_ms(self._ctx, ms, 6)
        names = ("generate", "trace", "sort", "shade", "shadow_trace", "accumulate_or_svgf")
        return {n: float(ms[i]) for i, n in enumerate(names)}

    def launch_geometry(self):
        g = (ctypes.c_int * 8)()
        lib().ref_launch_geometry(self._ctx, g)
        return dict(trace_grid=g[0], trace_block=g[1], trace_smem=g[2], shadow_grid=g[3], shadow_block=g[4], shadow_smem=g[5], acc_block=(g[6], g[7]))

    def close(self):
        if self._ctx.value:
            lib().ref_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()


class ReferenceAO(Reference):
    """The reference's AO integrator (oracle/_ref/ao_ref.cubin) behind the same surface: render_pass / get_aov / ray_stats."""

    def __init__(self, blob, config=None, device=0, ao_radius=1.0):
        from gpu_raytracer_b200 import pathtracer as pt
        self._pt = pt
        l = lib()
        self.ao_radius = float(ao_radius)
        self.width, self.height = int(blob["width"]), int(blob["height"])
        self.pitch = (self.width + 31) // 32 * 32
        self._ctx = ctypes.c_void_p()
        self._ck(l.ref_create(ctypes.byref(self._ctx), CUBIN_AO.encode(), device, self.width, self.height), "ref_create")
        keep = []
        scene = pt.fill_scene_struct(blob, keep)
        self._ck(l.ref_ao_upload_scene(self._ctx, ctypes.byref(scene)), "ref_ao_upload_scene")
        self.config = config or pt.default_config(num_bounces=1)
        self._ck(l.ref_set_config(self._ctx, ctypes.byref(self.config)), "ref_set_config")
        cam = pt.camera_struct(blob["camera"])
        self._ck(l.ref_set_camera(self._ctx, ctypes.byref(cam), None, None), "ref_set_camera")

    def render_pass(self, sample_index):
        self._ck(lib().ref_ao_render(self._ctx, int(sample_index), ctypes.c_float(self.ao_radius)), "ref_ao_render")
