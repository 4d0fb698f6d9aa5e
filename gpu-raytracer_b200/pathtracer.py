"""Host façade over the C ABI of include/ptb.h, shaped like the reference's Integrator / Pathtracer
(Src/Renderer/Integrators/Integrator.h:56-296, Pathtracer.h:146-286): same entry points (`update`, `render`),
same public state (`sample_index`, `invalidated_*`, `screen_width/height/pitch`, `get_aov`), same semantics:

  * `update()` advances `sample_index` exactly like Integrator::update (Integrator.cpp:518-526): any invalidation
    resets it to 0, otherwise it increments; `render()` then traces ONE pass with that sample index.
  * with SVGF off the accumulator holds the running mean of passes 1..sample_index (pass 0 is overwritten, AOV.h:35-46).

There is no CPU path: constructing a Pathtracer without the CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

AOV_RADIANCE, AOV_RADIANCE_DIRECT, AOV_RADIANCE_INDIRECT, AOV_ALBEDO, AOV_NORMAL, AOV_POSITION = range(6)
AOV_NAMES = ("radiance", "direct", "indirect", "albedo", "normal", "position")


class PtbConfig(ctypes.Structure):
    """GPUConfig, Src/CUDA/Common.h:39-67 (44 bytes)."""
    _fields_ = [("reconstruction_filter", ctypes.c_int32), ("aov_mask", ctypes.c_uint32), ("num_bounces", ctypes.c_int32),
                ("enable_mipmapping", ctypes.c_uint8), ("enable_next_event_estimation", ctypes.c_uint8),
                ("enable_multiple_importance_sampling", ctypes.c_uint8), ("enable_russian_roulette", ctypes.c_uint8),
                ("enable_svgf", ctypes.c_uint8), ("enable_spatial_variance", ctypes.c_uint8), ("enable_taa", ctypes.c_uint8), ("pad_", ctypes.c_uint8),
                ("alpha_colour", ctypes.c_float), ("alpha_moment", ctypes.c_float), ("num_atrous_iterations", ctypes.c_int32),
                ("sigma_z", ctypes.c_float), ("sigma_n", ctypes.c_float), ("sigma_l", ctypes.c_float)]


class PtbCamera(ctypes.Structure):
    _fields_ = [("position", ctypes.c_float * 3), ("bottom_left_corner", ctypes.c_float * 3), ("x_axis", ctypes.c_float * 3),
                ("y_axis", ctypes.c_float * 3), ("pixel_spread_angle", ctypes.c_float), ("aperture_radius", ctypes.c_float),
                ("focal_distance", ctypes.c_float)]


class PtbTexture(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int32), ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("num_levels", ctypes.c_int32),
                ("levels", ctypes.POINTER(ctypes.c_void_p)), ("lod_bias", ctypes.c_float)]


class PtbScene(ctypes.Structure):
    _fields_ = [("triangles", ctypes.c_void_p), ("triangle_count", ctypes.c_int32),
                ("bvh_nodes", ctypes.c_void_p), ("bvh_node_count", ctypes.c_int32), ("bvh_kind", ctypes.c_int32), ("tlas_node_count", ctypes.c_int32),
                ("mesh_count", ctypes.c_int32), ("mesh_bvh_root_indices", ctypes.c_void_p), ("mesh_material_ids", ctypes.c_void_p),
                ("mesh_transforms", ctypes.c_void_p), ("mesh_transforms_inv", ctypes.c_void_p), ("mesh_transforms_prev", ctypes.c_void_p),
                ("material_count", ctypes.c_int32), ("material_types", ctypes.c_void_p), ("materials", ctypes.c_void_p),
                ("medium_count", ctypes.c_int32), ("media", ctypes.c_void_p),
                ("texture_count", ctypes.c_int32), ("textures", ctypes.POINTER(PtbTexture)),
                ("sky", ctypes.c_void_p), ("sky_width", ctypes.c_int32), ("sky_height", ctypes.c_int32), ("sky_scale", ctypes.c_float),
                ("pmj_samples", ctypes.c_void_p), ("blue_noise", ctypes.c_void_p),
                ("lights_total_weight", ctypes.c_float), ("light_triangle_count", ctypes.c_int32),
                ("light_triangle_indices", ctypes.c_void_p), ("light_triangle_cumulative_probability", ctypes.c_void_p),
                ("light_mesh_count", ctypes.c_int32), ("light_mesh_cumulative_probability", ctypes.c_void_p),
                ("light_mesh_triangle_span", ctypes.c_void_p), ("light_mesh_transform_indices", ctypes.c_void_p)]


class PtbRayStats(ctypes.Structure):
    _fields_ = [("trace", ctypes.c_uint64 * 128), ("shadow", ctypes.c_uint64 * 128), ("shaded", ctypes.c_uint64 * 4), ("frames", ctypes.c_uint64)]


class PtbTraversalStats(ctypes.Structure):
    _fields_ = [("rays", ctypes.c_uint64 * 2), ("nodes", ctypes.c_uint64 * 2), ("triangles", ctypes.c_uint64 * 2),
                ("instance_transforms", ctypes.c_uint64 * 2), ("shadow_misses", ctypes.c_uint64)]


# every symbol include/ptb.h declares (tests check the built library exports exactly these)
ABI_SYMBOLS = ["ptb_create", "ptb_destroy", "ptb_upload_scene", "ptb_set_config", "ptb_set_camera", "ptb_update_instances", "ptb_refit_instances", "ptb_render", "ptb_reserve_wave", "ptb_set_ray_ordering", "ptb_set_static_merge", "ptb_set_integrator", "ptb_set_intersector", "ptb_resize", "ptb_render_frame",
               "ptb_measure_traversal", "ptb_sync", "ptb_get_aov", "ptb_get_display", "ptb_present", "ptb_download", "ptb_get_ray_stats", "ptb_set_pixel_query", "ptb_get_pixel_query", "ptb_get_stream", "ptb_export_rows",
               "ptb_assemble_rows", "ptb_exchange_create", "ptb_exchange_connect", "ptb_exchange_connect_ipc", "ptb_exchange_frame", "ptb_exchange_disconnect", "ptb_debug_read", "ptb_launch_count", "ptb_set_timing", "ptb_get_stage_ms", "ptb_stage_name",
               "ptb_error_string"]

_lib = None


def lib():
    """Loads libptb.so (building it if stale). Raises if it cannot be built or loaded -- there is no fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("PTB_LIB_PATH") or _build.build_cuda()      # PTB_LIB_PATH: A/B tuning variants only
        l = ctypes.CDLL(path)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        l.ptb_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, ci]
        l.ptb_destroy.argtypes = [vp]; l.ptb_destroy.restype = None
        l.ptb_upload_scene.argtypes = [vp, ctypes.POINTER(PtbScene)]
        l.ptb_set_config.argtypes = [vp, ctypes.POINTER(PtbConfig)]
        l.ptb_set_camera.argtypes = [vp, ctypes.POINTER(PtbCamera), vp, vp]
        l.ptb_update_instances.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
        l.ptb_refit_instances.argtypes = [vp, vp, vp, vp]
        l.ptb_render.argtypes = [vp, ci]
        l.ptb_render_frame.argtypes = [vp, ci, ci]
        l.ptb_reserve_wave.argtypes = [vp, ci]
        l.ptb_set_ray_ordering.argtypes = [vp, ci]
        l.ptb_set_static_merge.argtypes = [vp, ci]
        l.ptb_set_intersector.argtypes = [vp, ci]
        l.ptb_set_integrator.argtypes = [vp, ci, ctypes.c_float]
        l.ptb_resize.argtypes = [vp, ci, ci]
        l.ptb_measure_traversal.argtypes = [vp, ci, ctypes.POINTER(PtbTraversalStats)]
        l.ptb_sync.argtypes = [vp]
        l.ptb_get_aov.argtypes = [vp, ci, ci, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        l.ptb_get_display.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        l.ptb_download.argtypes = [vp, ci, ci, vp]
        l.ptb_present.argtypes = [vp, ctypes.POINTER(vp), vp]
        l.ptb_get_ray_stats.argtypes = [vp, ctypes.POINTER(PtbRayStats), ci]
        l.ptb_get_stream.argtypes = [vp, ctypes.POINTER(vp)]
        l.ptb_set_pixel_query.argtypes = [vp, ci, ci]
        l.ptb_get_pixel_query.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]
        l.ptb_exchange_create.argtypes = [vp, ctypes.POINTER(vp), vp]
        l.ptb_exchange_connect.argtypes = [vp, ctypes.POINTER(vp)]
        l.ptb_exchange_connect_ipc.argtypes = [vp, vp]
        l.ptb_exchange_frame.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        l.ptb_exchange_disconnect.argtypes = [vp]
        l.ptb_export_rows.argtypes = [vp, ci, vp, ctypes.POINTER(ci)]
        l.ptb_assemble_rows.argtypes = [vp, vp, ci, vp]
        l.ptb_debug_read.argtypes = [vp, ci, vp, ctypes.c_int64]
        l.ptb_launch_count.argtypes = [vp]; l.ptb_launch_count.restype = ctypes.c_int64
        l.ptb_set_timing.argtypes = [vp, ci]
        l.ptb_get_stage_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ci]
        l.ptb_stage_name.argtypes = [ci]; l.ptb_stage_name.restype = ctypes.c_char_p
        l.ptb_error_string.argtypes = [ci]; l.ptb_error_string.restype = ctypes.c_char_p
        _lib = l
    return _lib


class PtbError(RuntimeError):
    pass


def _check(code, what):
    if code != 0:
        raise PtbError(f"{what} failed: {code} ({lib().ptb_error_string(code).decode()})")


def default_config(**over) -> PtbConfig:
    c = PtbConfig(2, 1, 10, 1, 1, 1, 1, 0, 1, 1, 0, 0.1, 0.1, 6, 4.0, 16.0, 10.0)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def fill_scene_struct(blob, keep):
    """Builds a PtbScene pointing into the blob's numpy arrays (`keep` must outlive the call that consumes it)."""
    def arr(key, dtype=None):
        a = np.ascontiguousarray(blob[key] if dtype is None else np.asarray(blob[key]).astype(dtype, copy=False))
        keep.append(a)
        return a.ctypes.data
    s = PtbScene()
    node_bytes = {8: 80, 4: 128, 2: 32}[int(blob["bvh_kind"])]
    s.triangles = arr("triangles"); s.triangle_count = int(blob["triangles"].shape[0])
    s.bvh_nodes = arr("bvh_nodes"); s.bvh_node_count = int(blob["bvh_nodes"].size // node_bytes)
    s.bvh_kind = int(blob["bvh_kind"]); s.tlas_node_count = int(blob["tlas_node_count"])
    s.mesh_count = int(blob["mesh_bvh_root_indices"].size)
    s.mesh_bvh_root_indices = arr("mesh_bvh_root_indices"); s.mesh_material_ids = arr("mesh_material_ids")
    s.mesh_transforms = arr("mesh_transforms"); s.mesh_transforms_inv = arr("mesh_transforms_inv"); s.mesh_transforms_prev = arr("mesh_transforms_prev")
    s.material_count = int(blob["material_types"].size); s.material_types = arr("material_types"); s.materials = arr("materials")
    s.medium_count = int(blob["media"].shape[0]); s.media = arr("media")
    texs = blob["textures"]
    s.texture_count = len(texs)
    if texs:
        tarr = (PtbTexture * len(texs))()
        for i, t in enumerate(texs):
            levels = [np.ascontiguousarray(l, dtype=np.uint8) for l in t["levels"]]
            keep.extend(levels)
            ptrs = (ctypes.c_void_p * len(levels))(*[l.ctypes.data for l in levels])
            keep.append(ptrs)
            tarr[i].format = 1 if t["format"] == "bc1" else 0
            tarr[i].width, tarr[i].height, tarr[i].num_levels = int(t["width"]), int(t["height"]), len(levels)
            tarr[i].levels = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p)); tarr[i].lod_bias = float(t["lod_bias"])
        keep.append(tarr)
        s.textures = ctypes.cast(tarr, ctypes.POINTER(PtbTexture))
    sky = np.ascontiguousarray(blob["sky"], dtype=np.float32); keep.append(sky)
    s.sky = sky.ctypes.data; s.sky_height, s.sky_width = sky.shape[:2]; s.sky_scale = float(blob["sky_scale"])
    s.pmj_samples = arr("pmj"); s.blue_noise = arr("blue_noise")
    s.lights_total_weight = float(blob["lights_total_weight"])
    s.light_triangle_count = int(blob["light_triangle_indices"].size)
    s.light_triangle_indices = arr("light_triangle_indices"); s.light_triangle_cumulative_probability = arr("light_triangle_cdf")
    s.light_mesh_count = int(blob["light_mesh_cdf"].size); s.light_mesh_cumulative_probability = arr("light_mesh_cdf")
    s.light_mesh_triangle_span = arr("light_mesh_triangle_span"); s.light_mesh_transform_indices = arr("light_mesh_transform_indices")
    return s


def camera_struct(cam15) -> PtbCamera:
    c = PtbCamera()
    v = [float(x) for x in cam15]
    for i in range(3):
        c.position[i] = v[i]; c.bottom_left_corner[i] = v[3 + i]; c.x_axis[i] = v[6 + i]; c.y_axis[i] = v[9 + i]
    c.pixel_spread_angle, c.aperture_radius, c.focal_distance = v[12], v[13], v[14]
    return c


class Pathtracer:
    """Integrator-shaped handle: `update()` then `render()` once per frame, like Src/Main.cpp:137-138."""

    def __init__(self, blob, width=None, height=None, device=0, rank=0, world=1, band_rows=8, config: PtbConfig | None = None):
        l = lib()
        self.screen_width = int(width or blob["width"]); self.screen_height = int(height or blob["height"])
        self.screen_pitch = (self.screen_width + 31) // 32 * 32
        if (self.screen_width, self.screen_height) != (int(blob["width"]), int(blob["height"])):
            raise ValueError("the blob's camera block was built for a different film size; rebuild it with build_blob(width=, height=)")
        self._ctx = ctypes.c_void_p()
        _check(l.ptb_create(ctypes.byref(self._ctx), device, self.screen_width, self.screen_height, rank, world, band_rows), "ptb_create")
        self.rank, self.world, self.band_rows = rank, world, band_rows
        keep = []
        scene = fill_scene_struct(blob, keep)
        _check(l.ptb_upload_scene(self._ctx, ctypes.byref(scene)), "ptb_upload_scene")
        self.gpu_config = config or default_config(num_bounces=int(blob["num_bounces"]))
        self._instance_order = np.asarray(blob["instance_order"]) if "instance_order" in blob else None
        self._camera = camera_struct(blob["camera"])
        self._view_projection = np.ascontiguousarray(blob["view_projection"], dtype=np.float32)
        self._view_projection_prev = self._view_projection.copy()
        self.sample_index = 0
        self.invalidated_scene = False      # TLAS is uploaded with the scene; set + call update_instances() to change it
        self.invalidated_camera = True
        self.invalidated_gpu_config = True
        self._first = True

    # ---- Integrator::update (Integrator.cpp:432-528)
    def update(self, delta=0.0):
        l = lib()
        if self.invalidated_gpu_config and self.gpu_config.enable_svgf and self._camera.aperture_radius > 0.0:
            self._camera.aperture_radius = 0.0      # "SVGF and DoF cannot simultaneously be enabled" (Integrator.cpp:433-437)
            self.invalidated_camera = True
        camera_moved = self.invalidated_camera
        if self.invalidated_camera:
            _check(l.ptb_set_camera(self._ctx, ctypes.byref(self._camera), self._view_projection.ctypes.data, self._view_projection_prev.ctypes.data), "ptb_set_camera")
            if not self.gpu_config.enable_svgf:
                self.sample_index = 0
            self.invalidated_camera = False
        elif self.gpu_config.enable_svgf:
            # static camera: view_projection_prev catches up with view_projection (Camera.cpp:90)
            _check(l.ptb_set_camera(self._ctx, ctypes.byref(self._camera), self._view_projection.ctypes.data, self._view_projection.ctypes.data), "ptb_set_camera")
        if self.invalidated_gpu_config:
            self.invalidated_gpu_config = False
            self.sample_index = 0
            _check(l.ptb_set_config(self._ctx, ctypes.byref(self.gpu_config)), "ptb_set_config")
        elif camera_moved and not self.gpu_config.enable_svgf:
            self.sample_index = 0
        else:
            self.sample_index += 1

    def set_camera(self, cam15, view_projection=None):
        self._view_projection_prev = self._view_projection.copy()
        self._camera = camera_struct(cam15)
        if view_projection is not None:
            self._view_projection = np.ascontiguousarray(view_projection, dtype=np.float32)
        self.invalidated_camera = True

    # ---- Pathtracer::render (Pathtracer.cpp:738-855)
    def render(self):
        _check(lib().ptb_render(self._ctx, int(self.sample_index)), "ptb_render")

    def reserve_wave(self, samples):
        """Let a wave carry `samples` passes at once (ptb_reserve_wave); results stay bit-identical to pass-by-pass tracing."""
        _check(lib().ptb_reserve_wave(self._ctx, int(samples)), "ptb_reserve_wave")

    def render_frame(self, passes):
        """One displayed frame of the reference's `-N passes` mode: Integrator bookkeeping for a fresh accumulation
        (sample_index 0..passes) and the whole launch sequence replayed as one CUDA graph (ptb_render_frame)."""
        self.invalidated_gpu_config = True
        self.update()                                   # sample_index = 0, uploads camera/config if they changed
        _check(lib().ptb_render_frame(self._ctx, 0, int(passes) + 1), "ptb_render_frame")
        self.sample_index = int(passes)

    def render_pass(self, sample_index):
        """Direct control for tests: one pass with an explicit sample index (no Integrator bookkeeping)."""
        if self.invalidated_camera or self.invalidated_gpu_config:
            keep = self.sample_index
            self.update(); self.sample_index = keep
        _check(lib().ptb_render(self._ctx, int(sample_index)), "ptb_render")

    def sync(self):
        _check(lib().ptb_sync(self._ctx), "ptb_sync")

    def render_frames(self, passes):
        """sample_index 0..passes -> accumulator = mean of passes 1..passes (the reference's `-N passes` capture)."""
        for _ in range(passes + 1):
            self.update(); self.render()
        self.sync()

    # ---- readback
    def get_aov(self, aov_type, accumulated=True):
        out = np.empty((self.screen_height, self.screen_pitch, 4), dtype=np.float32)
        _check(lib().ptb_download(self._ctx, int(aov_type), int(accumulated), out.ctypes.data), "ptb_download")
        return out

    def get_display(self):
        out = np.empty((self.screen_height, self.screen_pitch, 4), dtype=np.float32)
        _check(lib().ptb_download(self._ctx, -1, 1, out.ctypes.data), "ptb_download")
        return out

    def present(self):
        """The frame as the reference's window shows it (post.frag: ACES tone mapping + gamma), 8-bit RGBA [height, pitch, 4],
        tone-mapped on the device (ptb_present)."""
        out = np.empty((self.screen_height, self.screen_pitch, 4), dtype=np.uint8)
        _check(lib().ptb_present(self._ctx, None, out.ctypes.data), "ptb_present")
        return out

    def aov_device_ptr(self, aov_type, accumulated=True):
        p, pitch = ctypes.c_void_p(), ctypes.c_int()
        _check(lib().ptb_get_aov(self._ctx, aov_type, int(accumulated), ctypes.byref(p), ctypes.byref(pitch)), "ptb_get_aov")
        return p.value, pitch.value

    def display_device_ptr(self):
        p, pitch = ctypes.c_void_p(), ctypes.c_int()
        _check(lib().ptb_get_display(self._ctx, ctypes.byref(p), ctypes.byref(pitch)), "ptb_get_display")
        return p.value, pitch.value

    def stream(self):
        p = ctypes.c_void_p()
        _check(lib().ptb_get_stream(self._ctx, ctypes.byref(p)), "ptb_get_stream")
        return p.value or 0

    def export_rows(self, device_dst, aov_type=AOV_RADIANCE):
        rows = ctypes.c_int()
        _check(lib().ptb_export_rows(self._ctx, aov_type, ctypes.c_void_p(device_dst), ctypes.byref(rows)), "ptb_export_rows")
        return rows.value

    def assemble_rows(self, device_src, max_rows, device_dst):
        _check(lib().ptb_assemble_rows(self._ctx, ctypes.c_void_p(device_src), int(max_rows), ctypes.c_void_p(device_dst)), "ptb_assemble_rows")

    # ---- frame exchange over peer memory (include/ptb.h: ptb_exchange_*)
    def set_ray_ordering(self, bins):
        """0 = trace queues in emission order, 8 / 64 = direction-binned (include/ptb.h: ptb_set_ray_ordering)."""
        _check(lib().ptb_set_ray_ordering(self._ctx, int(bins)), "ptb_set_ray_ordering")

    def set_pixel_query(self, x, y):
        """Integrator::set_pixel_query (Integrator.h:266-277): the next pass records what the primary ray of pixel (x, y) hits."""
        _check(lib().ptb_set_pixel_query(self._ctx, int(x), int(y)), "ptb_set_pixel_query")

    def get_pixel_query(self, scene_index=False):
        """(mesh_id, triangle_id) of the pending query, (-1, -1) for sky (or when another rank owns the pixel's row); clears the
        query.  mesh_id is the instance index in TLAS leaf order -- the index every device table uses; scene_index=True maps it
        back to the scene's mesh index through the TLAS leaf order like Integrator::update does (Integrator.cpp:486-488)."""
        m, t = ctypes.c_int(), ctypes.c_int()
        _check(lib().ptb_get_pixel_query(self._ctx, ctypes.byref(m), ctypes.byref(t)), "ptb_get_pixel_query")
        mesh = m.value
        if scene_index and mesh >= 0 and self._instance_order is not None:
            mesh = int(self._instance_order[mesh])
        return mesh, t.value

    def set_static_merge(self, enabled):
        """include/ptb.h: ptb_set_static_merge (identity-transform instances traced through one merged CWBVH)."""
        _check(lib().ptb_set_static_merge(self._ctx, int(enabled)), "ptb_set_static_merge")     # False/0 off, True/1 SBVH, 2 plain SAH

    def refit_instances(self, transforms, transforms_inv, transforms_prev=None):
        """Moving instances without a host TLAS rebuild (ptb_refit_instances): [mesh_count, 12] float32 object-to-world and
        world-to-object matrices in this context's table order (scene.instance_transforms(desc, blob["instance_order"])).
        The TLAS boxes are refitted on the device; accumulation restarts."""
        xf = np.ascontiguousarray(transforms, dtype=np.float32); xi = np.ascontiguousarray(transforms_inv, dtype=np.float32)
        xp = None if transforms_prev is None else np.ascontiguousarray(transforms_prev, dtype=np.float32)
        if xf.size != xi.size or (xp is not None and xp.size != xf.size):
            raise ValueError("transform tables differ in size")
        _check(lib().ptb_refit_instances(self._ctx, ctypes.c_void_p(xf.ctypes.data), ctypes.c_void_p(xi.ctypes.data),
                                         ctypes.c_void_p(xp.ctypes.data) if xp is not None else None), "ptb_refit_instances")
        if not self.gpu_config.enable_svgf:
            self.sample_index = 0            # like a moved camera: the accumulator starts over (Integrator.cpp:443-452)

    def resize(self, blob):
        """Pathtracer::resize_free + resize_init (Pathtracer.cpp:255-314): same scene, the film size and camera block of `blob`
        (scene.retarget_blob); accumulation restarts."""
        w, h = int(blob["width"]), int(blob["height"])
        _check(lib().ptb_resize(self._ctx, w, h), "ptb_resize")
        self.screen_width, self.screen_height, self.screen_pitch = w, h, (w + 31) // 32 * 32
        self._camera = camera_struct(blob["camera"])
        self._view_projection = np.ascontiguousarray(blob["view_projection"], dtype=np.float32)
        self._view_projection_prev = self._view_projection.copy()
        self.invalidated_camera = True; self.invalidated_gpu_config = True
        self.sample_index = 0

    def set_integrator(self, kind, ao_radius=1.0):
        """include/ptb.h: ptb_set_integrator -- "pathtracer" (default) or "ao" (the reference's ambient-occlusion integrator, AO.cu)."""
        _check(lib().ptb_set_integrator(self._ctx, {"pathtracer": 0, "ao": 1}[kind] if isinstance(kind, str) else int(kind), float(ao_radius)), "ptb_set_integrator")
        self.invalidated_gpu_config = True

    def set_intersector(self, kind):
        """include/ptb.h: ptb_set_intersector -- "mt" (reference's Moeller-Trumbore, default) or "woop" (merged BVH only)."""
        _check(lib().ptb_set_intersector(self._ctx, {"mt": 0, "woop": 1}[kind] if isinstance(kind, str) else int(kind)), "ptb_set_intersector")

    def exchange_create(self):
        """Allocates this rank's exchange block; returns (device base pointer, 64-byte CUDA IPC handle)."""
        base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _check(lib().ptb_exchange_create(self._ctx, ctypes.byref(base), handle), "ptb_exchange_create")
        return base.value, handle.raw

    def exchange_connect(self, peer_bases):
        """Peers in this process: list of `world` device base pointers, rank-major."""
        arr = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(b) for b in peer_bases])
        _check(lib().ptb_exchange_connect(self._ctx, arr), "ptb_exchange_connect")

    def exchange_connect_ipc(self, handles):
        """Peers in other processes: `world` 64-byte IPC handles (bytes), rank-major."""
        raw = b"".join(handles)
        assert len(raw) == 64 * self.world
        buf = ctypes.create_string_buffer(raw, len(raw))
        _check(lib().ptb_exchange_connect_ipc(self._ctx, buf), "ptb_exchange_connect_ipc")

    def exchange_disconnect(self):
        _check(lib().ptb_exchange_disconnect(self._ctx), "ptb_exchange_disconnect")

    def exchange_frame(self):
        """Device pointer of the complete frame (pitch x height float4) delivered by the last render_frame()."""
        p = ctypes.c_void_p()
        _check(lib().ptb_exchange_frame(self._ctx, ctypes.byref(p), None), "ptb_exchange_frame")
        return p.value

    def owned_rows(self):
        return sum(1 for y in range(self.screen_height) if (y // self.band_rows) % self.world == self.rank)

    def ray_stats(self, reset=False):
        st = PtbRayStats()
        _check(lib().ptb_get_ray_stats(self._ctx, ctypes.byref(st), int(reset)), "ptb_get_ray_stats")
        return dict(trace=np.array(st.trace[:], dtype=np.uint64), shadow=np.array(st.shadow[:], dtype=np.uint64),
                    shaded=np.array(st.shaded[:], dtype=np.uint64), frames=int(st.frames))

    def primary_hits(self):
        out = np.empty((self.screen_height, self.screen_pitch, 4), dtype=np.uint32)
        _check(lib().ptb_debug_read(self._ctx, 0, out.ctypes.data, out.nbytes), "ptb_debug_read")
        return out

    def last_pass_counters(self):
        out = np.empty(8 * 128, dtype=np.int32)
        _check(lib().ptb_debug_read(self._ctx, 1, out.ctypes.data, out.nbytes), "ptb_debug_read")
        o = out.reshape(8, 128)
        return dict(trace=o[0], diffuse=o[1], plastic=o[2], dielectric=o[3], conductor=o[4], shadow=o[5])

    def measure_traversal(self, sample_index):
        """One instrumented pass: node / triangle / instance-transform visit counts for closest-hit [0] and shadow [1] rays."""
        if self.invalidated_camera or self.invalidated_gpu_config:
            keep = self.sample_index
            self.update(); self.sample_index = keep
        st = PtbTraversalStats()
        _check(lib().ptb_measure_traversal(self._ctx, int(sample_index), ctypes.byref(st)), "ptb_measure_traversal")
        return dict(rays=list(st.rays), nodes=list(st.nodes), triangles=list(st.triangles),
                    instance_transforms=list(st.instance_transforms), shadow_misses=int(st.shadow_misses))

    SVGF_TAPS = {"history_normal_and_depth": 10, "history_direct": 11, "history_indirect": 12, "history_moment": 13,
                 "frame_buffer_moment": 14, "taa_frame_curr": 15, "taa_frame_prev": 16, "history_length": 17}

    def svgf_buffer(self, name):
        which = self.SVGF_TAPS[name]
        if which == 17:
            out = np.empty((self.screen_height, self.screen_pitch), dtype=np.int32)
        else:
            out = np.empty((self.screen_height, self.screen_pitch, 4), dtype=np.float32)
        _check(lib().ptb_debug_read(self._ctx, which, out.ctypes.data, out.nbytes), "ptb_debug_read")
        return out

    def tlas_nodes(self, node_count, node_bytes=80):
        """The TLAS the rays currently walk (front of the device node array), e.g. after refit_instances: uint8 [node_count, node_bytes]."""
        out = np.empty((int(node_count), int(node_bytes)), dtype=np.uint8)
        _check(lib().ptb_debug_read(self._ctx, 3, out.ctypes.data, out.nbytes), "ptb_debug_read")
        return out

    def lut_contents(self):
        n = 2 * 16 ** 3 + 2 * 16 ** 2 + 32 ** 2 + 32
        out = np.empty(n, dtype=np.float32)
        _check(lib().ptb_debug_read(self._ctx, 2, out.ctypes.data, out.nbytes), "ptb_debug_read")
        return out

    def launch_count(self):
        return int(lib().ptb_launch_count(self._ctx))

    def set_timing(self, on=True):
        _check(lib().ptb_set_timing(self._ctx, int(on)), "ptb_set_timing")

    def stage_ms(self):
        n = 0
        while lib().ptb_stage_name(n):
            n += 1
        ms = (ctypes.c_float * n)()
        _check(lib().ptb_get_stage_ms(self._ctx, ms, n), "ptb_get_stage_ms")
        return {lib().ptb_stage_name(i).decode(): float(ms[i]) for i in range(n)}

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            lib().ptb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
