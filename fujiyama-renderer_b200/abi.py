"""ctypes view of the drop-in C-ABI declared in include/fjgpu.h (libfjgpu.so).

The structs mirror include/fjgpu.h field for field; `load_fjgpu()` loads the CUDA
library built in-tree by `__graft_entry__.build()` and raises if it is missing —
there is no CPU fallback behind this ABI.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIBFJGPU = os.path.join(HERE, "csrc", "libfjgpu.so")
LIBFJSCENE = os.path.join(HERE, "host", "libfjscene.so")

FJGPU_MAX_SHADING_GROUPS = 8
SHADER_NONE, SHADER_CONSTANT, SHADER_PLASTIC, SHADER_PATHTRACING, SHADER_GLASS = 0, 1, 2, 3, 4
LIGHT_POINT, LIGHT_GRID, LIGHT_SPHERE, LIGHT_DOME = 0, 1, 2, 3
FLAG_FP64_BOXES, FLAG_MEGAKERNEL = 1, 2


class Instance(C.Structure):
    _fields_ = [("mesh_id", C.c_int32),
                ("shader_of_group", C.c_int32 * FJGPU_MAX_SHADING_GROUPS),
                ("reflect_target", C.c_int32), ("refract_target", C.c_int32),
                ("shadow_target", C.c_int32), ("_pad", C.c_int32),
                ("fwd", C.c_double * 16), ("inv", C.c_double * 16)]


class Shader(C.Structure):
    _fields_ = [("kind", C.c_int32), ("do_reflect", C.c_int32),
                ("do_color_filter", C.c_int32), ("texture", C.c_int32),
                ("diffuse", C.c_float * 3), ("reflect", C.c_float * 3),
                ("refract", C.c_float * 3), ("emission", C.c_float * 3),
                ("transmit", C.c_float * 3), ("ior", C.c_float), ("opacity", C.c_float),
                ("bump_texture", C.c_int32), ("bump_amplitude", C.c_float)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("nchannels", C.c_int32), ("tilesize", C.c_int32),
                ("tiles", C.POINTER(C.c_float))]


class Light(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sample_count", C.c_int32),
                ("double_sided", C.c_int32), ("dome_sample_count", C.c_int32),
                ("color", C.c_float * 3), ("intensity", C.c_float),
                ("translate", C.c_double * 3), ("fwd", C.c_double * 16),
                ("dome_dirs", C.POINTER(C.c_double)), ("dome_colors", C.POINTER(C.c_float))]


class Camera(C.Structure):
    _fields_ = [("fwd", C.c_double * 16), ("fov", C.c_double),
                ("znear", C.c_double), ("zfar", C.c_double)]


class RenderParams(C.Structure):
    _fields_ = [("xres", C.c_int32), ("yres", C.c_int32),
                ("xrate", C.c_int32), ("yrate", C.c_int32),
                ("xfwidth", C.c_double), ("yfwidth", C.c_double), ("jitter", C.c_double),
                ("max_diffuse_depth", C.c_int32), ("max_reflect_depth", C.c_int32),
                ("max_refract_depth", C.c_int32), ("cast_shadow", C.c_int32),
                ("target_group", C.c_int32), ("seed", C.c_uint32),
                ("flags", C.c_int32), ("_pad", C.c_int32)]


class Tile(C.Structure):
    _fields_ = [("id", C.c_int32), ("xmin", C.c_int32), ("ymin", C.c_int32),
                ("xmax", C.c_int32), ("ymax", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("rays_camera", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("rays_diffuse", C.c_uint64), ("rays_reflect", C.c_uint64),
                ("rays_refract", C.c_uint64), ("camera_samples", C.c_uint64),
                ("rays_hit", C.c_uint64), ("hit_mesh_levels", C.c_uint64),
                ("node_steps", C.c_uint64), ("tri_tests", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("trace_launches", C.c_uint64),
                ("ms_trace", C.c_float), ("ms_resolve", C.c_float),
                ("ms_total", C.c_float), ("ms_shade", C.c_float),
                ("batches", C.c_uint32), ("queue_regrows", C.c_uint32),
                ("first_regrow_batch", C.c_int32), ("_pad", C.c_uint32),
                ("leaf_phases", C.c_uint64), ("leaf_rounds", C.c_uint64)]

    @property
    def rays(self):
        return (self.rays_camera + self.rays_shadow + self.rays_diffuse +
                self.rays_reflect + self.rays_refract)


class SceneInfo(C.Structure):
    _fields_ = [("hbm_bytes", C.c_uint64), ("blas_nodes", C.c_uint64),
                ("blas_tris", C.c_uint64), ("tlas_nodes", C.c_uint64),
                ("instances", C.c_uint64), ("blas_max_depth", C.c_uint32),
                ("_pad", C.c_uint32), ("build_seconds", C.c_double),
                ("device_build_seconds", C.c_double)]


# every symbol include/fjgpu.h declares (tests check the built library exports all of them)
FJGPU_SYMBOLS = [
    "fjgpu_create", "fjgpu_destroy", "fjgpu_last_error", "fjgpu_api_version",
    "fjgpu_mesh_upload", "fjgpu_instances_set", "fjgpu_groups_set", "fjgpu_shaders_set",
    "fjgpu_lights_set", "fjgpu_camera_set", "fjgpu_render_tiles", "fjgpu_render_tiles_device",
    "fjgpu_render_tiles_resident", "fjgpu_trace_closest", "fjgpu_render_tile_samples",
    "fjgpu_scene_info_get", "fjgpu_scene_resend", "fjgpu_textures_set", "fjgpu_mesh_set_uv",
    "fjgpu_time_table", "fjgpu_instance_motion_set", "fjgpu_camera_motion_set",
    "fjgpu_mesh_upload_velocity", "fjgpu_shutter_set", "fjgpu_assemble_frame", "fjgpu_render_frame_multi",
]

_P = C.POINTER


def _proto(lib):
    vp, i32, f64p, i32p = C.c_void_p, C.c_int32, _P(C.c_double), _P(C.c_int32)
    lib.fjgpu_create.argtypes = [C.c_int, _P(vp)]
    lib.fjgpu_destroy.argtypes = [vp]
    lib.fjgpu_destroy.restype = None
    lib.fjgpu_last_error.argtypes = [vp]
    lib.fjgpu_last_error.restype = C.c_char_p
    lib.fjgpu_mesh_upload.argtypes = [vp, i32, f64p, f64p, i32, i32p, i32p, i32]
    lib.fjgpu_instances_set.argtypes = [vp, i32, _P(Instance)]
    lib.fjgpu_groups_set.argtypes = [vp, i32, i32p, i32p]
    lib.fjgpu_shaders_set.argtypes = [vp, i32, _P(Shader)]
    lib.fjgpu_lights_set.argtypes = [vp, i32, _P(Light)]
    lib.fjgpu_camera_set.argtypes = [vp, _P(Camera)]
    lib.fjgpu_render_tiles.argtypes = [vp, _P(RenderParams), _P(Tile), i32, _P(C.c_float), _P(Stats)]
    lib.fjgpu_render_tiles_device.argtypes = [vp, _P(RenderParams), _P(Tile), i32, i32, i32, vp, _P(Stats)]
    lib.fjgpu_render_tiles_resident.argtypes = [vp, _P(RenderParams), _P(Tile), i32, _P(Stats)]
    lib.fjgpu_trace_closest.argtypes = [vp, i32, i32, f64p, f64p, f64p, f64p, i32, f64p, f64p, f64p, i32p, i32p]
    lib.fjgpu_render_tile_samples.argtypes = [vp, _P(RenderParams), _P(Tile), i32, f64p, _P(C.c_float), i32p]
    lib.fjgpu_scene_info_get.argtypes = [vp, _P(SceneInfo)]
    lib.fjgpu_scene_resend.argtypes = [vp, _P(C.c_uint64)]
    lib.fjgpu_time_table.argtypes = [_P(RenderParams), _P(Tile), i32, C.c_double, C.c_double, f64p, i32]
    lib.fjgpu_instance_motion_set.argtypes = [vp, i32, i32, f64p, f64p]
    lib.fjgpu_camera_motion_set.argtypes = [vp, i32, f64p]
    lib.fjgpu_mesh_upload_velocity.argtypes = [vp, i32, f64p, f64p, i32, i32p, i32p, i32, f64p]
    lib.fjgpu_shutter_set.argtypes = [vp, C.c_double, C.c_double]
    lib.fjgpu_assemble_frame.argtypes = [vp, vp, i32, i32, i32, _P(Tile), i32, i32, i32, vp]
    lib.fjgpu_render_frame_multi.argtypes = [_P(vp), i32, _P(RenderParams), _P(Tile), i32, _P(C.c_float), _P(Stats)]
    return lib


_fjgpu = None


def load_fjgpu():
    """Loads libfjgpu.so (RTLD_GLOBAL so libfjscene resolves it too). Raises if not built."""
    global _fjgpu
    if _fjgpu is None:
        if not os.path.exists(LIBFJGPU):
            raise RuntimeError(
                "libfjgpu.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "the renderer has no CPU fallback" % LIBFJGPU)
        _fjgpu = _proto(C.CDLL(LIBFJGPU, mode=C.RTLD_GLOBAL))
    return _fjgpu
