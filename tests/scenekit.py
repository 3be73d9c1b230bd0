"""Test infrastructure: one scene description -> (a) the `.scn` command text the reference's
bin/scene executes, (b) the flat C-ABI structs of include/fjgpu.h for the oracle and the GPU.

Also loads/builds oracle/libfjoracle.so and locates the reference build in oracle/_ref.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libfjoracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


def pkg():
    """Imports the (hyphenated) package directory fujiyama-renderer_b200 as `fujiyama_renderer_b200`."""
    name = "fujiyama_renderer_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(REPO, "fujiyama-renderer_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


pkg()
abi = None


def _abi():
    global abi
    if abi is None:
        pkg()
        import fujiyama_renderer_b200.abi as a
        abi = a
    return abi


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "fj_oracle.cc")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-std=c++14",
                               "-I", os.path.join(REPO, "include"), src, "-o", ORACLE_SO])
    return ORACLE_SO


_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    a = _abi()
    lib = C.CDLL(build_oracle())
    P = C.POINTER
    vp, i32, f64p, i32p = C.c_void_p, C.c_int32, P(C.c_double), P(C.c_int32)
    lib.fjo_scene_new.restype = vp
    lib.fjo_scene_free.argtypes = [vp]
    lib.fjo_mesh.argtypes = [vp, i32, f64p, f64p, i32, i32p, i32p, i32]
    lib.fjo_compute_normals.argtypes = [f64p, i32, i32p, i32, f64p]
    lib.fjo_instances.argtypes = [vp, i32, P(a.Instance)]
    lib.fjo_groups.argtypes = [vp, i32, i32p, i32p]
    lib.fjo_shaders.argtypes = [vp, i32, P(a.Shader)]
    lib.fjo_mesh_set_uv.argtypes = [vp, i32, P(C.c_float), i32]
    lib.fjo_mesh_generate_velocity.argtypes = [vp, i32, f64p]
    lib.fjo_perlin3d.argtypes = [f64p, C.c_double, C.c_double, i32, f64p]
    lib.fjo_perlin3d.restype = None
    lib.fjo_smoothstep.argtypes = [C.c_double] * 3
    lib.fjo_smoothstep.restype = C.c_double
    lib.fjo_textures.argtypes = [vp, i32, P(a.Texture)]
    lib.fjo_dome_samples.argtypes = [P(a.Texture), i32, f64p, P(C.c_float)]
    lib.fjo_lights.argtypes = [vp, i32, P(a.Light)]
    lib.fjo_camera.argtypes = [vp, P(a.Camera)]
    lib.fjo_transform_samples.argtypes = [vp, i32, i32, i32, i32, f64p, i32, f64p, i32, f64p]
    lib.fjo_time_range.argtypes = [vp, C.c_double, C.c_double]
    lib.fjo_time_range.restype = None
    lib.fjo_sample_times.argtypes = [P(a.RenderParams), P(a.Tile), C.c_double, C.c_double, f64p, i32]
    lib.fjo_lerp_transform.argtypes = [i32, i32, i32, f64p, i32, f64p, i32, f64p, C.c_double, f64p, f64p]
    lib.fjo_build.argtypes = [vp]
    lib.fjo_render.argtypes = [vp, P(a.RenderParams), P(a.Tile), i32, P(C.c_float), i32, i32, P(a.Stats)]
    lib.fjo_render_tile_samples.argtypes = [vp, P(a.RenderParams), P(a.Tile), i32, f64p, P(C.c_float)]
    lib.fjo_trace_closest.argtypes = [vp, i32, i32, f64p, f64p, f64p, f64p, f64p, f64p, f64p, i32p, i32p]
    lib.fjo_xorshift_u32.argtypes = [P(C.c_uint32), i32]
    lib.fjo_xorshift_f01.argtypes = [f64p, i32]
    lib.fjo_tri_intersect.argtypes = [f64p] * 6
    lib.fjo_box_intersect.argtypes = [f64p, f64p, f64p, f64p, C.c_double, C.c_double, f64p]
    lib.fjo_make_transform.argtypes = [i32, i32, f64p, f64p, f64p, f64p, f64p]
    lib.fjo_camera_ray.argtypes = [P(a.Camera), i32, i32, C.c_double, C.c_double, f64p, f64p]
    lib.fjo_generate_samples.argtypes = [P(a.RenderParams), P(a.Tile), f64p, i32]
    lib.fjo_gaussian.argtypes = [C.c_double] * 4
    lib.fjo_gaussian.restype = C.c_double
    lib.fjo_fresnel.argtypes = [f64p, f64p, C.c_double]
    lib.fjo_fresnel.restype = C.c_double
    lib.fjo_reflect.argtypes = [f64p, f64p, f64p]
    lib.fjo_refract.argtypes = [f64p, f64p, C.c_double, f64p]
    lib.fjo_philox.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, f64p]
    _oracle = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "bin", "ref_probe"))


def make_transform(T=(0, 0, 0), R=(0, 0, 0), S=(1, 1, 1), torder=0, rorder=10):
    """Reference arithmetic (make_transform_matrix + MatInverse) via the oracle. ORDER_SRT=0, ORDER_ZXY=10."""
    fwd = np.zeros(16)
    inv = np.zeros(16)
    t, r, s = (np.asarray(x, dtype=np.float64) for x in (T, R, S))
    oracle().fjo_make_transform(torder, rorder, dptr(t), dptr(r), dptr(s), dptr(fwd), dptr(inv))
    return fwd, inv


def sample_rows(samples, default, initial=(0.0, 0.0, 0.0)):
    """Time samples of one transform channel as sorted (x, y, z, time) rows.  A PropertySampleList starts with one sample
    at time 0 (`initial`: zeros, ones for scale; PropInitSampleList / XfmInitTransformSampleList), PropPushSample keeps
    the list sorted by time and a sample at an existing time replaces it (src/fj_property.cc:284-312): SetProperty3
    (`default`, time 0) or the SetSampleProperty3 keys (`samples`) are applied on top of that."""
    rows = {0.0: [float(initial[0]), float(initial[1]), float(initial[2]), 0.0]}
    for r in (samples if samples else [tuple(default) + (0.0,)]):
        rows[float(r[3])] = [float(r[0]), float(r[1]), float(r[2]), float(r[3])]
    return np.ascontiguousarray([rows[t] for t in sorted(rows)], np.float64)


def lerp_rows(rows, time):
    """PropLerpSamples (src/fj_property.cc:317-345) on sample_rows output, in the reference's operation order."""
    if rows[0][3] >= time or len(rows) == 1:
        return rows[0][:3]
    if rows[-1][3] <= time:
        return rows[-1][:3]
    for i in range(len(rows)):
        if rows[i][3] == time:
            return rows[i][:3]
        if rows[i][3] > time:
            a, b = rows[i - 1], rows[i]
            t = (time - a[3]) / (b[3] - a[3])          # Fit(time, t0, t1, 0, 1) strictly inside the interval
            return np.array([(1 - t) * a[k] + t * b[k] for k in range(3)])


def motion_table(T4, R4, S4, times, torder=0, rorder=10):
    """Transform::matrix / inverse at every time of the frame's time table (XfmLerpTransformSample per entry)."""
    fwd = np.zeros((len(times), 16))
    inv = np.zeros((len(times), 16))
    for k, t in enumerate(times):
        fwd[k], inv[k] = make_transform(lerp_rows(T4, t), lerp_rows(R4, t), lerp_rows(S4, t), torder, rorder)
    return fwd, inv


def make_tiles(xres, yres, tile=32, region=None):
    """Tiler::GenerateTiles (src/fj_tiler.cc:56-113): row-major tiles clipped to the region."""
    x0, y0, x1, y1 = region if region else (0, 0, xres, yres)
    X0, Y0 = max(0, x0) // tile, max(0, y0) // tile
    X1, Y1 = -(-min(xres, x1) // tile), -(-min(yres, y1) // tile)
    out = []
    for y in range(Y0, Y1):
        for x in range(X0, X1):
            out.append((len(out), max(x * tile, x0), max(y * tile, y0), min((x + 1) * tile, x1), min((y + 1) * tile, y1)))
    return out


SHADER_PLUGIN = {"constant": ("constant_shader", "ConstantShader"), "plastic": ("plastic_shader", "PlasticShader"),
                 "pathtracing": ("pathtracing_shader", "PathtracingShader"), "glass": ("glass_shader", "GlassShader")}
LIGHT_NAME = {0: "PointLight", 1: "GridLight", 2: "SphereLight", 3: "DomeLight"}


class SceneDesc:
    """A static scene: meshes, instances (TRS), shaders, lights, camera, renderer properties."""

    def __init__(self):
        self.meshes = []      # (name, P float32 [V,3], idx int32 [F,3], ply_path or None)
        self.mesh_uv = {}     # name -> uv float32 [V,2] (PLY properties uv1 / uv2)
        self.mesh_velocity = set()   # names of meshes the reference's VelocityGeneratorProcedure runs on
        self.textures = []    # (name, image float32 [H,W,C]) written as .mip; shaders refer to them by name in the
                              # `texture` (constant) / `diffuse_map` (plastic, pathtracing) property
        self.shaders = []     # (name, kind str, props dict)
        self.instances = []   # dict(name, mesh, T, R, S, shader)
        self.lights = []      # dict(kind, T, R, S, intensity, color, sample_count, double_sided)
        self.cam = dict(T=(0, 0, 4.5), R=(0, 0, 0), fov=30.0, znear=.01, zfar=1000.0)
        self.ren = dict(resolution=(320, 240), pixelsamples=(3, 3), filterwidth=(2, 2), sample_jitter=1.0,
                        max_diffuse_depth=3, max_reflect_depth=3, max_refract_depth=3, cast_shadow=1, tilesize=32,
                        seed=1, sample_time_range=(0.0, 1.0))
        # motion blur: an instance dict / self.cam may hold "samples" = dict(T=[(x, y, z, time), ...], R=[...], S=[...]);
        # channels listed there are set with SetSampleProperty3 instead of SetProperty3

    # ---- construction
    def mesh(self, name, P, idx, ply_path=None, uv=None, velocity=False):
        self.meshes.append((name, np.ascontiguousarray(P, np.float32), np.ascontiguousarray(idx, np.int32), ply_path))
        if velocity:
            self.mesh_velocity.add(name)
        if uv is not None:
            self.mesh_uv[name] = np.ascontiguousarray(uv, np.float32)

    def texture(self, name, img):
        self.textures.append((name, np.ascontiguousarray(img, np.float32)))

    def shader(self, name, kind, **props):
        self.shaders.append((name, kind, props))

    def instance(self, name, mesh, shader, T=(0, 0, 0), R=(0, 0, 0), S=(1, 1, 1), samples=None):
        self.instances.append(dict(name=name, mesh=mesh, shader=shader, T=T, R=R, S=S, samples=samples))

    @staticmethod
    def channel_rows(d):
        """(T4, R4, S4, moving) of an instance / camera dict."""
        smp = d.get("samples") or {}
        T4 = sample_rows(smp.get("T"), d["T"])
        R4 = sample_rows(smp.get("R"), d["R"])
        S4 = sample_rows(smp.get("S"), d.get("S", (1, 1, 1)), (1.0, 1.0, 1.0))
        return T4, R4, S4, max(len(T4), len(R4), len(S4)) > 1

    @staticmethod
    def scn_transform(L, name, d, channels):
        smp = d.get("samples") or {}
        for key, prop in channels:
            if smp.get(key):
                for r in smp[key]:
                    L.append("SetSampleProperty3 %s %s %r %r %r %r" % ((name, prop) + tuple(float(x) for x in r)))
            else:
                L.append("SetProperty3 %s %s %r %r %r" % ((name, prop) + tuple(float(x) for x in d[key])))

    def light(self, kind=0, T=(0, 0, 0), R=(0, 0, 0), S=(1, 1, 1), intensity=1.0, color=(1, 1, 1), sample_count=16,
              double_sided=0, environment_map=None):
        self.lights.append(dict(kind=kind, T=T, R=R, S=S, intensity=intensity, color=color,
                                sample_count=sample_count, double_sided=double_sided, environment_map=environment_map))

    def tiles(self, region=None):
        return make_tiles(self.ren["resolution"][0], self.ren["resolution"][1], self.ren["tilesize"], region)

    # ---- (a) .scn text for the reference binary
    def to_scn(self, workdir, out_fb, threads=1, plugin_dir=None, region=None):
        s = pkg()
        plugin_dir = plugin_dir or os.path.join(REF_DIR, "lib")
        L = []
        kinds = sorted(set(k for _, k, _ in self.shaders))
        for k in kinds:
            L.append("OpenPlugin %s %s" % (SHADER_PLUGIN[k][0], os.path.join(plugin_dir, SHADER_PLUGIN[k][1])))
        L.append("OpenPlugin stanfordply_procedure %s" % os.path.join(plugin_dir, "StanfordPlyProcedure"))
        if self.mesh_velocity:
            L.append("OpenPlugin velocity_generator_procedure %s" % os.path.join(plugin_dir, "VelocityGeneratorProcedure"))
        L.append("NewCamera cam1 PerspectiveCamera")
        self.scn_transform(L, "cam1", self.cam, (("T", "translate"), ("R", "rotate")))
        L.append("SetProperty1 cam1 fov %r" % float(self.cam["fov"]))
        env_assign = []
        for i, lt in enumerate(self.lights):
            n = "light%d" % i
            L.append("NewLight %s %s" % (n, LIGHT_NAME[lt["kind"]]))
            L.append("SetProperty3 %s translate %r %r %r" % ((n,) + tuple(float(x) for x in lt["T"])))
            L.append("SetProperty3 %s rotate %r %r %r" % ((n,) + tuple(float(x) for x in lt["R"])))
            L.append("SetProperty3 %s scale %r %r %r" % ((n,) + tuple(float(x) for x in lt["S"])))
            L.append("SetProperty1 %s intensity %r" % (n, float(lt["intensity"])))
            L.append("SetProperty3 %s color %r %r %r" % ((n,) + tuple(float(x) for x in lt["color"])))
            L.append("SetProperty1 %s sample_count %d" % (n, lt["sample_count"]))
            L.append("SetProperty1 %s double_sided %d" % (n, lt["double_sided"]))
            if lt.get("environment_map"):
                env_assign.append("AssignTexture %s environment_map %s" % (n, lt["environment_map"]))
        from fujiyama_renderer_b200 import synth
        for tname, img in self.textures:
            mip = os.path.join(workdir, tname + ".mip")
            synth.write_mip(mip, img)
            L.append("NewTexture %s %s" % (tname, mip))
        L += env_assign
        for name, kind, props in self.shaders:
            L.append("NewShader %s %s" % (name, SHADER_PLUGIN[kind][0]))
            for k, v in props.items():
                if isinstance(v, str):
                    L.append("AssignTexture %s %s %s" % (name, k, v))
                elif np.isscalar(v):
                    L.append("SetProperty1 %s %s %r" % (name, k, float(v)))
                else:
                    L.append("SetProperty3 %s %s %r %r %r" % ((name, k) + tuple(float(x) for x in v)))
        from fujiyama_renderer_b200 import synth
        for name, P, idx, ply in self.meshes:
            if ply is None:
                ply = os.path.join(workdir, name + ".ply")
                synth.write_ply(ply, P, idx, self.mesh_uv.get(name))
            L.append("NewMesh %s" % name)
            L.append("NewProcedure %s_proc stanfordply_procedure" % name)
            L.append("AssignMesh %s_proc mesh %s" % (name, name))
            L.append("SetStringProperty %s_proc filepath %s" % (name, ply))
            L.append("SetStringProperty %s_proc io_mode r" % name)
            L.append("RunProcedure %s_proc" % name)
            if name in self.mesh_velocity:       # scenes/mesh_velocity_blur.py:70-73
                L += ["NewProcedure %s_velgen velocity_generator_procedure" % name, "AssignMesh %s_velgen mesh %s" % (name, name),
                      "RunProcedure %s_velgen" % name]
        for ins in self.instances:
            n = ins["name"]
            L.append("NewObjectInstance %s %s" % (n, ins["mesh"]))
            self.scn_transform(L, n, ins, (("T", "translate"), ("R", "rotate"), ("S", "scale")))
            if ins["shader"] is not None:
                L.append("AssignShader %s DEFAULT_SHADING_GROUP %s" % (n, ins["shader"]))
        r = self.ren
        L += ["NewFrameBuffer fb1 rgba", "NewRenderer ren1", "AssignCamera ren1 cam1", "AssignFrameBuffer ren1 fb1",
              "SetProperty2 ren1 resolution %d %d" % tuple(r["resolution"]),
              "SetProperty2 ren1 pixelsamples %d %d" % tuple(r["pixelsamples"]),
              "SetProperty2 ren1 filterwidth %r %r" % tuple(float(x) for x in r["filterwidth"]),
              "SetProperty2 ren1 tilesize %d %d" % (r["tilesize"], r["tilesize"]),
              "SetProperty1 ren1 sample_jitter %r" % float(r["sample_jitter"]),
              "SetProperty1 ren1 max_diffuse_depth %d" % r["max_diffuse_depth"],
              "SetProperty1 ren1 max_reflect_depth %d" % r["max_reflect_depth"],
              "SetProperty1 ren1 max_refract_depth %d" % r["max_refract_depth"],
              "SetProperty1 ren1 cast_shadow %d" % r["cast_shadow"],
              "SetProperty2 ren1 sample_time_range %r %r" % tuple(float(x) for x in r["sample_time_range"]),
              "SetProperty1 ren1 use_max_thread 0", "SetProperty1 ren1 thread_count %d" % threads]
        if region:
            L.append("SetProperty4 ren1 render_region %d %d %d %d" % tuple(region))
        L += ["RenderScene ren1"]
        if out_fb:
            L.append("SaveFrameBuffer fb1 %s" % out_fb)
        return "\n".join(L) + "\n"

    # ---- (b) flat structs of include/fjgpu.h
    def shader_struct(self, kind, props):
        a = _abi()
        sh = a.Shader()
        f32 = np.float32
        if kind == "constant":
            sh.kind = a.SHADER_CONSTANT
            d = props.get("diffuse", (1, 1, 1))
            sh.diffuse[:] = [max(0.0, float(f32(x))) for x in d]
        elif kind == "plastic":
            sh.kind = a.SHADER_PLASTIC
            sh.diffuse[:] = [max(0.0, float(f32(x))) for x in props.get("diffuse", (.8, .8, .8))]
            refl = [max(0.0, float(f32(x))) for x in props.get("reflect", (1, 1, 1))]
            sh.reflect[:] = refl
            sh.do_reflect = 1 if any(x > 0 for x in refl) else 0
            sh.ior = max(float(f32(.001)), float(f32(props.get("ior", 1.4))))
            sh.opacity = min(1.0, max(0.0, float(f32(props.get("opacity", 1)))))
            sh.bump_amplitude = float(f32(props.get("bump_amplitude", 1)))
        elif kind == "glass":
            sh.kind = a.SHADER_GLASS
            fc = [max(float(f32(.001)), float(f32(x))) for x in props.get("filter_color", (1, 1, 1))]
            sh.transmit[:] = fc
            sh.do_color_filter = 0 if all(x == 1 for x in fc) else 1
            sh.ior = max(0.0, float(f32(props.get("ior", 1.4))))
            sh.opacity = 1.0
        else:
            sh.kind = a.SHADER_PATHTRACING
            sh.emission[:] = [max(0.0, float(f32(x))) for x in props.get("emission", (0, 0, 0))]
            sh.diffuse[:] = [max(0.0, float(f32(x))) for x in props.get("diffuse", (.8, .8, .8))]
            sh.reflect[:] = [max(0.0, float(f32(x))) for x in props.get("reflect", (0, 0, 0))]
            sh.refract[:] = [max(0.0, float(f32(x))) for x in props.get("refract", (0, 0, 0))]
            tr = [max(float(f32(.001)), float(f32(x))) for x in props.get("transmit", (1, 1, 1))]
            sh.transmit[:] = tr
            sh.do_color_filter = 0 if all(x == 1 for x in tr) else 1
            sh.ior = max(float(f32(.001)), float(f32(props.get("ior", 1.4))))
            sh.opacity = 1.0
            sh.bump_amplitude = float(f32(props.get("bump_amplitude", 1)))
        return sh

    def to_structs(self):
        """Returns a dict of ctypes arrays / numpy buffers ready for fjo_* and fjgpu_* calls."""
        a = _abi()
        o = oracle()
        out = {}
        mesh_ids = {}
        meshes = []
        for mid, (name, P, idx, _ply) in enumerate(self.meshes):
            mesh_ids[name] = mid
            P64 = np.ascontiguousarray(P.astype(np.float64))
            idx = np.ascontiguousarray(idx.reshape(-1), np.int32)
            N64 = np.zeros_like(P64)
            o.fjo_compute_normals(dptr(P64), len(P64), iptr(idx), len(idx) // 3, dptr(N64))
            meshes.append((mid, P64, N64, idx))
        out["meshes"] = meshes
        out["velocity_meshes"] = sorted(mesh_ids[n] for n in self.mesh_velocity)
        # what VelocityGeneratorProcedure writes on those meshes: from the oracle's restatement of the generator, which is
        # pinned on vectors dumped from the reference (tests/test_oracle_golden.py); libfjscene's own copy is held to it too
        out["mesh_velocity"] = {}
        for mid in out["velocity_meshes"]:
            _, P64, N64, idx = meshes[mid]
            tmp = o.fjo_scene_new()
            o.fjo_mesh(tmp, mid, dptr(P64), dptr(N64), len(P64), iptr(idx), None, len(idx) // 3)
            vel = np.zeros_like(P64)
            assert o.fjo_mesh_generate_velocity(tmp, mid, dptr(vel)) == 0
            o.fjo_scene_free(tmp)
            out["mesh_velocity"][mid] = vel
        out["mesh_uv"] = {mesh_ids[n]: uv for n, uv in self.mesh_uv.items()}
        # textures: the tile arrays a .mip file of the image holds (synth.write_mip), as fjgpu_texture structs
        tex_ids = {}
        texs = (a.Texture * max(1, len(self.textures)))()
        for i, (tname, img) in enumerate(self.textures):
            tex_ids[tname] = i
            h, w, c = img.shape
            tiles = np.ascontiguousarray(img.reshape(h // 64, 64, w // 64, 64, c).transpose(0, 2, 1, 3, 4), np.float32)
            out.setdefault("_keep", []).append(tiles)
            texs[i].width, texs[i].height, texs[i].nchannels, texs[i].tilesize = w, h, c, 64
            texs[i].tiles = fptr(tiles)
        out["textures"] = texs
        out["ntextures"] = len(self.textures)
        sh_ids = {}
        shs = (a.Shader * max(1, len(self.shaders)))()
        for i, (name, kind, props) in enumerate(self.shaders):
            sh_ids[name] = i
            shs[i] = self.shader_struct(kind, {k: v for k, v in props.items() if not isinstance(v, str)})
            for k, v in props.items():
                if isinstance(v, str):
                    assert (kind, k) in (("constant", "texture"), ("plastic", "diffuse_map"), ("pathtracing", "diffuse_map"),
                                         ("plastic", "bump_map"), ("pathtracing", "bump_map"))
                    if k == "bump_map":
                        shs[i].bump_texture = tex_ids[v] + 1
                    else:
                        shs[i].texture = tex_ids[v] + 1
        out["shaders"] = shs
        out["nshaders"] = len(self.shaders)
        ins = (a.Instance * max(1, len(self.instances)))()
        samples = {}          # instance index (-1 = camera) -> (T4, R4, S4) of the time-sampled transforms
        for i, d in enumerate(self.instances):
            T4, R4, S4, moving = self.channel_rows(d)
            if moving:
                samples[i] = (T4, R4, S4)
            fwd, inv = make_transform(lerp_rows(T4, 0.0), lerp_rows(R4, 0.0), lerp_rows(S4, 0.0))
            ins[i].mesh_id = mesh_ids[d["mesh"]]
            for g in range(a.FJGPU_MAX_SHADING_GROUPS):
                ins[i].shader_of_group[g] = -1
            ins[i].shader_of_group[0] = sh_ids[d["shader"]] if d["shader"] is not None else -1
            ins[i].reflect_target = ins[i].refract_target = ins[i].shadow_target = 0
            ins[i].fwd[:] = list(fwd)
            ins[i].inv[:] = list(inv)
        out["instances"] = ins
        out["ninstances"] = len(self.instances)
        out["group_offsets"] = np.array([0, len(self.instances)], np.int32)
        out["group_ids"] = np.arange(max(1, len(self.instances)), dtype=np.int32)
        lts = (a.Light * max(1, len(self.lights)))()
        for i, d in enumerate(self.lights):
            fwd, _ = make_transform(d["T"], d["R"], d["S"])
            lts[i].kind = d["kind"]
            lts[i].sample_count = max(1, int(d["sample_count"]))
            lts[i].double_sided = d["double_sided"]
            lts[i].color[:] = [float(x) for x in d["color"]]
            lts[i].intensity = float(d["intensity"])
            lts[i].translate[:] = [float(x) for x in d["T"]]
            lts[i].fwd[:] = list(fwd)
            if d["kind"] == a.LIGHT_DOME and d.get("environment_map"):   # DomeLight::preprocess with an environment map (:78-96)
                n = lts[i].sample_count
                dirs = np.zeros((n, 3), np.float64)
                cols = np.zeros((n, 3), np.float32)
                assert o.fjo_dome_samples(C.byref(texs[tex_ids[d["environment_map"]]]), n, dptr(dirs), fptr(cols)) == 0
                out.setdefault("_keep", []).extend([dirs, cols])
                lts[i].dome_sample_count = n
                lts[i].dome_dirs = dptr(dirs)
                lts[i].dome_colors = fptr(cols)
            elif d["kind"] == a.LIGHT_DOME:   # DomeLight::preprocess without an environment map (fj_dome_light.cc:64-76)
                n = lts[i].sample_count
                v = np.array([1. / n, 1., 1. / n])
                inv_len = 1. / np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
                dirs = np.ascontiguousarray(np.tile(v * inv_len, (n, 1)))
                cols = np.ascontiguousarray(np.tile(np.array([1, .63, .63], np.float32), (n, 1)))
                out.setdefault("_keep", []).extend([dirs, cols])
                lts[i].dome_sample_count = n
                lts[i].dome_dirs = dptr(dirs)
                lts[i].dome_colors = fptr(cols)
        out["lights"] = lts
        out["nlights"] = len(self.lights)
        cam = a.Camera()
        T4, R4, S4, moving = self.channel_rows(self.cam)
        if moving:
            samples[-1] = (T4, R4, S4)
        fwd, _ = make_transform(lerp_rows(T4, 0.0), lerp_rows(R4, 0.0), (1, 1, 1))
        cam.fwd[:] = list(fwd)
        cam.fov, cam.znear, cam.zfar = self.cam["fov"], self.cam["znear"], self.cam["zfar"]
        out["camera"] = cam
        r = self.ren
        p = a.RenderParams()
        p.xres, p.yres = r["resolution"]
        p.xrate, p.yrate = r["pixelsamples"]
        p.xfwidth, p.yfwidth = r["filterwidth"]
        p.jitter = r["sample_jitter"]
        p.max_diffuse_depth, p.max_reflect_depth, p.max_refract_depth = (
            r["max_diffuse_depth"], r["max_reflect_depth"], r["max_refract_depth"])
        p.cast_shadow = r["cast_shadow"]
        p.target_group = 0
        p.seed = r["seed"]
        out["params"] = p
        out["samples"] = samples
        out["time_range"] = tuple(float(x) for x in r["sample_time_range"])
        if samples:
            # the frame's time table (include/fjgpu.h, motion blur) and every moving transform evaluated at its entries
            # with the reference's arithmetic: what a caller of the C-ABI hands to fjgpu_*_motion_set
            from fujiyama_renderer_b200 import abi
            lib = abi.load_fjgpu()
            ta = self.tile_array(self.tiles())
            n = lib.fjgpu_time_table(C.byref(p), ta, len(ta), out["time_range"][0], out["time_range"][1], None, 0)
            assert n > 0
            times = np.zeros(n)
            assert lib.fjgpu_time_table(C.byref(p), ta, len(ta), out["time_range"][0], out["time_range"][1], dptr(times), n) == n
            out["times"] = times
            out["motion"] = {i: motion_table(T4, R4, S4, times) for i, (T4, R4, S4) in samples.items()}
        return out

    @staticmethod
    def tile_array(tiles):
        a = _abi()
        arr = (a.Tile * len(tiles))()
        for i, t in enumerate(tiles):
            arr[i].id, arr[i].xmin, arr[i].ymin, arr[i].xmax, arr[i].ymax = t
        return arr


def oracle_scene(st):
    """Feeds the flat structs to the oracle; returns the fjo_scene handle."""
    o = oracle()
    sc = o.fjo_scene_new()
    for mid, P, N, idx in st["meshes"]:
        o.fjo_mesh(sc, mid, dptr(P), dptr(N), len(P), iptr(idx), None, len(idx) // 3)
    for mid, uv in st.get("mesh_uv", {}).items():
        assert o.fjo_mesh_set_uv(sc, mid, fptr(uv), len(uv)) == 0
    for mid in st.get("velocity_meshes", []):
        assert o.fjo_mesh_generate_velocity(sc, mid, None) == 0
    o.fjo_textures(sc, st.get("ntextures", 0), st.get("textures"))
    o.fjo_instances(sc, st["ninstances"], st["instances"])
    o.fjo_groups(sc, 1, iptr(st["group_offsets"]), iptr(st["group_ids"]))
    o.fjo_shaders(sc, st["nshaders"], st["shaders"])
    o.fjo_lights(sc, st["nlights"], st["lights"])
    o.fjo_camera(sc, C.byref(st["camera"]))
    for target, (T4, R4, S4) in st.get("samples", {}).items():
        assert o.fjo_transform_samples(sc, target, 0, 10, len(T4), dptr(T4), len(R4), dptr(R4), len(S4), dptr(S4)) == 0
    if "time_range" in st:
        o.fjo_time_range(sc, st["time_range"][0], st["time_range"][1])
    o.fjo_build(sc)
    return sc


def oracle_render(desc, rng_mode=0, threads=8, region=None, st=None, tiles=None):
    """Renders with the oracle.  Returns (image [H,W,4] float32, Stats).  `tiles`: explicit (id, xmin, ymin, xmax, ymax)
    list — e.g. a subset of the full frame's tiles WITH their full-frame ids (the counter RNG is keyed by tile id)."""
    a = _abi()
    st = st or desc.to_structs()
    sc = oracle_scene(st)
    tiles = tiles if tiles is not None else desc.tiles(region)
    ta = SceneDesc.tile_array(tiles)
    p = st["params"]
    img = np.zeros((p.yres, p.xres, 4), np.float32)
    stats = a.Stats()
    rc = oracle().fjo_render(sc, C.byref(p), ta, len(tiles), fptr(img), rng_mode, threads, C.byref(stats))
    oracle().fjo_scene_free(sc)
    assert rc == 0
    return img, stats


def reference_render(desc, workdir, threads=1, region=None, timeout=3600):
    """Runs the unmodified reference (oracle/_ref/bin/ref_probe run).  Returns (image, frame_seconds)."""
    from fujiyama_renderer_b200 import fbio
    os.makedirs(workdir, exist_ok=True)
    fb = os.path.join(workdir, "ref_out.fb")
    scn = os.path.join(workdir, "scene.scn")
    with open(scn, "w") as f:
        f.write(desc.to_scn(workdir, fb, threads=threads, region=region))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(REF_DIR, "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    res = subprocess.run([os.path.join(REF_DIR, "bin", "ref_probe"), "run", scn], env=env, capture_output=True,
                         text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError("reference failed: %s\n%s" % (res.stdout[-2000:], res.stderr[-2000:]))
    secs = [float(l.split()[1]) for l in res.stdout.split("\n") if l.startswith("FJ_FRAME_SECONDS")]
    return fbio.read_fb(fb), (secs[-1] if secs else None)


# ---- canned scenes (BASELINE.json configs, scaled) -------------------------------------------------
def scene_cube(res=256, rate=1):
    """Config 1: a 24-vertex/12-triangle cube like scenes/cube.ply, constant shader (values of scenes/cube.cc:97-98,136)."""
    from fujiyama_renderer_b200 import synth
    d = SceneDesc()
    P, idx = synth.cube()
    d.mesh("mesh1", P, idx)
    d.shader("sh1", "constant", diffuse=(.2, .4, .8))
    d.instance("obj1", "mesh1", "sh1", R=(0, 10, 0))
    d.cam.update(T=(3, 3, 3), R=(-35.264389682754654, 45, 0))
    d.ren.update(resolution=(res, res), pixelsamples=(rate, rate))
    return d


def scene_blob_plastic(n=24, res=(160, 90), rate=2, reflect=(1, 1, 1), nlights=1):
    """Config 2 stand-in: bumpy sphere, plastic shader, point light(s)."""
    from fujiyama_renderer_b200 import synth
    d = SceneDesc()
    P, idx = synth.blob(n)
    d.mesh("blob", P, idx)
    d.shader("sh1", "plastic", diffuse=(.7, .5, .3), reflect=reflect)
    d.instance("obj1", "blob", "sh1", R=(20, 30, 0))
    pos = [(5, 12, 5), (-6, 8, 3), (2, -7, 6), (0, 10, -8)]
    for i in range(nlights):
        d.light(0, T=pos[i % 4], intensity=1.0 / nlights)
    d.ren.update(resolution=res, pixelsamples=(rate, rate))
    return d


def scene_blob_pathtracing(n=24, res=(96, 54), rate=2, depth=3, reflect=(0, 0, 0), refract=(0, 0, 0)):
    """North-star stand-in: bumpy sphere inside an emissive shell, pathtracing shader."""
    from fujiyama_renderer_b200 import synth
    d = SceneDesc()
    P, idx = synth.blob(n)
    d.mesh("blob", P, idx)
    Ps, idxs = synth.blob(max(8, n // 3))
    d.mesh("shell", Ps, idxs)
    d.shader("sh1", "pathtracing", diffuse=(.8, .6, .4), emission=(.05, .05, .05), reflect=reflect, refract=refract)
    d.shader("sh2", "pathtracing", diffuse=(.2, .2, .2), emission=(1.0, .9, .8))
    d.instance("obj1", "blob", "sh1", R=(20, 30, 0))
    d.instance("shell1", "shell", "sh2", S=(8, 8, 8))
    d.ren.update(resolution=res, pixelsamples=(rate, rate), max_diffuse_depth=depth)
    return d
