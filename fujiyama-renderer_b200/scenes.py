"""Scene-description text (`.scn`, the grammar of tools/scene_parser) for the BASELINE.json configurations,
built on seeded synthetic meshes (the reference's own assets are not shipped).

The same text drives both renderers: libfjscene (GPU) and the unmodified reference's bin/scene — only the
plugin directory differs (the reference dlopens the DSOs; libfjscene keys its device shaders on the name).
"""
import os

from . import synth

PLUGINS = {"constant": ("constant_shader", "ConstantShader"), "plastic": ("plastic_shader", "PlasticShader"),
           "pathtracing": ("pathtracing_shader", "PathtracingShader"), "glass": ("glass_shader", "GlassShader")}


def _mesh_cmds(name, ply):
    return ["NewMesh %s" % name, "NewProcedure %s_proc stanfordply_procedure" % name,
            "AssignMesh %s_proc mesh %s" % (name, name), "SetStringProperty %s_proc filepath %s" % (name, ply),
            "SetStringProperty %s_proc io_mode r" % name, "RunProcedure %s_proc" % name]


def _renderer_cmds(res, rate, depth, threads):
    return ["NewFrameBuffer fb1 rgba", "NewRenderer ren1", "AssignCamera ren1 cam1", "AssignFrameBuffer ren1 fb1",
            "SetProperty2 ren1 resolution %d %d" % tuple(res), "SetProperty2 ren1 pixelsamples %d %d" % (rate, rate),
            "SetProperty1 ren1 max_diffuse_depth %d" % depth,
            "SetProperty1 ren1 use_max_thread 0", "SetProperty1 ren1 thread_count %d" % threads]


def ensure_ply(workdir, name, maker):
    os.makedirs(workdir, exist_ok=True)
    path = os.path.join(workdir, name + ".ply")
    if not os.path.exists(path):
        P, idx = maker()
        synth.write_ply(path + ".tmp", P, idx)
        os.replace(path + ".tmp", path)
    return path


def pathtracing_blob(workdir, plugin_dir, n=707, res=(1920, 1080), rate=8, depth=3, threads=1, shell_n=64, motion=False):
    """North star: ~1M-triangle bumpy sphere (S-blob(707) = 999 698 triangles) with pathtracing_shader inside an
    emissive shell (the shader ignores lights; scenes/pathtracing.py:78-88 does the same),
    1920x1080, 8x8 = 64 spp, max_diffuse_depth 3.  Returns the set-up commands (no RenderScene)."""
    blob = ensure_ply(workdir, "blob_%d" % n, lambda: synth.blob(n))
    shell = ensure_ply(workdir, "shell_%d" % shell_n, lambda: synth.blob(shell_n))
    L = ["OpenPlugin pathtracing_shader %s" % os.path.join(plugin_dir, "PathtracingShader"),
         "OpenPlugin stanfordply_procedure %s" % os.path.join(plugin_dir, "StanfordPlyProcedure"),
         "NewCamera cam1 PerspectiveCamera", "SetProperty3 cam1 translate 0 0 4.5",
         "NewShader sh1 pathtracing_shader", "SetProperty3 sh1 diffuse 0.8 0.6 0.4", "SetProperty3 sh1 emission 0.05 0.05 0.05",
         "NewShader sh2 pathtracing_shader", "SetProperty3 sh2 diffuse 0.2 0.2 0.2", "SetProperty3 sh2 emission 1.0 0.9 0.8"]
    L += _mesh_cmds("blob", blob) + _mesh_cmds("shell", shell)
    L += ["NewObjectInstance obj1 blob", "SetProperty3 obj1 rotate 20 30 0", "AssignShader obj1 DEFAULT_SHADING_GROUP sh1",
          "NewObjectInstance shell1 shell", "SetProperty3 shell1 scale 8 8 8", "AssignShader shell1 DEFAULT_SHADING_GROUP sh2"]
    if motion:       # motion blur (scenes/transform_motion_blur.py): the blob spins 15 degrees and the camera dollies during the shutter
        L += ["SetSampleProperty3 obj1 rotate 20 30 0 0", "SetSampleProperty3 obj1 rotate 20 45 0 1",
              "SetSampleProperty3 cam1 translate 0 0 4.5 0", "SetSampleProperty3 cam1 translate 0.1 0 4.45 1"]
    L += _renderer_cmds(res, rate, depth, threads)
    return "\n".join(L) + "\n"


def plastic_blob(workdir, plugin_dir, n=187, res=(1280, 720), rate=4, threads=1):
    """BASELINE config 2 stand-in: 69 938-triangle bumpy sphere, plastic_shader (mirror bounce on), one point light."""
    blob = ensure_ply(workdir, "blob_%d" % n, lambda: synth.blob(n))
    L = ["OpenPlugin plastic_shader %s" % os.path.join(plugin_dir, "PlasticShader"),
         "OpenPlugin stanfordply_procedure %s" % os.path.join(plugin_dir, "StanfordPlyProcedure"),
         "NewCamera cam1 PerspectiveCamera", "SetProperty3 cam1 translate 0 0 4.5",
         "NewLight light1 PointLight", "SetProperty3 light1 translate 5 12 5",
         "NewShader sh1 plastic_shader", "SetProperty3 sh1 diffuse 0.7 0.5 0.3"]
    L += _mesh_cmds("blob", blob)
    L += ["NewObjectInstance obj1 blob", "SetProperty3 obj1 rotate 20 30 0", "AssignShader obj1 DEFAULT_SHADING_GROUP sh1"]
    L += _renderer_cmds(res, rate, 3, threads)
    return "\n".join(L) + "\n"


def instanced_blobs(workdir, plugin_dir, n=740, res=(1920, 1080), rate=8, threads=1, grid_samples=16):
    """BASELINE config 4 stand-in: 16 instances of one ~1.09 M-triangle mesh on the 4x4 grid of scenes/happy_buddhas.scn
    (translate -1.5 i, 0, -1.5 j; rotate 0, 30 k, 0; scale .6) over a floor, plastic_shader with the mirror bounce on, one
    GridLight with `grid_samples` samples: a TLAS with inner nodes over a shared BLAS, 16 shadow rays per hit."""
    blob = ensure_ply(workdir, "blob_%d" % n, lambda: synth.blob(n))
    floor = ensure_ply(workdir, "floor_12", lambda: synth.quad(12.0, -0.7))
    L = ["OpenPlugin plastic_shader %s" % os.path.join(plugin_dir, "PlasticShader"),
         "OpenPlugin stanfordply_procedure %s" % os.path.join(plugin_dir, "StanfordPlyProcedure"),
         "NewCamera cam1 PerspectiveCamera", "SetProperty3 cam1 translate 0 4 9", "SetProperty3 cam1 rotate -25 0 0", "SetProperty1 cam1 fov 40",
         "NewLight light1 GridLight", "SetProperty3 light1 translate 0 8 2", "SetProperty3 light1 rotate 180 0 0", "SetProperty3 light1 scale 4 1 4",
         "SetProperty1 light1 intensity 1.5", "SetProperty1 light1 sample_count %d" % grid_samples,
         "NewShader sh1 plastic_shader", "SetProperty3 sh1 diffuse 0.7 0.5 0.3",
         "NewShader sh2 plastic_shader", "SetProperty3 sh2 diffuse 0.6 0.6 0.6", "SetProperty3 sh2 reflect 0 0 0"]
    L += _mesh_cmds("blob", blob) + _mesh_cmds("floor", floor)
    k = 0
    for i in range(4):
        for j in range(4):
            L += ["NewObjectInstance obj%d blob" % k, "SetProperty3 obj%d translate %r 0 %r" % (k, 2.25 - 1.5 * i, 2.25 - 1.5 * j),
                  "SetProperty3 obj%d rotate 0 %d 0" % (k, 30 * k), "SetProperty3 obj%d scale 0.6 0.6 0.6" % k,
                  "AssignShader obj%d DEFAULT_SHADING_GROUP sh1" % k]
            k += 1
    L += ["NewObjectInstance floor1 floor", "AssignShader floor1 DEFAULT_SHADING_GROUP sh2"]
    L += _renderer_cmds(res, rate, 3, threads)
    return "\n".join(L) + "\n"


def pathtracing_soup(workdir, plugin_dir, ntris=10_000_000, res=(3840, 2160), rate=16, depth=8, threads=1, shell_n=64):
    """BASELINE config 5 stand-in: S-random triangle soup (10 M triangles) with pathtracing_shader, 8 diffuse bounces, inside
    the emissive shell of the north star; 3840x2160, 16x16 = 256 spp."""
    soup = ensure_ply(workdir, "soup_%d" % ntris, lambda: synth.random_tris(ntris, seed=1234))
    shell = ensure_ply(workdir, "shell_%d" % shell_n, lambda: synth.blob(shell_n))
    L = ["OpenPlugin pathtracing_shader %s" % os.path.join(plugin_dir, "PathtracingShader"),
         "OpenPlugin stanfordply_procedure %s" % os.path.join(plugin_dir, "StanfordPlyProcedure"),
         "NewCamera cam1 PerspectiveCamera", "SetProperty3 cam1 translate 0 0 2.2",
         "NewShader sh1 pathtracing_shader", "SetProperty3 sh1 diffuse 0.8 0.6 0.4", "SetProperty3 sh1 emission 0.05 0.05 0.05",
         "NewShader sh2 pathtracing_shader", "SetProperty3 sh2 diffuse 0.2 0.2 0.2", "SetProperty3 sh2 emission 1.0 0.9 0.8"]
    L += _mesh_cmds("soup", soup) + _mesh_cmds("shell", shell)
    L += ["NewObjectInstance obj1 soup", "SetProperty3 obj1 rotate 20 30 0", "AssignShader obj1 DEFAULT_SHADING_GROUP sh1",
          "NewObjectInstance shell1 shell", "SetProperty3 shell1 scale 8 8 8", "AssignShader shell1 DEFAULT_SHADING_GROUP sh2"]
    L += _renderer_cmds(res, rate, depth, threads)
    return "\n".join(L) + "\n"


def center_region(res, tile, ntiles_x, ntiles_y):
    """A tile-aligned block of ntiles_x x ntiles_y tiles around the image centre, as render_region (xmin ymin xmax ymax)."""
    tx, ty = -(-res[0] // tile), -(-res[1] // tile)
    nx, ny = min(ntiles_x, tx), min(ntiles_y, ty)
    x0, y0 = (tx - nx) // 2, (ty - ny) // 2
    return (x0 * tile, y0 * tile, min((x0 + nx) * tile, res[0]), min((y0 + ny) * tile, res[1]))


def region_camera_samples(region, rate, tile=32, fwidth=2.0):
    """Camera samples (incl. filter-margin samples) the sampler generates for the tiles of `region`
    (count_samples_in_margin, src/fj_fixed_grid_sampler.cc:131-146)."""
    import math
    m = int(math.ceil((fwidth - 1) * rate * .5))
    x0, y0, x1, y1 = region
    n = 0
    for ty in range(y0 // tile, -(-y1 // tile)):
        for tx in range(x0 // tile, -(-x1 // tile)):
            w = min((tx + 1) * tile, x1) - max(tx * tile, x0)
            h = min((ty + 1) * tile, y1) - max(ty * tile, y0)
            n += (rate * w + 2 * m) * (rate * h + 2 * m)
    return n
