"""The BASELINE.json configurations as test-kit scene descriptions: the SAME scenes `fujiyama-renderer_b200/scenes.py`
writes as `.scn` text for bench.py (same seeded synthetic meshes, transforms, shader and renderer properties), so the
oracle can render a region of exactly the frame the bench times.  Used by the full-size parity tests
(tests/test_fullsize_gpu.py) and by bench.py's `parity` leg (checker only, never the thing measured)."""
import scenekit as sk


def _synth():
    sk.pkg()
    from fujiyama_renderer_b200 import synth
    return synth


def pathtracing_blob(n=707, res=(1920, 1080), rate=8, depth=3, shell_n=64, motion=False):
    """scenes.pathtracing_blob: S-blob(n) with pathtracing_shader inside the emissive shell (north star, config 3)."""
    synth = _synth()
    d = sk.SceneDesc()
    d.mesh("blob", *synth.blob(n))
    d.mesh("shell", *synth.blob(shell_n))
    d.shader("sh1", "pathtracing", diffuse=(.8, .6, .4), emission=(.05, .05, .05))
    d.shader("sh2", "pathtracing", diffuse=(.2, .2, .2), emission=(1.0, .9, .8))
    d.instance("obj1", "blob", "sh1", R=(20, 30, 0))
    d.instance("shell1", "shell", "sh2", S=(8, 8, 8))
    if motion:
        d.instances[0]["samples"] = dict(R=[(20, 30, 0, 0.0), (20, 45, 0, 1.0)])
        d.cam["samples"] = dict(T=[(0, 0, 4.5, 0.0), (0.1, 0, 4.45, 1.0)])
    d.ren.update(resolution=tuple(res), pixelsamples=(rate, rate), max_diffuse_depth=depth)
    return d


def plastic_blob(n=187, res=(1280, 720), rate=4):
    """scenes.plastic_blob: BASELINE config 2 (69 938 triangles, plastic_shader with the mirror bounce, one point light)."""
    synth = _synth()
    d = sk.SceneDesc()
    d.mesh("blob", *synth.blob(n))
    d.shader("sh1", "plastic", diffuse=(.7, .5, .3))
    d.instance("obj1", "blob", "sh1", R=(20, 30, 0))
    d.light(0, T=(5, 12, 5))
    d.ren.update(resolution=tuple(res), pixelsamples=(rate, rate))
    return d


def instanced_blobs(n=740, res=(1920, 1080), rate=8, grid_samples=16):
    """scenes.instanced_blobs: BASELINE config 4 (16 instances of a 1.09 M-triangle mesh + floor, plastic, GridLight)."""
    synth = _synth()
    d = sk.SceneDesc()
    d.mesh("blob", *synth.blob(n))
    d.mesh("floor", *synth.quad(12.0, -0.7))
    d.shader("sh1", "plastic", diffuse=(.7, .5, .3))
    d.shader("sh2", "plastic", diffuse=(.6, .6, .6), reflect=(0, 0, 0))
    k = 0
    for i in range(4):
        for j in range(4):
            d.instance("obj%d" % k, "blob", "sh1", T=(2.25 - 1.5 * i, 0, 2.25 - 1.5 * j), R=(0, 30 * k, 0), S=(.6, .6, .6))
            k += 1
    d.instance("floor1", "floor", "sh2")
    d.light(1, T=(0, 8, 2), R=(180, 0, 0), S=(4, 1, 4), intensity=1.5, sample_count=grid_samples)
    d.cam.update(T=(0, 4, 9), R=(-25, 0, 0), fov=40.0)
    d.ren.update(resolution=tuple(res), pixelsamples=(rate, rate))
    return d


def pathtracing_soup(ntris=10_000_000, res=(3840, 2160), rate=16, depth=8, shell_n=64):
    """scenes.pathtracing_soup: BASELINE config 5 (S-random triangle soup, 8 diffuse bounces, inside the shell)."""
    synth = _synth()
    d = sk.SceneDesc()
    d.mesh("soup", *synth.random_tris(ntris, seed=1234))
    d.mesh("shell", *synth.blob(shell_n))
    d.shader("sh1", "pathtracing", diffuse=(.8, .6, .4), emission=(.05, .05, .05))
    d.shader("sh2", "pathtracing", diffuse=(.2, .2, .2), emission=(1.0, .9, .8))
    d.instance("obj1", "soup", "sh1", R=(20, 30, 0))
    d.instance("shell1", "shell", "sh2", S=(8, 8, 8))
    d.cam.update(T=(0, 0, 2.2))
    d.ren.update(resolution=tuple(res), pixelsamples=(rate, rate), max_diffuse_depth=depth)
    return d


BUILDERS = {"pathtracing_blob": pathtracing_blob, "plastic_blob": plastic_blob, "instanced_blobs": instanced_blobs,
            "pathtracing_soup": pathtracing_soup}


def desc_for(builder, kw):
    return BUILDERS[builder](**kw)
