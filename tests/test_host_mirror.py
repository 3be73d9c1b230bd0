"""libfjscene.so — the host mirror of fj_scene_interface: (CPU) its matrices and normals are bit-identical to
the reference's (golden vectors dumped from the reference's libscene.so), the `.scn` grammar and error
behaviour follow tools/scene_parser; (GPU) a `.scn` file written for the reference's bin/scene renders through
it to the same frame as the reference's own .fb."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import golden_scenes
import scenekit as sk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fuji():
    import __graft_entry__ as g
    g.build_fjgpu()
    g.build_host()
    sk.pkg()
    from fujiyama_renderer_b200 import fujiyama
    fujiyama.load_fjscene()
    return fujiyama


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLD, "ref_vectors.json")) as f:
        return json.load(f)


def test_transform_bit_exact_vs_reference(fuji, vec):
    lib = fuji.load_fjscene()
    for c in vec["xfm"]:
        fwd, inv = np.zeros(16), np.zeros(16)
        T, R, S = (np.asarray(c[k], np.float64) for k in "TRS")
        lib.fjscene_make_transform(c["torder"], c["rorder"], sk.dptr(T), sk.dptr(R), sk.dptr(S), sk.dptr(fwd), sk.dptr(inv))
        assert fwd.tolist() == c["fwd"] and inv.tolist() == c["inv"]


def test_compute_normals_bit_exact_vs_reference(fuji, vec):
    c = vec["normals"]
    with fuji.Session() as s:
        s.run("NewMesh m\n")
        P = np.asarray(c["P"], np.float64)
        s.set_mesh("m", P, np.asarray(c["idx"], np.int32))
        N = np.zeros_like(P)
        assert s.lib.fjscene_mesh_normals(s.id("m"), sk.dptr(N), len(P)) == 0
    assert N.tolist() == c["N"]


def test_ids_and_errors_follow_the_reference(fuji):
    with fuji.Session() as s:
        s.run("NewMesh mesh1\nNewObjectInstance obj1 mesh1\nNewRenderer ren1\nNewCamera cam1 PerspectiveCamera\n")
        # ID = type * 10000000 + index with the reference's type numbering (src/fj_scene_interface.cc:42-62)
        assert s.id("mesh1") == 16 * 10000000 and s.id("obj1") == 1 * 10000000 and s.id("ren1") == 8 * 10000000
        assert s.id("cam1") == 10 * 10000000
        for bad, msg in [("Frobnicate x", "unknown command"), ("NewMesh", "too few arguments"),
                         ("NewMesh a b", "too many arguments"), ("NewMesh mesh1", "entry name already exists"),
                         ("RenderScene nope", "entry name not found"), ("SetProperty1 ren1 cast_shadow abc", "bad number arguments"),
                         ("NewLight l1 LaserLight", "bad enum arguments"),
                         ("OpenPlugin p /x/HairShader.so", "plugin not found"),
                         ("NewVolume v", "new entry failed"),
                         ("SetProperty1 ren1 no_such_property 1", "command failed")]:
            with pytest.raises(fuji.SceneError, match=msg):
                s.run(bad + "\n")
        s.run("# a comment\n\nSetProperty3 obj1 rotate_order ORDER_XYZ 0 0\nSetProperty2 ren1 resolution 64 48\n")
        # rendering without a camera / framebuffer fails like the reference's SI_FAIL paths
        with pytest.raises(fuji.SceneError):
            s.run("RenderScene ren1\n")


def test_ply_reader_matches_python_reader(fuji, tmp_path):
    from fujiyama_renderer_b200 import synth
    P, idx = synth.blob(9)
    path = str(tmp_path / "b.ply")
    synth.write_ply(path, P, idx)
    ascii_path = str(tmp_path / "quad.ply")
    with open(ascii_path, "w") as f:      # ascii polygon (fan triangulation, ply2mesh.cc:129-136) with an extra property
        f.write("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nproperty float q\n"
                "element face 1\nproperty list uchar int vertex_indices\nend_header\n"
                "0 0 0 9\n1 0 0 9\n1 1 0 9\n0 1 0 9\n4 0 1 2 3\n")
    with fuji.Session() as s:
        s.run("OpenPlugin ply /any/where/StanfordPlyProcedure\nNewMesh m\nNewProcedure p ply\nAssignMesh p mesh m\n"
              "SetStringProperty p filepath %s\nSetStringProperty p io_mode r\nRunProcedure p\n" % path)
        N = np.zeros((len(P), 3))
        assert s.lib.fjscene_mesh_normals(s.id("m"), sk.dptr(N), len(P)) == 0
        ref = np.zeros_like(N)
        P64 = np.ascontiguousarray(P.astype(np.float64))
        sk.oracle().fjo_compute_normals(sk.dptr(P64), len(P64), sk.iptr(np.ascontiguousarray(idx.reshape(-1))), len(idx), sk.dptr(ref))
        assert np.array_equal(N, ref)
        s.run("NewMesh q\nNewProcedure p2 ply\nAssignMesh p2 mesh q\nSetStringProperty p2 filepath %s\nRunProcedure p2\n" % ascii_path)
        Nq = np.zeros((4, 3))
        assert s.lib.fjscene_mesh_normals(s.id("q"), sk.dptr(Nq), 4) == 0
        assert np.allclose(Nq, [[0, 0, 1]] * 4)
        with pytest.raises(fuji.SceneError):
            s.run("SetStringProperty p2 filepath /no/such/file.ply\nRunProcedure p2\n")


def test_python_shim_emits_reference_grammar(fuji):
    si = fuji.SceneInterface(argv=["-R", "64", "48"])
    si.OpenPlugin("PlasticShader", "${FJ}/PlasticShader")
    si.NewMesh("mesh1")
    si.SetProperty3("obj1", "rotate", 0, 10, 0)
    si.AssignShader("obj1", "DEFAULT_SHADING_GROUP", "sh1")
    si.RenderScene("ren1")
    assert si.commands == ["OpenPlugin PlasticShader ${FJ}/PlasticShader.so", "NewMesh mesh1",
                           "SetProperty3 obj1 rotate 0 10 0", "AssignShader obj1 DEFAULT_SHADING_GROUP sh1",
                           "SetProperty2 ren1 resolution 64 48", "RenderScene ren1"]
    with pytest.raises(TypeError):
        si.NewMesh("a", "b")


@pytest.mark.parametrize("name", list(golden_scenes.SCENES))
def test_flattened_scene_matches_the_test_kits(fuji, tmp_path, name):
    """What SiRenderScene would hand to libfjgpu (fjscene_flatten, no device involved) against the test kit's independent
    flattening of the same scene description (SceneDesc.to_structs, which feeds the oracle): instances, shaders, lights
    incl. the dome sample tables, camera, frame parameters and tiles, field by field."""
    abi = sk.abi
    desc = golden_scenes.SCENES[name]()
    st = desc.to_structs()
    scn = desc.to_scn(str(tmp_path), None, plugin_dir="/x")
    scn = "\n".join(l for l in scn.split("\n") if not l.startswith(("RenderScene", "SaveFrameBuffer"))) + "\n"
    with fuji.Session() as s:
        s.run(scn)
        lib = s.lib
        n = [C.c_int32() for _ in range(4)]
        assert lib.fjscene_flatten(C.c_long(s.id("ren1")), *[C.byref(x) for x in n]) == 0
        ninst, nlights, nshaders, ntiles = (x.value for x in n)
        assert (ninst, nlights, nshaders) == (st["ninstances"], st["nlights"], st["nshaders"])
        for i in range(ninst):
            o, r = abi.Instance(), st["instances"][i]
            assert lib.fjscene_flat_instance(i, C.byref(o)) == 0
            assert o.mesh_id == r.mesh_id and o.shader_of_group[0] == r.shader_of_group[0]
            assert (o.reflect_target, o.refract_target, o.shadow_target) == (r.reflect_target, r.refract_target, r.shadow_target)
            assert list(o.fwd) == list(r.fwd) and list(o.inv) == list(r.inv)
        for i in range(nshaders):
            o, r = abi.Shader(), st["shaders"][i]
            assert lib.fjscene_flat_shader(i, C.byref(o)) == 0
            assert bytes(o) == bytes(r), (name, i)
        for i in range(nlights):
            o, r = abi.Light(), st["lights"][i]
            assert lib.fjscene_flat_light(i, C.byref(o)) == 0
            assert (o.kind, o.sample_count, o.double_sided, o.dome_sample_count) == (r.kind, r.sample_count, r.double_sided, r.dome_sample_count)
            assert list(o.color) == list(r.color) and o.intensity == r.intensity
            assert list(o.translate) == list(r.translate) and list(o.fwd) == list(r.fwd)
            k = o.dome_sample_count
            if k:
                assert [o.dome_dirs[j] for j in range(3 * k)] == [r.dome_dirs[j] for j in range(3 * k)]
                assert [o.dome_colors[j] for j in range(3 * k)] == [r.dome_colors[j] for j in range(3 * k)]
        cam, p = abi.Camera(), abi.RenderParams()
        assert lib.fjscene_flat_frame(C.byref(cam), C.byref(p)) == 0
        rc, rp = st["camera"], st["params"]
        assert list(cam.fwd) == list(rc.fwd) and (cam.fov, cam.znear, cam.zfar) == (rc.fov, rc.znear, rc.zfar)
        for f in ("xres", "yres", "xrate", "yrate", "xfwidth", "yfwidth", "jitter", "max_diffuse_depth", "max_reflect_depth",
                  "max_refract_depth", "cast_shadow", "target_group"):
            assert getattr(p, f) == getattr(rp, f), f
        tiles = desc.tiles()
        assert ntiles == len(tiles)
        for i, t in enumerate(tiles):
            o = abi.Tile()
            assert lib.fjscene_flat_tile(i, C.byref(o)) == 0
            assert (o.id, o.xmin, o.ymin, o.xmax, o.ymax) == tuple(t)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cube_c1", "plastic", "multi", "dome_light", "glass", "textured", "dome_envmap", "motion_blur"])
def test_scn_file_renders_like_the_reference(fuji, tmp_path, name):
    """The exact command text the reference's bin/scene was given for the golden .fb, fed to libfjscene."""
    from fujiyama_renderer_b200 import fbio
    desc = golden_scenes.SCENES[name]()
    out = str(tmp_path / "out.fb")
    scn = desc.to_scn(str(tmp_path), out, threads=1, plugin_dir="/opt/fujiyama/lib")
    with fuji.Session() as s:
        s.run(scn)
        img = s.framebuffer("fb1")
        stats, info, _ = s.stats()
    ref = np.load(os.path.join(GOLD, "ref_images.npz"))[name]
    assert img.shape == ref.shape
    d = img.astype(np.float64) - ref
    assert np.sqrt((d * d).reshape(-1, 4).mean(0)).max() < 1e-4
    assert stats.rays_camera == stats.camera_samples > 0
    # the .fb it saved parses back to the same pixels at 6 significant digits
    fb = fbio.read_fb(out)
    assert np.allclose(fb, img, rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_scn_pathtracing_matches_oracle_and_groups(fuji, tmp_path):
    desc = golden_scenes.SCENES["pt_branching"]()
    scn = desc.to_scn(str(tmp_path), None, plugin_dir="/x")
    with fuji.Session() as s:
        s.run(scn)
        img = s.framebuffer("fb1")
    ref, _ = sk.oracle_render(desc, rng_mode=0, threads=8)
    d = img.astype(np.float64) - ref
    assert np.sqrt((d * d).mean()) < 1e-4


@pytest.mark.gpu
def test_tile_sharding_ranks_partition_the_frame(fuji, tmp_path):
    desc = golden_scenes.SCENES["plastic"]()
    scn = desc.to_scn(str(tmp_path), None, plugin_dir="/x")
    with fuji.Session() as s:
        s.run(scn)
        full = s.framebuffer("fb1")
    acc = np.zeros_like(full)
    for rank in range(3):
        with fuji.Session(rank=rank, world_size=3) as s:
            s.run(scn)
            part = s.framebuffer("fb1")
        assert not (acc.astype(bool) & part.astype(bool)).any()
        acc += part
    assert np.array_equal(acc, full)


@pytest.mark.gpu
def test_fjscene_cli_renders_on_several_gpus_without_python(tmp_path):
    """`FJ_GPU_COUNT=N fjscene file.scn`: ONE process, one context per GPU, fjgpu_render_frame_multi (tiles round-robin, one
    ncclAllGather, rank 0 assembles) — the same .fb as one GPU, for a deterministic and a path-traced scene.  Needs >= 2 GPUs."""
    import subprocess
    torch = pytest.importorskip("torch")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from fujiyama_renderer_b200 import fbio
    exe = os.path.join(sk.REPO, "fujiyama-renderer_b200", "host", "fjscene")
    for name in ("multi", "pt_branching"):
        desc = golden_scenes.SCENES[name]()
        frames = {}
        for count in (1, 2, min(n, 4)):
            out = str(tmp_path / ("%s_%d.fb" % (name, count)))
            scn = str(tmp_path / ("%s_%d.scn" % (name, count)))
            with open(scn, "w") as f:
                f.write(desc.to_scn(str(tmp_path), out, threads=1, plugin_dir="/opt/fujiyama/lib"))
            env = dict(os.environ, FJ_GPU_COUNT=str(count))
            res = subprocess.run([exe, scn], env=env, capture_output=True, text=True, timeout=600)
            assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
            if count > 1:
                assert "one all-gather" in res.stdout
            frames[count] = fbio.read_fb(out)
        for count, img in frames.items():
            assert np.array_equal(img, frames[1]), (name, count)
