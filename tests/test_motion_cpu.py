"""Motion blur, host side — no GPU needed: the frame's time table of libfjgpu, the transform
interpolation of libfjscene and of the test kit, all bit for bit against the oracle's restatement of
FixedGridSampler::generate_samples (time draw) and XfmLerpTransformSample.  The oracle itself is pinned on the
reference's frame of the golden scene `motion_blur` (tests/test_oracle_cpu.py)."""
import ctypes as C

import numpy as np
import pytest

import golden_scenes
import scenekit as sk


@pytest.fixture(scope="module")
def libs():
    import __graft_entry__ as g
    g.build_fjgpu()
    g.build_host()
    sk.pkg()
    from fujiyama_renderer_b200 import abi, fujiyama
    return abi, abi.load_fjgpu(), fujiyama


def params(abi, res, rate, fw):
    p = abi.RenderParams()
    p.xres, p.yres = res
    p.xrate, p.yrate = rate
    p.xfwidth, p.yfwidth = fw
    p.jitter = 1.0
    return p


@pytest.mark.parametrize("res,rate,fw,tile,rng", [((64, 48), (2, 2), (2, 2), 32, (0.0, 1.0)), ((70, 50), (3, 1), (2.5, 1.0), 32, (0.1, 0.9)),
                                                  ((33, 17), (1, 4), (4, 2), 16, (-2.0, 3.5)), ((40, 40), (2, 2), (2, 2), 64, (0.5, 0.5))])
def test_time_table_is_the_samplers_time_draw(libs, res, rate, fw, tile, rng):
    abi, lib, _ = libs
    p = params(abi, res, rate, fw)
    tiles = sk.make_tiles(res[0], res[1], tile)
    ta = sk.SceneDesc.tile_array(tiles)
    n = lib.fjgpu_time_table(C.byref(p), ta, len(tiles), rng[0], rng[1], None, 0)
    times = np.full(n + 3, -7.0)
    assert lib.fjgpu_time_table(C.byref(p), ta, len(tiles), rng[0], rng[1], sk.dptr(times), n) == n
    assert (times[n:] == -7.0).all()
    o = sk.oracle()
    biggest = 0
    for i in range(len(tiles)):          # the k-th sample of EVERY tile draws entry k; the table is as long as the largest tile
        buf = np.zeros(n)
        m = o.fjo_sample_times(C.byref(p), C.byref(ta[i]), rng[0], rng[1], sk.dptr(buf), n)
        assert 0 < m <= n and np.array_equal(buf[:m], times[:m])
        biggest = max(biggest, m)
    assert biggest == n
    assert (times[:n] >= rng[0]).all() and (times[:n] <= rng[1]).all()
    # a short buffer is filled as far as it goes, the count is still returned
    short = np.full(5, -7.0)
    assert lib.fjgpu_time_table(C.byref(p), ta, len(tiles), rng[0], rng[1], sk.dptr(short), 3) == n
    assert np.array_equal(short[:3], times[:3]) and (short[3:] == -7.0).all()
    assert lib.fjgpu_time_table(None, ta, len(tiles), 0.0, 1.0, None, 0) < 0


def test_transform_interpolation_bit_exact(libs):
    """libfjscene (what SiRenderScene tabulates) and the test kit (what the C-ABI tests tabulate) against the oracle."""
    _, _, fuji = libs
    o = sk.oracle()
    rs = np.random.RandomState(5)
    keysets = [[0.0], [0.0, 1.0], [0.0, 0.4, 1.0], [-1.0, 0.25, 0.5, 2.0], [0.3, 0.31]]
    with fuji.Session() as s:
        s.run("NewMesh m\nNewCamera cam PerspectiveCamera\nNewLight l PointLight\n")
        for case in range(12):
            kT, kR, kS = (keysets[rs.randint(len(keysets))] for _ in range(3))
            T = [tuple(rs.uniform(-3, 3, 3)) + (t,) for t in kT]
            R = [tuple(rs.uniform(-180, 180, 3)) + (t,) for t in kR]
            S = [tuple(rs.uniform(.3, 2, 3)) + (t,) for t in kS]
            name = "o%d" % case
            s.run("NewObjectInstance %s m\n" % name)
            for prop, rows in (("translate", T), ("rotate", R), ("scale", S)):
                for r in rows[::-1]:            # pushed in reverse: PropPushSample keeps them sorted by time
                    s.run("SetSampleProperty3 %s %s %r %r %r %r\n" % ((name, prop) + tuple(float(x) for x in r)))
            T4, R4, S4 = sk.sample_rows(T, None), sk.sample_rows(R, None), sk.sample_rows(S, None, (1.0, 1.0, 1.0))   # incl. the initial time-0 sample
            for time in [-2.0, 0.0, 0.1, 0.25, 0.3, 0.305, 0.4, 0.77, 1.0, 2.5] + list(rs.uniform(-1, 2, 6)):
                fo, io = np.zeros(16), np.zeros(16)
                assert o.fjo_lerp_transform(0, 10, len(T4), sk.dptr(T4), len(R4), sk.dptr(R4), len(S4), sk.dptr(S4), time, sk.dptr(fo), sk.dptr(io)) == 0
                fh, ih = np.zeros(16), np.zeros(16)
                assert s.lib.fjscene_lerp_transform(s.id(name), time, sk.dptr(fh), sk.dptr(ih)) == 0
                assert np.array_equal(fh, fo) and np.array_equal(ih, io)
                fk, ik = sk.motion_table(T4, R4, S4, [time])
                assert np.array_equal(fk[0], fo) and np.array_equal(ik[0], io)
        # a later sample at an existing time replaces it; a ninth key is refused (8 per channel, src/fj_property.h:117)
        s.run("NewObjectInstance q m\nSetSampleProperty3 q translate 1 2 3 0.5\nSetSampleProperty3 q translate 4 5 6 0.5\n")
        f, i = np.zeros(16), np.zeros(16)
        assert s.lib.fjscene_lerp_transform(s.id("q"), 0.5, sk.dptr(f), sk.dptr(i)) == 0
        assert f.reshape(4, 4)[:3, 3].tolist() == [4.0, 5.0, 6.0]
        for k in range(6):                      # initial sample at 0, the one at 0.5, six more: the list is full
            s.run("SetSampleProperty3 q translate 0 0 0 %d\n" % (k + 1))
        for full in ("SetSampleProperty3 q translate 0 0 0 99\n", "SetSampleProperty3 q translate 7 7 7 0.5\n"):
            with pytest.raises(fuji.SceneError):   # PropPushSample checks the count before it looks for an equal time
                s.run(full)
        assert s.lib.fjscene_lerp_transform(s.id("m"), 0.0, sk.dptr(f), sk.dptr(i)) == -1


def test_motion_scene_description_is_consistent(tmp_path):
    """The golden scene's three descriptions agree: `.scn` text carries the sample keys, the flat structs carry tables
    as long as the time table, static instances carry none."""
    d = golden_scenes.SCENES["motion_blur"]()
    scn = d.to_scn(str(tmp_path), None, plugin_dir="/x")
    assert scn.count("SetSampleProperty3 a rotate") == 3 and scn.count("SetSampleProperty3 cam1 translate") == 2
    assert "SetProperty2 ren1 sample_time_range 0.1 0.9" in scn
    st = d.to_structs()
    assert sorted(st["motion"]) == [-1, 0, 1]
    n = len(st["times"])
    assert n == (2 * 32 + 2) ** 2
    for fwd, inv in st["motion"].values():
        assert fwd.shape == (n, 16) and inv.shape == (n, 16)
        prod = np.einsum("nij,njk->nik", fwd.reshape(n, 4, 4), inv.reshape(n, 4, 4))
        assert np.abs(prod - np.eye(4)).max() < 1e-12
    assert "motion" not in golden_scenes.SCENES["multi"]().to_structs()


def test_velocity_generator_matches_oracle_and_reaches_the_device_structs(libs, tmp_path):
    """VelocityGeneratorProcedure through libfjscene writes the oracle's velocities bit for bit (the oracle's are pinned
    on the reference: tests/test_oracle_golden.py), and the flattened scene hands them on (rendering needs a GPU:
    tests/test_host_mirror.py / test_gpu_parity.py)."""
    _, _, fuji = libs
    d = golden_scenes.SCENES["velocity_blur"]()
    st = d.to_structs()
    scn = d.to_scn(str(tmp_path), None, plugin_dir="/x")
    head = "\n".join(l for l in scn.split("\n") if not l.startswith("RenderScene")) + "\n"
    o = sk.oracle()
    mid, P, N, idx = st["meshes"][st["velocity_meshes"][0]]
    sc = o.fjo_scene_new()                       # the generator runs once, on the static mesh
    o.fjo_mesh(sc, mid, sk.dptr(P), sk.dptr(N), len(P), sk.iptr(idx), None, len(idx) // 3)
    ref = np.zeros_like(P)
    assert o.fjo_mesh_generate_velocity(sc, mid, sk.dptr(ref)) == 0
    o.fjo_scene_free(sc)
    assert np.abs(ref).max() > 0.01
    with fuji.Session() as s:
        s.run(head)
        vel = np.zeros_like(P)
        assert s.lib.fjscene_mesh_velocity(s.id("blob"), sk.dptr(vel), len(P)) == 0
        assert np.array_equal(vel, ref)
        assert s.lib.fjscene_mesh_velocity(s.id("floor"), sk.dptr(vel), 4) == -1          # a mesh without velocities
        assert s.lib.fjscene_flatten(C.c_long(s.id("ren1")), None, None, None, None) == 0       # no refusal any more
    assert np.array_equal(st["mesh_velocity"][mid], ref)          # what the test kit hands to fjgpu_mesh_upload_velocity
