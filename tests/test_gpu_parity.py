"""Parity of the CUDA path (through the C-ABI of include/fjgpu.h) against the oracle and the committed
reference renders.  Everything here needs a B200: run with `pytest -m gpu`.

Bars: closest-hit t/u/v bit-exact in FP64 (same operation order, no FMA), sample positions bit-exact,
frames within 1e-4 per-channel RMSE of the oracle (BASELINE.json north_star) — and of the reference's own
.fb output for the deterministic shaders.  Stochastic shaders share the counter RNG with the oracle and are
compared sample for sample."""
import ctypes as C

import numpy as np
import pytest

import golden_scenes

pytestmark = pytest.mark.gpu

RMSE_BAR = 1e-4          # per-channel RMSE bar of BASELINE.json's north_star


@pytest.fixture(scope="module")
def device(sk):
    from fujiyama_renderer_b200 import device as d
    return d


def gpu_render(device, desc, st=None, region=None):
    st = st or desc.to_structs()
    dev = device.Device(0)
    try:
        dev.load_structs(st)
        return dev.render(st["params"], desc.tiles(region))
    finally:
        dev.close()


def rmse(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d * d).reshape(-1, a.shape[-1]).mean(0))


def random_rays(n, seed, radius=4.0):
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * radius * (0.2 + rng.random((n, 1)))
    tgt = rng.normal(size=(n, 3)) * 0.7
    d = tgt - o
    d[: n // 2] /= np.linalg.norm(d[: n // 2], axis=1, keepdims=True)     # half unnormalised
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


@pytest.mark.parametrize("scene", ["cube_c1", "plastic", "multi"])
@pytest.mark.parametrize("flags", [0, 1, 2])      # wavefront extend kernel, megakernel FP64 boxes, megakernel FP32 boxes
def test_trace_closest_bit_exact(sk, device, scene, flags):
    desc = golden_scenes.SCENES[scene]()
    st = desc.to_structs()
    n = 20000
    o, d = random_rays(n, 7)
    tmin = np.full(n, 1e-3)
    tmax = np.full(n, 1000.0)
    tmax[::5] = 3.0                                     # clipped rays (RayInRange upper bound)
    sc = sk.oracle_scene(st)
    rt, ru, rv = np.zeros(n), np.zeros(n), np.zeros(n)
    rp, ri = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert sk.oracle().fjo_trace_closest(sc, 0, n, sk.dptr(o), sk.dptr(d), sk.dptr(tmin), sk.dptr(tmax),
                                         sk.dptr(rt), sk.dptr(ru), sk.dptr(rv), sk.iptr(rp), sk.iptr(ri)) == 0
    sk.oracle().fjo_scene_free(sc)
    dev = device.Device(0)
    dev.load_structs(st)
    t, u, v, p, i = dev.trace_closest(0, o, d, tmin, tmax, flags)
    dev.close()
    nhit = int((ri >= 0).sum())
    assert 0.05 * n < nhit < 0.98 * n
    assert np.array_equal(i, ri)
    assert np.array_equal(t, rt)                        # bit-exact FP64
    same_prim = p == rp
    # an exact tie in t between two triangles sharing an edge may pick the other triangle: (u, v) then differ
    assert same_prim.mean() > 0.999
    assert np.array_equal(u[same_prim], ru[same_prim]) and np.array_equal(v[same_prim], rv[same_prim])


def test_empty_and_degenerate_inputs(sk, device):
    desc = golden_scenes.SCENES["cube_c1"]()
    st = desc.to_structs()
    dev = device.Device(0)
    dev.load_structs(st)
    # zero rays, zero tiles
    t, u, v, p, i = dev.trace_closest(0, np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0), np.zeros(0))
    assert len(t) == 0
    img, stats = dev.render(st["params"], [])
    assert not img.any() and stats.rays == 0
    # axis-parallel rays (zero direction components) and rays starting inside the mesh
    o = np.array([[0, 0, 5.], [0, 0, 0.], [5, 0, 0.], [0.1, 0.1, 5.]])
    d = np.array([[0, 0, -1.], [0, 1, 0.], [-1, 0, 0.], [0, 0, -1.]])
    tmin, tmax = np.full(4, 1e-3), np.full(4, 1000.)
    sc = sk.oracle_scene(st)
    rt, ru, rv = np.zeros(4), np.zeros(4), np.zeros(4)
    rp, ri = np.zeros(4, np.int32), np.zeros(4, np.int32)
    sk.oracle().fjo_trace_closest(sc, 0, 4, sk.dptr(o), sk.dptr(d), sk.dptr(tmin), sk.dptr(tmax),
                                  sk.dptr(rt), sk.dptr(ru), sk.dptr(rv), sk.iptr(rp), sk.iptr(ri))
    sk.oracle().fjo_scene_free(sc)
    t, u, v, p, i = dev.trace_closest(0, o, d, tmin, tmax)
    assert np.array_equal(t, rt) and np.array_equal(i, ri)
    # a mesh with no faces and an instance of it
    dev.mesh(7, np.zeros((3, 3)), None, np.zeros((0, 3), np.int32))
    dev.close()


def test_tri64_path_matches_tri32(sk, device, monkeypatch):
    """Vertices that are not FP32-representable take the FP64 triangle packets; same hits."""
    desc = golden_scenes.SCENES["plastic"]()
    st = desc.to_structs()
    n = 5000
    o, d = random_rays(n, 11)
    dev = device.Device(0)
    dev.load_structs(st)
    a = dev.trace_closest(0, o, d, 1e-3, 1000.)
    dev.close()
    monkeypatch.setenv("FJGPU_FORCE_TRI64", "1")
    dev = device.Device(0)
    dev.load_structs(st)
    b = dev.trace_closest(0, o, d, 1e-3, 1000.)
    dev.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_sample_positions_bit_exact(sk, device):
    desc = golden_scenes.SCENES["plastic"]()
    st = desc.to_structs()
    p = st["params"]
    dev = device.Device(0)
    dev.load_structs(st)
    for tile in (desc.tiles()[0], desc.tiles()[-1]):
        uv, rgba = dev.tile_samples(p, tile)
        t = sk.abi.Tile(*tile)
        ref = np.zeros_like(uv)
        n = sk.oracle().fjo_generate_samples(C.byref(p), C.byref(t), sk.dptr(ref), len(ref))
        assert n == len(uv)
        assert np.array_equal(uv, ref)
    dev.close()


@pytest.mark.parametrize("name", ["plastic", "pt_branching", "grid_light", "sphere_light", "motion_blur", "velocity_blur"])
def test_tile_samples_match_oracle(sk, device, name):
    """Per-sample radiance before the pixel filter (Sample::data), same counter RNG on both sides."""
    desc = golden_scenes.SCENES[name]()
    st = desc.to_structs()
    p = st["params"]
    tiles = desc.tiles()
    tile = tiles[len(tiles) // 2]
    dev = device.Device(0)
    dev.load_structs(st)
    uv, rgba = dev.tile_samples(p, tile)
    dev.close()
    sc = sk.oracle_scene(st)
    ruv, rrgba = np.zeros_like(uv), np.zeros_like(rgba)
    t = sk.abi.Tile(*tile)
    assert sk.oracle().fjo_render_tile_samples(sc, C.byref(p), C.byref(t), 0, sk.dptr(ruv), sk.fptr(rrgba)) == 0
    sk.oracle().fjo_scene_free(sc)
    assert np.array_equal(uv, ruv)
    assert np.array_equal(rgba[:, 3], rrgba[:, 3])
    err = np.abs(rgba - rrgba)
    assert err.max() <= 2e-5 * max(1.0, float(np.abs(rrgba).max())), float(err.max())


@pytest.mark.parametrize("name", list(golden_scenes.SCENES))
def test_frame_matches_oracle(sk, device, name):
    desc = golden_scenes.SCENES[name]()
    st = desc.to_structs()
    ref, rstats = sk.oracle_render(desc, rng_mode=0, threads=8, st=st)
    img, stats = gpu_render(device, desc, st)
    e = rmse(img, ref)
    assert e.max() < RMSE_BAR, e
    assert np.abs(img - ref).max() < 1e-4
    # identical ray trees: the counts per ray type agree exactly
    for k in ("rays_camera", "rays_shadow", "rays_diffuse", "rays_reflect", "rays_refract", "camera_samples"):
        assert getattr(stats, k) == getattr(rstats, k), k
    assert stats.kernel_launches >= 2


@pytest.mark.parametrize("name", ["plastic_4l", "multi", "pt_branching", "sphere_light", "motion_blur", "velocity_blur"])
@pytest.mark.parametrize("flags", [1, 2])
def test_megakernel_cross_check(sk, device, name, flags):
    """The two independent device implementations (wavefront rounds over ray queues vs one sample per lane with a
    private ray stack, FP32 or FP64 box culling) trace identical ray trees and agree to float rounding."""
    desc = golden_scenes.SCENES[name]()
    st = desc.to_structs()
    a, sa = gpu_render(device, desc, st)
    st["params"].flags = flags
    b, sb = gpu_render(device, desc, st)
    st["params"].flags = 0
    assert np.abs(a - b).max() < 2e-6 * max(1.0, float(np.abs(a).max()))
    # (hit_mesh_levels is not compared: the wavefront's shadow rays stop at ANY occluder when every shader is opaque, which may
    # be another mesh than the closest one the megakernel reports — occluded or not, and therefore every colour, is the same)
    for k in ("rays_camera", "rays_shadow", "rays_diffuse", "rays_reflect", "rays_refract", "rays_hit"):
        assert getattr(sa, k) == getattr(sb, k), k


@pytest.mark.parametrize("name", golden_scenes.DETERMINISTIC)
def test_frame_matches_reference_fb(sk, device, name):
    """Against the unmodified reference's own .fb output (tests/golden/ref_images.npz)."""
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_images.npz"))[name]
    img, _ = gpu_render(device, golden_scenes.SCENES[name]())
    assert img.shape == ref.shape
    assert rmse(img, ref).max() < RMSE_BAR


@pytest.mark.parametrize("name", golden_scenes.STOCHASTIC)
def test_stochastic_frame_statistics_vs_reference(sk, device, name):
    """pathtracing_shader / grid / sphere lights: the reference's XorShift streams cannot be reproduced in
    parallel — coverage is exact, the image mean agrees, the result is deterministic."""
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_images.npz"))[name].astype(np.float64)
    desc = golden_scenes.SCENES[name]()
    img, _ = gpu_render(device, desc)
    img2, _ = gpu_render(device, desc)
    assert np.array_equal(img, img2)
    assert np.array_equal(img[..., 3] > 0, ref[..., 3] > 0)
    assert abs(img[..., :3].mean() - ref[..., :3].mean()) < 0.01 * max(ref[..., :3].mean(), 1e-3) + 1e-3


def test_region_and_tile_subsets(sk, device):
    """Tiles are independent units: rendering a subset writes only those pixels and gives the same values
    (render_region, src/internal/fj_property_list_include.cc:239-246)."""
    desc = golden_scenes.SCENES["multi"]()
    st = desc.to_structs()
    full, _ = gpu_render(device, desc, st)
    tiles = desc.tiles()
    dev = device.Device(0)
    dev.load_structs(st)
    frame = np.full_like(full, -1.0)
    dev.render(st["params"], tiles[1::2], frame)
    dev.close()
    mask = np.zeros(full.shape[:2], bool)
    for _, x0, y0, x1, y1 in tiles[1::2]:
        mask[y0:y1, x0:x1] = True
    assert np.array_equal(frame[mask], full[mask])
    assert (frame[~mask] == -1).all()
    # ragged region (not tile aligned)
    region = (10, 5, 100, 70)
    part, _ = gpu_render(device, desc, st, region=region)
    ref, _ = sk.oracle_render(desc, rng_mode=0, threads=8, region=region, st=st)
    assert rmse(part, ref).max() < RMSE_BAR


def test_batching_is_invisible(sk, device, monkeypatch):
    desc = golden_scenes.SCENES["plastic_4l"]()
    st = desc.to_structs()
    a, _ = gpu_render(device, desc, st)
    monkeypatch.setenv("FJGPU_SAMPLE_MB", "1")        # forces several tile batches
    b, sb = gpu_render(device, desc, st)
    assert np.array_equal(a, b)
    assert sb.kernel_launches > 2


def test_queue_overflow_in_a_later_batch_grows_and_retries(sk, device, monkeypatch):
    """A branching ray tree (diffuse + mirror + refracted lobes: up to 3 children per hit) starts with an optimistic queue of
    2 records per sample slot.  With tiny tile batches the first batches see only background — they fit — and the overflow
    happens in a LATER batch: that batch alone is rendered again with a larger queue, the earlier batches' ray counts are
    kept, and the frame equals the single-batch frame bit for bit (round 1 failed such frames)."""
    desc = sk.scene_blob_pathtracing(n=24, res=(128, 160), rate=4, reflect=(.3, .3, .3), refract=(.4, .4, .4))
    desc.cam.update(T=(0, 1.6, 4.5))                   # the blob sits in the lower half of the frame: the top tile rows are background
    st = desc.to_structs()
    a, sa = gpu_render(device, desc, st)
    ref, rstats = sk.oracle_render(desc, rng_mode=0, threads=8, st=st)
    assert rstats.rays_reflect > 0 and rstats.rays_refract > 0
    assert sa.batches == 1 and sa.queue_regrows >= 1 and sa.first_regrow_batch == 0      # one batch: the overflow is in batch 0
    monkeypatch.setenv("FJGPU_SAMPLE_MB", "1")         # one or two tiles per batch
    b, sb = gpu_render(device, desc, st)
    assert np.array_equal(a, b)
    for k in ("rays_camera", "rays_diffuse", "rays_reflect", "rays_refract", "camera_samples"):
        assert getattr(sa, k) == getattr(sb, k) == getattr(rstats, k), k
    assert sa.rays_hit == sb.rays_hit and sa.hit_mesh_levels == sb.hit_mesh_levels
    assert sb.batches > 4 and sb.queue_regrows >= 1 and sb.first_regrow_batch > 0        # the case round 1 refused
    assert rmse(b, ref).max() < RMSE_BAR


def test_failed_mesh_upload_leaves_no_record(sk, device):
    """A mesh upload that fails (face index out of range is caught before anything is touched; a too-deep tree after the
    record exists) must not leave a half-built mesh behind: the scene still renders and the id can be uploaded again."""
    desc = golden_scenes.SCENES["cube_c1"]()
    st = desc.to_structs()
    dev = device.Device(0)
    dev.load_structs(st)
    ref, _ = dev.render(st["params"], desc.tiles())
    mid, P, N, idx = st["meshes"][0]
    bad = idx.copy(); bad[0] = len(P) + 5
    with pytest.raises(device.FjGpuError):
        dev.mesh(mid, P, N, bad)
    img, _ = dev.render(st["params"], desc.tiles())     # the earlier, valid upload is untouched
    assert np.array_equal(img, ref)
    dev.mesh(mid, P, N, idx)
    img, _ = dev.render(st["params"], desc.tiles())
    assert np.array_equal(img, ref)
    dev.close()


def test_resident_and_device_block_outputs(sk, device):
    torch = pytest.importorskip("torch")
    desc = golden_scenes.SCENES["multi"]()
    st = desc.to_structs()
    full, _ = gpu_render(device, desc, st)
    tiles = desc.tiles()
    dev = device.Device(0)
    dev.load_structs(st)
    blocks = torch.full((len(tiles), 32, 32, 4), -1.0, device="cuda:0")
    torch.cuda.synchronize()
    stats = dev.render_to_device_blocks(st["params"], tiles, 32, 32, blocks.data_ptr())
    rs = dev.render_resident(st["params"], tiles)
    dev.close()
    hb = blocks.cpu().numpy()
    for k, (_, x0, y0, x1, y1) in enumerate(tiles):
        assert np.array_equal(hb[k, : y1 - y0, : x1 - x0], full[y0:y1, x0:x1])
    assert stats.rays == rs.rays and rs.ms_trace > 0


def test_assemble_gathered_blocks_single_gpu(sk, device):
    """fjgpu_assemble_frame (the last step of the multi-GPU frame) on ONE GPU: the tiles of R = 1, 2, 3, 8 simulated ranks
    rendered into packed blocks, laid out as the all-gather leaves them (rank-major, short ranks padded), un-permuted on the
    device and copied to the host once — the full frame bit for bit, ragged last tiles and a render_region included."""
    torch = pytest.importorskip("torch")
    from fujiyama_renderer_b200 import sharding
    desc = golden_scenes.SCENES["multi"]()
    st = desc.to_structs()
    dev = device.Device(0)
    dev.load_structs(st)
    try:
        for region in (None, (10, 5, 100, 70)):
            tiles = desc.tiles(region)
            full, _ = dev.render(st["params"], tiles)
            p = st["params"]
            for R in (1, 2, 3, 8):
                per = sharding.blocks_per_rank(len(tiles), R)
                gathered = torch.full((R * per, 32, 32, 4), -7.0, device="cuda:0")
                for r in range(R):
                    mine = sharding.rank_tiles(tiles, r, R)
                    if mine:
                        dev.render_to_device_blocks(p, mine, 32, 32, gathered[r * per:].data_ptr())
                torch.cuda.synchronize()
                pinned = torch.empty((p.yres, p.xres, 4), dtype=torch.float32, pin_memory=True)
                for host in (None, pinned.numpy()):                      # pageable and pinned destinations
                    out = dev.assemble_frame(gathered.data_ptr(), R, 32, 32, tiles, p.xres, p.yres, host)
                    assert np.array_equal(out, full), (region, R)
    finally:
        dev.close()


def test_render_frame_multi_in_one_process(sk, device):
    """fjgpu_render_frame_multi: one process, one context per GPU, ncclAllGather of the tile blocks — the same frame as one GPU
    bit for bit (the counter RNG is keyed by tile id).  Needs >= 2 GPUs; with one, the call must reduce to fjgpu_render_tiles."""
    torch = pytest.importorskip("torch")
    desc = golden_scenes.SCENES["pt_branching"]()
    st = desc.to_structs()
    tiles = desc.tiles()
    dev = device.Device(0)
    dev.load_structs(st)
    ref, rs = dev.render(st["params"], tiles)
    one, s1 = device.render_frame_multi([dev], st["params"], tiles)
    assert np.array_equal(one, ref) and s1[0].rays == rs.rays
    n = torch.cuda.device_count()
    if n >= 2:
        devs = [dev] + [device.Device(k) for k in range(1, min(n, 4))]
        for d in devs[1:]:
            d.load_structs(st)
        img, stats = device.render_frame_multi(devs, st["params"], tiles)
        assert np.array_equal(img, ref)
        assert sum(s.rays for s in stats) == rs.rays and all(s.rays > 0 for s in stats)
        for d in devs[1:]:
            d.close()
    dev.close()
    if n < 2:
        pytest.skip("one GPU: the multi-context path itself needs >= 2 (covered by tools/r2 multi-GPU runs)")


def test_big_mesh_property(sk, device):
    """Full-size property check (no oracle render at this size): a 1M-triangle blob, closest-hit rays cross-checked
    between FP32 and FP64 box culling, and hit points verified against the analytic surface radius."""
    from fujiyama_renderer_b200 import synth
    P, idx = synth.blob(synth.BLOB_N["1M"])
    desc = sk.SceneDesc()
    desc.mesh("blob", P, idx)
    desc.shader("s", "constant")
    desc.instance("o", "blob", "s", R=(20, 30, 0))
    st = desc.to_structs()
    dev = device.Device(0)
    dev.load_structs(st)
    n = 200000
    o, d = random_rays(n, 3)
    a = dev.trace_closest(0, o, d, 1e-3, 1000., 0)
    b = dev.trace_closest(0, o, d, 1e-3, 1000., 1)
    info = dev.info()
    dev.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert info.blas_tris == len(idx)
    t, u, v, prim, inst = a
    hit = inst >= 0
    assert 0.1 < hit.mean() < 0.95
    Ph = o[hit] + t[hit, None] * d[hit]
    r = np.linalg.norm(Ph, axis=1)
    assert r.min() > 0.85 and r.max() < 1.15       # radius 1 +- .11 bumps
    # oracle spot check on a subset
    sub = np.arange(0, n, 200)
    sc = sk.oracle_scene(st)
    m = len(sub)
    rt, ru, rv = np.zeros(m), np.zeros(m), np.zeros(m)
    rp, ri = np.zeros(m, np.int32), np.zeros(m, np.int32)
    os_, ds_ = np.ascontiguousarray(o[sub]), np.ascontiguousarray(d[sub])
    tmin, tmax = np.full(m, 1e-3), np.full(m, 1000.)
    sk.oracle().fjo_trace_closest(sc, 0, m, sk.dptr(os_), sk.dptr(ds_), sk.dptr(tmin), sk.dptr(tmax),
                                  sk.dptr(rt), sk.dptr(ru), sk.dptr(rv), sk.iptr(rp), sk.iptr(ri))
    sk.oracle().fjo_scene_free(sc)
    assert np.array_equal(t[sub], rt) and np.array_equal(inst[sub], ri)


# ---------------------------------------------------------------------------------------------- extend-kernel variants
def instanced16(sk):
    """BASELINE config 4 in miniature: 16 instances of one mesh on the 4x4 grid of scenes/happy_buddhas.scn
    (translate -1.5 i, 0, -1.5 j; rotate 0, 30 k, 0; scale .6) + a floor: a TLAS with inner nodes over a shared BLAS."""
    from fujiyama_renderer_b200 import synth
    d = sk.SceneDesc()
    P, idx = synth.blob(24)
    d.mesh("blob", P, idx)
    Pq, iq = synth.quad(12.0, -0.7)
    d.mesh("floor", Pq, iq)
    d.shader("s", "plastic", diffuse=(.5, .5, .5))
    k = 0
    for i in range(4):
        for j in range(4):
            d.instance("o%d" % k, "blob", "s", T=(2.25 - 1.5 * i, 0, 2.25 - 1.5 * j), R=(0, 30 * k, 0), S=(.6, .6, .6))
            k += 1
    d.instance("f", "floor", "s")
    d.light(0, T=(5, 12, 5), intensity=1.0)
    d.cam.update(T=(0, 4, 9), R=(-25, 0, 0), fov=40)
    d.ren.update(resolution=(96, 64), pixelsamples=(2, 2))
    return d


def soup(sk):
    """BASELINE config 5 in miniature: S-random triangle soup (overlapping leaf boxes, no surface coherence)."""
    from fujiyama_renderer_b200 import synth
    d = sk.SceneDesc()
    P, idx = synth.random_tris(5000, seed=1234)
    d.mesh("soup", P * np.float32(3.0), idx)
    d.shader("s", "constant")
    d.instance("o", "soup", "s", R=(10, 20, 30))
    d.cam.update(T=(0, 0, 6), fov=35)
    d.ren.update(resolution=(64, 64), pixelsamples=(1, 1))
    return d


EXTEND_VARIANTS = {
    "v1_registers": {"FJGPU_EXTEND": "1"},
    "v2_float_nodes": {"FJGPU_EXTEND": "2", "FJGPU_QUANT": "0"},
    "v2_quantised_6": {"FJGPU_EXTEND": "2", "FJGPU_QUANT": "1", "FJGPU_COOP": "0", "FJGPU_EXTEND_MINBLOCKS": "6"},
    "v2_quantised_8": {"FJGPU_EXTEND": "2", "FJGPU_QUANT": "1", "FJGPU_COOP": "0", "FJGPU_EXTEND_MINBLOCKS": "8", "FJGPU_REFILL": "4", "FJGPU_PHASE_A_MIN": "4"},
    "v2_cooperative_leaves": {"FJGPU_EXTEND": "2", "FJGPU_QUANT": "1", "FJGPU_COOP": "1"},
    "v2_cooperative_leaves_8": {"FJGPU_EXTEND": "2", "FJGPU_QUANT": "1", "FJGPU_COOP": "1", "FJGPU_EXTEND_MINBLOCKS": "8", "FJGPU_REFILL": "4", "FJGPU_PHASE_A_MIN": "4"},
    # shared-memory traversal stack of 8 / 16 entries per lane (default 12; deeper entries spill to local memory: the soup's
    # overlapping boxes do reach them)
    "v2_stack_smem_8": {"FJGPU_EXTEND": "2", "FJGPU_STACK_SMEM": "8"},
    "v2_stack_smem_16": {"FJGPU_EXTEND": "2", "FJGPU_STACK_SMEM": "16"},
    # the top of the largest tree staged in shared memory by one cp.async.bulk per CTA (7 and 6 CTAs per SM)
    "v2_top_staged_64": {"FJGPU_EXTEND": "2", "FJGPU_TOP_NODES": "64"},
    "v2_top_staged_341": {"FJGPU_EXTEND": "2", "FJGPU_TOP_NODES": "341", "FJGPU_EXTEND_MINBLOCKS": "6"},
    "v2_unchunked_queue": {"FJGPU_EXTEND": "2", "FJGPU_QUEUE_CHUNK": "0"},
    # shadow rays walked to their CLOSEST occluder (the default stops at the first hit when every shader is opaque)
    "v2_closest_hit_shadows": {"FJGPU_EXTEND": "2", "FJGPU_ANYHIT": "0"},
    # rays of the next queue sorted by (direction octant, origin cell) between bounces (frames only; the probe has no bounces)
    "v2_sorted_rays": {"FJGPU_EXTEND": "2", "FJGPU_SORT_BITS": "4"},
    "v2_sorted_rays_unchunked": {"FJGPU_EXTEND": "2", "FJGPU_SORT_BITS": "3", "FJGPU_QUEUE_CHUNK": "0"},
    # k_extend_ring: a per-warp ring of prepared rays between the queue and the lanes (fj_extend_ring.cuh), at several ring
    # thresholds, stack depths and occupancies; with sorted rays (the ring is filled through `perm`) and closest-hit shadows
    "v3_ring": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_B1_MIN": "1", "FJGPU_B2_MIN": "1"},
    "v3_ring_eager": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_REFILL": "1", "FJGPU_PHASE_A_MIN": "4"},
    "v3_ring_lazy_sd12": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_REFILL": "32", "FJGPU_STACK_SMEM": "12"},
    "v3_ring_8": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_EXTEND_MINBLOCKS": "8", "FJGPU_REFILL": "8"},
    "v3_ring_6": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_EXTEND_MINBLOCKS": "6"},
    "v3_default_sorted": {"FJGPU_SORT_BITS": "4"},
    "v3_default": {},
    "v3_default_closest_hit_shadows": {"FJGPU_ANYHIT": "0"},
    "v3_ring_scalar_fma": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_FMA2": "0", "FJGPU_STACK_SMEM": "8"},
    # the new node loop with the direct refill of k_extend2 instead of the ring
    "v3_direct_ungated": {"FJGPU_EXTEND": "3", "FJGPU_RING": "0", "FJGPU_B1_MIN": "1", "FJGPU_B2_MIN": "1", "FJGPU_REFILL": "12"},
    "v3_direct_refill_sd8": {"FJGPU_EXTEND": "3", "FJGPU_RING": "0", "FJGPU_STACK_SMEM": "8", "FJGPU_REFILL": "4"},
    # heavy phases (leaf tests / instance entries) deferred until enough lanes wait for them
    "v3_gated": {"FJGPU_EXTEND": "3", "FJGPU_RING": "0", "FJGPU_B1_MIN": "24", "FJGPU_B2_MIN": "8"},
    "v3_gated_ring": {"FJGPU_EXTEND": "3", "FJGPU_RING": "1", "FJGPU_B1_MIN": "28", "FJGPU_B2_MIN": "12"},
    "v3_gated_extreme": {"FJGPU_EXTEND": "3", "FJGPU_RING": "0", "FJGPU_B1_MIN": "200", "FJGPU_B2_MIN": "32", "FJGPU_PHASE_A_MIN": "4"},
}
VARIANT_KEYS = ("FJGPU_EXTEND", "FJGPU_QUANT", "FJGPU_EXTEND_MINBLOCKS", "FJGPU_REFILL", "FJGPU_PHASE_A_MIN", "FJGPU_COOP", "FJGPU_QUEUE_CHUNK",
                "FJGPU_STACK_SMEM", "FJGPU_TOP_NODES", "FJGPU_SORT_BITS", "FJGPU_ANYHIT", "FJGPU_FMA2", "FJGPU_RING", "FJGPU_B1_MIN", "FJGPU_B2_MIN")


@pytest.mark.parametrize("scene", ["multi", "instanced16", "soup"])
def test_extend_variants_bit_exact(sk, device, scene, monkeypatch):
    """Every closest-hit kernel (register-resident, shared-memory state with FP32 or 8-bit quantised nodes, quad-per-ray)
    returns the oracle's hits bit for bit, on a multi-instance scene, a 17-instance TLAS and a triangle soup, with
    clipped, unnormalised and axis-parallel rays."""
    desc = {"multi": golden_scenes.SCENES["multi"], "instanced16": lambda: instanced16(sk), "soup": lambda: soup(sk)}[scene]()
    st = desc.to_structs()
    n = 30000
    o, d = random_rays(n, 21, radius=5.0)
    d[::7, 0] = 0.0                                     # axis-parallel components
    d[::11, 1] = 0.0
    tmin = np.full(n, 1e-3)
    tmax = np.full(n, 1000.0)
    tmax[::5] = 4.0
    sc = sk.oracle_scene(st)
    rt, ru, rv = np.zeros(n), np.zeros(n), np.zeros(n)
    rp, ri = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert sk.oracle().fjo_trace_closest(sc, 0, n, sk.dptr(o), sk.dptr(d), sk.dptr(tmin), sk.dptr(tmax),
                                         sk.dptr(rt), sk.dptr(ru), sk.dptr(rv), sk.iptr(rp), sk.iptr(ri)) == 0
    sk.oracle().fjo_scene_free(sc)
    assert 0.02 * n < int((ri >= 0).sum()) < 0.99 * n
    dev = device.Device(0)
    dev.load_structs(st)
    try:
        for name, env in EXTEND_VARIANTS.items():
            for k in VARIANT_KEYS:
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            try:
                t, u, v, p, i = dev.trace_closest(0, o, d, tmin, tmax, 0)
            except Exception as e:
                raise AssertionError("variant %s: %s" % (name, e))
            assert np.array_equal(i, ri), name
            assert np.array_equal(t, rt), name
            same = p == rp
            assert same.mean() > 0.999, name
            assert np.array_equal(u[same], ru[same]) and np.array_equal(v[same], rv[same]), name
    finally:
        dev.close()


@pytest.mark.parametrize("name", ["multi", "pt_branching", "motion_blur", "velocity_blur"])
def test_extend_variants_same_frame(sk, device, name, monkeypatch):
    """Whole frames (shadow rays, mirror bounces, branching path trees) are bit-identical whichever extend kernel traces
    them, and so are the ray counts."""
    desc = golden_scenes.SCENES[name]()
    frames = {}
    for vname, env in EXTEND_VARIANTS.items():
        for k in VARIANT_KEYS:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        img, stats = gpu_render(device, desc)
        frames[vname] = (img, stats.rays)
    ref_img, ref_rays = frames["v1_registers"]
    for vname, (img, rays) in frames.items():
        assert rays == ref_rays, vname
        assert np.array_equal(img, ref_img), vname


def test_chunked_queue_same_frame(sk, device, monkeypatch):
    """k_shade reserving queue slots in per-warp chunks (FJGPU_QUEUE_CHUNK=1, the default: fillers in the queue, another ray
    order) renders the same frame bit for bit as one atomic per spawn (FJGPU_QUEUE_CHUNK=0) and counts the same rays, on a branching path tree large enough that every warp
    goes through many chunks."""
    desc = sk.scene_blob_pathtracing(n=32, res=(320, 180), rate=4, reflect=(.3, .3, .3), refract=(.4, .4, .4))
    st = desc.to_structs()
    monkeypatch.setenv("FJGPU_QUEUE_CHUNK", "0")
    a, sa = gpu_render(device, desc, st)
    monkeypatch.setenv("FJGPU_QUEUE_CHUNK", "1")
    b, sb = gpu_render(device, desc, st)
    assert np.array_equal(a, b)
    for k in ("rays_camera", "rays_shadow", "rays_diffuse", "rays_reflect", "rays_refract", "rays_hit", "hit_mesh_levels"):
        assert getattr(sa, k) == getattr(sb, k), k
    assert sa.rays_diffuse > 0 and sa.rays_reflect > 0 and sa.rays_refract > 0


# ---------------------------------------------------------------------------------------------- device BVH build (§8f row 2)
@pytest.mark.parametrize("scene", ["plastic", "multi", "soup"])
def test_device_built_bvh_gives_the_same_hits_and_frames(sk, device, scene, monkeypatch):
    """FJGPU_BUILD=device (linear BVH built in HBM, fj_build.cu): every traversal kernel still returns the oracle's closest
    hits bit for bit, and whole frames are bit-identical to the frames of the host-built (binned SAH) trees."""
    desc = {"plastic": golden_scenes.SCENES["plastic"], "multi": golden_scenes.SCENES["multi"], "soup": lambda: soup(sk)}[scene]()
    st = desc.to_structs()
    img_host, stats_host = gpu_render(device, desc, st=st)
    n = 20000
    o, d = random_rays(n, 33, radius=5.0)
    tmin = np.full(n, 1e-3)
    tmax = np.full(n, 1000.0)
    tmax[::4] = 4.0
    sc = sk.oracle_scene(st)
    rt, ru, rv = np.zeros(n), np.zeros(n), np.zeros(n)
    rp, ri = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert sk.oracle().fjo_trace_closest(sc, 0, n, sk.dptr(o), sk.dptr(d), sk.dptr(tmin), sk.dptr(tmax),
                                         sk.dptr(rt), sk.dptr(ru), sk.dptr(rv), sk.iptr(rp), sk.iptr(ri)) == 0
    sk.oracle().fjo_scene_free(sc)
    monkeypatch.setenv("FJGPU_BUILD", "device")
    monkeypatch.setenv("FJGPU_BUILD_DEVICE_MIN", "2")
    dev = device.Device(0)
    try:
        dev.load_structs(st)
        assert dev.info().device_build_seconds > 0            # the device builder really ran
        for flags in (0, 1, 2):                               # wavefront kernel, megakernel with FP64 / FP32 boxes
            t, u, v, p, i = dev.trace_closest(0, o, d, tmin, tmax, flags)
            assert np.array_equal(i, ri) and np.array_equal(t, rt), flags
            same = p == rp
            assert same.mean() > 0.999
            assert np.array_equal(u[same], ru[same]) and np.array_equal(v[same], rv[same])
        for name, env in EXTEND_VARIANTS.items():
            for k in VARIANT_KEYS:
                monkeypatch.delenv(k, raising=False)
            for k, v_ in env.items():
                monkeypatch.setenv(k, v_)
            t, u, v, p, i = dev.trace_closest(0, o, d, tmin, tmax, 0)
            assert np.array_equal(i, ri) and np.array_equal(t, rt), name
        for k in VARIANT_KEYS:
            monkeypatch.delenv(k, raising=False)
        img_dev, stats_dev = dev.render(st["params"], desc.tiles())
    finally:
        dev.close()
    assert stats_dev.rays == stats_host.rays
    assert np.array_equal(img_dev, img_host)


def test_device_build_of_a_million_triangles(sk, device, monkeypatch):
    """Full-size build on the device: same hits as the host-built tree on 200 k rays, and the build itself is timed."""
    from fujiyama_renderer_b200 import synth
    P, idx = synth.blob(synth.BLOB_N["1M"])
    desc = sk.SceneDesc()
    desc.mesh("blob", P, idx)
    desc.shader("s", "constant")
    desc.instance("o", "blob", "s", R=(20, 30, 0))
    st = desc.to_structs()
    n = 200000
    o, d = random_rays(n, 5)
    dev = device.Device(0)
    dev.load_structs(st)
    a = dev.trace_closest(0, o, d, 1e-3, 1000., 0)
    host_s = dev.info().build_seconds
    dev.close()
    monkeypatch.setenv("FJGPU_BUILD", "device")
    dev = device.Device(0)
    dev.load_structs(st)
    b = dev.trace_closest(0, o, d, 1e-3, 1000., 0)
    info = dev.info()
    dev.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert 0 < info.device_build_seconds < host_s
    print("BVH build of %d triangles: host %.3f s, device %.4f s (%.0fx)" % (len(idx), host_s, info.device_build_seconds,
                                                                             host_s / info.device_build_seconds))


# ---------------------------------------------------------------------------------------------- full-size configurations
def test_config4_instanced_million_triangle_meshes(sk, device):
    """BASELINE config 4 at full size: 16 instances of a ~1.09 M-triangle mesh on the 4x4 grid of scenes/happy_buddhas.scn
    plus a floor (17.5 M instanced triangles behind a TLAS with inner nodes).  No oracle render at this size: the
    wavefront kernel is cross-checked against the megakernel with FP64 boxes (independent traversal code, binary BVH), hit
    points must lie on the analytic surface of the instance they report, and an oracle spot check pins a subset."""
    from fujiyama_renderer_b200 import synth
    P, idx = synth.blob(740)                              # 1 095 200 triangles
    d = sk.SceneDesc()
    d.mesh("blob", P, idx)
    Pq, iq = synth.quad(12.0, -0.7)
    d.mesh("floor", Pq, iq)
    d.shader("s", "constant")
    xf = []
    k = 0
    for i in range(4):
        for j in range(4):
            T, R, S = (2.25 - 1.5 * i, 0, 2.25 - 1.5 * j), (0, 30 * k, 0), (.6, .6, .6)
            d.instance("o%d" % k, "blob", "s", T=T, R=R, S=S)
            xf.append(T)
            k += 1
    d.instance("f", "floor", "s")
    st = d.to_structs()
    n = 200000
    o, dr = random_rays(n, 9, radius=6.0)
    dev = device.Device(0)
    dev.load_structs(st)
    a = dev.trace_closest(0, o, dr, 1e-3, 1000., 0)
    b = dev.trace_closest(0, o, dr, 1e-3, 1000., 1)
    info = dev.info()
    dev.close()
    assert info.instances == 17 and info.blas_tris == len(idx) + 2
    t, u, v, prim, inst = a
    assert np.array_equal(t, b[0]) and np.array_equal(inst, b[4])
    same = prim == b[3]
    assert same.mean() > 0.999 and np.array_equal(u[same], b[1][same])
    hit = (inst >= 0) & (inst < 16)
    assert hit.sum() > 0.2 * n
    Ph = o[hit] + t[hit, None] * dr[hit]
    r = np.linalg.norm(Ph - np.asarray(xf)[inst[hit]], axis=1) / 0.6
    assert r.min() > 0.85 and r.max() < 1.15              # radius 1 +- .11 bumps around the reported instance's centre


def test_config5_ten_million_triangle_soup(sk, device, monkeypatch):
    """BASELINE config 5's mesh at full size: S-random with 10 M triangles, BVH built on the device (host build: ~10 s).
    Wavefront kernel (8-bit quantised 4-wide nodes) against the megakernel with FP64 boxes on the binary tree."""
    from fujiyama_renderer_b200 import synth
    P, idx = synth.random_tris(10_000_000, seed=1234)
    d = sk.SceneDesc()
    d.mesh("soup", P, idx)
    d.shader("s", "constant")
    d.instance("o", "soup", "s")
    st = d.to_structs()
    monkeypatch.setenv("FJGPU_BUILD", "device")
    n = 100000
    o, dr = random_rays(n, 13, radius=2.0)
    dev = device.Device(0)
    dev.load_structs(st)
    a = dev.trace_closest(0, o, dr, 1e-3, 1000., 0)
    b = dev.trace_closest(0, o, dr, 1e-3, 1000., 1)
    info = dev.info()
    dev.close()
    assert info.blas_tris == 10_000_000 and info.device_build_seconds > 0
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[4], b[4])
    same = a[3] == b[3]
    assert same.mean() > 0.999 and np.array_equal(a[1][same], b[1][same]) and np.array_equal(a[2][same], b[2][same])
    assert 0.3 < (a[4] >= 0).mean() <= 1.0
    print("10 M triangles: device build %.3f s, scene %.2f GB in HBM" % (info.device_build_seconds, info.hbm_bytes / 1e9))
