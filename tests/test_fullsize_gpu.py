"""Frame-level parity at the sizes the numbers are quoted on ("What's weak" 1): BASELINE config 2 as a whole
frame against the oracle and against a region the unmodified reference rendered (tests/golden/config2_region.npz, made by
oracle/gen_golden.py), and >= 8-tile regions of the north-star frame, of config 3 (7 M triangles) and of config 4
(16 instances of a 1.09 M-triangle mesh, 16 grid-light samples) against the oracle — same counter RNG, full-frame tile ids.
Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest

import workloads

pytestmark = pytest.mark.gpu

RMSE_BAR = 1e-4
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rmse(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d * d).reshape(-1, a.shape[-1]).mean(0))


def block_of_tiles(desc, tx0, ty0, nx, ny):
    """nx x ny tiles of the full frame's tile list starting at tile column tx0, row ty0 — with their full-frame ids."""
    res, ts = desc.ren["resolution"], desc.ren["tilesize"]
    per_row = -(-res[0] // ts)
    all_tiles = desc.tiles()
    return [all_tiles[(ty0 + j) * per_row + tx0 + i] for j in range(ny) for i in range(nx)]


def region_parity(sk, desc, tiles, threads=16):
    from fujiyama_renderer_b200 import device
    st = desc.to_structs()
    ref, rstats = sk.oracle_render(desc, rng_mode=0, threads=threads, st=st, tiles=tiles)
    dev = device.Device(0)
    try:
        dev.load_structs(st)
        img, stats = dev.render(st["params"], tiles)
    finally:
        dev.close()
    mask = np.zeros(img.shape[:2], bool)
    for _, x0, y0, x1, y1 in tiles:
        mask[y0:y1, x0:x1] = True
    e = rmse(img[mask], ref[mask])
    assert e.max() < RMSE_BAR, e
    assert np.abs(img[mask] - ref[mask]).max() < 1e-4 * max(1.0, float(np.abs(ref[mask]).max()))
    for k in ("rays_camera", "rays_shadow", "rays_diffuse", "rays_reflect", "rays_refract", "camera_samples"):
        assert getattr(stats, k) == getattr(rstats, k), k
    assert ref[mask][..., :3].max() > 0.05          # the region does show the object
    return float(e.max()), stats.rays


def test_config2_full_frame(sk):
    """BASELINE config 2 exactly as bench.py --workload config2 renders it: 69 938 triangles, plastic_shader with the mirror
    bounce, one point light, 1280x720, 16 spp — the WHOLE frame against the oracle (identical ray counts per type), and the
    centre region against the unmodified reference's own .fb."""
    from fujiyama_renderer_b200 import device
    desc = workloads.plastic_blob()
    st = desc.to_structs()
    ref, rstats = sk.oracle_render(desc, rng_mode=0, threads=16, st=st)
    dev = device.Device(0)
    try:
        dev.load_structs(st)
        img, stats = dev.render(st["params"], desc.tiles())
    finally:
        dev.close()
    e = rmse(img, ref)
    assert e.max() < RMSE_BAR, e
    assert np.abs(img - ref).max() < 1e-4
    for k in ("rays_camera", "rays_shadow", "rays_diffuse", "rays_reflect", "rays_refract", "camera_samples"):
        assert getattr(stats, k) == getattr(rstats, k), k
    g = np.load(os.path.join(GOLDEN, "config2_region.npz"))
    x0, y0, x1, y1 = (int(v) for v in g["region"])
    fb = g["image"]                                   # the reference's pixels of that region (float32 [y1-y0, x1-x0, 4])
    assert fb.shape == (y1 - y0, x1 - x0, 4)
    assert rmse(img[y0:y1, x0:x1], fb).max() < RMSE_BAR
    assert 0.5 < fb[..., 3].mean() < 0.99             # the region shows the object and both of its silhouettes
    print("config 2 full frame: rmse vs oracle %.3g, vs reference region %.3g, %d rays" % (
        e.max(), rmse(img[y0:y1, x0:x1], fb).max(), stats.rays))


def test_north_star_region(sk):
    """The north-star frame (999 698 triangles, pathtracing depth 3, 1920x1080, 64 spp): a 4x2-tile block across the blob's
    silhouette, rendered by the device with the tiles' full-frame ids, against the oracle sample stream for sample stream."""
    desc = workloads.pathtracing_blob(n=707)
    e, rays = region_parity(sk, desc, block_of_tiles(desc, 42, 15, 4, 2))
    print("north star, 8 tiles: rmse %.3g, %d rays" % (e, rays))


def test_config3_region(sk):
    """Config 3 (S-blob(1871) = 7.0 M triangles, the DRAM-resident tree): 8 tiles against the oracle."""
    desc = workloads.pathtracing_blob(n=1871)
    e, rays = region_parity(sk, desc, block_of_tiles(desc, 28, 12, 4, 2))
    print("config 3, 8 tiles: rmse %.3g, %d rays" % (e, rays))


def test_config4_region(sk):
    """Config 4 (16 instances of a 1.09 M-triangle mesh + floor, plastic, GridLight with 16 samples — 16 shadow rays per
    hit through the wavefront, TLAS with inner nodes): 8 tiles against the oracle."""
    desc = workloads.instanced_blobs(n=740)
    e, rays = region_parity(sk, desc, block_of_tiles(desc, 26, 16, 4, 2))
    print("config 4, 8 tiles: rmse %.3g, %d rays" % (e, rays))
