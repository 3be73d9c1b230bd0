"""The shipped-scene runner on a B200 (the reference's tests/run_all_scenes.py:30-60): every in-scope
shipped scene — 3 `.scn` files and 13 `.py` scripts executed unchanged on the py3 `fujiyama` module — with seeded synthetic
stand-in assets at 160x120, rendered by (a) the unmodified reference binary on the host cores, (b) libfjscene -> libfjgpu,
(c) the bridge build (the reference's own host + plugin DSOs on libfjgpu), and compared: deterministic scenes < 1e-4 per-channel
RMSE against the reference's .fb; scenes with grid / sphere lights or the path tracer (per-thread XorShift streams no parallel
renderer can reproduce) exact coverage + image mean.  `pytest -m gpu`."""
import os
import subprocess

import numpy as np
import pytest

import scenekit as sk
import shipped_scenes as sh
from test_bridge_cpu import BR

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not sh.streams(), reason="oracle/_ref/shipped absent (needs /root/reference at build time)")]


@pytest.fixture(scope="module")
def assets(tmp_path_factory):
    return sh.make_assets(str(tmp_path_factory.mktemp("shipped_assets")))


def rmse(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d * d).reshape(-1, a.shape[-1]).mean(0))


def run_binary(root, text, workdir, env_extra=None):
    from fujiyama_renderer_b200 import fbio
    scn = os.path.join(workdir, "scene.scn")
    with open(scn, "w") as f:
        f.write(text)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(root, "lib"))
    env.pop("FJ_DEVICE", None)
    env.update(env_extra or {})
    res = subprocess.run([os.path.join(root, "bin", "scene"), scn], env=env, capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    return res


@pytest.mark.parametrize("path", sh.streams(), ids=[os.path.basename(p)[:-4] for p in sh.streams()])
def test_shipped_scene_on_both_builds(assets, tmp_path, path):
    sk.pkg()
    from fujiyama_renderer_b200 import fbio, fujiyama
    raw = open(path).read()
    threads = min(os.cpu_count() or 1, 32)
    # (a) the unmodified reference
    wa = str(tmp_path / "ref"); os.makedirs(wa)
    run_binary(sk.REF_DIR, sh.prepare(raw, assets, os.path.join(sk.REF_DIR, "lib"), os.path.join(wa, "out"), threads=threads), wa)
    ref = fbio.read_fb(os.path.join(wa, "out.fb"))
    # (b) libfjscene -> libfjgpu
    wb = str(tmp_path / "gpu"); os.makedirs(wb)
    with fujiyama.Session() as s:
        s.run(sh.prepare(raw, assets, "/opt/fujiyama/lib", os.path.join(wb, "out")))
        stats, _, _ = s.stats()
    img = fbio.read_fb(os.path.join(wb, "out.fb"))
    assert stats.rays_camera > 0 and img.shape == ref.shape == (120, 160, 4)
    # (c) the bridge: the reference's own host and plugin DSOs on libfjgpu
    imgs = {"libfjscene": img}
    if os.path.exists(os.path.join(BR, "bin", "scene")):
        wc = str(tmp_path / "bridge"); os.makedirs(wc)
        res = run_binary(BR, sh.prepare(raw, assets, os.path.join(BR, "lib"), os.path.join(wc, "out"), threads=threads), wc,
                         {"FJ_DEVICE": "0", "FJ_DEVICE_VERBOSE": "1"})
        assert "# fjgpu bridge: device 0 rendered" in res.stderr, res.stderr[-1500:]
        imgs["bridge"] = fbio.read_fb(os.path.join(wc, "out.fb"))
    for who, im in imgs.items():
        if sh.is_stochastic(raw):
            assert np.array_equal(im[..., 3] > 0, ref[..., 3] > 0), who
            m, mr = float(im[..., :3].mean()), float(ref[..., :3].mean())
            assert abs(m - mr) < 0.03 * max(mr, 1e-3) + 2e-3, (who, m, mr)
        else:
            e = rmse(im, ref)
            assert e.max() < 1e-4, (who, e)
    if "bridge" in imgs:                # both hosts hand libfjgpu the same scene: the same frame (to the .fb's six digits)
        assert np.abs(imgs["bridge"] - img).max() <= 2e-6 * max(1.0, float(np.abs(img).max()))
