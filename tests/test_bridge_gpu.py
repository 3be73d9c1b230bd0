"""Integration A on a B200: the UNMODIFIED reference binary `scene` and the UNMODIFIED shader / procedure plugin DSOs
(ConstantShader.so, PlasticShader.so, PathtracingShader.so, GlassShader.so, StanfordPlyProcedure.so,
VelocityGeneratorProcedure.so — dlopen'ed by the reference's own SiOpenPlugin, src/fj_plugin.cc:28-70) running on the bridge
build of libscene.so with FJ_DEVICE=0: every golden `.scn` scene renders on libfjgpu and matches the reference's own .fb
(tests/golden/ref_images.npz); a plugin without a device kind, the adaptive sampler or an unset FJ_DEVICE leave the frame
to the reference's CPU workers.  `pytest -m gpu`."""
import os

import numpy as np
import pytest

import golden_scenes
import scenekit as sk
from test_bridge_cpu import BR, run_bridge_scene

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(BR, "bin", "scene")),
                                                  reason="bridge build absent (needs /root/reference at build time)")]
DEV = {"FJ_DEVICE": "0", "FJ_DEVICE_VERBOSE": "1"}


def rmse(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d * d).reshape(-1, a.shape[-1]).mean(0))


def golden(name):
    return np.load(os.path.join(sk.REPO, "tests", "golden", "ref_images.npz"))[name]


@pytest.mark.parametrize("name", golden_scenes.DETERMINISTIC)
def test_reference_binary_and_plugins_render_on_the_device(tmp_path, name):
    img, res = run_bridge_scene(golden_scenes.SCENES[name](), str(tmp_path), DEV)
    assert "# fjgpu bridge: device 0 rendered" in res.stderr, res.stderr[-1500:]      # no silent CPU frame
    ref = golden(name)
    assert img.shape == ref.shape
    assert rmse(img, ref).max() < 1e-4


@pytest.mark.parametrize("name", golden_scenes.STOCHASTIC)
def test_stochastic_scenes_on_the_device(tmp_path, name):
    img, res = run_bridge_scene(golden_scenes.SCENES[name](), str(tmp_path), DEV)
    assert "# fjgpu bridge: device 0 rendered" in res.stderr, res.stderr[-1500:]
    ref = golden(name).astype(np.float64)
    assert np.array_equal(img[..., 3] > 0, ref[..., 3] > 0)
    assert abs(img[..., :3].mean() - ref[..., :3].mean()) < 0.01 * max(ref[..., :3].mean(), 1e-3) + 1e-3


def test_bridge_frame_equals_the_host_mirrors(tmp_path):
    """The same scene through the reference's host (bridge) and through libfjscene (the host mirror) reaches libfjgpu as the
    same description: identical frames."""
    from fujiyama_renderer_b200 import device
    desc = golden_scenes.SCENES["multi"]()
    img, _ = run_bridge_scene(desc, str(tmp_path), DEV)
    st = desc.to_structs()
    dev = device.Device(0)
    dev.load_structs(st)
    ours, _ = dev.render(st["params"], desc.tiles())
    dev.close()
    assert np.abs(img - ours).max() <= 1e-6 * max(1.0, float(np.abs(ours).max()))      # .fb text keeps 6 significant digits


def test_fallbacks_live_in_the_reference(tmp_path):
    # FJ_DEVICE unset: the reference renders on its CPU workers, bit for bit its own frame
    img, res = run_bridge_scene(golden_scenes.SCENES["plastic"](), str(tmp_path / "a"), {})
    assert "fjgpu bridge" not in res.stderr and np.array_equal(img, golden("plastic"))
    # a shader plugin without a device kind (MaterialShader): CPU workers
    d = golden_scenes.SCENES["plastic"]()
    scn_dir = str(tmp_path / "b")
    os.makedirs(scn_dir, exist_ok=True)
    fb = os.path.join(scn_dir, "out.fb")
    txt = d.to_scn(scn_dir, fb, threads=2, plugin_dir=os.path.join(BR, "lib"))
    txt = txt.replace("OpenPlugin plastic_shader %s" % os.path.join(BR, "lib", "PlasticShader"),
                      "OpenPlugin plastic_shader %s" % os.path.join(BR, "lib", "MaterialShader"))
    lines = [l for l in txt.split("\n") if not (l.startswith("SetProperty") and " sh1 " in l)]
    import subprocess
    scn = os.path.join(scn_dir, "scene.scn")
    open(scn, "w").write("\n".join(lines))
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(BR, "lib"), **DEV)
    res = subprocess.run([os.path.join(BR, "bin", "scene"), scn], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert "without a device kind; rendering on the CPU workers" in res.stderr
    # the adaptive sampler: CPU workers
    d = golden_scenes.SCENES["cube_3x3"]()
    os.makedirs(str(tmp_path / "c"), exist_ok=True)
    txt = d.to_scn(str(tmp_path / "c"), os.path.join(str(tmp_path / "c"), "out.fb"), threads=2, plugin_dir=os.path.join(BR, "lib"))
    txt = txt.replace("RenderScene ren1", "SetProperty1 ren1 sampler_type 1\nRenderScene ren1")
    scn = os.path.join(str(tmp_path / "c"), "scene.scn")
    open(scn, "w").write(txt)
    res = subprocess.run([os.path.join(BR, "bin", "scene"), scn], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert "adaptive sampler; rendering on the CPU workers" in res.stderr
