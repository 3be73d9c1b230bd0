"""The shipped example scenes on the host side, without a GPU: every in-scope scene's command stream — the shipped `.scn`
files and the shipped `.py` scripts run UNCHANGED on the py3 `fujiyama` module (oracle/gen_shipped.py) — is accepted by
libfjscene's parser command for command and flattens to a device scene description; the unmodified reference renders the same
stream with the same stand-in assets (so the GPU comparison in tests/test_shipped_scenes_gpu.py has a reference frame)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import scenekit as sk
import shipped_scenes as sh

pytestmark = pytest.mark.skipif(not sh.streams(), reason="oracle/_ref/shipped absent (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def assets(tmp_path_factory):
    return sh.make_assets(str(tmp_path_factory.mktemp("shipped_assets")))


@pytest.fixture(scope="module")
def fuji():
    sk.pkg()
    import __graft_entry__ as entry
    entry.build_fjgpu(); entry.build_host()
    from fujiyama_renderer_b200 import fujiyama
    return fujiyama


def test_all_sixteen_in_scope_scenes_are_present():
    names = [os.path.basename(p) for p in sh.streams()]
    assert len(names) == 16 and "teapot_scn.scn" in names and "pathtracing_py.scn" in names and "mesh_velocity_blur_py.scn" in names


@pytest.mark.parametrize("path", sh.streams(), ids=[os.path.basename(p)[:-4] for p in sh.streams()])
def test_stream_parses_and_flattens(fuji, assets, tmp_path, path):
    text = sh.prepare(open(path).read(), assets, "/opt/fujiyama/lib", str(tmp_path / "out"))
    head = "\n".join(l for l in text.split("\n") if not l.startswith(("RenderScene", "SaveFrameBuffer"))) + "\n"
    with fuji.Session() as s:
        s.run(head)                                              # every command of the shipped scene is understood
        n = [C.c_int32() for _ in range(4)]
        assert s.lib.fjscene_flatten(C.c_long(s.id("ren1")), *[C.byref(x) for x in n]) == 0, s.lib.fjscene_last_message()
        ninst, nlights, nshaders, ntiles = (x.value for x in n)
        assert ninst >= 2 and nshaders >= 1 and ntiles == 5 * 4           # 160x120 in 32x32 tiles


def test_reference_renders_a_shipped_stream_with_the_stand_ins(assets, tmp_path):
    sk.pkg()
    from fujiyama_renderer_b200 import fbio
    path = [p for p in sh.streams() if p.endswith("teapot_scn.scn")][0]
    text = sh.prepare(open(path).read(), assets, os.path.join(sk.REF_DIR, "lib"), str(tmp_path / "ref"), res=(80, 60), spp=(1, 1), threads=4)
    scn = tmp_path / "s.scn"
    scn.write_text(text)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(sk.REF_DIR, "lib"))
    res = subprocess.run([os.path.join(sk.REF_DIR, "bin", "scene"), str(scn)], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    img = fbio.read_fb(str(tmp_path / "ref.fb"))
    assert img.shape == (60, 80, 4) and img[..., 3].min() == 1.0 and img[..., :3].std() > 0.01      # dome everywhere, an image in it
